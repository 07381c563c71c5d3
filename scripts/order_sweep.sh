#!/bin/bash
# A/B of the Gram work-list order at C4 on one GPU (biobank_check: one job, no warm-up; compare gram ms)
for v in "1 4 4" "0 4 4" "0 8 8" "0 8 16" "0 16 8" "0 12 12"; do
  set -- $v
  echo "order=$1 l2=$2x$3 mxf4: $(TPG_GRAM_ORDER=$1 TPG_GRAM_L2_ROWS=$2 TPG_GRAM_L2_COLS=$3 python scripts/biobank_check.py --indiv 50000 --loci 500000 --pairs 8 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms"],1), d["stage_ms_rank0"], d["ok"])')"
done
for v in "1 4 4" "0 4 4" "0 8 8" "0 6 6"; do
  set -- $v
  echo "order=$1 l2=$2x$3 int8: $(TPG_GRAM_I8=1 TPG_GRAM_ORDER=$1 TPG_GRAM_L2_ROWS=$2 TPG_GRAM_L2_COLS=$3 python scripts/biobank_check.py --indiv 50000 --loci 500000 --pairs 8 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms"],1), d["stage_ms_rank0"], d["ok"])')"
done
