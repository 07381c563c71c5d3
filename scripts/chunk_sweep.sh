#!/bin/bash
# A/B: device-resident C2 job with decode(c+1) overlapping gram(c) at several chunk sizes.
for lc in 0 65536 131072 217088 325120; do
  if [ "$lc" = "0" ]; then unset TPG_DEVICE_CHUNK_LOCI; else export TPG_DEVICE_CHUNK_LOCI=$lc; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk_loci=$lc', 'ms_per_step=%.4f'%d['ms_per_step'], d['stage_ms_per_step'], 'e2e_ms=%.3f'%d['e2e']['ms_per_step'])
"
done
