#!/bin/bash
# e2e (host bytes -> host KING) at several host chunk sizes, with and without the hybrid
# link/pack schedule and with packing off (round-1 behaviour).
run() {
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['streaming'] or {}
print('$1', 'e2e_ms=%.3f'%d['e2e']['ms_per_step'], 'match', d['e2e']['matches_device_run'], 'file_ms=%.2f'%s.get('ms_total',0), 'staging=%.2f h2d=%.2f chunks=%s match=%s'%(s.get('ms_staging_sum',0), s.get('ms_h2d_sum',0), s.get('chunks'), s.get('matches_device_run')))
"
}
for mb in 16 32 48 96; do TPG_HOST_CHUNK_MB=$mb run "hybrid chunk_mb=$mb"; done
TPG_NO_HYBRID=1 run "pack-only chunk_mb=48"
TPG_NO_PACK=1 run "no-pack chunk_mb=48"
TPG_NO_PACK=1 TPG_HOST_CHUNK_MB=256 run "no-pack chunk_mb=256 (round-1 setting)"
