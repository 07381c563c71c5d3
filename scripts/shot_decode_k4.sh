#!/bin/bash
# One GPU call: parity of the rebuilt decode_planes_k4_kernel, bench A/B against the first build
# (TPG_DECODE_K4_V1=1), one ncu --set full capture of the new kernel, smoke, then the whole GPU suite.
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "decode_planes_k4 or raw_mxf4 or hgdp_shape" > gpurun_out/t_targeted.log 2>&1
echo "targeted exit $? at $((SECONDS-t0)) s"; tail -3 gpurun_out/t_targeted.log
timeout 170 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $? at $((SECONDS-t0)) s"
TPG_DECODE_K4_V1=1 timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming > gpurun_out/bench_n1_k4v1.json 2> gpurun_out/bench_n1_k4v1.err
echo "bench v1 exit $? at $((SECONDS-t0)) s"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:decode_planes_k4_kernel -c 1 -f -o gpurun_out/decode_k4 \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-streaming > gpurun_out/ncu.log 2>&1
echo "ncu exit $? at $((SECONDS-t0)) s"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? at $((SECONDS-t0)) s"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout ${FULL_SUITE_TIMEOUT:-240} python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/t_full.log 2>&1
echo "full suite exit $? at $((SECONDS-t0)) s"; tail -20 gpurun_out/t_full.log
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_k4v1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("stage_ms_per_step"), d.get("e2e", {}).get("ms_per_step"), d["roofline_decode"]["frac"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
