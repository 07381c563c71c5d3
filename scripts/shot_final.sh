#!/bin/bash
# Round-end evidence in one GPU call: whole GPU suite, smoke, bench (both routes, reference arm),
# ncu launch list and one --set full capture of the step's three kernels.
mkdir -p gpurun_out
t0=$SECONDS
timeout 240 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/t_full.log 2>&1
echo "full suite exit $? at $((SECONDS-t0)) s"; tail -14 gpurun_out/t_full.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? at $((SECONDS-t0)) s"
timeout 170 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $? at $((SECONDS-t0)) s"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm exit $? at $((SECONDS-t0)) s"
TPG_GRAM_I8=1 timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming > gpurun_out/bench_n1_int8.json 2> gpurun_out/bench_n1_int8.err
echo "bench int8 exit $? at $((SECONDS-t0)) s"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-streaming > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $? at $((SECONDS-t0)) s"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'decode_planes_k4_kernel|gram_mxf4_pair_kernel|finish_kernel' -c 3 -f -o gpurun_out/step_full \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-streaming > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $? at $((SECONDS-t0)) s"
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_int8", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("stage_ms_per_step"), d.get("e2e", {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_decode") or {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
