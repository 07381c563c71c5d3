#!/bin/bash
# Round-2 profiler shots (one gpurun call, one GPU).  Every ncu run follows a plain run of the same command.
#   1. launch list (+ DRAM bytes) of the bench's device-timed part at the full C4 shape
#   2. ncu --set full of decode_planes_k4 / gram_mxf4_wide / finish at a C4-like shape whose memory ncu can save and restore
#   3. ncu --set full of gather_pack / packed decode (C5 subset) and of the table-free scan (C2)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
$B > gpurun_out/r02_prof_plain_c4.json 2> gpurun_out/r02_prof_plain_c4.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:'decode_planes|gram_|finish_kernel|zero_counts' --csv --log-file gpurun_out/r02_launches_c4.csv $B > gpurun_out/r02_ncu1.log 2>&1
echo "launch list rc=$?"
D="python bench.py --indiv 20000 --loci 131072 --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline"
$D > gpurun_out/r02_prof_plain_dev.json 2> gpurun_out/r02_prof_plain_dev.err &&
ncu --set full --clock-control none --import-source on -k regex:'decode_planes_k4_kernel|gram_mxf4_wide_kernel|finish_kernel|gram_i8_pair_kernel' \
    -s 3 -c 8 -o gpurun_out/r02_c4like_full -f $D > gpurun_out/r02_ncu2.log 2>&1
echo "c4-like full rc=$?"
G="python scripts/configs_bench.py"
$G > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err &&
ncu --set full --clock-control none --import-source on -k regex:'gather_pack_kernel|decode_planes_k4_kernel<1>|decode_planes_k4_kernel<\(bool\)1>' \
    -c 2 -o gpurun_out/r02_gather_full -f $G > gpurun_out/r02_ncu3.log 2>&1
echo "gather full rc=$?"
S="python scripts/consumers_bench.py"
$S > gpurun_out/r02_consumers.jsonl 2> gpurun_out/r02_consumers.err &&
ncu --set full --clock-control none --import-source on -k regex:'scan_counts' -s 3 -c 1 -o gpurun_out/r02_scan_full -f $S > gpurun_out/r02_ncu4.log 2>&1
echo "scan full rc=$?"
ls -la gpurun_out/*.ncu-rep
