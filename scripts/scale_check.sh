#!/bin/bash
# Multi-GPU checks on a box with N GPUs (default 8): NCCL parity, bench, biobank shape, in-process driver.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 500 $TR --master-port 29541 scripts/dist_check.py 2>&1 | tail -2
timeout 400 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2600 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 900 $TR --master-port 29543 scripts/biobank_check.py --indiv 50000 --loci 500000 2> gpurun_out/biobank_n$N.err | tail -1 | tee gpurun_out/biobank_n$N.json
tail -3 gpurun_out/biobank_n$N.err
timeout 600 python scripts/multi_check.py 6000 200000 2> gpurun_out/multi_n$N.err | tail -1 | tee gpurun_out/multi_n$N.json
tail -3 gpurun_out/multi_n$N.err
