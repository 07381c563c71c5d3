#!/bin/bash
# What the GPU box offers besides the GPU: host cores, RAM, scratch space, page cache behaviour.
{
echo "== nproc: $(nproc)"; lscpu | grep -E "Model name|Socket|Thread|Core|NUMA|L3" 
echo "== mem"; free -g
echo "== df"; df -h /tmp /dev/shm "$PWD" 2>/dev/null
echo "== ulimit -l: $(ulimit -l)"
echo "== gpus"; nvidia-smi --query-gpu=index,name,memory.total --format=csv
echo "== cgroup mem"; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/cpu.max 2>/dev/null
echo "== dd write 4G to /tmp"; dd if=/dev/zero of=/tmp/ddtest bs=1M count=4096 oflag=direct 2>&1 | tail -1
echo "== dd read direct"; dd if=/tmp/ddtest of=/dev/null bs=4M iflag=direct 2>&1 | tail -1
echo "== dd read cached"; dd if=/tmp/ddtest of=/dev/null bs=4M 2>&1 | tail -1; dd if=/tmp/ddtest of=/dev/null bs=4M 2>&1 | tail -1
rm -f /tmp/ddtest
mount | grep -E " /tmp | / " | head
} > gpurun_out/box_probe.txt 2>&1
