#!/bin/bash
# what the GPU box's host looks like (cores, NUMA, memory, PCIe, disks) -> gpurun_out/box.txt
mkdir -p gpurun_out
{
lscpu | head -40
echo ---; numactl -H 2>&1 | head -30
echo ---; free -g
echo ---; df -h /tmp /dev/shm /root 2>&1
echo ---; nvidia-smi topo -m 2>&1 | head -40
echo ---; nvidia-smi -q -d PCIE 2>/dev/null | grep -i -A3 "link\|gen" | head -40
echo ---; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/memory.max 2>/dev/null
echo ---; nproc; python -c "import os; print(os.cpu_count(), len(os.sched_getaffinity(0)))"
echo ---; ls /sys/devices/system/node/ 2>&1; for n in /sys/devices/system/node/node*; do echo $n; cat $n/cpulist; done
echo ---; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x0302 $d/class 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
echo ---; ulimit -a
} > gpurun_out/box.txt 2>&1
