#!/bin/bash
# host topology of the GPU box (for the N>1 host-path affinity): NUMA nodes, cpuset, GPU <-> CPU affinity
mkdir -p gpurun_out
{
  echo "== nproc / affinity"; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
  echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|CPU\(s\)"
  echo "== numa nodes"; ls /sys/devices/system/node/ | grep node; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
  echo "== cpuset"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; cat /proc/self/status | grep -E "Cpus_allowed_list|Mems_allowed_list"
  echo "== nvidia-smi topo"; nvidia-smi topo -m
  echo "== gpu pci numa"; for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo $d $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) $(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null) $(cat /sys/bus/pci/devices/$b/current_link_speed 2>/dev/null) x$(cat /sys/bus/pci/devices/$b/current_link_width 2>/dev/null); done
  echo "== mem"; free -g | head -2
} > gpurun_out/topo.txt 2>&1
cat gpurun_out/topo.txt
