#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
nvidia-smi topo -m
echo ---; lscpu | head -30
echo ---; cat /proc/self/status | grep -i "allowed"
echo ---; ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo ---; python - <<'PY'
import os, pynvml
pynvml.nvmlInit()
print("affinity", sorted(os.sched_getaffinity(0)))
for i in range(pynvml.nvmlDeviceGetCount()):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        print(i, pynvml.nvmlDeviceGetPciInfo(h).busId, [hex(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, 8)], "numa", open(f"/sys/bus/pci/devices/{pynvml.nvmlDeviceGetPciInfo(h).busId.lower()[4:]}/numa_node").read().strip())
    except Exception as e:
        print(i, "ERR", e)
PY
} > gpurun_out/topo.log 2>&1
echo done
