#!/bin/bash
# Round 2, call 1 (1 GPU): every GPU test with the round-1 opt-in gates open, plus facts about the box (NUMA, caps, mbind).
mkdir -p gpurun_out; O=gpurun_out/r2c1
export B200COLL_TIMEOUT_MS=5000
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; grep -i cap /proc/self/status; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc; free -g | head -2; } > ${O}_box.txt 2>&1
python - > ${O}_mbind.txt 2>&1 <<'PY'
import ctypes, os, mmap
libc = ctypes.CDLL(None, use_errno=True)
n = 1 << 22
buf = mmap.mmap(-1, n)
addr = ctypes.addressof(ctypes.c_char.from_buffer(buf))
mask = ctypes.c_ulong(1)
r = libc.syscall(237, ctypes.c_void_p(addr), ctypes.c_ulong(n), 2, ctypes.byref(mask), ctypes.c_ulong(64), 0)   # mbind(MPOL_BIND, node 0)
print("mbind rc", r, "errno", ctypes.get_errno())
print("affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
try:
    os.sched_setaffinity(0, {0}); print("setaffinity ok"); 
except Exception as e: print("setaffinity failed", e)
for d in sorted(os.listdir("/sys/bus/pci/devices")):
    p = f"/sys/bus/pci/devices/{d}"
    try:
        if open(p + "/vendor").read().strip() == "0x10de" and open(p + "/class").read().startswith("0x0302"):
            print(d, "numa", open(p + "/numa_node").read().strip(), "cpus", open(p + "/local_cpulist").read().strip())
    except Exception as e: print(d, e)
PY
B200_RUN_UNVALIDATED=1 timeout 1000 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 40 ${O}_pytest.log
B200_RUN_FAULT_INJECTION=1 timeout 200 python -m pytest tests/test_zz_tools_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider > ${O}_pytest_tools.log 2>&1; echo "tools rc=$?"; tail -n 15 ${O}_pytest_tools.log
nvidia-smi --query-gpu=index,name,mig.mode.current,persistence_mode --format=csv > ${O}_smi.txt 2>&1
