"""Host topology of the GPU box: NUMA nodes, the node of every GPU's PCIe function, this process's CPU set, and whether the
memory-policy system calls are permitted here (development aid for the multi-GPU host-call numbers)."""
import ctypes
import glob
import os
import subprocess

print("nodes online:", open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else "n/a")
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        cpus = open(n + "/cpulist").read().strip()
        mem = [l for l in open(n + "/meminfo") if "MemTotal" in l or "MemFree" in l]
        print(os.path.basename(n), "cpus", cpus, " ".join(" ".join(l.split()[2:]) for l in mem))
    except OSError as e:
        print(n, e)
print("cpu affinity of this process:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus")
try:
    print(open("/proc/self/status").read().split("Mems_allowed_list:")[1].splitlines()[0].strip(), "= Mems_allowed_list")
except Exception as e:
    print("mems_allowed:", e)
out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
for line in out.strip().splitlines():
    idx, bus = [x.strip() for x in line.split(",")]
    path = "/sys/bus/pci/devices/" + bus.lower()[4:] + "/numa_node"
    try:
        print("gpu", idx, bus, "numa_node", open(path).read().strip())
    except OSError as e:
        print("gpu", idx, bus, "numa_node unreadable:", e)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy, SYS_get_mempolicy = 238, 239
mode = ctypes.c_int(-1)
rc = libc.syscall(SYS_get_mempolicy, ctypes.byref(mode), None, 0, None, 0)
print("get_mempolicy rc", rc, "errno", ctypes.get_errno(), "mode", mode.value)
mask = ctypes.c_ulong(1)
rc = libc.syscall(SYS_set_mempolicy, 1, ctypes.byref(mask), 64)       # MPOL_PREFERRED, node 0
print("set_mempolicy(PREFERRED, node 0) rc", rc, "errno", ctypes.get_errno())
rc = libc.syscall(SYS_set_mempolicy, 0, None, 0)
print("set_mempolicy(DEFAULT) rc", rc, "errno", ctypes.get_errno())
