"""Does NUMA placement of the pinned host buffers bound the host<->device link?  Topology of the
box, then H2D / D2H / duplex copy bandwidth with the process (and so its first-touch pinned pages)
bound to each NUMA node in turn."""
import glob, os, subprocess, sys, time
import torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"<{e}>"

print(sh("nvidia-smi topo -m | head -20"))
print("nodes:", sh("ls -d /sys/devices/system/node/node* | xargs -n1 basename | tr '\\n' ' '"))
for nd in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(os.path.basename(nd), "cpus", open(nd + "/cpulist").read().strip(), "| mem", sh(f"grep MemTotal {nd}/meminfo"))
print("affinity now:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus")
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
print("gpu0 pci:", sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0"))
pci = sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0").lower()
if pci.startswith("00000000:"):
    pci = "0000:" + pci.split(":", 1)[1]
print("gpu0 numa_node:", sh(f"cat /sys/bus/pci/devices/{pci}/numa_node"))
print("mempolicy tools:", sh("which numactl"), "| cgroup cpuset:", sh("cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null"), "| mems:", sh("cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null"))

dev = torch.device("cuda", 0)
torch.cuda.init()
n = 64 << 20
d_in, d_out = torch.empty(n, device=dev), torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
by = 4 * n


def measure(tag):
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    h_in.fill_(1.0); h_out.fill_(0.0)                       # touch

    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    a = by / timed(lambda: d_in.copy_(h_in, non_blocking=True)) / 1e9
    b = by / timed(lambda: h_out.copy_(d_out, non_blocking=True)) / 1e9
    c = by / timed(both) / 1e9
    print(f"{tag}: h2d {a:.1f}  d2h {b:.1f}  duplex each way {c:.1f} GB/s", flush=True)


all_cpus = sorted(os.sched_getaffinity(0))
measure("default affinity")
for nd in sorted(glob.glob("/sys/devices/system/node/node*")):
    cl = open(nd + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        if "-" in part:
            lo, hi = part.split("-"); cpus |= set(range(int(lo), int(hi) + 1))
        elif part:
            cpus.add(int(part))
    cpus &= set(all_cpus)
    if not cpus:
        print(os.path.basename(nd), "no allowed cpus"); continue
    os.sched_setaffinity(0, cpus)
    measure(f"bound to {os.path.basename(nd)} ({len(cpus)} cpus)")
os.sched_setaffinity(0, all_cpus)
