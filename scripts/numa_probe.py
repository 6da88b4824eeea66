"""Developer probe: NUMA placement of pinned host memory vs H2D bandwidth (run on a GPU box)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)'; nproc; cat /proc/self/status | grep -i allowed", shell=True, capture_output=True, text=True).stdout)
dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(dev)
bus = torch.cuda.get_device_properties(dev)
pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
node = open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip() if os.path.exists(f"/sys/bus/pci/devices/{pci}/numa_node") else "?"
print("gpu", dev, "pci", pci, "numa_node", node, "affinity now", len(os.sched_getaffinity(0)), "cpus")

def h2d_gbs(nbytes=1 << 30, reps=5):
    h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    d = torch.empty_like(h, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

print("H2D default placement: %.1f GB/s" % h2d_gbs())
nodes = sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node"))
for n in nodes:
    cl = open(f"/sys/devices/system/node/{n}/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        if not part: continue
        a, _, b = part.partition("-"); cpus.update(range(int(a), int(b or a) + 1))
    allowed = cpus & os.sched_getaffinity(0)
    if not allowed:
        print(n, "no allowed cpus"); continue
    old = os.sched_getaffinity(0)
    os.sched_setaffinity(0, allowed)
    print(f"H2D with the process bound to {n} ({len(allowed)} cpus) before allocating: %.1f GB/s" % h2d_gbs())
    os.sched_setaffinity(0, old)
