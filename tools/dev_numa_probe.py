"""Developer probe: host topology of the GPU box and pinned host -> HBM copy bandwidth with the allocating thread bound to
each NUMA node in turn (is the e2e copy sensitive to where the pinned buffers live?)."""
import glob, os, subprocess, sys, time
import torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:
        return f"<{e}>"

print(sh("lscpu | grep -i -E 'model name|socket|numa|^cpu\\(s\\)|thread'"))
print(sh("nvidia-smi topo -m"))
print("affinity now:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("numa nodes:", nodes)
for d in range(torch.cuda.device_count()):
    bus = torch.cuda.get_device_properties(d).pci_bus_id if hasattr(torch.cuda.get_device_properties(d), "pci_bus_id") else None
    print("gpu", d, "bus", bus)
print(sh("for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ \"$(cat $d/class 2>/dev/null)\" = 0x030200 ]; then echo $d $(cat $f) $(cat $d/local_cpulist); fi; done"))
dev = torch.device("cuda:0")
dst = torch.empty(640 << 20, dtype=torch.uint8, device=dev)
full = os.sched_getaffinity(0)
for node in nodes:
    cpus = sh(f"cat /sys/devices/system/node/node{node}/cpulist")
    ids = set()
    for part in cpus.split(","):
        if "-" in part:
            a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
        elif part:
            ids.add(int(part))
    ids &= full
    if not ids:
        print("node", node, "no allowed cpus"); continue
    os.sched_setaffinity(0, ids)
    src = torch.empty(640 << 20, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"pinned on node {node} ({len(ids)} cpus): {src.numel() / dt / 1e9:.1f} GB/s", flush=True)
    del src
os.sched_setaffinity(0, full)
