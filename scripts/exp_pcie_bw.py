import torch, time, os
pr = torch.cuda.get_device_properties(0)
pci = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
cpus=set()
for part in open("/sys/bus/pci/devices/%s/local_cpulist" % pci).read().strip().split(","):
    lo,_,hi=part.partition("-"); cpus.update(range(int(lo), int(hi or lo)+1))
os.sched_setaffinity(0, cpus)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (src, dst) in {"H2D": (h, d), "D2H": (d, h)}.items():
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, "%.1f GB/s" % (n/dt/1e9))
