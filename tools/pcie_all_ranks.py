"""Run under torchrun: every rank copies pinned host <-> its GPU at the same time; prints per-rank and total GB/s.
Tells whether the host-resident (e2e) rate at N GPUs is limited by the box's PCIe / memory fabric."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_emulator_b200 import sharding
cpus = None if os.environ.get("GPE_NO_NUMA_BIND") else sharding.bind_host_to_gpu(lr)
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
NB = 512 << 20
hin = torch.empty(NB // 8, dtype=torch.float64).pin_memory(); hout = torch.empty(NB // 8, dtype=torch.float64).pin_memory()
din = torch.empty(NB // 8, dtype=torch.float64, device="cuda"); dout = torch.empty(NB // 8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    return reps * NB / (time.perf_counter() - t0) / 1e9
for name, a, b in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
    run(a, b, 2)
    r = run(a, b)
    t = torch.tensor([r], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        vals = [float(x) for x in lst]
    else:
        vals = [r]
    if rank == 0:
        print("%-9s per-direction GB/s per rank: %s | sum %.1f%s" % (name, " ".join("%.1f" % v for v in vals), sum(vals),
              " (x2 directions)" if name == "both" else ""), flush=True)
if rank == 0:
    print("cpus per rank after binding:", None if cpus is None else len(cpus), "of", os.cpu_count(), flush=True)
    os.system("nvidia-smi topo -m | head -14; lscpu | grep -i 'numa\\|socket\\|^CPU(s)'")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
