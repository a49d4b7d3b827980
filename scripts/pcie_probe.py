"""Plain pinned H2D / D2H copies on N ranks at once: what N concurrent cudaMemcpyAsync streams reach on this box,
the ceiling dsdtm_pair_batch_e2e is compared with. Run alone (N = 1) or under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/pcie_probe.py [--bind]

Each rank copies a 1 GiB pinned buffer to / from its GPU 8 times; all ranks start together (barrier); the aggregate is total bytes over
the slowest rank's time. --bind pins each rank to its GPU's NUMA node before the pinned allocation (bench.py does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bench.bind_to_gpu_numa(local) if "--bind" in sys.argv else {"bound": False}
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    res = {}
    for name, f in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        f(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(8):
            f()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res[name] = (8 * n / dt / 1e9, world * 8 * n / float(tt.item()) / 1e9)
    nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
    print("rank %d/%d gpu %d numa %s | H2D %.1f GB/s D2H %.1f GB/s (this rank)%s" % (rank, world, local, numa, res["H2D"][0], res["D2H"][0],
          (" | aggregate H2D %.1f GB/s D2H %.1f GB/s over %d ranks; host cpus %d, numa nodes %s" % (res["H2D"][1], res["D2H"][1], world, os.cpu_count(), nodes)) if rank == 0 else ""), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
