"""CPU suite: the N > 1 path of bench.py (replica sharding of independent frame pairs + max-over-ranks timing) with
world_size 2 on gloo. No collective touches the data path; the only exchange is the timing reduction."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, l = bench.dist_setup(world)
    lo, hi = bench.shard(n_total, w, r)
    # each rank "processes" its block; pretend rank 1 is slower: the job time is the max over ranks
    t = torch.tensor([10.0 + 5.0 * r], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(units)
    q.put((r, lo, hi, float(t.item()), int(units.item())))
    dist.destroy_process_group()


def test_shard_partition_properties():
    import bench
    for n in (4096, 4097, 7, 1, 0):
        for world in (1, 2, 4, 8):
            blocks = [bench.shard(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_and_max_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, 4097, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, u0), (r1, lo1, hi1, t1, u1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 2049, 2049, 4097)
    assert t0 == t1 == 15.0 and u0 == u1 == 4097


def test_circle_offsets_match_cv2_golden(golden):
    """workload.select_features paints OpenCV's filled midpoint circle; its span table must equal the cv2 golden discs."""
    from dsdtm_b200 import workload as W
    g = golden["circle_cv2"]
    H, Wd = g["shape"]
    for (cx, cy, r), bits in zip(g["cases"], g["masks"]):
        m = np.zeros((H, Wd), bool)
        for dy, hw in W.circle_mask_offsets(int(r)).items():
            y = cy + dy
            if 0 <= y < H and cx + hw >= 0 and cx - hw < Wd:
                m[y, max(cx - hw, 0):min(cx + hw, Wd - 1) + 1] = True
        assert (m == np.unpackbits(bits)[:H * Wd].reshape(H, Wd).astype(bool)).all(), (cx, cy, r)
