"""world_size-2 test of the multi-GPU host logic on the CPU (gloo): point-sharded MSM with all-gathered partial sums and
proof-sharded batch verification with all-gathered verdicts.  The per-rank compute is played by the oracle
here (no GPU in this container); on the GPU box the same functions run over NCCL in bench.py."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from bulletproofs_amcl_b200 import sharding
    from oracle.curves import BLS12_381 as C
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 37
        G = C.from_affine(C.g)
        pts = [C.mul(G, 3 + i) for i in range(n)]
        sc = C.synth_scalars(21, n)
        lo, hi = sharding.shard_bounds(n, world, rank)
        partial = C.g1_xy_bytes(C.msm(pts[lo:hi], sc[lo:hi]))
        parts = sharding.allgather_bytes(dist, partial)
        total = C.INF
        for p in parts:
            total = C.add(total, C.g1_from_xy_bytes(p))
        ok_msm = C.g1_xy_bytes(total) == C.g1_xy_bytes(C.msm(pts, sc)) and len(parts) == world
        # the product's own combine step (host-side bph_g1_sum; needs no GPU) gives the same bytes on every rank
        ok_msm = ok_msm and sharding.combine_partials(0, parts) == C.g1_xy_bytes(total)
        # proof-sharded verdicts: rank 1 holds the failing proof
        count = 6
        lo, hi = sharding.shard_bounds(count, world, rank)
        local = [(-4 if i == 4 else 0) for i in range(lo, hi)]
        verdicts = sharding.sharded_verdicts(dist, local)
        ok_v = verdicts == [0, 0, 0, 0, -4, 0]
        q.put((rank, ok_msm, ok_v))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from bulletproofs_amcl_b200.sharding import shard_bounds
    for total in (0, 1, 7, 8, 4096, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            cuts = [shard_bounds(total, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
