"""Multi-GPU plumbing of the two paths that shard: one process per GPU, torch.distributed for the
exchange step (NCCL over NVLink on the GPU box; gloo in the CPU tests).

 * MSM shards by POINTS: rank g owns terms [lo_g, hi_g), computes one partial sum on its GPU; the affine partials
   (2*MODBYTES bytes each) are all-gathered and summed on every rank.  The affine form of the total is unique, so the
   result is bit-identical for every world size.
 * Batch verification shards by PROOFS: rank g verifies proofs [lo_g, hi_g); the per-proof verdict bytes are all-gathered.
   Verdicts stay per proof (the reference has no batch API: verifier.rs:267).
A single IPP / R1CS proof does not shard: replicas only.
"""


def shard_bounds(total, world, rank):
    """contiguous, balanced [lo, hi) of `total` units for `rank` of `world` (first total % world ranks get one more)"""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_bytes(dist, payload, device=None):
    """all-gather equal-length byte strings; returns the list ordered by rank.  `dist` is torch.distributed or None
    (single process)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(payload)]
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, t)
    raw = out.cpu().numpy().tobytes()
    k = len(payload)
    return [raw[i * k:(i + 1) * k] for i in range(world)]


def combine_partials(ctx, partials):
    """sum of the per-rank affine partial sums (2*MODBYTES bytes each): at most 8 additions, done on the host
    (bph_g1_sum) -- a device launch pipeline would cost more than the additions.  `ctx` is a Context or a curve id."""
    if len(partials) == 1:
        return partials[0]
    from .binding import g1_sum
    return g1_sum(ctx if isinstance(ctx, int) else ctx.curve, b"".join(partials))


def sharded_msm(ctx, dist, local_msm):
    """`local_msm()` returns this rank's partial sum (X||Y bytes); returns the total on every rank."""
    return combine_partials(ctx, allgather_bytes(dist, local_msm()))


def sharded_verdicts(dist, local_verdicts):
    """`local_verdicts`: this rank's list of int verdicts (0 = Ok, negative = error); ranks must hold equal counts.
    Returns the verdicts of all proofs in rank order."""
    payload = bytes((v & 0xFF) for v in local_verdicts)
    parts = allgather_bytes(dist, payload)
    return [b - 256 if b > 127 else b for p in parts for b in p]
