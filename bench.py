#!/usr/bin/env python3
"""bench.py — headline benchmark: BLS12-381 G1 MSM points/s at 2^20 points (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--lg 20] [--impl ours|reference]

One "step" = one multi-scalar multiplication sum_i s_i*P_i over `2^lg` synthetic points per GPU
(what G1Vector::multi_scalar_mul_var_time / inner_product_var_time_with_ref_vecs compute:
/root/reference/src/ipp.rs:251-253, src/r1cs/verifier.rs:451), through libbpgpu's C ABI.

 * value  : whole-job points/s with points and scalars already resident in HBM (bpgpu_msm_device).
 * e2e    : the same metric through bpgpu_msm_refs with HOST (pinned) point and scalar buffers:
            H2D of both inputs every step, D2H of the per-window sums, host finish.
 * N > 1  : one process per GPU (torchrun).  The MSM shards by points: every rank computes the
            partial sum of its own 2^lg points (weak scaling), the 96-byte affine partials are
            all-gathered over NCCL and summed on every rank.  Timed with barrier + sync on both
            sides, max over ranks.
 * roofline: the dominant kernel (k_chunk_acc, bucket accumulation) against the INTEGER pipe:
            algorithmic IMADs = points * windows * 10 Fq-mul * (2*12^2+12) per launch, divided by
            the kernel's average duration (CUDA events on the ctx stream, inside the timed
            region); peak = IMAD.WIDE.U32 throughput measured in this same run by
            bpgpu_int_pipe_bench (MEASURED_PEAKS.json carries no integer peak).  The HBM view of the
            same kernel is reported next to it.
 * cpu_baseline / --impl reference: the reference's own CPU algorithm (Straus, wNAF-5) restated in
            C (oracle/c), timed on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "G1 MSM points/s at 2^20 (BLS12-381)"
UNIT = "points/s"
FQ_LIMBS = 12
IMAD_PER_FQMUL = 2 * FQ_LIMBS * FQ_LIMBS + FQ_LIMBS      # 300, SURVEY.md 8d
FQMUL_PER_MADD = 10                                      # XYZZ mixed add 8M + 2S


def synth_scalars_be(n, seed, modbytes=48):
    """n uniform 254-bit scalars (< r for BLS12-381), big-endian MODBYTES each."""
    import numpy as np
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 0] &= 0x3F
    out = np.zeros((n, modbytes), dtype=np.uint8)
    out[:, modbytes - 32:] = raw
    return out


BLS_R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def make_config(lg, world):
    """the workload description both arms print (the driver compares it between the arms)"""
    import bulletproofs_amcl_b200 as bp
    n = 1 << lg
    c = bp.lib().bpgpu_msm_window_bits(n)
    W = (256 + c - 1) // c
    return {"workload": f"G1 MSM, BLS12-381, 2^{lg} random points x 255-bit scalars per GPU, general bases "
                        f"(no precomputation), signed {c}-bit windows x {W}",
            "l2": f"2 input sets alternated (2x{(n * (2 * 48 + 32)) >> 20} MiB > 126 MB L2), no explicit flush",
            "parallelism": f"points sharded over {world} GPU(s); partial sums all-gathered over NCCL" if world > 1 else "1 GPU"}


def be_ints(block, n, mb):
    """n big-endian MODBYTES scalars (bytes or numpy block) -> Python ints"""
    b = block if isinstance(block, (bytes, bytearray)) else block.tobytes()
    return [int.from_bytes(b[i * mb:(i + 1) * mb], "big") for i in range(n)]


def expected_point(total_scalar):
    """(sum s_i * k_i mod r) * G by the ORACLE (checker only): what an MSM over the bench's points k_i * G must equal"""
    from oracle.curves import BLS12_381 as C
    return C.g1_xy_bytes(C.mul(C.from_affine(C.g), total_scalar % C.r))


def gen_points(ctx, n, seed, with_multipliers=False):
    """n synthetic G1 points k_i*G, computed on the device (selftest_group op 2)."""
    from bulletproofs_amcl_b200 import lib
    mb = ctx.modbytes
    gx = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
    gy = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
    g = gx.to_bytes(mb, "big") + gy.to_bytes(mb, "big")
    ks = synth_scalars_be(n, seed, mb).tobytes()
    xy = ctx.selftest_group(2, g * n, g * n, ks)
    return (xy, ks) if with_multipliers else xy


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every ~5 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception as e:           # pragma: no cover
            self.err = str(e)
            self.th = None

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if not self.th:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.th.join(timeout=1)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in self.samples))
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(s[1] for s in self.samples) if sm else None,
                "reasons": reasons, "samples": len(sm)}


TRAFFIC_CAPTURES = ["k_chunk_acc_r02_raw.csv", "k_chunk_acc_r01_raw.csv"]


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_chunk_acc per launch, READ FROM A STORED `ncu --set full` capture of
    this same command (profiles/, newest round first) -- not measured in this run; (None, None) if no capture is there."""
    import csv
    path = next((os.path.join(ROOT, "profiles", f) for f in TRAFFIC_CAPTURES if os.path.exists(os.path.join(ROOT, "profiles", f))), "")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(vals[i]) * scale[units[i]]
        return tot, "profiles/" + os.path.basename(path)
    except Exception:
        return None, None


def proof_section(bp, ctx, local, rank, world, dist, torch, no_cpu=False):
    """IPP / R1CS prove+verify proofs/s (BASELINE.json metric, second half) through the host layer (include/bphost.h).
    Independent proofs are spread over `nctx` contexts (one host thread + one CUDA stream each) of this rank's GPU; with N
    ranks the batch-verification job is sharded by proofs and the verdict bytes are all-gathered over NCCL."""
    import numpy as np
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # one proof per context is round-synchronous (a host transcript step per IPP round): the GPU is kept busy by contexts, not
    # by cores.  With a core per context the waits spin; with fewer cores than that (8 ranks on a 16-core box) the contexts
    # sleep on blocking events instead (bpgpu_ctx_set_blocking_sync) and 16 of them share this rank's cores.
    cores = max(1, ncpu // max(1, world))
    nctx = 16
    oversub = cores < nctx
    hthreads = max(1, ncpu // max(1, world))          # host threads of the batch calls: this rank's share of the cores
    out = {"contexts_per_gpu": nctx, "host_cpus": ncpu, "host_threads_per_rank": hthreads, "contexts_sleep_on_events": oversub}

    def run(curve, m, bits, count, tag, verify_reps=1, batch_call=False, cpu_base=0, weak=4096):
        n = m * bits
        # sleeping contexts pay ~30 us per wake-up: worth it for proofs of milliseconds (n >= 1024), not for the 1.8 ms of a
        # 64-bit proof with its ~30 waits -- those keep one spinning context per core
        sleepers = oversub and n >= 1024
        ctxs = [bp.Context(curve, local) for _ in range(nctx if (sleepers or not oversub) else cores)]
        for c in ctxs if sleepers else []:
            c.set_blocking_sync(True)
        c0 = ctxs[0]
        gx, hx = c0.g1_from_msg_hash(b"g"), c0.g1_from_msg_hash(b"h")
        pre = True                 # window tables in HBM (2 x 8.4 GB on BN254 at n = 2^14): table sums for the commitments and the first
                                   # IPP rounds, then the folded generators are materialised (csrc/ipp.cu hybrid): 223 -> 440 proofs/s
        G, H = c0.get_generators("G", n, precompute=pre), c0.get_generators("H", n, precompute=pre)
        rng = np.random.default_rng(4242 + rank)
        vals = [int(x) for x in rng.integers(0, 1 << 63, size=count * m, dtype=np.uint64)]
        bp.range_prove_many(ctxs, b"bench", gx, hx, G, H, vals[:m * min(count, 2 * nctx)], m, bits)      # warm-up
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        proofs, stride, comms = bp.range_prove_many(ctxs, b"bench", gx, hx, G, H, vals, m, bits)
        tp = time.perf_counter() - t0
        bp.range_verify_many(ctxs, b"bench", gx, hx, G, H, min(count, 2 * nctx), m, bits, proofs, stride, comms)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        ok = True
        for _ in range(verify_reps):
            v = bp.range_verify_many(ctxs, b"bench", gx, hx, G, H, count, m, bits, proofs, stride, comms)
            ok = ok and v == [0] * count
        tv = time.perf_counter() - t0
        tb = tpb = tb1 = twv = twp = None
        weak_n = 0
        if batch_call:
            # the same statement proved in lock-step on ONE context (bph_range_prove_batch): every prover stage and IPP round
            # is one device call for the whole batch, the transcripts run on the host threads in between
            nb = count * verify_reps
            bvals = [int(x) for x in rng.integers(0, 1 << 63, size=nb * m, dtype=np.uint64)]
            bp.range_prove_batch(c0, b"bench", gx, hx, G, H, bvals, m, bits, nthreads=hthreads)                # warm-up (scratch sized)
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            bproofs, bstride, bcomms = bp.range_prove_batch(c0, b"bench", gx, hx, G, H, bvals, m, bits, nthreads=hthreads)
            tpb = time.perf_counter() - t0
            vals_per_call = range(nb)
            v = bp.range_verify_batch(c0, b"bench", gx, hx, G, H, nb, m, bits, bproofs, bstride, bcomms, nthreads=hthreads)
            ok = ok and v == [0] * nb
        if batch_call:
            # config 5: the whole batch through bph_range_verify_batch: host threads build the scalars slab by slab, the
            # device evaluates every proof's verification MSM of a slab in one group of launches (one verdict byte per
            # proof), overlapped with the host work on the next slab
            reps = verify_reps
            big_p, big_c = proofs * reps, comms * reps
            bp.range_verify_batch(c0, b"bench", gx, hx, G, H, count * reps, m, bits, big_p, stride, big_c, nthreads=hthreads)   # warm-up
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            v = bp.range_verify_batch(c0, b"bench", gx, hx, G, H, count * reps, m, bits, big_p, stride, big_c, nthreads=hthreads)
            tb = time.perf_counter() - t0
            ok = ok and v == [0] * (count * reps)
            # the same call with the transcripts replayed on this rank's host threads (mode 1) instead of on the device
            bp.range_verify_batch(c0, b"bench", gx, hx, G, H, count * reps, m, bits, big_p, stride, big_c, nthreads=hthreads, mode=1)   # warm-up
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            v = bp.range_verify_batch(c0, b"bench", gx, hx, G, H, count * reps, m, bits, big_p, stride, big_c, nthreads=hthreads, mode=1)
            tb1 = time.perf_counter() - t0
            ok = ok and v == [0] * (count * reps)
            # WEAK scaling of the batch calls: `weak` proofs on EVERY GPU (config 5's literal 4096 proofs split 8 ways leave 512
            # per GPU, where the per-slab latency -- transcript threads, the 252-doubling Horner chains -- is all there is)
            wreps = max(1, weak // len(vals_per_call))
            wp, wc = bproofs[:len(vals_per_call) * bstride] * wreps, bcomms[:len(vals_per_call) * m * 2 * c0.modbytes] * wreps
            wn = len(vals_per_call) * wreps
            bp.range_verify_batch(c0, b"bench", gx, hx, G, H, wn, m, bits, wp, bstride, wc)
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            v = bp.range_verify_batch(c0, b"bench", gx, hx, G, H, wn, m, bits, wp, bstride, wc)
            twv = time.perf_counter() - t0
            ok = ok and v == [0] * wn
            wvals = [int(x) for x in rng.integers(0, 1 << 63, size=wn * m, dtype=np.uint64)]
            bp.range_prove_batch(c0, b"bench", gx, hx, G, H, wvals, m, bits)
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            bp.range_prove_batch(c0, b"bench", gx, hx, G, H, wvals, m, bits)
            twp = time.perf_counter() - t0
            weak_n = wn
        if dist is not None:
            # the exchange step of the sharded batch verification: verdict bytes of every rank, all-gathered
            t = torch.tensor([1 if ok else 0] * count, dtype=torch.uint8, device="cuda")
            allv = torch.empty(world * count, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allv, t)
            ok = bool(allv.min().item() == 1)
            tt = torch.tensor([tp, tv, tb or 0.0, tpb or 0.0, tb1 or 0.0, twv or 0.0, twp or 0.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            twv = float(tt[5].item()) if twv is not None else None
            twp = float(tt[6].item()) if twp is not None else None
            tp, tv = float(tt[0].item()), float(tt[1].item())
            tb = float(tt[2].item()) if tb is not None else None
            tpb = float(tt[3].item()) if tpb is not None else None
            tb1 = float(tt[4].item()) if tb1 is not None else None
        for c in ctxs:
            c.close()
        out[tag] = {"multipliers": n, "committed_values": m, "proofs": count * world, "prove_per_s": count * world / tp,
                    "verify_per_s": count * world * verify_reps / tv, "verifications": count * world * verify_reps,
                    "prove_verify_per_s": count * world / (tp + tv / verify_reps), "all_verified": ok, "proof_bytes": stride}
        if tb is not None:
            out[tag]["verify_batch_call_per_s"] = count * world * verify_reps / tb
            out[tag]["verify_batch_call"] = "bph_range_verify_batch: transcripts, per-proof scalars and MSMs on the device; host uploads bytes"
        if tb1 is not None:
            out[tag]["verify_batch_host_transcripts_per_s"] = count * world * verify_reps / tb1
        if twv is not None:
            out[tag]["weak_scaling_batch"] = {"proofs_per_gpu": weak_n, "verify_batch_per_s": weak_n * world / twv,
                                              "prove_batch_per_s": weak_n * world / twp}
        if tpb is not None:
            out[tag]["prove_batch_call_per_s"] = count * world * verify_reps / tpb
        if cpu_base and rank == 0 and world == 1:
            # the reference's algorithm for the same proof on the host cores (oracle/fast.py: ipp.rs / prover.rs /
            # verifier.rs restated, group operations in oracle/c), one independent proof stream per core
            import subprocess
            cname = "BLS12_381" if curve == bp.BLS12_381 else "BN254"
            r = subprocess.run([sys.executable, "-m", "oracle.fast", cname, str(m), str(bits), str(cpu_base), str(ncpu)],
                               capture_output=True, text=True, cwd=ROOT)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                out[tag]["cpu_baseline"] = {"prove_per_s": d["prove_per_s"], "verify_per_s": d["verify_per_s"],
                                            "prove_verify_per_s": d["prove_verify_per_s"], "unit": "proofs/s", "cores": ncpu,
                                            "kind": "port", "sample": f"{d['proofs']} proofs, {ncpu} processes x {cpu_base}, "
                                            f"{d['wall_s']:.1f} s wall incl. generator set-up (not in the rates)"}
            except Exception:
                out[tag]["cpu_baseline"] = {"error": (r.stderr or r.stdout)[-300:]}

    def run_ipp(n, count, tag, cpu_base=0):
        """BASELINE config 1: the standalone inner-product argument (IPP::create_ipp / verify_ipp, ipp.rs:35-260, the shape of
        the reference's own ipp tests at length n) through bph_ipp_create / bph_ipp_verify, one proof per context in flight"""
        from concurrent.futures import ThreadPoolExecutor
        ctxs = [bp.Context(bp.BLS12_381, local) for _ in range(min(nctx, cores))]     # n = 64: one spinning context per core
        nc = len(ctxs)
        c0 = ctxs[0]
        mb = c0.modbytes
        G, H = c0.get_generators("g", n, precompute=True), c0.get_generators("h", n, precompute=True)
        Q = c0.g1_from_msg_hash(b"Q")
        ones = b"".join([(1).to_bytes(mb, "big")] * n)
        jobs = []
        for i in range(count):
            ab = synth_scalars_be(n, 7000 + 131 * rank + 2 * i, mb).tobytes()
            bb = synth_scalars_be(n, 7001 + 131 * rank + 2 * i, mb).tobytes()
            cb = (sum(x * y for x, y in zip(be_ints(ab, n, mb), be_ints(bb, n, mb))) % BLS_R).to_bytes(mb, "big")
            P = c0.msm_parts([(G, ab, n), (H, bb, n), (Q, cb, 1)])
            jobs.append((ab, bb, P))

        def prove(k):
            ctx = ctxs[k % nc]
            return [ctx.ipp_create(b"ipp", G, H, Q, ones, ones, jobs[i][0], jobs[i][1], n) for i in range(k, count, nc)]

        def verify(k):
            ctx = ctxs[k % nc]
            return all(ctx.ipp_verify(b"ipp", n, ones, ones, jobs[i][2], Q, G, H, proofs[i]) for i in range(k, count, nc))
        with ThreadPoolExecutor(nc) as ex:
            list(ex.map(prove, range(nc)))                                         # warm-up
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            res = list(ex.map(prove, range(nc)))
            tp = time.perf_counter() - t0
            proofs = [None] * count
            for k, lst in enumerate(res):
                for j, pr in enumerate(lst):
                    proofs[k + j * nc] = pr
            list(ex.map(verify, range(nc)))
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            ok = all(ex.map(verify, range(nc)))
            tv = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([tp, tv, 0.0 if ok else 1.0], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tp, tv, ok = float(tt[0].item()), float(tt[1].item()), tt[2].item() == 0.0
        for c in ctxs:
            c.close()
        out[tag] = {"n": n, "proofs": count * world, "prove_per_s": count * world / tp, "verify_per_s": count * world / tv,
                    "prove_verify_per_s": count * world / (tp + tv), "all_verified": ok, "proof_bytes": len(proofs[0])}
        if cpu_base and rank == 0 and world == 1:
            r = subprocess.run([sys.executable, "-m", "oracle.fast", "ipp", "BLS12_381", str(n), str(cpu_base), str(ncpu)],
                               capture_output=True, text=True, cwd=ROOT)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                out[tag]["cpu_baseline"] = {"prove_per_s": d["prove_per_s"], "verify_per_s": d["verify_per_s"],
                                            "prove_verify_per_s": d["prove_verify_per_s"], "unit": "proofs/s", "cores": ncpu, "kind": "port",
                                            "sample": f"{d['proofs']} proofs, {ncpu} processes x {cpu_base}, {d['wall_s']:.1f} s wall"}
            except Exception:
                out[tag]["cpu_baseline"] = {"error": (r.stderr or r.stdout)[-300:]}

    # config 1: IPP prove + verify at n = 64
    run_ipp(64, max(nctx, 512 // world), "ipp_bls12_381_n64", cpu_base=0 if no_cpu else 8)
    # config 5 unit / config 1 size: one 64-bit range proof = 64 multipliers, IPP of length 64; 4096 verifications in total
    per_rank = max(nctx, 512 // world)
    run(bp.BLS12_381, 1, 64, per_rank, "range64_bls12_381_n64", verify_reps=max(1, 4096 // (per_rank * world)),
        batch_call=True, cpu_base=0 if no_cpu else 4)
    # config 2: 16 x 64-bit values in one constraint system, 1024 generators
    run(bp.BLS12_381, 16, 64, max(nctx, 64 // world), "range64x16_bls12_381_n1024", verify_reps=4, batch_call=True,
        cpu_base=0 if no_cpu else 1, weak=1024)
    # config 3: 2^14 multipliers on BN254 (256 x 64-bit values)
    run(bp.BN254, 256, 64, max(2 * nctx, 32 // world), "range64x256_bn254_n16384", cpu_base=0 if no_cpu else 1)
    return out


def cpu_reference(lg_sample, threads, curve_id=0, seed=5):
    """Time the oracle's Straus/wNAF-5 MSM (the reference's CPU algorithm) on 2^lg_sample points."""
    from oracle import cref
    from oracle.curves import BLS12_381 as C
    n = 1 << lg_sample
    xy = cref.multiples(curve_id, C.g1_xy_bytes(C.mul(C.from_affine(C.g), 12345)), n)
    sb = synth_scalars_be(n, seed).tobytes()
    t0 = time.perf_counter()
    cref.msm(curve_id, xy, sb, n, threads)
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args):
    cores = os.cpu_count() or 1
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lg_s = args.ref_lg
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_reference(lg_s, cores, seed=5 + i)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(1 << lg_s for _ in vals) / sum(dt for _, dt in vals)
    sample = f"2^{lg_s} of the 2^{args.lg} points per step, Straus wNAF-5 (oracle/c), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (host: u64)", "data": "synthetic",
        "config": make_config(args.lg, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lg", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-lg", type=int, default=18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-proofs", action="store_true", help="skip the IPP / R1CS proofs-per-second section")
    ap.add_argument("--no-sweep", action="store_true", help="skip BASELINE config 4 (size sweep at N = 1, fixed-size strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: anything libraries print while we run (NCCL's version banner goes to
    # stdout) is sent to stderr instead, and the real stdout is restored just before the result is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    # 16 contexts (streams) per GPU submit small launches concurrently: give them their own hardware queues (default 8)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import bulletproofs_amcl_b200 as bp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.lg
    ctx = bp.Context(bp.BLS12_381, local)
    mb = ctx.modbytes

    # ---- synthetic inputs: two independent input sets (2 x (96+32) MB at 2^20 > 126 MB L2), alternated
    # between steps so no step finds its inputs in L2
    NSETS = 2
    gen = [gen_points(ctx, n, 1000 + 17 * rank + s, with_multipliers=True) for s in range(NSETS)]
    pts_xy = [g[0] for g in gen]
    ks_be = [g[1] for g in gen]
    del gen
    sc_np = [synth_scalars_be(n, 2000 + 17 * rank + s, mb) for s in range(NSETS)]
    dpts = [ctx.upload_points(x) for x in pts_xy]
    dsc = [ctx.upload_scalars(s.tobytes()) for s in sc_np]
    # pinned host copies for the end-to-end legs (hs: 48-byte big-endian scalars; hs_le: 32-byte little-endian)
    hp, hs, hs_le = [], [], []
    for s in range(NSETS):
        le = sc_np[s][:, mb - 32:][:, ::-1].copy()
        q = ctx.host_alloc(le.nbytes)
        ctypes.memmove(q, le.ctypes.data, le.nbytes)
        hs_le.append(q)
        p = ctx.host_alloc(len(pts_xy[s]))
        ctypes.memmove(p, pts_xy[s], len(pts_xy[s]))
        hp.append(p)
        q = ctx.host_alloc(sc_np[s].nbytes)
        ctypes.memmove(q, sc_np[s].ctypes.data, sc_np[s].nbytes)
        hs.append(q)

    from bulletproofs_amcl_b200 import sharding

    def combine(partial_xy):
        """N>1: all-gather the 2*MODBYTES-byte partial sums over NCCL, add them on every rank (sharding.py)."""
        if world == 1:
            return partial_xy
        return sharding.combine_partials(ctx, sharding.allgather_bytes(dist, partial_xy))

    def step_resident(i):
        return combine(ctx.msm_device(dpts[i % NSETS], dsc[i % NSETS], n=n))

    def step_e2e(i):
        return combine(ctx.msm_refs(hp[i % NSETS], hs[i % NSETS], n=n))

    def step_e2e_resident(i):            # bases cached on the device (fixed generators), scalars from the host every step
        return combine(ctx.msm(dpts[i % NSETS], hs[i % NSETS], n=n))

    def step_e2e_resident_le32(i):
        return combine(ctx.msm_le32(dpts[i % NSETS], hs_le[i % NSETS], n=n))

    stream = torch.cuda.ExternalStream(ctx.stream)

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        if profile:
            ctx.set_profile(n)          # only the 2^lg-term MSMs, not the N-term combine of the multi-GPU path
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), wall_ms)     # the host finish sits between kernels: never under-report
        stages = ctx.msm_stage_ms() if profile else None
        if profile:
            ctx.set_profile(0)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launches - l0, stages

    # integer-pipe peak, measured live
    imad_peak, _ = ctx.int_pipe_bench(0, 2000)
    fqmul_peak, _ = ctx.int_pipe_bench(1, 200)

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, (runs, stages) = timed(step_resident, args.steps, args.warmup, profile=True)
    clocks = sampler.stop()
    ms_e2e_seq, _, _ = timed(step_e2e, args.steps, args.warmup)

    # ---- the headline passes: TWO MSMs in flight on two contexts of the GPU (bpgpu_msm_*_begin / bpgpu_msm_finish).  The sort
    # stages, the host-to-device copies and the host finish (Horner + inversion) of one step run under the bucket
    # accumulation of the other; every step still returns its own point, and the steps are the same independent MSMs over the
    # alternating input sets as in the sequential pass above.
    ctx2 = bp.Context(bp.BLS12_381, local)
    pair = [ctx, ctx2]

    def timed_pipelined(begin, steps, warmup):
        def run(count, base):
            begin(pair[0], base)
            for i in range(1, count):
                begin(pair[i % 2], base + i)
                combine(pair[(i - 1) % 2].msm_finish())
            return combine(pair[(count - 1) % 2].msm_finish())
        run(warmup, 0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches + ctx2.launches
        t0 = time.perf_counter()
        run(steps, warmup)
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, ctx.launches + ctx2.launches - l0

    ms_seq, launches_seq = ms, launches
    ms, launches = timed_pipelined(lambda c, i: c.msm_device_begin(dpts[i % NSETS], dsc[i % NSETS], n=n), args.steps, args.warmup)
    ms_e2e, _ = timed_pipelined(lambda c, i: c.msm_refs_begin(hp[i % NSETS], hs[i % NSETS], n=n), args.steps, args.warmup)
    ms_e2e_res, _ = timed_pipelined(lambda c, i: c.msm_begin(dpts[i % NSETS], hs[i % NSETS], n=n), args.steps, args.warmup)
    ms_e2e_le, _ = timed_pipelined(lambda c, i: c.msm_le32_begin(dpts[i % NSETS], hs_le[i % NSETS], n=n), args.steps, args.warmup)
    # the pipelined entry points against the oracle-checked sequential ones
    pair[1].msm_refs_begin(hp[1], hs[1], n=n)
    pair[0].msm_device_begin(dpts[0], dsc[0], n=n)
    if pair[0].msm_finish() != ctx.msm_device(dpts[0], dsc[0], n=n) or pair[1].msm_finish() != ctx.msm_refs(hp[1], hs[1], n=n):
        sys.stderr.write(f"PARITY FAILURE rank {rank}: begin/finish result differs from the one-call result\n")
        sys.exit(3)

    # ---- parity guard against the ORACLE, on every rank and on the combined result at every N: the bench points are
    # k_i * G with known k_i, so the MSM must equal (sum s_i * k_i mod r) * G -- one oracle scalar multiplication.
    # A mismatch aborts the run: no bench line is printed.
    def shard_total(set_idx):
        ks = be_ints(ks_be[set_idx], n, mb)
        ss = be_ints(sc_np[set_idx], n, mb)
        return sum(a * b for a, b in zip(ks, ss)) % BLS_R

    def all_totals(local_total):
        if world == 1:
            return [local_total]
        t = torch.tensor(list(local_total.to_bytes(32, "big")), dtype=torch.uint8, device="cuda")
        allt = torch.empty(world * 32, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allt, t)
        raw = bytes(allt.cpu().tolist())
        return [int.from_bytes(raw[i * 32:(i + 1) * 32], "big") for i in range(world)]

    parity = {}
    for set_idx in range(NSETS):
        tot = shard_total(set_idx)
        exp_local = expected_point(tot)
        for name, got in (("msm_device", ctx.msm_device(dpts[set_idx], dsc[set_idx], n=n)),
                          ("msm_refs", ctx.msm_refs(hp[set_idx], hs[set_idx], n=n)),
                          ("msm", ctx.msm(dpts[set_idx], hs[set_idx], n=n)),
                          ("msm_le32", ctx.msm_le32(dpts[set_idx], hs_le[set_idx], n=n))):
            if got != exp_local:
                sys.stderr.write(f"PARITY FAILURE rank {rank}: {name} on input set {set_idx} differs from the oracle\n")
                sys.exit(3)
        exp_all = expected_point(sum(all_totals(tot)))
        if step_resident(set_idx) != exp_all or step_e2e(set_idx) != exp_all:
            sys.stderr.write(f"PARITY FAILURE rank {rank}: combined {world}-GPU result of input set {set_idx} differs from the oracle\n")
            sys.exit(3)
    parity = {"checked_against": "oracle: (sum s_i*k_i mod r)*G, one scalar multiplication (oracle/curves.py)",
              "entries": ["bpgpu_msm_device", "bpgpu_msm_refs", "bpgpu_msm", "bpgpu_msm_le32"], "input_sets": NSETS,
              "per_rank": True, "combined_over_ranks": world, "ok": True}

    # ---- config 4: size sweep on one GPU, and a FIXED-size MSM split by points over the ranks (strong scaling)
    def time_msm(dp, ds, nn, reps, comb):
        for _ in range(2):
            r0 = ctx.msm_device(dp, ds, n=nn)
            if comb:
                combine(r0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r0 = ctx.msm_device(dp, ds, n=nn)
            if comb:
                r0 = combine(r0)
        dt = (time.perf_counter() - t0) * 1e3 / reps
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, r0

    sweep, strong = None, None
    if not args.no_sweep:
        # release the headline inputs first: the 2^22 case needs 0.4 GB of points + scratch
        if world == 1:
            sweep = []
            for lg in range(10, 23, 2):
                nn = 1 << lg
                xy, kb = gen_points(ctx, nn, 7000 + lg, with_multipliers=True)
                sb = synth_scalars_be(nn, 8000 + lg, mb)
                dp, ds = ctx.upload_points(xy), ctx.upload_scalars(sb.tobytes())
                reps = 50 if lg <= 14 else (20 if lg <= 18 else 8)
                dt, res = time_msm(dp, ds, nn, reps, False)
                tot = sum(a * b for a, b in zip(be_ints(kb, nn, mb), be_ints(sb, nn, mb))) % BLS_R
                if res != expected_point(tot):
                    sys.stderr.write(f"PARITY FAILURE: sweep 2^{lg} differs from the oracle\n")
                    sys.exit(3)
                sweep.append({"lg_n": lg, "ms": dt, "points_per_s": nn / (dt * 1e-3), "window_bits": bp.lib().bpgpu_msm_window_bits(nn),
                              "oracle_checked": True})
                dp.free()
                ds.free()
        strong = {}
        for lg_tot in (20, 22):
            nn = (1 << lg_tot) // world
            xy, kb = gen_points(ctx, nn, 9000 + lg_tot + 31 * rank, with_multipliers=True)
            sb = synth_scalars_be(nn, 9500 + lg_tot + 31 * rank, mb)
            dp, ds = ctx.upload_points(xy), ctx.upload_scalars(sb.tobytes())
            dt, res = time_msm(dp, ds, nn, 8, True)
            tot = sum(a * b for a, b in zip(be_ints(kb, nn, mb), be_ints(sb, nn, mb))) % BLS_R
            if res != expected_point(sum(all_totals(tot))):
                sys.stderr.write(f"PARITY FAILURE rank {rank}: strong-scaling 2^{lg_tot} over {world} ranks differs from the oracle\n")
                sys.exit(3)
            strong[f"total_2^{lg_tot}"] = {"points_total": 1 << lg_tot, "points_per_gpu": nn, "gpus": world, "ms": dt,
                                            "points_per_s": (1 << lg_tot) / (dt * 1e-3), "oracle_checked": True,
                                            "includes": "per-rank MSM + NCCL all-gather of the partial sums + host add on every rank"}
            dp.free()
            ds.free()

    proofs = None if args.no_proofs else proof_section(bp, ctx, local, rank, world, dist, torch, no_cpu=args.no_cpu_baseline)

    if rank == 0:
        c = bp.lib().bpgpu_msm_window_bits(n)
        W = (256 + c - 1) // c
        total_points = n * world * args.steps
        value = total_points / (ms * 1e-3)
        e2e_value = total_points / (ms_e2e * 1e-3)
        k_ms = stages["chunk_acc"]
        alg_imad = n * W * FQMUL_PER_MADD * IMAD_PER_FQMUL
        achieved = alg_imad / (k_ms * 1e-3)
        alg_bytes = n * W * (2 * mb + 8) + n * 32      # gathered affine points + sorted (index,key) pairs, once per window
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "pipelining": "two MSMs in flight on two contexts of each GPU (bpgpu_msm_*_begin / bpgpu_msm_finish): sort stages, copies and the "
                          "host finish of one step run under the bucket accumulation of the other; every step returns its own point",
            "sequential": {"value": total_points / (ms_seq * 1e-3), "ms_per_step": ms_seq / args.steps, "gpu_launches": int(launches_seq),
                           "e2e_ms_per_step": ms_e2e_seq / args.steps, "e2e_value": total_points / (ms_e2e_seq * 1e-3),
                           "note": "one MSM at a time on one context (round 1's timed region); the stage times and the roofline below "
                                   "are measured in this pass"},
            "dtype": "u32 limbs (Montgomery Fq 12x32, Fr 8x32)", "data": "synthetic",
            "config": make_config(args.lg, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 3 * mb, "d2h_bytes_per_step": 2 * W * 4 * mb,
                    "ms_per_step": ms_e2e / args.steps, "api": "bpgpu_msm_refs (host points + host scalars, pinned)"},
            # the reference's call sites multiply FIXED generators (ipp.rs:91-104, prover.rs:347-362, verifier.rs:451): bases stay
            # resident, only the scalars travel
            "e2e_resident_bases": {"value": total_points / (ms_e2e_res * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * mb,
                                   "ms_per_step": ms_e2e_res / args.steps, "api": "bpgpu_msm (cached bases, 48-byte big-endian host scalars)",
                                   "le32": {"value": total_points / (ms_e2e_le * 1e-3), "h2d_bytes_per_step": n * 32,
                                            "ms_per_step": ms_e2e_le / args.steps,
                                            "api": "bpgpu_msm_le32 (cached bases, 32-byte little-endian host scalars, no conversion pass)"}},
            "parity": parity,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int32-imad", "kernel": "k_chunk_acc", "achieved": achieved, "peak": imad_peak,
                         "unit": "IMAD.WIDE/s", "frac": achieved / imad_peak, "traffic": None,
                         "peak_source": "bpgpu_int_pipe_bench(IMAD.WIDE.U32) measured in this run",
                         "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (ms_seq / args.steps),
                         "measured_in": "the sequential pass (CUDA events on the ctx stream between the stages of every MSM)",
                         "algorithmic_imad_per_launch": alg_imad, "fq_mul_per_s_peak_measured": fqmul_peak,
                         "hbm": {"achieved_GBps": alg_bytes / (k_ms * 1e-3) / 1e9, "algorithmic_bytes": alg_bytes}},
            "stages_ms": stages,
        }
        if proofs is not None:
            out["proofs"] = proofs
        if sweep is not None:
            out["sweep"] = sweep
        if strong is not None:
            out["strong_scaling"] = strong
        out["roofline"]["traffic"], out["roofline"]["traffic_source"] = ncu_traffic_bytes()
        if out["roofline"]["traffic_source"]:
            out["roofline"]["traffic_source"] += " (stored ncu --set full capture, not measured in this run)"
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v1, dt1 = cpu_reference(16, 1)
            vN, dtN = cpu_reference(args.lg, cores)
            out["cpu_baseline"] = {"value": vN, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"Straus wNAF-5 (oracle/c) on all 2^{args.lg} points, {cores} threads, {dtN:.1f} s",
                                   "single_core": {"value": v1, "cores": 1, "sample": f"2^16 points, {dt1:.1f} s"}}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
