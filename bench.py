#!/usr/bin/env python
"""bench.py — create_proof-schedule benchmark for the B200 back end (BASELINE.json metric:
"create_proof ms + MSM G1-pairs/s at k=19 ECDSA").

One "step" = one pass of the prover hot path for ONE proof of the halo2-ecc secp256k1 ECDSA circuit at k=19
(BASELINE.json configs[2]; column shape 1 advice / q_lookup / 1 fixed, halo2-ecc/configs/secp256k1/bench_ecdsa.config:1):
the witness-column assignment, the 12 MSMs of size 2^19 and the (coset) NTTs create_proof issues for that
constraint system (restated — the prover crate is not vendored).  `value` is the MSM
throughput of the whole step (G1 pairs / step time); `ms_per_step` is the create_proof-schedule time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--k 19]
"""
from __future__ import annotations
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

# ---- the restated create_proof schedule for the ECDSA circuit (1 advice, lookup on the same column, d = 5)
# (basis, scalar class): witness-like columns are dominated by 0/1 bits and <= 88-bit limbs
MSM_SCHEDULE = (
    [("lagrange", "witness")] * 1      # advice column commitment                       (§3.3 step 2)
    + [("lagrange", "witness")] * 2    # lookup permuted input / table                  (step 3)
    + [("lagrange", "uniform")] * 2    # permutation grand product, lookup grand product (step 4)
    + [("monomial", "uniform")] * 1    # vanishing argument random polynomial           (step 5)
    + [("monomial", "uniform")] * 4    # h(X) pieces, d - 1 = 4                         (step 6)
    + [("monomial", "uniform")] * 2    # SHPLONK                                        (step 8)
)
# prover phases: the MSMs inside one phase are independent (batched over the library's lanes); a phase can only start
# when the commitments of the previous one have been hashed into the transcript (challenge dependency), so phases
# are serialised on the stream.  Indices into MSM_SCHEDULE.
MSM_PHASES = [[0], [1, 2], [3, 4, 5], [6, 7, 8, 9], [10], [11]]
N_INTT = 5        # lagrange_to_coeff: advice, 2 permuted, 2 grand products          (steps 3,4,6)
N_COSET = 5       # coeff_to_extended (2^k -> 2^(k+2)) of the same five polynomials  (step 6)
N_COSET_INV = 1   # extended_to_coeff of h(X)                                        (step 6)
QUOTIENT_J = 5    # cs.degree() with the q_lookup lookup: extended_k = k + 2


def witness_like(rng, n):
    """canonical ints: 35% zero, 25% one, 30% < 2^88, 10% uniform Fr (SURVEY.md §8d distribution W)"""
    cls = rng.random(n)
    out = np.zeros((n, 4), dtype=np.uint64)
    one = (cls >= 0.35) & (cls < 0.60)
    small = (cls >= 0.60) & (cls < 0.90)
    full = cls >= 0.90
    out[one, 0] = 1
    k = int(small.sum())
    out[small, 0] = rng.integers(0, 1 << 63, size=k, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rng.integers(0, 2, size=k, dtype=np.int64).astype(np.uint64)
    out[small, 1] = rng.integers(0, 1 << 24, size=k, dtype=np.int64).astype(np.uint64)
    k = int(full.sum())
    out[full] = uniform_residues(rng, k)
    return out


def uniform_residues(rng, n):
    """n uniform values < 2^252 < r as 4 x u64 limbs (valid Montgomery residues and valid canonical values)"""
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    a[:, :3] = a[:, :3] * np.uint64(2) + rng.integers(0, 2, size=(n, 3), dtype=np.int64).astype(np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(k: int, threads: int | None = None):
    """Times the CPU restatement (oracle/, OpenMP, all host threads) on one op of each class of the schedule and
    composes the step time: sum(count_i * t_i).  ~10-30 s of CPU work on a typical host."""
    from oracle import oracle as orc
    try:  # a -march=native build for the host it runs on (the shipped .so is x86-64-v3)
        so = os.path.join(ROOT, "oracle", "_build", "liboracle_native.so")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "MARCH=native", f"OUT={so}"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        import ctypes
        orc._lib = None
        orc._SO = so
    except Exception:
        pass
    if threads is None:  # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
    # SMT siblings hurt this integer code on some hosts: every op is timed with all logical CPUs and with half of
    # them, and the faster one is kept (a tuned CPU run would do the same)
    cand = sorted({threads, max(1, threads // 2)}, reverse=True)
    orc.set_threads(threads)
    orc.use_fast_ntt(True)  # many-core four-step NTT (oracle/bn254_oracle.c: orc_ntt_fast)
    n = 1 << k
    ext_k = k + 2
    rng = np.random.default_rng(0xB2000000 + k)
    # bases: any valid curve points time the same; build n points a_i * G cheaply with the oracle itself
    from util import affine_to_limbs, mont
    from oracle import pyref
    g = affine_to_limbs([pyref.G1])[0]
    small = np.zeros((n, 4), dtype=np.uint64)
    small[:, 0] = np.arange(3, 3 + 5 * n, 5, dtype=np.uint64)  # canonical small scalars
    t0 = time.perf_counter()
    bases = orc.g1_fixed_base_mul(orc.to_mont(orc.FR, small), g)
    t_setup = time.perf_counter() - t0
    s_uni = uniform_residues(rng, n)
    s_wit = orc.to_mont(orc.FR, witness_like(rng, n))
    # warm the OpenMP pool and the code paths on a tiny instance before timing anything
    orc.msm_pippenger(s_uni[:256], bases[:256], threads)
    orc.extended_to_coeff(orc.coeff_to_extended(orc.lagrange_to_coeff(s_uni[:256], 8, threads), 10, threads), 10, threads)
    times, used = {}, {}
    a = uniform_residues(rng, n)

    def best(name, fn, reps=1):
        res = None
        for th in cand:
            for _ in range(reps):
                t0 = time.perf_counter(); r = fn(th); dt = time.perf_counter() - t0
                if name not in times or dt < times[name]:
                    times[name], used[name] = dt, th
                res = r
        return res
    best("msm_uniform", lambda th: orc.msm_pippenger(s_uni, bases, th))
    best("msm_witness", lambda th: orc.msm_pippenger(s_wit, bases, th))
    coeffs = best("intt", lambda th: orc.lagrange_to_coeff(a, k, th), reps=2)
    ext = best("coset_ntt", lambda th: orc.coeff_to_extended(coeffs, ext_k, th), reps=2)
    best("coset_intt", lambda th: orc.extended_to_coeff(ext, ext_k, th), reps=2)
    best("assign", lambda th: orc.assign_witnesses(a[: n - 20], np.zeros(0, dtype=np.uint64), k, 1), reps=2)
    orc.use_fast_ntt(False)
    n_wit = sum(1 for _, c in MSM_SCHEDULE if c == "witness")
    n_uni = len(MSM_SCHEDULE) - n_wit
    step_s = (n_uni * times["msm_uniform"] + n_wit * times["msm_witness"] + N_INTT * times["intt"] + N_COSET * times["coset_ntt"]
              + N_COSET_INV * times["coset_intt"] + times["assign"])
    pairs = len(MSM_SCHEDULE) * n
    return {
        "value": pairs / step_s,
        "unit": "G1 pairs/s",
        "cores": threads,
        "threads_used": used,
        "kind": "port",
        "sample": (f"oracle/bn254_oracle.c (restated CPU path, OpenMP, best of {cand} threads per op; the Rust reference cannot be built here): one MSM(2^{k}) per scalar "
                   f"class + one iNTT(2^{k}) + one coeff_to_extended/extended_to_coeff(2^{ext_k}) + one assignment, composed by the schedule counts"),
        "step_ms": step_s * 1e3,
        "op_ms": {kk: v * 1e3 for kk, v in times.items()},
        "setup_s": t_setup,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_sample(args.k)
        if i >= args.warmup:
            vals.append(last)
    step_ms = float(np.mean([v["step_ms"] for v in vals]))
    pairs = len(MSM_SCHEDULE) * (1 << args.k)
    value = pairs / (step_ms / 1e3)
    cb = dict(last)
    cb["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": "msm_g1_pairs_per_s (create_proof schedule, ECDSA k=%d)" % args.k, "value": value, "unit": "G1 pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": workload_config(args.k, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "G1 pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "create_proof_schedule_ms": step_ms,
    }))


def workload_config(k, gpus):
    return {
        "workload": f"halo2-ecc secp256k1 ECDSA verify circuit, k={k} (BASELINE.json configs[2]): create_proof schedule restated in SURVEY.md §3.3/§8",
        "k": k, "columns": "1 advice / q_lookup / 1 fixed", "msm": f"{len(MSM_SCHEDULE)} x 2^{k} ({sum(1 for b, _ in MSM_SCHEDULE if b == 'lagrange')} lagrange + {sum(1 for b, _ in MSM_SCHEDULE if b == 'monomial')} monomial basis)",
        "ntt": f"{N_INTT} x iNTT(2^{k}) + {N_COSET} x coeff_to_extended(2^{k + 2}) + {N_COSET_INV} x extended_to_coeff(2^{k + 2})",
        "assignment": f"1 column x 2^{k} rows", "scalars": "3 witness-like + 9 uniform columns (SURVEY.md §8d)",
        "overlap": "MSMs of a transcript phase run on 3 lanes; the iNTT + coset NTT of a polynomial run on a side stream from the moment the polynomial exists and are joined before extended_to_coeff / the h(X) commitments",
        "parallelism": f"msm point-range sharded x{gpus} + fused NVLink peer all-reduce of the partial sums (one kernel per phase); NTT one polynomial per device" if gpus > 1 else "single GPU",
        "l2_policy": "inputs larger than L2: 12 distinct scalar columns + two 15-level base tables (~1 GB) + NTT buffers (~0.5 GB) per step vs 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import ctypes as C
    import halo2_lib_b200 as h
    from halo2_lib_b200._capi import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    k = args.k
    n = 1 << k
    ext_k = k + 2
    begin, n_loc = h.shard_range(n, rank, world)
    ctx = h.Context(local_rank)
    # a dedicated non-default stream: the library treats a NULL stream as "use the context's own stream", and the
    # CUDA events below must be recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    # second context + stream on the same GPU: the polynomial transforms of a column run beside the commitment phases
    # that do not depend on them (see step_resident)
    ctx_ntt = h.Context(local_rank)
    stream_ntt = torch.cuda.Stream(device=dev)
    ctx_ntt.set_stream(stream_ntt.cuda_stream)
    rng = np.random.default_rng(0xB2000000 + k)

    def dev_u64(arr):
        return torch.from_numpy(arr.view(np.int64)).to(dev)

    # ---- setup (untimed): SRS-like bases on the GPU, this rank's shard only
    gbase = np.array([0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f,  # x = 1 (Montgomery)
                      0xa6ba871b8b1e1b3a, 0x14f1d651eb8e167b, 0xccdd46def0f28c58, 0x1c14ef83340fbe5e], dtype=np.uint64)  # y = 2
    tables = {}
    for name, (a0, d) in {"monomial": (3, 5), "lagrange": (7, 11)}.items():
        sc = np.zeros((n_loc, 4), dtype=np.uint64)
        sc[:, 0] = (a0 + d * (begin + np.arange(n_loc, dtype=np.uint64))).astype(np.uint64)
        sc_m = ctx.field_op(1, 5, sc)  # to Montgomery form on the GPU
        d_sc = dev_u64(sc_m)
        d_pts = torch.empty((n_loc, 8), dtype=torch.int64, device=dev)
        ctx.check(lib.h2b_g1_fixed_base_mul_dev(ctx.h, C.c_void_p(gbase.ctypes.data), C.c_void_p(d_sc.data_ptr()), n_loc, C.c_void_p(d_pts.data_ptr())))
        tables[name] = d_pts
    torch.cuda.synchronize()
    if world > 1:
        h.connect_peers(ctx)  # NVLink mailboxes for the fused all-reduce of partial commitments (csrc/peer.cu)
    params = h.ParamsKZG(ctx, k, g=tables["monomial"].data_ptr(), g_lagrange=tables["lagrange"].data_ptr(), begin=begin, count=n_loc, device_ptrs=True)
    del tables
    # ---- inputs: 12 scalar columns (host pinned + device resident), NTT polynomials, the virtual witness column
    cols_host, cols_dev = [], []
    for basis, cls in MSM_SCHEDULE:
        full = uniform_residues(rng, n) if cls == "uniform" else ctx.field_op(1, 5, witness_like(rng, n))
        shard = np.ascontiguousarray(full[begin:begin + n_loc])
        th = torch.from_numpy(shard.view(np.int64)).pin_memory()
        cols_host.append(th)
        cols_dev.append(th.to(dev))
    basis_id = [0 if b == "monomial" else 1 for b, _ in MSM_SCHEDULE]
    my_ntt = lambda i: h.ntt_owner(i, world) == rank  # one polynomial per device, round-robin
    polys_host = [torch.from_numpy(uniform_residues(rng, n).view(np.int64)).pin_memory() for _ in range(N_INTT)]
    polys_dev = [p.to(dev) for p in polys_host]
    ext_host = [torch.empty((1 << ext_k, 4), dtype=torch.int64).pin_memory() for _ in range(N_COSET)]
    ext_dev = [torch.empty((1 << ext_k, 4), dtype=torch.int64, device=dev) for _ in range(N_COSET)]
    n_cells = n - 20
    vcol_host = torch.from_numpy(ctx.field_op(1, 5, witness_like(rng, n_cells)).view(np.int64)).pin_memory()
    vcol_dev = vcol_host.to(dev)
    acol_host = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    acol_dev = torch.empty((n, 4), dtype=torch.int64, device=dev)
    outs_dev = torch.zeros((len(MSM_SCHEDULE), 12), dtype=torch.int64, device=dev)
    outs_host = np.zeros((len(MSM_SCHEDULE), 12), dtype=np.uint64)
    vp = C.c_void_p

    # which polynomials exist when a commitment phase starts (their iNTT + coset NTT may then run on the side stream):
    # the advice column after the assignment, the two permuted lookup columns with phase 1, the two grand products with
    # phase 2.  Everything must be back before h(X) is formed (extended_to_coeff, then the h-piece commitments).
    NTT_READY = {0: [0], 1: [1, 2], 2: [3, 4]}
    ev_fork = [torch.cuda.Event() for _ in range(4)]
    ev_join = torch.cuda.Event()

    def ntt_side(polys):
        for i in polys:
            if my_ntt(i):
                ctx_ntt.check(lib.h2b_lagrange_to_coeff_dev(ctx_ntt.h, vp(polys_dev[i].data_ptr()), k))
            if my_ntt(N_INTT + i):
                ctx_ntt.check(lib.h2b_coeff_to_extended_dev(ctx_ntt.h, vp(polys_dev[i].data_ptr()), n, ext_k, vp(ext_dev[i].data_ptr())))

    def step_resident(overlap=True):
        if rank == 0:
            ctx.check(lib.h2b_assign_columns_dev(ctx.h, vp(vcol_dev.data_ptr()), n_cells, None, 0, k, 1, vp(acol_dev.data_ptr())))
        for pi, phase in enumerate(MSM_PHASES):
            if pi in NTT_READY:
                if overlap:
                    ev_fork[pi].record(stream)
                    stream_ntt.wait_event(ev_fork[pi])
                    ntt_side(NTT_READY[pi])
            elif pi == 3:
                if overlap:
                    ev_join.record(stream_ntt)
                    stream.wait_event(ev_join)
                else:
                    for q in (0, 1, 2):
                        for i in NTT_READY[q]:
                            if my_ntt(i):
                                ctx.check(lib.h2b_lagrange_to_coeff_dev(ctx.h, vp(polys_dev[i].data_ptr()), k))
                            if my_ntt(N_INTT + i):
                                ctx.check(lib.h2b_coeff_to_extended_dev(ctx.h, vp(polys_dev[i].data_ptr()), n, ext_k, vp(ext_dev[i].data_ptr())))
                if my_ntt(N_INTT + N_COSET):
                    ctx.check(lib.h2b_extended_to_coeff_dev(ctx.h, vp(ext_dev[0].data_ptr()), ext_k))
            j0 = phase[0]
            params.commit_batch_dev([basis_id[j] for j in phase], [cols_dev[j].data_ptr() for j in phase], n_loc, outs_dev[j0].data_ptr())
            if world > 1:  # all-reduce under EC addition: one fused kernel over NVLink peer memory, no NCCL call
                h.allreduce_points(ctx, outs_dev[j0].data_ptr(), len(phase))

    from concurrent.futures import ThreadPoolExecutor
    side_pool = ThreadPoolExecutor(max_workers=1)  # the host thread that drives the transform context (ctypes drops the GIL)

    def ntt_side_host(polys):
        """h2b_lagrange_to_coeff_and_extended_batch on the second context: pinned host buffers in and out, the coefficients
        stay on the device between the two transforms, PCIe legs pipelined against the kernels inside the call"""
        both = [i for i in polys if my_ntt(i) and my_ntt(N_INTT + i)]
        if both:
            pa = (C.c_void_p * len(both))(*[polys_host[i].data_ptr() for i in both])
            pe = (C.c_void_p * len(both))(*[ext_host[i].data_ptr() for i in both])
            ctx_ntt.check(lib.h2b_lagrange_to_coeff_and_extended_batch(ctx_ntt.h, pa, len(both), k, ext_k, pe))
        a = [i for i in polys if my_ntt(i) and i not in both]
        if a:
            ptrs = (C.c_void_p * len(a))(*[polys_host[i].data_ptr() for i in a])
            ctx_ntt.check(lib.h2b_lagrange_to_coeff_batch(ctx_ntt.h, ptrs, len(a), k))
        b = [i for i in polys if my_ntt(N_INTT + i) and i not in both]
        if b:
            pin = (C.c_void_p * len(b))(*[polys_host[i].data_ptr() for i in b])
            pout = (C.c_void_p * len(b))(*[ext_host[i].data_ptr() for i in b])
            ctx_ntt.check(lib.h2b_coeff_to_extended_batch(ctx_ntt.h, pin, len(b), n, ext_k, pout))

    trace = [] if os.environ.get("H2B_E2E_TRACE") else None  # host wall-clock marks of one e2e step (diagnostic, stderr)

    def mark(label):
        if trace is not None:
            trace.append((label, time.perf_counter()))

    def step_e2e(overlap=True):
        """the same step through the host-pointer C ABI: pinned host buffers in, host buffers out.  With `overlap` a second
        host thread drives the transforms of the polynomials that already exist (same dependency model as step_resident)
        through a second context, so their PCIe traffic runs beside the commitment phases."""
        mark("start")
        if rank == 0:
            ctx.check(lib.h2b_assign_columns(ctx.h, vp(vcol_host.data_ptr()), n_cells, None, 0, k, 1, vp(acol_host.data_ptr())))
        mark("assign")
        pending = []
        for pi, phase in enumerate(MSM_PHASES):
            if pi in NTT_READY:
                if overlap:
                    pending.append(side_pool.submit(ntt_side_host, NTT_READY[pi]))
            elif pi == 3:
                if overlap:
                    for f in pending:
                        f.result()
                else:
                    ntt_side_host([0, 1, 2, 3, 4])
                mark("join_side")
                if my_ntt(N_INTT + N_COSET):
                    ctx.check(lib.h2b_extended_to_coeff(ctx.h, vp(ext_host[0].data_ptr()), ext_k))
                mark("ext_to_coeff")
            m = len(phase)
            ptrs = (C.c_void_p * m)(*[cols_host[j].data_ptr() for j in phase])
            bs = (C.c_int * m)(*[basis_id[j] for j in phase])
            out = np.empty((m, 12), dtype=np.uint64)
            ctx.check(lib.h2b_msm_g1_batch(ctx.h, params.h, bs, ptrs, m, n_loc, vp(out.ctypes.data)))
            if world > 1:
                t = torch.from_numpy(out.view(np.int64)).to(dev)
                g = h.all_gather_points(t)
                for jj in range(m):
                    ctx.check(lib.h2b_g1_sum_dev(ctx.h, vp(g[jj].data_ptr()), world, vp(t[jj].data_ptr())))
                out = t.cpu().numpy().view(np.uint64)
            outs_host[phase] = out
            mark("phase%d" % pi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, prof=None):
        for _ in range(warmup):
            fn()
        barrier()
        if prof:
            ctx.profile_reset()
            ctx.profile_enable(prof)
        l0 = ctx.kernel_launches + ctx_ntt.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if prof:
            ctx.profile_enable(None)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, ctx.kernel_launches + ctx_ntt.kernel_launches - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    params_windows = params.windows
    ms_step, launches = timed(step_resident, args.steps, args.warmup, prof="k_accumulate")
    ms_step_seq, _ = timed(lambda: step_resident(False), max(1, min(args.steps, 5)), 1)
    acc_ms, acc_cnt = ctx.profile_read("k_accumulate")
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 5)), 1)
    ms_e2e_seq, _ = timed(lambda: step_e2e(False), max(1, min(args.steps, 3)), 1)
    if trace is not None and rank == 0:
        trace.clear()
        step_e2e(True)
        torch.cuda.synchronize()
        t0 = trace[0][1]
        print("e2e trace (ms since start): " + ", ".join("%s=%.2f" % (l, 1e3 * (t - t0)) for l, t in trace[1:]), file=sys.stderr)

    # per-op device timings (context for the headline; same CUDA-event method, 3 reps each)
    def time_op(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    iu = next(j for j, (_, c) in enumerate(MSM_SCHEDULE) if c == "uniform")
    iw = next(j for j, (_, c) in enumerate(MSM_SCHEDULE) if c == "witness")
    op_ms = {
        "msm_uniform": time_op(lambda: params.commit_dev(basis_id[iu], cols_dev[iu].data_ptr(), n_loc, outs_dev[iu].data_ptr())),
        "msm_witness": time_op(lambda: params.commit_dev(basis_id[iw], cols_dev[iw].data_ptr(), n_loc, outs_dev[iw].data_ptr())),
        "intt": time_op(lambda: ctx.check(lib.h2b_lagrange_to_coeff_dev(ctx.h, vp(polys_dev[0].data_ptr()), k))),
        "coset_ntt": time_op(lambda: ctx.check(lib.h2b_coeff_to_extended_dev(ctx.h, vp(polys_dev[0].data_ptr()), n, ext_k, vp(ext_dev[0].data_ptr())))),
        "coset_intt": time_op(lambda: ctx.check(lib.h2b_extended_to_coeff_dev(ctx.h, vp(ext_dev[0].data_ptr()), ext_k))),
        "assign": time_op(lambda: ctx.check(lib.h2b_assign_columns_dev(ctx.h, vp(vcol_dev.data_ptr()), n_cells, None, 0, k, 1, vp(acol_dev.data_ptr())))),
    }

    # the same 12 columns once more, one MSM at a time (no lane overlap): k_accumulate timed alone
    ctx.profile_reset()
    ctx.profile_enable("k_accumulate")
    for j in range(len(MSM_SCHEDULE)):
        params.commit_dev(basis_id[j], cols_dev[j].data_ptr(), n_loc, outs_dev[j].data_ptr())
    torch.cuda.synchronize()
    iso_ms, iso_cnt = ctx.profile_read("k_accumulate")
    ctx.profile_enable(None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pairs = len(MSM_SCHEDULE) * n
    value = pairs / (ms_step / 1e3)
    peak, peak_src = measured_hbm_peak()
    acc_avg_ms = acc_ms / max(acc_cnt, 1)
    iso_avg_ms = iso_ms / max(iso_cnt, 1)
    achieved = 96.0 * n_loc / (acc_avg_ms / 1e3) / 1e9  # algorithmic 96 B per pair (32 B scalar + 64 B base), SURVEY.md §8d
    achieved_iso = 96.0 * n_loc / (iso_avg_ms / 1e3) / 1e9
    # field products of the average launch (estimate): entries = pairs * windows * P(non-zero digit); witness-like columns
    # keep ~23% of their digits (35% zeros, 25% ones, 30% 88-bit limbs, 10% full width)
    entries = n_loc * params_windows * (sum(1.0 if c == "uniform" else 0.23 for _, c in MSM_SCHEDULE) / len(MSM_SCHEDULE))
    products = 9.06 * entries  # a mixed addition = 1160 wide multiplies = 9.06 x the 128 of one full product
    roofline = {"bound": "hbm", "kernel": "k_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch on a uniform column at k=19, single GPU, from the
                # `ncu --set full` capture summarised in profiles/ (979.2 + 37.7 MB)
                "traffic": 1016869632 if (k == 19 and world == 1) else None, "peak_source": peak_src,
                "avg_launch_ms": acc_avg_ms, "launches_timed": acc_cnt, "algorithmic_bytes_per_launch": 96 * n_loc,
                "timed_region_note": "launches of the timed region overlap with the kernels of the other two MSM lanes, which stretches each launch",
                "isolated": {"avg_launch_ms": iso_avg_ms, "launches": iso_cnt, "achieved": achieved_iso, "frac": achieved_iso / peak,
                             "integer_multiplier": {"products_per_s": products / (iso_avg_ms / 1e3), "peak_products_per_s": 66.9e9,
                                                    "frac": products / (iso_avg_ms / 1e3) / 66.9e9,
                                                    "note": "estimate: 9.06 Montgomery-product equivalents per XYZZ mixed add (6 products + 2 squarings at 100/128 + one fused two-product at 192/128) x non-zero digits; peak = tools/latbench.cu (profiles/r01_pipe_microbench.txt)"}},
                "note": "bucket accumulation is bound by the integer multiplier (IMAD.WIDE), not by HBM: traffic is ~10% of HBM peak; see DESIGN.md 4.1/4.2"}
    # the coefficients of a polynomial go up once for both of its transforms (fused batch call) when one rank owns both
    coset_up = 0 if world == 1 else N_COSET * n * 32
    h2d = (len(MSM_SCHEDULE) * n_loc * 32 + (N_INTT * n * 32 + coset_up + N_COSET_INV * (1 << ext_k) * 32) // world + n_cells * 32)
    d2h = (len(MSM_SCHEDULE) * 96 + (N_INTT * n * 32 + N_COSET * (1 << ext_k) * 32 + N_COSET_INV * (1 << ext_k) * 32) // world + n * 32)
    cpu = cpu_sample(k) if world == 1 and not args.no_cpu else None
    line = {
        "metric": "msm_g1_pairs_per_s (create_proof schedule, ECDSA k=%d)" % k, "value": value, "unit": "G1 pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
        "config": workload_config(k, world),
        "create_proof_schedule_ms": ms_step,
        "create_proof_schedule_ms_no_ntt_overlap": ms_step_seq,
        "ntt_fr_elements_per_s": (1 << ext_k) / (op_ms["coset_ntt"] / 1e3),
        "msm_only_pairs_per_s": n / (op_ms["msm_uniform"] / 1e3),
        "op_ms": op_ms,
        "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "G1 pairs/s", "ms_per_step": ms_e2e, "ms_per_step_sequential_calls": ms_e2e_seq,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "h2b_assign_columns / h2b_msm_g1_batch / h2b_lagrange_to_coeff_and_extended_batch / h2b_extended_to_coeff with pinned host buffers; transforms driven by a second host thread + context beside the commitment phases (dependency model of step_resident)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=19)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
