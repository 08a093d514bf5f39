#!/usr/bin/env python
"""bench.py -- VB E-step throughput (points/s) of one VB iteration at fixed K.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W      # the reference's CPU path

A "step" is one loop body of vbem() (src/cluster.cpp:203-226) over the whole
synthetic matrix at fixed K: sufficient statistics of the current
responsibilities -> posterior update -> E-step (new responsibilities) -> F.
Workload (BASELINE.json metric): N = 50M rows, D = 128, K = 64 full-covariance
(GaussWish) clusters, Dirichlet weights (learnBGMM's pair); rows are sharded
over the ranks (strong scaling: N is the total), one all-reduce of the packed
statistics per iteration.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 250_000          # rows per generation chunk; shards are whole chunks
SEED = 20260925
SOFT_SPREADS = [0.65, 0.3]  # means U(-s, s)^D of the extra soft-regime points: about 4 and 17 candidate pairs per row
                            # (profiles/soft_sweep_r02_4m.log)
# fallback only if MEASURED_PEAKS.json is absent (/opt/skills/guides/B200_PROFILING.md)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-points", type=int, default=50_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--clusters", type=int, default=64)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--model", default="bgmm", choices=["bgmm", "dgmm", "gmc"],
                    help="bgmm: full covariance (headline); dgmm: diagonal (config 5); gmc: grouped, GDirichlet weights per "
                         "group and shared GaussWish clusters, whole groups per rank (config 4: --groups 256 --n-points "
                         "25600000 --dim 64 --clusters 64)")
    ap.add_argument("--groups", type=int, default=256, help="groups of --model gmc (rows are split evenly over them)")
    ap.add_argument("--fit", action="store_true",
                    help="config 3: time a whole learnVDP fit from K = 1 with greedy splits (lcb_learn) instead of "
                         "steady-state iterations; prints fit seconds, final K and F")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --n-points rows per GPU instead of in total")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--spread", type=float, default=10.0,
                    help="cluster means ~ U(-spread, spread)^D (SURVEY 8d: 10); small values make clusters overlap")
    ap.add_argument("--soft-spreads", default="auto",
                    help="comma-separated spreads of the extra soft-regime points (N=1 only; 'none' to skip)")
    ap.add_argument("--soft-rows", type=int, default=8_000_000, help="rows of each soft-regime point")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            d["_source"] = "measured"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


# ------------------------------------------------------------------ data ---
def mixture_params(D, K, diag=False, spread=10.0):
    """means U(-10,10)^D, covariances A A^T / D + 0.5 I (or diag(U(0.5,2))), weights Dirichlet(5).
    `spread` replaces the 10: the same unit draws scaled, so the covariances and weights do not change with it."""
    rng = np.random.default_rng(SEED)
    mu = rng.uniform(-10, 10, size=(K, D)) * (spread / 10.0)
    if diag:
        L = np.sqrt(rng.uniform(0.5, 2.0, size=(K, D)))     # per-dimension standard deviations
    else:
        L = np.empty((K, D, D))
        for k in range(K):
            A = rng.normal(size=(D, D))
            L[k] = np.linalg.cholesky(A @ A.T / D + 0.5 * np.eye(D))
    w = rng.dirichlet(5.0 * np.ones(K))
    return mu, L, w


def gen_chunk_torch(torch, dev, c, rows, D, K, mu_t, L_t, w_t):
    g = torch.Generator(device=dev).manual_seed(SEED + 1 + c)
    z = torch.multinomial(w_t, rows, replacement=True, generator=g).to(torch.int32)
    e = torch.randn(rows, D, device=dev, generator=g)
    x = torch.empty(rows, D, device=dev, dtype=torch.float32)
    zl = z.long()
    if L_t.dim() == 2:                                      # diagonal covariances
        return (mu_t[zl] + e * L_t[zl]).contiguous(), z
    order = torch.argsort(zl)
    counts = torch.bincount(zl, minlength=K).tolist()
    o = 0
    for k in range(K):
        n = counts[k]
        if n:
            idx = order[o:o + n]
            x[idx] = mu_t[k] + e[idx] @ L_t[k].T
            o += n
    return x, z


def gen_rows_numpy(first_rows, D, K):
    """The same generator family on the host for the CPU sample (independent stream)."""
    mu, L, w = mixture_params(D, K)
    rng = np.random.default_rng(SEED + 7)
    z = rng.choice(K, size=first_rows, p=w)
    X = mu[z] + np.einsum("nd,ned->ne", rng.normal(size=(first_rows, D)), L[z])
    return X, z


# ------------------------------------------------------- clocks sampling ---
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------- CPU baseline ---
def _ref_available():
    try:
        from oracle import pyref
        return pyref.available()
    except Exception:
        return False


def cpu_iteration_rate(D, K, budget_s, reps=1, want="auto"):
    """One vbem iteration (src/cluster.cpp:203-226) of the reference's CPU path on a bounded sample of the same
    mixture.  kind "reference": the reference's own sources (oracle/_ref, built against the Eigen/Boost stand-ins of
    oracle/refshim -- unoptimised GEMM/solve, so slower than real Eigen would be); kind "port": oracle/vb_oracle.c.
    The J=1 entry points are effectively single-threaded in the reference: only
    clusters[k].update() runs under OpenMP, so `cores` is the thread count offered, not a speed-up claim."""
    from oracle import pyoracle as po
    kind = "reference" if (want in ("auto", "reference") and _ref_available()) else "port"
    threads = os.cpu_count() or 1

    def one(X, q):
        if kind == "reference":
            from oracle import pyref
            t0 = time.perf_counter()
            r = pyref.vbem(po.BGMM, [X], q, maxit=0, nthreads=threads)
            return time.perf_counter() - t0, r.qZ[0]
        m = po.Model(po.BGMM, [X])
        t0 = time.perf_counter()
        m.vbem(q, maxit=0)
        return time.perf_counter() - t0, m.qZ()

    Xs, zs = gen_rows_numpy(1024, D, K)
    q0 = np.zeros((1024, K)); q0[np.arange(1024), zs] = 1.0
    dt, _ = one(Xs, q0)
    per_row = dt / 1024
    rows = int(min(1 << 17, max(2048, budget_s / max(per_row, 1e-9) / max(reps, 1))))
    X, z = gen_rows_numpy(rows, D, K)
    q = np.zeros((rows, K)); q[np.arange(rows), z] = 1.0
    _, q = one(X, q)                          # responsibilities now soft, like a steady-state iteration
    times = []
    for _ in range(reps):
        dt, q = one(X, q)
        times.append(dt)
    return rows, times, kind, (threads if kind == "reference" else 1)


def blas_estimate(D, K, rows=8192):
    """What the stand-in costs: the two O(N K D^2) halves of one iteration (src/distributions.cpp:301-313 addobs GEMM,
    src/probutils.cpp:113-138 mahaldist solve) for `rows` rows with numpy's BLAS/LAPACK on all host threads -- the speed
    a build of the reference against real Eigen (+ a threaded BLAS) could approach.  Context for cpu_baseline only."""
    import scipy.linalg as sla
    rng = np.random.default_rng(0)
    X = rng.normal(size=(rows, D))
    q = rng.uniform(size=(rows, K))
    A = rng.normal(size=(D, D))
    Lc = np.linalg.cholesky(A @ A.T / D + np.eye(D))
    t0 = time.perf_counter()
    for k in range(K):
        (X * q[:, k:k + 1]).T @ X                                  # xx_s += qZkX^T X
        y = sla.solve_triangular(Lc, (X - 0.1).T, lower=True)      # LDLT solve of mahaldist
        (y * y).sum(0)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "points/s", "rows": rows,
            "what": "numpy/scipy BLAS timing of the GEMM and triangular-solve halves of one iteration (not the reference)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, K = a.dim, a.clusters
    total = a.steps + a.warmup
    rows, times, kind, cores = cpu_iteration_rate(D, K, budget_s=150.0, reps=total)
    t = times[a.warmup:] if len(times) > a.warmup else times
    sec = float(np.mean(t))
    val = rows / sec
    what = ("reference sources (src/cluster.cpp vbem<Dirichlet,GaussWish>) built against the Eigen/Boost stand-ins"
            if kind == "reference" else "restated reference (oracle/vb_oracle.c)")
    line = {
        "impl": "reference", "metric": "VB E-step points/sec at N=50M D=128 K=64", "value": val, "unit": "points/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "VB iteration (SS + M + E + F), BGMM full-cov, N=%d D=%d K=%d" % (a.n_points, D, K),
                   "sample_rows": rows, "note": what + "; the J=1 entry points run the O(N) loops on one core "
                   "(SURVEY.md section 2)"},
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": kind,
                         "sample": "%d rows of the same synthetic mixture, one vbem iteration each step" % rows},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------- soft regime ---
def soft_point(torch, lc, dev, D, K, rows, spread, steps=3, warmup=3):
    """One extra measured point on a mixture whose clusters overlap (means U(-spread, spread)^D): the two-level E pass
    and the S pass cost grow with the candidate pairs per row, which the headline mixture (spread 10) keeps at 1.0;
    the reference's cost does not depend on the data (src/cluster.cpp:75-79,120-121)."""
    mu, L, w = mixture_params(D, K, False, spread)
    mu_t = torch.tensor(mu, dtype=torch.float32, device=dev)
    L_t = torch.tensor(L, dtype=torch.float32, device=dev)
    w_t = torch.tensor(w, dtype=torch.float32, device=dev)
    X = torch.empty(rows, D, dtype=torch.float32, device=dev)
    z = torch.empty(rows, dtype=torch.int32, device=dev)
    for c, o in enumerate(range(0, rows, CHUNK)):
        n = min(CHUNK, rows - o)
        xc, zc = gen_chunk_torch(torch, dev, 100000 + c, n, D, K, mu_t, L_t, w_t)
        X[o:o + n] = xc
        z[o:o + n] = zc
    torch.cuda.synchronize()
    eng = lc.Engine(dev.index or 0, lc.F32)
    eng.set_data_device(X.data_ptr(), rows, D, D)
    eng.model_init(lc.BGMM)
    eng.set_labels_device(z.data_ptr(), K)
    for _ in range(warmup):
        eng.vbem_step()
    ms = s_ms = e_ms = 0.0
    pairs, paths, F = 0, [], None
    for _ in range(steps):
        F = eng.vbem_step()
        t = eng.step_timing()
        ms += t["step_ms"]; s_ms += t["sstat_ms"]; e_ms += t["estep_ms"]
        d = eng.estep_detail()
        pairs += d["pairs"]
        paths.append(d["path"])
    eng.close()
    del X, z
    torch.cuda.empty_cache()
    name = {0: "dense", 1: "two-level", 2: "two-level abandoned -> dense"}
    return {"spread": spread, "rows": rows, "value": rows / (ms / steps * 1e-3), "ms_per_step": ms / steps,
            "sstat_ms": s_ms / steps, "estep_ms": e_ms / steps, "candidate_pairs_per_row": pairs / steps / rows,
            "estep_path": name.get(paths[-1], str(paths[-1])), "F_last": F}


# ----------------------------------------------------------------- main ---
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist

    import libcluster_b200 as lc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    N, D, K = a.n_points, a.dim, a.clusters
    if a.weak:
        N *= world
    prec = lc.F32 if a.precision == "f32" else lc.F64
    grouped = a.model == "gmc"
    J = a.groups if grouped else 1
    chunk = (N // J) if grouped else CHUNK          # grouped: one generation chunk per group, whole groups per rank
    if grouped and (N % J or J % world):
        raise SystemExit("--model gmc needs n-points divisible by --groups and --groups divisible by the GPUs")
    nchunks = (N + chunk - 1) // chunk
    c0, c1 = lc.shard_rows(nchunks, rank, world)
    r0, r1 = c0 * chunk, min(N, c1 * chunk)
    nloc = r1 - r0

    eng = lc.Engine(local, prec)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lc.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)

    diag = a.model == "dgmm"
    mu, L, w = mixture_params(D, K, diag, a.spread)
    mu_t = torch.tensor(mu, dtype=torch.float32, device=dev)
    L_t = torch.tensor(L, dtype=torch.float32, device=dev)
    w_t = torch.tensor(w, dtype=torch.float32, device=dev)
    X = torch.empty(nloc, D, dtype=torch.float32, device=dev)
    z = torch.empty(nloc, dtype=torch.int32, device=dev)
    gid = torch.empty(nloc, dtype=torch.int32, device=dev) if grouped else None
    for c in range(c0, c1):
        rows = min(chunk, N - c * chunk)
        wc = w_t
        if grouped:   # per-group mixing weights Dirichlet(0.5) over the shared clusters
            wc = torch.tensor(np.random.default_rng(SEED + 1000 + c).dirichlet(0.5 * np.ones(K)), dtype=torch.float32, device=dev)
        xc, zc = gen_chunk_torch(torch, dev, c, rows, D, K, mu_t, L_t, wc)
        o = c * chunk - r0
        X[o:o + rows] = xc
        z[o:o + rows] = zc
        if grouped:
            gid[o:o + rows] = c
    torch.cuda.synchronize()

    MODEL = lc.DGMM if diag else lc.GMC if grouped else lc.BGMM
    if a.fit:
        return run_fit(torch, dist, lc, eng, X, N, nloc, D, world, rank, dev, a)
    eng.set_data_device(X.data_ptr(), nloc, D, D, gid.data_ptr() if grouped else None, J)
    eng.model_init(MODEL)
    eng.set_labels_device(z.data_ptr(), K)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        eng.vbem_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, s_ms, e_ms, launches, Fs = 0.0, 0.0, 0.0, 0, []
    lv = {"coarse_ms": 0.0, "lists_ms": 0.0, "refine_ms": 0.0, "finalize_ms": 0.0, "pairs": 0, "two_level_steps": 0}
    for _ in range(a.steps):
        Fs.append(eng.vbem_step())
        t = eng.step_timing()
        dev_ms += t["step_ms"]; s_ms += t["sstat_ms"]; e_ms += t["estep_ms"]; launches += t["launches"]
        d = eng.estep_detail()
        if d["path"] == 1:
            lv["two_level_steps"] += 1
            for k_ in ("coarse_ms", "lists_ms", "refine_ms", "finalize_ms", "pairs"):
                lv[k_] += d[k_]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([dev_ms, wall * 1e3, s_ms, e_ms, lv["coarse_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, s_ms, e_ms, coarse_ms = [float(v) for v in tt.tolist()]
    two_level = lv["two_level_steps"] == a.steps
    ms_per_step = dev_ms / a.steps
    value = N / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (device-event time, this run) -------
    pk = peaks()
    tc = (prec == lc.F32 and D in (64, 128) and not diag and not os.environ.get("LCB_DISABLE_TC"))
    flops_half = (3.5 * K * D * nloc) if diag else float(K) * D * D * nloc          # algorithmic flops of either half per launch (2KD^2 total)
    levels = None
    if two_level and coarse_ms >= s_ms:
        # two-level E pass: level 1 evaluates all K*D^2 algorithmic flops of every point (one fp16 product, rigorous
        # bound); levels 2-3 only touch the candidate pairs.  The dominant kernel is level 1.
        kname, kms = "estep_coarse_tc128_kernel", coarse_ms / a.steps
        # ncu --set full at N=4M (profiles/ncu_r01_coarse_v4_raw.csv): dram read 2.12 GB + write 1.07 GB per launch
        # = 799 B / point (algorithmic: 512 B of X, 256 B of level-1 bounds, 8 B of candidate mask)
        # dram__bytes_read + dram__bytes_write of one ncu --set full capture of this kernel at N = 4 M
        # (profiles/ncu_r02_coarse128_summary.txt: 2.1165 GB + 1.0651 GB = 795 B / point; algorithmic: 512 B of X,
        # 256 B of level-1 bounds parked in q, 8 B of candidate mask), scaled to this run's rows
        traffic = (2.116509e9 + 1.065072e9) / 4.0e6 * nloc if D == 128 else None
        levels = {k_: (lv[k_] / a.steps) for k_ in ("coarse_ms", "lists_ms", "refine_ms", "finalize_ms")}
        levels["candidate_pairs_per_row"] = lv["pairs"] / a.steps / max(nloc, 1)
    elif e_ms >= s_ms:
        kname, kms = ("estep_tc128_kernel" if tc else "estep_full_kernel"), e_ms / a.steps
        # measured with ncu --set full at N=2M (profiles/ncu_r01_tc_summary.md): dram read+write per point
        traffic = 756.0 * nloc if (tc and D == 128) else None
    else:
        kname, kms = ("nz_count+nz_fill+sstat_tc128_kernel" if tc else "sstat pass"), s_ms / a.steps
        traffic = None
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    ach_tf = flops_half / (kms * 1e-3) / 1e12
    if diag:
        # diagonal models are bound by the read of X (4 D bytes per point and pass); the dominant kernel is
        # the SIMT E pass (DESIGN.md section 5, config 5 analysis)
        kname, kms = ("estep_diag_kernel", e_ms / a.steps) if e_ms >= s_ms else ("nz lists + sstat_gather_diag_kernel", s_ms / a.steps)
        ach_gbs = 4.0 * D * nloc / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm" if diag else "tensor", "kernel": kname, "achieved": ach_gbs if diag else ach_tf,
                "peak": pk["hbm_gbs"] if diag else peak_tf, "unit": "GB/s" if diag else "TFLOP/s",
                "frac": (ach_gbs / pk["hbm_gbs"]) if diag else ach_tf / peak_tf, "traffic": traffic,
                "peak_source": pk["_source"] + (" hbm_gbs" if diag else " bf16_tflops_sustained"),
                "kernel_ms": kms, "hbm_frac": (4.0 * (D + K) * nloc / (kms * 1e-3) / 1e9) / pk["hbm_gbs"],
                "sstat_ms": s_ms / a.steps, "estep_ms": e_ms / a.steps, "estep_levels": levels,
                "note": ("achieved = algorithmic K*D^2 flop/point (triangular whitening) x points / event time of the "
                         "level-1 kernel of the two-level E pass (one fp16 product per pair, 0.69 x 2 executed tensor "
                         "flop per algorithmic flop incl. the centring chunk); exact logits are recomputed only for "
                         "the candidate pairs (estep_levels)" if levels else
                         ("achieved = algorithmic 4*D bytes/point x points / event time of the dominant pass (diagonal "
                          "model: HBM-bound in principle, SIMT E pass in practice, DESIGN.md section 5)") if diag else
                         "achieved = algorithmic K*D^2 flop/point (triangular whitening) x points / event time of the "
                         "E-pass kernel; the fp16 hi/lo scheme executes 3 x 0.56 x 2 = 3.4 tensor flop per algorithmic "
                         "flop, so tensor-pipe utilisation is higher than frac (ncu: profiles/)")}

    # ---- soft regime: the same step on overlapping mixtures (N=1 only) -------
    soft = None
    if rank == 0 and world == 1 and tc and D == 128 and K == 64 and a.soft_spreads != "none":
        spreads = SOFT_SPREADS if a.soft_spreads == "auto" else [float(v) for v in a.soft_spreads.split(",") if v]
        soft = []
        for sp in spreads:
            try:
                soft.append(soft_point(torch, lc, dev, D, K, min(a.soft_rows, N), sp))
            except Exception as ex:  # noqa: BLE001
                soft.append({"spread": sp, "error": str(ex)[:200]})

    # ---- e2e: the same step through the C ABI from HOST buffers ---------------
    e2e = None
    if not a.no_e2e and not grouped:
        try:
            e2e = run_e2e(torch, dist, lc, eng, X, z, N, nloc, D, K, world, dev, a)
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": "points/s", "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:   # reported at N=1 only; the other ranks would idle in NCCL teardown
        rows, times, kind, cores = cpu_iteration_rate(D, K, budget_s=15.0, reps=1)
        cpu = {"value": rows / times[0], "unit": "points/s", "cores": cores, "kind": kind,
               "sample": "%d rows of the same synthetic mixture, one vbem iteration (fp64; %s)" % (
                   rows, "reference sources + Eigen/Boost stand-ins, oracle/_ref" if kind == "reference"
                   else "oracle/vb_oracle.c")}
        if kind == "reference":
            r2, t2, _, _ = cpu_iteration_rate(D, K, budget_s=10.0, reps=1, want="port")
            cpu["port_value"] = r2 / t2[0]
        try:
            cpu["blas_estimate"] = blas_estimate(D, K)
        except Exception as ex:  # noqa: BLE001
            cpu["blas_estimate"] = {"error": str(ex)[:100]}

    if rank == 0:
        line = {
            "metric": "VB E-step points/sec at N=%s D=%d K=%d" % (
                "50M" if N == 50_000_000 else str(N), D, K), "value": value, "unit": "points/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak" if a.weak else "strong", "vs_baseline": None,
            "dtype": "f32" if prec == lc.F32 else "f64", "data": "synthetic",
            "config": {"workload": "VB iteration (SS + M + E + F), %s, N=%d D=%d K=%d" % (
                "learnDGMM pair (Dirichlet, NormGamma diagonal)" if diag else
                ("learnGMC pair (GDirichlet per group, shared GaussWish), J=%d groups, whole groups per rank" % J) if grouped else
                "learnBGMM pair (Dirichlet, GaussWish full-cov)", N, D, K), "rows_per_gpu": nloc, "parallelism": "rows sharded x%d" % world,
                       "l2": "inputs (%.1f GB/GPU) far larger than L2" % (nloc * D * 4 / 1e9),
                       "timing": "CUDA events on the engine stream around each step, summed, max over ranks",
                       "wall_ms_per_step": wall_ms / a.steps, "F_last": Fs[-1]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            # flat copies of what characterises the data-dependent part of the step
            "candidate_pairs_per_row": (levels or {}).get("candidate_pairs_per_row"),
            "estep_path": "two-level" if two_level else "dense",
            "spread": a.spread, "soft_regime": soft,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_fit(torch, dist, lc, eng, X, N, nloc, D, world, rank, dev, a):
    """Config 3: a whole learnVDP fit (cluster<StickBreak,GaussWish>, src/cluster.cpp:564-629) on the resident rows:
    K = 1, VB to convergence, greedy split search, ... until no split lowers F.  Wall time of lcb_learn."""
    eng.set_data_device(X.data_ptr(), nloc, D, D)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    F = eng.learn(lc.VDP, maxclusters=a.clusters if a.clusters > 0 else -1)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    Ftr, Ktr = eng.trace()
    if rank == 0:
        print(json.dumps({
            "metric": "learnVDP fit seconds at N=%d D=%d (greedy split from K=1)" % (N, D), "value": sec, "unit": "s",
            "n_gpus": world, "higher_is_better": False, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "learnVDP full fit, N=%d D=%d, true clusters %d, maxclusters %d" % (N, D, a.clusters, a.clusters),
                       "rows_per_gpu": nloc},
            "final_K": int(eng.K), "F": F, "vb_iterations": int(len(Ftr)), "points_per_s_per_iteration": N * len(Ftr) / sec,
        }), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, dist, lc, eng, X, z, N, nloc, D, K, world, dev, a):
    """Every step, through the C ABI from HOST buffers: host fp64 X (page-locked) -> lcb_set_data -> labels H2D ->
    one VB iteration -> F on the host -> qZ (fp64, row-major, the caller's matrix: what every learnXXX returns,
    src/cluster.cpp:661,692,723) back on the host.  `value` includes the qZ emit; `iteration_only_value` stops at F."""
    import psutil
    need = nloc * D * 8
    need_q = nloc * K * 8
    if psutil.virtual_memory().available < 1.3 * (need + need_q) + (8 << 30):
        raise RuntimeError("not enough host memory for page-locked copies of X (%d bytes) and qZ (%d bytes)" % (need, need_q))
    Xh = torch.empty(nloc, D, dtype=torch.float64, pin_memory=True)
    step = 1 << 20
    for r in range(0, nloc, step):
        Xh[r:r + step].copy_(X[r:r + step].double())
    zh = torch.empty(nloc, dtype=torch.int32, pin_memory=True)
    zh.copy_(z)
    qh = torch.empty(nloc, K, dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    Xn, qn = Xh.numpy(), qh.numpy()
    zd = torch.empty(nloc, dtype=torch.int32, device=dev)
    times, times_it, times_up = [], [], []
    for i in range(1 + a.e2e_steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        eng.set_data(Xn)
        tu = time.perf_counter()
        eng.model_init(lc.DGMM if a.model == "dgmm" else lc.BGMM)
        zd.copy_(zh, non_blocking=True)
        torch.cuda.synchronize()
        eng.set_labels_device(zd.data_ptr(), K)
        F = eng.vbem_step()
        t1 = time.perf_counter()
        eng.qZ(0, out=qn)
        t2 = time.perf_counter()
        if i > 0:
            times.append(t2 - t0)
            times_it.append(t1 - t0)
            times_up.append(tu - t0)
    rowsum_err = float(np.abs(qn[: 1 << 16].sum(1) - 1.0).max())
    t = torch.tensor([float(np.mean(times)), float(np.mean(times_it))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec, sec_it = [float(v) for v in t.tolist()]
    return {"value": N / sec, "unit": "points/s", "h2d_bytes_per_step": int(need + nloc * 4),
            "d2h_bytes_per_step": int(need_q + 8), "ms_per_step": sec * 1e3, "steps": a.e2e_steps,
            "iteration_only_value": N / sec_it, "iteration_only_ms": sec_it * 1e3, "iteration_only_d2h_bytes": 8,
            "upload_ms": float(np.mean(times_up)) * 1e3, "qz_emit_ms": (sec - sec_it) * 1e3,
            "qz_rowsum_max_err": rowsum_err,
            "path": "Engine.set_data(host fp64) + set_labels + lcb_vbem_step + lcb_get_qz (fp64, row-major, "
                    "page-locked destination) through the C ABI", "F": F}


if __name__ == "__main__":
    main()
