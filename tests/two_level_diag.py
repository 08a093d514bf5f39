"""Diagnostics of the two-level E step on a GPU box (not a pytest file): prints, never asserts.

    python tests/two_level_diag.py [--big]

For a few data sets: candidates per row, tightness of the level-1 bound, agreement of every level with the dense
tensor-core kernel and with the fp64 engine; with --big also the per-level device times at N = 4M, K = 64.
"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import libcluster_b200 as lc  # noqa: E402
from conftest import make_blobs, soft_labels  # noqa: E402

D = 128


def engine(two_level, stage=None, prec=lc.F32):
    os.environ["LCB_TC_TWO_LEVEL"] = "1" if two_level else "0"
    if stage:
        os.environ["LCB_TC_STAGE"] = stage
    try:
        return lc.Engine(0, prec)
    finally:
        os.environ.pop("LCB_TC_TWO_LEVEL", None)
        os.environ.pop("LCB_TC_STAGE", None)


def run(X, q0, **kw):
    eng = engine(**kw)
    eng.set_data(X)
    eng.model_init(lc.BGMM)
    eng.set_qz(q0)
    F, _ = eng.vbem(maxit=0)
    q = eng.qZ(0)
    det = eng.estep_detail()
    eng.close()
    return F, q, det


def case(N, K, spread):
    print("=== N=%d K=%d spread=%g" % (N, K, spread), flush=True)
    X, z = make_blobs(N, D, K, seed=N + K, spread=spread)
    q0 = soft_labels(z, K, seed=K)
    F_d, q_d, _ = run(X, q0, two_level=False)
    _, q64, _ = run(X, q0, two_level=False, prec=lc.F64)
    print("dense vs fp64: max|dq| = %.3g" % np.abs(q_d - q64).max(), flush=True)
    try:
        _, ub, det1 = run(X, q0, two_level=True, stage="coarse")
        cand = np.isfinite(ub)
        print("level 1:", det1, "cand/row mean %.3f max %d rows-without %d nan %d" % (
            cand.sum(1).mean(), cand.sum(1).max(), (~cand.any(1)).sum(), np.isnan(ub).sum()), flush=True)
        with np.errstate(divide="ignore"):
            rel64 = np.log(q64) - np.log(q64.max(1, keepdims=True))
        if (~cand).any():
            print("  non-candidates: max log(q/qbest) = %.3f (must be <= -24)" % rel64[~cand].max(), flush=True)
        _, lg, det2 = run(X, q0, two_level=True, stage="refine")
        same = (np.isfinite(lg) == cand).all()
        both = cand & np.isfinite(lg)
        slack = ub[both] - lg[both]
        print("level 2:", det2, "mask same", same, "UB-exact: min %.4g median %.4g max %.4g" % (
            slack.min(), np.median(slack), slack.max()), flush=True)
        L = np.where(both, lg, -np.inf)
        m = L.max(1, keepdims=True)
        e = np.exp(L - m)
        sm = e / e.sum(1, keepdims=True)
        print("  softmax(level 2) vs dense: max|dq| = %.3g" % np.abs(sm - q_d).max(), flush=True)
        # how good is the one-product distance itself?  (UB is the centre of the bracket plus the bound)
        F_t, q_t, det = run(X, q0, two_level=True)
        print("all levels:", det, "max|dq| vs dense %.3g vs fp64 %.3g  dF/F %.3g rowsum err %.3g" % (
            np.abs(q_t - q_d).max(), np.abs(q_t - q64).max(), abs(F_t - F_d) / abs(F_d), np.abs(q_t.sum(1) - 1).max()),
            flush=True)
    except Exception:  # noqa: BLE001
        traceback.print_exc()


def big(N=4_000_000, K=64):
    import torch
    print("=== timing N=%d K=%d" % (N, K), flush=True)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    mu = (torch.rand(K, D, device=dev, generator=g) * 20 - 10)
    z = torch.randint(0, K, (N,), device=dev, generator=g).to(torch.int32)
    A = torch.randn(K, D, D, device=dev, generator=g)
    Lc = torch.linalg.cholesky(A @ A.transpose(1, 2) / D + 0.5 * torch.eye(D, device=dev))
    X = torch.empty(N, D, device=dev)
    for k in range(K):
        idx = (z == k).nonzero().squeeze(1)
        X[idx] = mu[k] + torch.randn(idx.numel(), D, device=dev, generator=g) @ Lc[k].T
    torch.cuda.synchronize()
    res = {}
    for two in (False, True):
        eng = engine(two)
        eng.set_data_device(X.data_ptr(), N, D, D)
        eng.model_init(lc.BGMM)
        eng.set_labels_device(z.data_ptr(), K)
        Fs = []
        for i in range(4):
            t0 = time.perf_counter()
            Fs.append(eng.vbem_step())
            dt = time.perf_counter() - t0
            t, d = eng.step_timing(), eng.estep_detail()
            print("two_level=%s step %d: F=%.10g wall %.1f ms | S %.2f E %.2f step %.2f launches %d | "
                  "coarse %.3f lists %.3f refine %.3f finalize %.3f pairs %d path %d" % (
                      two, i, Fs[-1], dt * 1e3, t["sstat_ms"], t["estep_ms"], t["step_ms"], t["launches"],
                      d["coarse_ms"], d["lists_ms"], d["refine_ms"], d["finalize_ms"], d["pairs"], d["path"]), flush=True)
        res[two] = Fs
        eng.close()
    base = res[False]
    for key, Fs in res.items():
        print(key, "F", Fs, "rel diff vs dense", [abs(a - b) / abs(a) for a, b in zip(base, Fs)], flush=True)


if __name__ == "__main__":
    for c in [(5000, 16, 4.0), (4099, 10, 1.0), (20000, 33, 3.0), (3000, 64, 6.0), (1500, 8, 0.3)]:
        case(*c)
    if "--big" in sys.argv:
        try:
            big()
        except Exception:  # noqa: BLE001
            traceback.print_exc()
