"""CPU, world_size 2 over gloo: the multi-rank path of one VB iteration without a GPU.

Each rank owns a contiguous row shard (lcb_shard_rows), produces the packed sufficient
statistics of its rows (here with the oracle standing in for the CUDA pass, which needs a
GPU), all-reduces the packed buffer -- the one exchange step of the design -- and runs the
product's host M-step (lcb_host_mstep).  Every rank must end with bit-identical posteriors
that match the single-process oracle on the full data."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_blobs, soft_labels


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, model, omodel, X, q0, groups_cut, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import libcluster_b200 as lc
    from oracle import pyoracle as po

    N, D = X.shape
    K = q0.shape[1]
    J = len(groups_cut) - 1
    diag = omodel in (po.DGMM, po.DGMC)
    S = D if diag else D * D
    b, e = lc.shard_rows(N, rank, world)
    packed = np.zeros(lc.packed_len(model, J, K, D))
    blk = 1 + D + S
    ckind = po.C_NORMGAMMA if diag else po.C_GAUSSWISH
    for j in range(J):                       # groups may straddle the rank boundary
        lo, hi = max(b, groups_cut[j]), min(e, groups_cut[j + 1])
        if hi > lo:
            packed[j * K:(j + 1) * K] = q0[lo:hi].sum(0)
    for k in range(K):
        c = po.Cluster(ckind, 1.0, D)
        if e > b:
            c.addobs(q0[b:e, k], X[b:e])
        st = c.state()
        o = J * K + k * blk
        packed[o] = st["N_s"]
        packed[o + 1:o + 1 + D] = st["x_s"]
        packed[o + 1 + D:o + blk] = st["xx_s"].ravel()
    t = torch.from_numpy(packed)
    dist.all_reduce(t)                        # the single exchange step per VB iteration
    F, elogw, means, covs = lc.host_mstep(model, packed, J, K, D)
    ret[rank] = (F, elogw.copy(), means.copy(), covs.copy(), (b, e))
    dist.destroy_process_group()


@pytest.mark.parametrize("mname", ["BGMM", "VDP", "DGMM", "GMC"])
def test_two_rank_iteration_matches_single_process_oracle(mname):
    import libcluster_b200 as lc
    from oracle import pyoracle as po
    model, omodel = {"BGMM": (lc.BGMM, po.BGMM), "VDP": (lc.VDP, po.VDP), "DGMM": (lc.DGMM, po.DGMM),
                     "GMC": (lc.GMC, po.GMC)}[mname]
    X, z = make_blobs(501, 4, 3, seed=8, diag=(mname == "DGMM"))
    q0 = soft_labels(z, 3, seed=8)
    cuts = [0, 200, 333, 501] if mname == "GMC" else [0, 501]
    groups = [X[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    m = po.Model(omodel, groups)
    m.vbem(q0, maxit=0)
    Fref = sum(m.weights_fenergy(j) for j in range(len(groups))) + sum(m.cluster(k)["fenergy"] for k in range(3))

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), model, omodel, X, q0, cuts, ret), nprocs=world, join=True)
    assert set(ret.keys()) == {0, 1}
    (F0, e0, m0, c0, s0), (F1, e1, m1, c1, s1) = ret[0], ret[1]
    assert s0[1] == s1[0] and s0[0] == 0 and s1[1] == 501
    # replicated M-step: bit-identical on every rank
    assert F0 == F1 and np.array_equal(m0, m1) and np.array_equal(c0, c1) and np.array_equal(e0, e1)
    assert F0 == pytest.approx(Fref, rel=1e-10)
    for k in range(3):
        assert np.allclose(m0[k], m.cluster(k)["m"], rtol=1e-10)
    for j in range(len(groups)):
        assert np.allclose(e0[j], m.weights(j)[0], rtol=1e-10, atol=1e-12)
