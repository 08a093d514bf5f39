"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multirank.py launches it when the box has them): row-sharded VB
iterations with the NCCL all-reduce must reproduce the single-GPU F trace, posteriors and qZ -- flat models, a
grouped model with whole groups per rank (SURVEY.md 8e, config 4's partitioning) and a full learn() with splits.
LCB_HOST_MSTEP=1 LCB_DIST_MSTEP=1 in the environment runs the same checks through the host M step with the
factorisations split over the ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import libcluster_b200 as lc  # noqa: E402
from conftest import make_blobs, soft_labels  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (model, D, K, N, diag) in [(lc.BGMM, 128, 6, 30011, False), (lc.BGMM, 128, 16, 60011, False), (lc.VDP, 16, 4, 9001, False), (lc.DGMM, 24, 5, 12345, True)]:
        X, z = make_blobs(N, D, K, seed=N, spread=4.0, diag=diag)
        q0 = soft_labels(z, K, seed=3)
        ref = None
        if rank == 0:
            e1 = lc.Engine(local, lc.F32)
            e1.set_data(X); e1.model_init(model); e1.set_qz(q0); e1.vbem(maxit=3)
            ref = (e1.trace()[0], np.stack([e1.cluster(k)["mean"] for k in range(K)]), e1.qZ(0))
            e1.close()
        eng = lc.Engine(local, lc.F32)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lc.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)
        b, e = lc.shard_rows(N, rank, world)
        eng.set_data(X[b:e]); eng.model_init(model); eng.set_qz(q0[b:e]); eng.vbem(maxit=3)
        F = eng.trace()[0]
        means = np.stack([eng.cluster(k)["mean"] for k in range(K)])
        q = eng.qZ(0)
        if rank == 0:
            good = (len(F) == len(ref[0]) and np.allclose(F, ref[0], rtol=1e-6) and np.allclose(means, ref[1], atol=1e-5)
                    and np.abs(q - ref[2][b:e]).max() < 1e-5)
            print("model", model, "D", D, "F", F[-1], "ref", ref[0][-1], "OK" if good else "MISMATCH", flush=True)
            ok = ok and good
        eng.close()
    # grouped model (learnGMC: GDirichlet weights per group, shared GaussWish clusters), whole groups per rank:
    # every rank passes all J groups, the ones it does not own with zero rows
    for (D, K, J, N) in [(128, 9, 6, 24000), (5, 4, 7, 7000)]:
        X, z = make_blobs(N, D, K, seed=J, spread=4.0)
        cuts = np.linspace(0, N, J + 1).astype(int)
        groups = [X[cuts[j]:cuts[j + 1]] for j in range(J)]
        q0 = soft_labels(z, K, seed=4)
        ref = None
        if rank == 0:
            e1 = lc.Engine(local, lc.F32)
            e1.set_data(groups); e1.model_init(lc.GMC); e1.set_qz(q0); e1.vbem(maxit=3)
            ref = (e1.trace()[0], [e1.group_weights(j)[1] for j in range(J)], e1.qZ())
            e1.close()
        eng = lc.Engine(local, lc.F32)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lc.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)
        jb, je = lc.shard_rows(J, rank, world)
        mine = [groups[j] if jb <= j < je else groups[j][:0] for j in range(J)]
        q0m = np.concatenate([q0[cuts[j]:cuts[j + 1]] for j in range(jb, je)] + [q0[:0]], 0)
        eng.set_data(mine); eng.model_init(lc.GMC); eng.set_qz(q0m); eng.vbem(maxit=3)
        F = eng.trace()[0]
        ew = [eng.group_weights(j)[1] for j in range(J)]
        qs = eng.qZ()
        good = True
        if rank == 0:
            dew = max(np.abs(a - b).max() for a, b in zip(ew, ref[1]))
            dq = max([np.abs(qs[j] - ref[2][j]).max() for j in range(jb, je) if qs[j].size] + [0.0])
            # two fp32 runs with different summation orders: each is within 1e-5 of the oracle
            good = good and len(F) == len(ref[0]) and np.allclose(F, ref[0], rtol=1e-6) and dew < 2e-5 and dq < 2e-5
            print("GMC D", D, "K", K, "J", J, "F", F[-1], "ref", ref[0][-1], "max dElogw %.2e max dq %.2e" % (dew, dq),
                  "OK" if good else "MISMATCH", flush=True)
            ok = ok and good
        eng.close()
    # full learn with splits, sharded
    X, _ = make_blobs(4000, 3, 4, seed=5, spread=8.0)
    eng = lc.Engine(local, lc.F32)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(lc.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    eng.comm_init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)
    b, e = lc.shard_rows(4000, rank, world)
    eng.set_data(X[b:e])
    F = eng.learn(lc.BGMM)
    if rank == 0:
        e1 = lc.Engine(local, lc.F32)
        e1.set_data(X)
        F1 = e1.learn(lc.BGMM)
        good = eng.K == e1.K and abs(F - F1) <= 1e-5 * abs(F1)
        print("learn sharded K", eng.K, "F", F, "single K", e1.K, "F", F1, "OK" if good else "MISMATCH", flush=True)
        ok = ok and good
        e1.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL", "world", world, "host_mstep", os.environ.get("LCB_HOST_MSTEP", "0"),
              "dist_mstep", os.environ.get("LCB_DIST_MSTEP", "auto"), flush=True)


if __name__ == "__main__":
    main()
