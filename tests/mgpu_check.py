"""Run under torchrun on >= 2 GPUs (not collected by pytest): row-sharded VB iterations with the
NCCL all-reduce must reproduce the single-GPU F trace and statistics."""
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
        print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)


if __name__ == "__main__":
    main()
