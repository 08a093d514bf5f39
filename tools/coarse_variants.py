"""Time the level-1 kernel variants on one synthetic problem: LCB_COARSE_VARIANT = 2 (two accumulators, three tile
slots), 3 (three accumulators, two tile slots), 4 (3 + split-halves MMA order)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libcluster_b200 as lc  # noqa: E402


def main(N=4_000_000, K=64, D=128):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    mu = torch.rand(K, D, device=dev, generator=g) * 20 - 10
    z = torch.randint(0, K, (N,), device=dev, generator=g).to(torch.int32)
    X = mu[z.long()] + torch.randn(N, D, device=dev, generator=g)
    for variant in sys.argv[1:] or ["2", "3", "4"]:
        os.environ["LCB_COARSE_VARIANT"] = variant
        eng = lc.Engine(0, lc.F32)
        eng.set_data_device(X.data_ptr(), N, D, D)
        eng.model_init(lc.BGMM)
        eng.set_labels_device(z.data_ptr(), K)
        out = []
        for _ in range(2):
            F = eng.vbem_step()
            d = eng.estep_detail()
            out.append("%.3f" % d["coarse_ms"])
        print("variant", variant, "coarse ms", out, "F %.10g" % F, "pairs", d["pairs"], flush=True)
        eng.close()


if __name__ == "__main__":
    main()
