"""GPU (-m gpu), needs >= 2 devices: tests/mgpu_check.py under torch.distributed.run with one rank per GPU over NCCL --
the row-sharded iteration (device M step: one statistics all-reduce + 16 bytes per iteration), a grouped model with
whole groups per rank, a full learn() with splits, and the same through the host M step with the factorisations split
over the ranks (LCB_HOST_MSTEP=1 LCB_DIST_MSTEP=1).  Skipped on a single-GPU box; logs of 2/4/8-GPU runs are kept
under profiles/."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["device_mstep", "host_mstep_split"])
def test_sharded_fits_match_single_gpu(mode):
    n = _gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    env = dict(os.environ)
    env.pop("LCB_HOST_MSTEP", None)
    env.pop("LCB_DIST_MSTEP", None)
    if mode == "host_mstep_split":
        env["LCB_HOST_MSTEP"] = "1"
        env["LCB_DIST_MSTEP"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert "MGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
