"""Data-parallel step on real GPUs (needs >= 2 visible B200s; skipped on a one-GPU box): the
row-sharded step of training.backward_and_step (reduce-scatter of dW, AdamW on the owned rows,
all-gather of the bf16 weights) against the single-GPU step on the same global batch.
tools/dp_check.py is the worker; this test launches it under torchrun."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["nvls-side", "nvls-side-bf16", "peer-side-bf16", "peer", "peer-side", "nvls", "nccl"])
def test_row_sharded_data_parallel_step_equals_single_gpu_step(mode, world):
    """nvls-side is the default mode bench.py / the Trainer pick on an NVSwitch box (bench.py also
    runs this comparison at its own world size before timing and prints it as `dp_parity`).
    peer: gradient rows / bf16 weights move through NVLink peer memory inside the AdamW kernel, on
    the compute stream on all SMs; peer-side: the same kernel on a side stream on a few SMs;
    nvls: the same with the gradient summed inside the NVSwitch (multimem.ld_reduce) and the bf16
    rows multicast (multimem.st); nccl: reduce-scatter + all-gather."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and mode not in ("nvls-side", "nvls-side-bf16", "peer-side", "nccl"):
        pytest.skip("inline modes are covered at 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tools", "dp_check.py"), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("dp_check")]
    assert res.returncode == 0 and lines and lines[-1].endswith("-> OK"), res.stdout[-2000:] + res.stderr[-2000:]
