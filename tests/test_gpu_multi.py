"""Multi-GPU correctness (needs >= 2 GPUs, skipped otherwise): the NCCL data-parallel gradient of the discriminator step
equals the single-GPU gradient of the whole batch (ref gigagan_pytorch/distributed.py + accelerate's DDP wrapping at
gigagan_pytorch.py:1902-1908: per-rank mean losses, gradients averaged over ranks)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_gradient_equals_single_rank_full_batch(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import ddp_grad_check as H
    out = str(tmp_path / "ddp.pt")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tests", "helpers", "ddp_grad_check.py"), out],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = torch.load(out)
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gan = H.build(dev)
    real, fake = H.batch()
    loss = H.d_grad(gan, real, fake, dev)
    ref = gan.D_opt.grad.cpu()
    assert got["world"] == 2
    # per-rank mean over B/2 samples averaged over 2 ranks == mean over B samples (hinge, multiscale hinge and penalty are
    # all batch means); fp32 with different reduction orders
    assert abs(got["loss"].item() - loss.item()) <= 1e-5 * abs(loss.item())
    rel = (got["grad"] - ref).abs().max().item() / ref.abs().max().item()
    assert rel < 1e-4, rel
