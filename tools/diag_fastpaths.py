"""Per-parameter deviation of a plain discriminator step's fast paths from the composed forms (the body of
tests/test_gpu_readme256.py::test_plain_step_fast_paths_match_composed), to locate a failing tensor.
usage: python tools/diag_fastpaths.py [fp32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_readme256 as T
import gigagan_pytorch_b200 as g
from gigagan_pytorch_b200 import ops
from gigagan_pytorch_b200.modules import img_cpad

dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
fx = T.fixture()
G, D = T.build_models(dtype)
D.train()
dev = T.dev()
img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(fx["seeds"]["img"])).to(dev)
fake = (torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(fx["seeds"]["fake"])) * 2 - 1).to(dev)
gan = g.GigaGAN(generator=G, discriminator=D, amp=(dtype == torch.bfloat16), mixed_precision_type="bf16",
                log_steps_every=10 ** 9, create_ema_generator_at_init=False, discr_aux_recon_loss_weight=0.).to(dev)
gan._ensure_optimizers()
fake_n = ops.to_nhwc(fake, img_cpad(3), dtype)
gan._generate = lambda noise, real_n=None, text=None: (fake_n, D.real_images_to_rgbs_nhwc(fake_n))
names = {id(p): n for n, p in D.named_parameters()}


def run(composed, sinks):
    gan._force_composed = composed
    for p in gan.D_opt.params:
        p._gg_sink, p._gg_sink1 = (p.ndim == 4 and sinks), (p.ndim == 1 and sinks)
    gan._begin_work(gan._stale_banks())
    gan.D_opt.zero_grad()
    total, _ = gan._d_objective(img, None, False, True)
    total.backward(inputs=gan.D_opt.params)
    torch.cuda.synchronize()
    return total.detach().float().clone(), gan.D_opt.grad.clone()


ref = run(True, False)
for label, composed, sinks in (("fast paths + sinks", False, True), ("fast paths, no sinks", False, False),
                               ("composed + sinks", True, True), ("composed again", True, False)):
    l, gr = run(composed, sinks)
    gmax = ref[1].abs().max().item()
    rows = []
    for p, off in zip(gan.D_opt.params, gan.D_opt._offsets()):
        a, b = gr[off:off + p.numel()], ref[1][off:off + p.numel()]
        rows.append(((a - b).abs().max().item() / gmax, (a - b).abs().max().item() / (b.abs().max().item() + 1e-30),
                     names.get(id(p), "?"), tuple(p.shape)))
    rows.sort(reverse=True)
    print(f"== {label}: loss {l.item():.6f} vs {ref[0].item():.6f}; worst tensors (diff/global max, diff/own max, name, shape)")
    for r in rows[:8]:
        print("   %.3e  %.3e  %s %s" % r)
