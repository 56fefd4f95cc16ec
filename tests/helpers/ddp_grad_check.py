"""Helper of tests/test_gpu_multi.py (launched by torchrun, one rank per GPU): every rank runs the discriminator objective
(hinge + multiscale hinge + gradient penalty) of trainer.GigaGAN on ITS SLICE of a fixed batch with fixed fake images,
all-reduces the flat gradient buffer over NCCL exactly as train_discriminator_step does (SUM, then 1/world folded into
AdamW's grad_scale) and rank 0 writes the averaged gradient.  The single-process test compares it with the gradient of
the whole batch on one GPU (the data-parallel identity of SURVEY.md section 4)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def build(dev):
    import gigagan_pytorch_b200 as g
    g.set_compute_dtype(torch.float32)
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=64, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8,
               self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=64, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(32, 16, 8))
    torch.manual_seed(0)
    gan = g.GigaGAN(generator=gen, discriminator=disc, log_steps_every=10 ** 9, create_ema_generator_at_init=False,
                    discr_aux_recon_loss_weight=0., save_and_sample_every=0).to(dev)
    return gan


def batch(n=8):
    real = torch.rand(n, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    fake = torch.rand(n, 3, 64, 64, generator=torch.Generator().manual_seed(4)) * 2 - 1
    return real, fake


def d_grad(gan, real, fake, dev):
    """flat D gradient of the D objective on (real, fake), through the trainer's own _d_objective"""
    from gigagan_pytorch_b200 import ops
    gan._ensure_optimizers()
    fake_n = ops.to_nhwc(fake.to(dev), 3, torch.float32)
    gan._generate = lambda noise, real_n=None, text=None: (fake_n, gan.D.real_images_to_rgbs_nhwc(fake_n))
    gan._begin_work(gan._stale_banks())
    gan.D_opt.zero_grad()
    total, _ = gan._d_objective(real.to(dev), None, True, True)
    total.backward(inputs=gan.D_opt.params)
    return total.detach()


if __name__ == "__main__":
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gan = build(dev)                       # _ensure_optimizers broadcasts rank 0's weights like DDP
    real, fake = batch()
    per = real.shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    loss = d_grad(gan, real[sl], fake[sl], dev)
    gan.D_opt.all_reduce_grads()
    g = gan.D_opt.grad / world
    lt = loss.clone()
    dist.all_reduce(lt)
    if rank == 0:
        torch.save(dict(grad=g.cpu(), loss=(lt / world).cpu(), world=world), out)
    dist.barrier()
    dist.destroy_process_group()
