"""CPU checks of host-side logic that needs no GPU: weight-bank tile tables, lazy bank bookkeeping of the trainer, the
bench.py reference arm contract (bounded CPU run of the oracle port)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_weight_bank_tile_table_covers_every_element_once():
    """WeightBank's (entry, o0, i0, TI) blocks tile [O] x [Ipad] exactly once per filter and the forward / flipped
    layout offsets do not overlap (table construction is pure host code; the re-layout kernel itself is GPU-tested)"""
    from gigagan_pytorch_b200 import ops
    shapes = [(3, 64, 1, 1), (32, 3, 3, 3), (40, 3, 7, 7), (64, 32, 3, 3), (2, 33, 16, 3, 3), (96, 160, 2, 2)]
    n = sum(int(torch.tensor(s).prod()) for s in shapes)
    flat = torch.zeros(n)
    params, off = [], 0
    for s in shapes:
        k = int(torch.tensor(s).prod())
        params.append(flat[off:off + k].view(s))
        off += k
    pad = lambda c: 16 if c < 16 else (c + 15) // 16 * 16
    bank = ops.WeightBank(flat, params, torch.bfloat16, pad)
    ent, chunks = bank.entries.tolist(), bank.chunks.tolist()
    assert len(ent) == 4 + 2 + 1                                    # every filter of the 5-D bank is its own entry
    seen = {}
    for e, o0, i0, ti in chunks:
        src, O, I, KK, ipad, fo, bo, _ = ent[e]
        assert 0 <= o0 < O and 0 <= i0 < ipad and ti * KK <= 380 and o0 % 32 == 0 and i0 % ti == 0
        key = (e, o0, i0)
        assert key not in seen
        seen[key] = True
    for e, (src, O, I, KK, ipad, fo, bo, _) in enumerate(ent):
        ti = max(1, min(32, 380 // KK))
        assert sum(1 for k in seen if k[0] == e) == ((O + 31) // 32) * ((ipad + ti - 1) // ti)
    spans = sorted((e[5], e[5] + e[1] * e[3] * e[4]) for e in ent)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))


def test_trainer_refreshes_only_stale_banks():
    from gigagan_pytorch_b200.trainer import GigaGAN

    class Bank:
        def __init__(self):
            self.dirty, self.refreshed = True, 0

        def refresh(self):
            self.refreshed += 1

    class Shell:
        pass

    t = Shell()
    t._banks = [Bank(), Bank()]
    stale = GigaGAN._stale_banks(t)
    assert stale == (True, True) and not any(b.dirty for b in t._banks)
    GigaGAN._begin_work(t, stale)
    assert [b.refreshed for b in t._banks] == [1, 1]
    t._banks[1].dirty = True                                         # e.g. D_opt.step()
    stale = GigaGAN._stale_banks(t)
    assert stale == (False, True)
    GigaGAN._begin_work(t, stale)
    assert [b.refreshed for b in t._banks] == [1, 2]
    assert GigaGAN._stale_banks(t) == (False, False)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unmodified reference from baseline/_ref on the host cores when installed, else
    the oracle port; bounded) at a tiny size: one JSON line with the contract keys; never touches CUDA"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--image-size", "64",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=280, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    has_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gigagan_pytorch"))
    assert line["cpu_baseline"]["kind"] == ("reference" if has_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_optimizer_state_dict_interchanges_with_the_reference_optimizer():
    """FlatAdamW.state_dict()/load_state_dict() speak torch.optim.AdamW's format as the reference builds it
    (optimizer.py:10-34: decayed parameters first, then the rest) - checkpoints move both ways (gp.py:2039-2107)"""
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import import_reference
    ref = import_reference()
    from gigagan_pytorch.optimizer import get_optimizer as ref_get_optimizer
    from gigagan_pytorch_b200.trainer import FlatAdamW
    import gigagan_pytorch_b200 as g
    cfg = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    torch.manual_seed(0)
    Gr = ref.Generator(**cfg)
    torch.manual_seed(0)
    Gm = g.Generator(**cfg)
    opt_r = ref_get_optimizer(Gr.parameters(), lr=2e-4, betas=(0.5, 0.9), weight_decay=0.)   # as gp.py:1982 calls it
    for step in range(3):
        for k, p in enumerate(Gr.parameters()):
            p.grad = torch.randn(p.shape, generator=torch.Generator().manual_seed(100 * step + k))
        opt_r.step()
    opt_m = FlatAdamW(Gm, lr=2e-4, betas=(0.5, 0.9))
    opt_m.load_state_dict(opt_r.state_dict())                       # reference -> this trainer
    assert int(opt_m.step_t) == 3
    off = 0
    by_param = {id(p): st for grp in opt_r.param_groups for p in grp["params"] for st in [opt_r.state[p]]}
    for pr, pm in zip(Gr.parameters(), opt_m.params):
        st = by_param[id(pr)]
        n = pm.numel()
        assert torch.equal(opt_m.m[off:off + n].view(pm.shape), st["exp_avg"])
        assert torch.equal(opt_m.v[off:off + n].view(pm.shape), st["exp_avg_sq"])
        off += n
    opt_r2 = ref_get_optimizer(Gr.parameters(), lr=2e-4, betas=(0.5, 0.9), weight_decay=0.)
    opt_r2.load_state_dict(opt_m.state_dict())                      # this trainer -> reference
    for pr in Gr.parameters():
        a, b = opt_r.state[pr], opt_r2.state[pr]
        assert torch.equal(a["exp_avg"], b["exp_avg"]) and torch.equal(a["exp_avg_sq"], b["exp_avg_sq"])
        assert float(a["step"]) == float(b["step"]) == 3.0
    assert [grp["weight_decay"] for grp in opt_r2.param_groups] == [grp["weight_decay"] for grp in opt_r.param_groups]
    opt_r2.step()                                                   # the restored reference optimiser is usable


def test_checkpoints_interchange_with_the_reference_trainer(tmp_path):
    """GigaGAN.save() here -> reference GigaGAN.load() and back (gp.py:2033-2107): same dictionary schema, weights
    bit-identical, optimizer state accepted by torch.optim.AdamW as the reference configures it"""
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import import_reference
    ref = import_reference()
    import gigagan_pytorch_b200 as g
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8))
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    mine = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9, create_ema_generator_at_init=False)
    mine._ensure_optimizers()
    mine.G_opt.m.normal_(generator=torch.Generator().manual_seed(1))          # a non-trivial optimizer state
    mine.G_opt.v.uniform_(generator=torch.Generator().manual_seed(2))
    mine.G_opt.step_t.fill_(7)
    p1 = str(tmp_path / "mine.pt")
    mine.save(p1)
    pkg = torch.load(p1, weights_only=False)
    assert {"G", "D", "G_opt", "D_opt", "steps", "version"} <= set(pkg)
    torch.manual_seed(1)
    theirs = ref.GigaGAN(generator=dict(gen), discriminator=dict(disc), amp=False, create_ema_generator_at_init=False)
    theirs.load(p1)
    for a, b in ((mine.G, theirs.G), (mine.D, theirs.D)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    st = theirs.G_opt.state_dict()["state"]
    assert len(st) == len(mine.G_opt.params) and all(float(v["step"]) == 7.0 for v in st.values())
    p2 = str(tmp_path / "theirs.pt")
    theirs.save(p2)
    torch.manual_seed(2)
    back = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9, create_ema_generator_at_init=False)
    back.load(p2)
    assert all(torch.equal(a, b) for a, b in zip(mine.G.state_dict().values(), back.G.state_dict().values()))
    assert back.G_opt is None                 # optimiser state is held until the flat buffers are built (first step),
    back._ensure_optimizers()                 # so that load() followed by .cuda()/.to() stays valid
    assert torch.equal(back.G_opt.m, mine.G_opt.m) and torch.equal(back.G_opt.v, mine.G_opt.v) and int(back.G_opt.step_t) == 7


def test_checkpoints_with_ema_interchange_with_the_reference_trainer(tmp_path):
    """with the default create_ema_generator_at_init=True the checkpoint carries G_ema in ema_pytorch's
    wrapper schema ('ema_model.*', 'initted', 'step'); save -> load restores the EMA weights and its step counter, and
    checkpoints move between this trainer and the reference trainer in both directions with EMA enabled"""
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import import_reference
    ref = import_reference()
    import gigagan_pytorch_b200 as g
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8))
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    mine = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9)
    mine._ensure_optimizers()
    assert mine.has_ema_generator
    with torch.no_grad():                                   # EMA weights that differ from the online generator
        for pe in mine.G_ema.parameters():
            pe.add_(0.25)
    mine._ema_step, mine._ema_initted = 137, True
    p1 = str(tmp_path / "mine.pt")
    mine.save(p1)
    pkg = torch.load(p1, weights_only=False)
    assert {"initted", "step"} <= set(pkg["G_ema"]) and int(pkg["G_ema"]["step"]) == 137
    assert all(("ema_model." + k) in pkg["G_ema"] for k in mine.G.state_dict())
    # this trainer -> this trainer: EMA weights and counter survive (they used to be dropped)
    torch.manual_seed(3)
    again = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9)
    again.load(p1)
    assert again._ema_step == 137 and again._ema_initted
    for a, b, o in zip(mine.G_ema.parameters(), again.G_ema.parameters(), again.G.parameters()):
        assert torch.equal(a, b) and not torch.equal(b, o)
    # this trainer -> reference trainer (strict load of G_ema in the reference, gp.py:2081-2082)
    torch.manual_seed(1)
    theirs = ref.GigaGAN(generator=dict(gen), discriminator=dict(disc), amp=False)
    theirs.load(p1)
    for a, b in zip(mine.G_ema.parameters(), theirs.G_ema.ema_model.parameters()):
        assert torch.equal(a, b)
    assert int(theirs.G_ema.step) == 137
    # reference trainer -> this trainer
    with torch.no_grad():
        for pe in theirs.G_ema.ema_model.parameters():
            pe.mul_(0.5)
        theirs.G_ema.step.fill_(211)
    p2 = str(tmp_path / "theirs.pt")
    theirs.save(p2)
    torch.manual_seed(2)
    back = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9)
    back.load(p2)
    assert back._ema_step == 211
    for a, b in zip(theirs.G_ema.ema_model.parameters(), back.G_ema.parameters()):
        assert torch.equal(a, b)


def test_moving_the_trainer_after_load_keeps_the_optimizer_state():
    """`gan.load(ckpt); gan.to(device)` (nn.Module._apply replaces every p.data) must not detach the
    parameters from the flat AdamW buffers: the flat buffers are torn down with their state carried over and rebuilt"""
    import gigagan_pytorch_b200 as g
    gen = dict(dim_capacity=2, style_network=dict(dim=16, depth=2), image_size=32, dim_max=16, dim_latent=16,
               num_skip_layers_excite=2, unconditional=True, self_attn_resolutions=(16,), self_attn_dim_head=8, self_attn_heads=2)
    disc = dict(dim_capacity=2, dim_max=16, image_size=32, num_skip_layers_excite=2, unconditional=True,
                attn_resolutions=(8,), attn_dim_head=8, attn_heads=2, multiscale_input_resolutions=(16, 8))
    g.set_compute_dtype(torch.float32)
    torch.manual_seed(0)
    gan = g.GigaGAN(generator=gen, discriminator=disc, amp=False, log_steps_every=10 ** 9, create_ema_generator_at_init=False)
    gan._ensure_optimizers()
    gan.D_opt.m.normal_(generator=torch.Generator().manual_seed(1))
    gan.D_opt.v.uniform_(generator=torch.Generator().manual_seed(2))
    gan.D_opt.step_t.fill_(9)
    m, v = gan.D_opt.m.clone(), gan.D_opt.v.clone()
    w = [p.detach().clone() for p in gan.D.parameters()]
    gan.double().float()                                     # any _apply: what .cuda() / .to(device) do
    assert gan.D_opt is None and "D_opt" in gan._pending_opt_state
    gan._ensure_optimizers()
    assert torch.equal(gan.D_opt.m, m) and torch.equal(gan.D_opt.v, v) and int(gan.D_opt.step_t) == 9
    off = 0
    for p, w0 in zip(gan.D_opt.params, w):                   # parameters are views of the NEW flat buffer, values intact
        assert p.data_ptr() == gan.D_opt.flat.data_ptr() + 4 * off and torch.equal(p.detach(), w0)
        off += p.numel()


def test_ema_decay_warmup_matches_ema_pytorch_formula():
    from gigagan_pytorch_b200.trainer import ema_current_decay
    # EMA.get_current_decay of ema-pytorch (inv_gamma 1, power 2/3, min 0): 0 until update_after_step + 1, then
    # 1 - (1 + epoch)^(-2/3) capped at beta
    assert ema_current_decay(100, 100, 0.995) == 0.0 and ema_current_decay(101, 100, 0.995) == 0.0
    assert abs(ema_current_decay(102, 100, 0.995) - (1 - 2 ** (-2 / 3))) < 1e-12
    assert abs(ema_current_decay(111, 100, 0.995) - (1 - 11 ** (-2 / 3))) < 1e-12
    assert ema_current_decay(10 ** 6, 100, 0.995) == 0.995


def test_flat_adamw_slices_cover_the_buffer_exactly():
    """FlatAdamW._part_bounds (the slices of the pipelined all-reduce + AdamW): contiguous, ordered, every element once, cut on
    chunk-table rows - also when there are fewer chunks than parts or one parameter dominates"""
    import torch
    from gigagan_pytorch_b200.trainer import FlatAdamW
    old = FlatAdamW.CHUNK
    try:
        for chunk, layers in ((8, [(3, 5), (7,)]), (1 << 16, [(3, 5)]), (4, [(100,), (2,), (3,)])):
            FlatAdamW.CHUNK = chunk
            net = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in layers])
            opt = FlatAdamW(net)
            rows = opt.chunks.tolist()
            for parts in (1, 2, 4, 7):
                b = opt._part_bounds(parts)
                assert 1 <= len(b) <= parts
                assert b[0][0] == 0 and b[0][2] == 0 and b[-1][1] == len(rows) and b[-1][3] == opt.flat.numel()
                for (c0, c1, lo, hi), nxt in zip(b, b[1:] + [None]):
                    assert c0 < c1 and lo == rows[c0][0] and hi == rows[c1 - 1][0] + rows[c1 - 1][1]
                    if nxt is not None:
                        assert nxt[0] == c1 and nxt[2] == hi
    finally:
        FlatAdamW.CHUNK = old
