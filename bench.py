"""Benchmark of the GigaGAN G+D training step (BASELINE.json configs; default = configs[1], unconditional 256x256).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the UNMODIFIED reference on the box's host cores
    python bench.py --config cfg3|cfg4|cfg5 ...               # the other BASELINE configs (same JSON contract)

Prints ONE JSON line.  `value` = images/s with inputs resident in HBM; `e2e` = images/s through the public GigaGAN
API with pinned-host batches copied in and a loss read back every step; `roofline` = the WHOLE step's useful
algorithmic FLOPs / time against the measured bf16 tensor peak (the convolution class is listed under
`roofline.kernels`); `cpu_baseline` = the reference on the host cores; `gpu_eager_reference` = the reference's own
PyTorch-eager path on the same GPU (bf16 autocast, same batch) - the denominator of north_star's 10x target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_CFG = dict(dim_capacity=8, style_network=dict(dim=64, depth=4), image_size=256, dim_max=512,
             num_skip_layers_excite=4, unconditional=True)
D_CFG = dict(dim_capacity=16, dim_max=512, image_size=256, num_skip_layers_excite=4, unconditional=True)
TEXT_ENC = dict(dim=64, depth=4)              # learned transformer over pre-encoded CLIP tokens (77 x 512)
CLIP_TOKENS, CLIP_DIM = 77, 512

# useful algorithmic flops per image (counted on the reference with torch's flop registry; the
# reference's discarded D weight gradient in the G step is NOT counted): plain step 988 G, penalty step 1761 G
FLOP_PER_IMG_PLAIN = 988e9
FLOP_PER_IMG_GP = 1761e9


def configs(name, size=None):
    """-> dict(G, D, batch (per GPU), size, text, upsampler, label)"""
    if name == "cfg2":
        s = size or 256
        return dict(G=dict(G_CFG, image_size=s), D=dict(D_CFG, image_size=s), batch=16, size=s, text=False,
                    upsampler=False, label=f"GigaGAN unconditional {s}x{s} G+D step (README config: G dim_capacity 8, "
                    "D dim_capacity 16, dim_max 512, 4 skip-layer-excite), gradient penalty every 4th step")
    if name in ("cfg3", "cfg5"):
        s = size or (256 if name == "cfg3" else 512)
        g = dict(G_CFG, image_size=s, unconditional=False, text_encoder=dict(TEXT_ENC),
                 style_network=dict(dim=64, depth=4, dim_text_latent=TEXT_ENC["dim"]))
        d = dict(D_CFG, image_size=s, unconditional=False, text_encoder=dict(TEXT_ENC))
        if name == "cfg5":
            d["multiscale_input_resolutions"] = (256, 128)
        return dict(G=g, D=d, batch=8 if name == "cfg3" else 4, size=s, text=True, upsampler=False,
                    label=f"GigaGAN text-conditional {s}x{s} G+D step on pre-encoded CLIP tokens (77x512), cross attention, "
                    "text-modulated predictors" + (", multiscale inputs (256,128)" if name == "cfg5" else ""))
    if name == "cfg4":
        s = size or 256
        g = dict(style_network=dict(dim=64, depth=4), dim=32, image_size=s, input_image_size=s // 4, unconditional=True)
        d = dict(D_CFG, image_size=s, multiscale_input_resolutions=(s // 2,))
        return dict(G=g, D=d, batch=4, size=s, text=False, upsampler=True,
                    label=f"GigaGAN UnetUpsampler {s // 4}->{s} G+D step (README upsampler config)")
    raise ValueError(name)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "ref-worker", "cpu-worker"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--image-size", type=int, default=None)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-global-batch (strong scaling) row at N>1")
    ap.add_argument("--ref-batch", type=int, default=2)
    ap.add_argument("--ref-device", default="cpu")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"])
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    burst=d["bf16_tflops"], src="MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, burst=1650.0, src="fallback of B200_PROFILING.md (no MEASURED_PEAKS.json)")


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) > 8:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        mx = [int(float(r[2])) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def real_batch(step, world, rank, batch, size, pin=False):
    import torch
    g = torch.Generator().manual_seed(1000 + step * world + rank)
    t = torch.rand(batch, 3, size, size, generator=g)
    return t.pin_memory() if pin else t


def text_batch(step, world, rank, batch, pin=False):
    """random pre-encoded CLIP tokens (b, 77, 512) with a seeded valid length per sample, zeros beyond it (the mask is
    derived from != 0, ref gigagan_pytorch.py:852) - SURVEY 8d"""
    import torch
    g = torch.Generator().manual_seed(5000 + step * world + rank)
    t = torch.randn(batch, CLIP_TOKENS, CLIP_DIM, generator=g)
    lens = torch.randint(4, CLIP_TOKENS + 1, (batch,), generator=g)
    t[torch.arange(CLIP_TOKENS)[None, :] >= lens[:, None]] = 0.
    return t.pin_memory() if pin else t


# ------------------------------------------------------------------------------------------- reference arms
def reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gigagan_pytorch"))


def import_reference():
    """the UNMODIFIED reference installed into baseline/_ref (pip --no-deps --target), with the stand-ins
    of oracle/ref_shims for its five third-party imports that are not installable offline"""
    for p in (os.path.join(ROOT, "oracle", "ref_shims"), os.path.join(ROOT, "baseline", "_ref")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gigagan_pytorch
    return gigagan_pytorch


def ref_worker(args):
    """Subprocess: G+D steps of the reference through its own public API (GigaGAN.train_discriminator_step /
    train_generator_step), on the CPU (fp32, all host threads) or on cuda:0 (bf16 autocast = README `amp=True` with
    mixed_precision_type='bf16').  One JSON line per finished step so the parent can use whatever completed."""
    import torch
    dev = args.ref_device
    cfg = configs(args.config, args.image_size)
    assert not cfg["text"], "the reference's text path needs OpenCLIP weights (not available offline)"
    b = args.ref_batch
    cores = os.cpu_count() or 1
    if dev == "cpu":
        torch.set_num_threads(cores)
    ref = import_reference()
    from torch.utils.data import DataLoader, Dataset
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        gan = ref.GigaGAN(generator=dict(cfg["G"]), discriminator=dict(cfg["D"]), train_upsampler=cfg["upsampler"],
                          amp=(dev != "cpu"), mixed_precision_type="bf16", model_folder="/tmp/gg_ref_models",
                          results_folder="/tmp/gg_ref_results")
        if dev != "cpu":
            gan = gan.to(torch.device(dev))
    if cfg["upsampler"]:               # the trainer passes lowres_image=, the upsampler names it differently
        fwd = gan.G.forward
        gan.G.forward = lambda *a, lowres_image=None, **k: fwd(lowres_image, *a, **k) if lowres_image is not None else fwd(*a, **k)

    class DS(Dataset):
        def __init__(self):
            self.pool = torch.cat([real_batch(s, 1, 0, b, cfg["size"]) for s in range(4)])

        def __len__(self):
            return 1 << 20

        def __getitem__(self, i):
            return self.pool[i % self.pool.shape[0]]

    from gigagan_pytorch.gigagan_pytorch import cycle
    it = cycle(DataLoader(DS(), batch_size=b, shuffle=False, num_workers=0))
    print(json.dumps({"event": "ready", "cores": cores if dev == "cpu" else 0, "threads": torch.get_num_threads(),
                      "device": dev}), flush=True)
    for s in range(args.warmup + args.steps):
        gp = (s + 1) % 4 == 0
        if dev != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            dl = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp)
            gl = gan.train_generator_step(batch_size=b, dl_iter=it)
        if dev != "cpu":
            torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        print(json.dumps({"event": "step", "index": s, "timed": s >= args.warmup, "gp": gp, "seconds": secs,
                          "d_divergence": float(dl.divergence), "g_divergence": float(gl.divergence)}), flush=True)


def _oracle_trainer(size):
    import torch
    from oracle import gigagan_oracle as O
    import gigagan_pytorch_b200 as g            # constructors only (CPU tensors): same seeded init as the reference
    torch.manual_seed(0)
    G, D = g.Generator(**dict(G_CFG, image_size=size)), g.Discriminator(**dict(D_CFG, image_size=size))
    return O.OracleTrainer(dict(G.state_dict()), O.generator_plan(size, 8, 512, num_skip_layers_excite=4),
                           dict(D.state_dict()), O.discriminator_plan(size, 16, 512, num_skip_layers_excite=4))


def cpu_worker(args):
    """Fallback when baseline/_ref is absent: the oracle port of the reference algorithm on the host cores."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size = args.image_size or 256
    tr = _oracle_trainer(size)
    b = args.ref_batch
    print(json.dumps({"event": "ready", "cores": cores, "threads": torch.get_num_threads(), "device": "cpu"}), flush=True)
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        gp = (s + 1) % 4 == 0
        tr.step(real_batch(s, 1, 0, b, size), gp)
        print(json.dumps({"event": "step", "index": s, "timed": s >= args.warmup, "gp": gp,
                          "seconds": time.perf_counter() - t0}), flush=True)


def run_ref_steps(args, device, batch, warmup, steps, budget_s, kind="auto"):
    """-> dict(value img/s, cores, used steps, kind, sample).  Subprocess bounded by budget_s of wall clock."""
    use_ref = reference_available() if kind == "auto" else kind == "reference"
    if args.config != "cfg2" and not use_ref:
        return None
    impl = "ref-worker" if use_ref else "cpu-worker"
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", impl, "--config", args.config, "--ref-batch", str(batch),
           "--warmup", str(warmup), "--steps", str(steps), "--ref-device", device]
    if args.image_size:
        cmd += ["--image-size", str(args.image_size)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    t_start, rows, errs = time.perf_counter(), [], []

    def reader():
        for line in proc.stdout:
            try:
                rows.append(json.loads(line))
            except Exception:
                pass

    def err_reader():
        for line in proc.stderr:
            errs.append(line)
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    threading.Thread(target=err_reader, daemon=True).start()
    while proc.poll() is None and time.perf_counter() - t_start < budget_s:
        time.sleep(0.5)
    if proc.poll() is None:
        proc.kill()
    th.join(timeout=5)
    ready = next((r for r in rows if r.get("event") == "ready"), {})
    cores = ready.get("threads", os.cpu_count() or 1)
    timed = [r for r in rows if r.get("event") == "step" and r["timed"]]
    anyst = [r for r in rows if r.get("event") == "step"]
    use = timed if timed else anyst
    size = configs(args.config, args.image_size)["size"]
    what = ("unmodified reference (baseline/_ref) via GigaGAN.train_discriminator_step/train_generator_step" if use_ref
            else "oracle port of the reference")
    prec = "fp32 torch CPU" if device == "cpu" else "bf16 autocast, PyTorch eager (cuDNN/cuBLAS/ATen)"
    if not use:
        dt = time.perf_counter() - t_start
        tail = "".join(errs[-3:]).strip().replace("\n", " | ")[-300:]
        return dict(value=batch / dt, cores=cores, used=0, kind="reference" if use_ref else "port", failed=True,
                    sample=f"no G+D step of batch {batch} finished within {budget_s:.0f} s ({what}); upper bound. {tail}")
    secs = sum(r["seconds"] for r in use)
    desc = (f"{len(use)} {'timed' if timed else 'warm-up'} G+D step(s) of batch {batch} at {size}x{size}, {what}, {prec}, "
            f"{cores} threads, {sum(1 for r in use if r['gp'])} with gradient penalty; wall budget {budget_s:.0f} s")
    out = dict(value=batch * len(use) / secs, cores=cores, used=len(use), kind="reference" if use_ref else "port",
               sample=desc, ms_per_step=1e3 * secs / len(use))
    if "d_divergence" in use[-1]:          # the reference's own losses after the same number of steps (plausibility only:
        out["losses_last"] = [use[-1]["d_divergence"], use[-1]["g_divergence"]]     # the noise streams are not shared)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = configs(args.config, args.image_size)
    if cfg["text"]:
        print(json.dumps({"impl": "reference", "unavailable": "the reference's text-conditional path needs pretrained "
                          "OpenCLIP weights (third-party, no network)"}))
        return
    r = run_ref_steps(args, "cpu", args.ref_batch, min(args.warmup, 1), args.steps, budget_s=240.0, kind=args.ref_kind)
    val = r["value"]
    print(json.dumps({
        "impl": "reference", "metric": "images/sec, 256x256 unconditional G+D training step", "value": val,
        "unit": "images/s", "n_gpus": args.gpus, "steps": r["used"], "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * args.ref_batch / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"] + "; reference CPU path on the host cores", "global_batch": args.ref_batch},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------- this repo's arm
def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel class and per step, from the committed ncu launch list of this
    round (profiles/r02_ncu_step_launches.json, written by tools/ncu_summarize.py); None when absent."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_step_launches.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gigagan_pytorch_b200 as g
    from gigagan_pytorch_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = configs(args.config, args.image_size)
    size, B = cfg["size"], (args.batch or cfg["batch"])
    torch.manual_seed(0)
    extra = dict(matching_awareness_loss_weight=0., generator_contrastive_loss_weight=0.) if cfg["text"] else {}
    gan = g.GigaGAN(generator=dict(cfg["G"]), discriminator=dict(cfg["D"]), train_upsampler=cfg["upsampler"], amp=True,
                    mixed_precision_type="bf16", log_steps_every=10 ** 9, create_ema_generator_at_init=True,
                    save_and_sample_every=0, **extra).to(dev)
    gan.use_cuda_graphs = not args.no_graphs
    torch.manual_seed(1234 + rank)
    n_pool = 8

    def make_pool(batch, pin, base=0):
        items = []
        for s in range(n_pool):
            img = real_batch(base + s, world, rank, batch, size, pin)
            if cfg["text"]:
                items.append((img if pin else img.to(dev), text_batch(base + s, world, rank, batch, pin) if pin
                              else text_batch(base + s, world, rank, batch).to(dev)))
            else:
                items.append(img if pin else img.to(dev))
        return items

    class Pool:
        def __init__(self, items, batch):
            self.items, self.batch_size = items, batch

        def __iter__(self):
            return iter(self.items)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, first_step, it, batch):
        last = None
        for s in range(n):
            step = first_step + s
            d = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=step % 4 == 0)
            gl = gan.train_generator_step(batch_size=batch, dl_iter=it)
            last = (d, gl)
        return last

    from gigagan_pytorch_b200.trainer import cycle

    def timed(batch, steps, warmup, prime):
        """-> (ms max over ranks, own launches, #penalty steps).  One-time setup independent of --warmup: every step
        variant is seen twice so that its CUDA graph is captured before anything is timed (a capture inside the timed
        region would be a ~0.2 s host stall); step numbering continues so the penalty cadence is unchanged."""
        it = cycle(Pool(make_pool(batch, False), batch))
        run_steps(prime, 1, it, batch)
        run_steps(warmup, prime + 1, it, batch)
        sync_all()
        l0, g0 = _lib.launch_count, gan.graph_kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(steps, prime + warmup + 1, it, batch)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_gp = sum(1 for s in range(steps) if (prime + warmup + 1 + s) % 4 == 0)
        return t.item(), _lib.launch_count - l0 + gan.graph_kernel_launches - g0, n_gp, it

    # ---- device-resident inputs: `value`
    PRIME = 0 if args.no_graphs else 8
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, n_gp, it = timed(B, args.steps, args.warmup, PRIME)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API: pinned host batches, H2D inside the timed region, loss read back
    host = Pool(make_pool(B, True, base=100), B)
    gan.train_dl = None
    gan.set_dataloader(host)
    gan._host_steps = 1
    gan(steps=max(1, min(args.warmup, 4)))
    sync_all()
    d2h = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    it2 = cycle(host)
    for s in range(args.steps):
        step = s + 1
        d = gan.train_discriminator_step(dl_iter=it2, apply_gradient_penalty=step % 4 == 0)
        gl = gan.train_generator_step(batch_size=B, dl_iter=it2)
        losses.append((float(d.divergence), float(gl.divergence)))      # device -> host read of the step result
        d2h += 8
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms2 = t.item()
    e2e = world * B * args.steps / (ms2 / 1e3)
    h2d = B * 3 * size * size * 4 * (2 if (cfg["text"] or cfg["upsampler"]) else 1)
    if cfg["text"]:
        h2d += 2 * B * CLIP_TOKENS * CLIP_DIM * 4

    # ---- strong scaling row (fixed global batch = the 1-GPU batch), SURVEY 8e
    strong = None
    if world > 1 and not args.no_strong and B % world == 0:
        bs = B // world
        ms_s, _, _, _ = timed(bs, args.steps, max(args.warmup, 3), PRIME)
        strong = {"global_batch": B, "per_gpu_batch": bs, "value": B * args.steps / (ms_s / 1e3), "unit": "images/s",
                  "ms_per_step": ms_s / args.steps}

    # ---- roofline: whole step against the tensor peak; the convolution launches of one extra (eager) plain step
    pk = peaks()
    prof = ops.ConvProfiler()
    gan.use_cuda_graphs = False
    with prof:
        run_steps(1, 5, it, B)                 # a plain (non gradient-penalty) step
    torch.cuda.synchronize()
    tc = prof.summary()
    n_plain = args.steps - n_gp
    useful = (n_plain * FLOP_PER_IMG_PLAIN + n_gp * FLOP_PER_IMG_GP) * B * world
    step_tflops = useful / (ms / 1e3) / 1e12 / world          # per GPU
    ncu = ncu_traffic() or {}
    roof = None
    if args.config == "cfg2" and size == 256:
        roof = {"bound": "tensor", "scope": "whole G+D step (every kernel, AdamW and all-reduce included)",
                "achieved": step_tflops, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": step_tflops / pk["tflops"],
                "peak_source": pk["src"],
                "flops_per_step": useful / args.steps / world,
                "flops_source": "SURVEY.md 8d useful algorithmic flops per image (plain 988 G, penalty step 1761 G) x batch",
                "traffic": ncu.get("dram_bytes_per_step"), "traffic_source": ncu.get("source"),
                "kernels": [{"kernel": "convolution fprop+dgrad launches of one plain step: conv_fprop_tc_kernel (tcgen05 "
                                       "implicit GEMM) + conv_thin_tc_kernel (128^2/256^2 layers)",
                             "bound": "tensor", "achieved": tc["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": tc["tflops"] / pk["tflops"] if tc["tflops"] else None, "launches": tc["launches"],
                             "ms_per_step": tc["ms"], "traffic": ncu.get("conv_dram_bytes_per_launch")}]}

    cpu = gpu_ref = None
    if rank == 0 and world == 1 and not cfg["text"]:
        if not args.no_gpu_reference and reference_available():
            del prof
            torch.cuda.empty_cache()
            r = run_ref_steps(args, f"cuda:{local}", B, 4, 8, budget_s=150.0, kind="reference")
            if r is not None:
                gpu_ref = {"value": r["value"], "unit": "images/s", "batch": B, "ms_per_step": r.get("ms_per_step"),
                           "losses_last": r.get("losses_last"),
                           "steps": r["used"], "sample": r["sample"], "speedup_e2e": e2e / r["value"] if not r.get("failed") else None}
        if not args.no_cpu_baseline:
            r = run_ref_steps(args, "cpu", 2, 1, 4, budget_s=100.0)      # one untimed step first: the cold one is 5-8x slower
            if r is not None:
                cpu = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if rank == 0:
        print(json.dumps({
            "metric": "images/sec, 256x256 unconditional G+D training step" if args.config == "cfg2" else
                      f"images/sec, {cfg['label']}",
            "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["label"], "name": args.config,
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "l2": "per-step activations (several GB) exceed the 126 MB L2; no explicit flush",
                       "cuda_graphs": bool(not args.no_graphs), "gradient_penalty_steps_timed": n_gp},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": ms2 / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "gpu_eager_reference": gpu_ref, "strong_scaling": strong, "losses_last": losses[-1]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "cpu-worker":
        cpu_worker(a)
    elif a.impl == "ref-worker":
        ref_worker(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
