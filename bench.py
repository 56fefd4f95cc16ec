"""Benchmark of the GigaGAN 256x256 unconditional G+D training step (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm (oracle port) on host cores

Prints ONE JSON line (see README/DESIGN for the keys).  `value` = images/s with inputs resident in HBM, `e2e` =
images/s through the public GigaGAN API with pinned-host batches copied in and a loss read back every step.
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
FLOP_PER_IMG_PLAIN = 988e9      # useful algorithmic flops per image, plain step
FLOP_PER_IMG_CYCLE = 1181e9     # 4-step cycle mean (every 4th step carries the gradient penalty)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-worker"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=1)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured (sustained)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, src="fallback")


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


# ------------------------------------------------------------------------------------------- reference arm (CPU)
CPU_THREADS_CAP = 32
# DRAM traffic of the dominant kernel class per launch, from the ncu launch list of one plain step
# (profiles/r01_ncu_step_launches.txt: 503 x 22.23 MB conv_fprop_tc + 36 x 43.18 MB conv_thin_tc)
NCU_CONV_DRAM_BYTES_PER_LAUNCH = 23.6e6


def _oracle_trainer(size):
    import torch
    from oracle import gigagan_oracle as O
    import gigagan_pytorch_b200 as g            # constructors only (CPU tensors): same seeded init as the reference
    torch.manual_seed(0)
    G, D = g.Generator(**dict(G_CFG, image_size=size)), g.Discriminator(**dict(D_CFG, image_size=size))
    return O.OracleTrainer(dict(G.state_dict()), O.generator_plan(size, 8, 512, num_skip_layers_excite=4),
                           dict(D.state_dict()), O.discriminator_plan(size, 16, 512, num_skip_layers_excite=4))


def cpu_worker(args):
    """Runs in a subprocess (hard wall-clock bound from the parent): oracle G+D steps on the host cores, one JSON
    progress line per finished step so the parent can use whatever completed."""
    import torch
    cores = min(os.cpu_count() or 1, CPU_THREADS_CAP)
    torch.set_num_threads(cores)
    tr = _oracle_trainer(args.image_size)
    b = args.ref_batch
    print(json.dumps({"event": "ready", "cores": cores}), flush=True)
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        gp = s >= args.warmup and (s - args.warmup + 1) % 4 == 0
        tr.step(real_batch(s, 1, 0, b, args.image_size), gp)
        print(json.dumps({"event": "step", "index": s, "timed": s >= args.warmup, "gp": gp,
                          "seconds": time.perf_counter() - t0}), flush=True)


def run_cpu_steps(size, batch, warmup, steps, budget_s):
    """-> (images/s, cores, steps used, description).  Bounded by budget_s of wall clock."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "cpu-worker", "--image-size", str(size), "--ref-batch",
           str(batch), "--warmup", str(warmup), "--steps", str(steps)]
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    t_start, rows, cores = time.perf_counter(), [], None

    def reader():
        for line in proc.stdout:
            try:
                rows.append(json.loads(line))
            except Exception:
                pass
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    while proc.poll() is None and time.perf_counter() - t_start < budget_s:
        time.sleep(0.5)
    if proc.poll() is None:
        proc.kill()
    th.join(timeout=5)
    cores = next((r["cores"] for r in rows if r.get("event") == "ready"), min(os.cpu_count() or 1, CPU_THREADS_CAP))
    timed = [r for r in rows if r.get("event") == "step" and r["timed"]]
    anyst = [r for r in rows if r.get("event") == "step"]
    use = timed if timed else anyst
    if not use:
        dt = time.perf_counter() - t_start
        return batch / dt, cores, 0, f"no G+D step of batch {batch} finished within the {budget_s:.0f} s budget (upper bound)"
    secs = sum(r["seconds"] for r in use)
    desc = (f"{len(use)} {'timed' if timed else 'warm-up'} G+D step(s) of batch {batch} at {size}x{size}, fp32 torch CPU, "
            f"{cores} threads, {sum(1 for r in use if r['gp'])} with gradient penalty; wall budget {budget_s:.0f} s")
    return batch * len(use) / secs, cores, len(use), desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, used, desc = run_cpu_steps(args.image_size, args.ref_batch, min(args.warmup, 1), args.steps, budget_s=240.0)
    print(json.dumps({
        "impl": "reference", "metric": "images/sec, 256x256 unconditional G+D training step", "value": val,
        "unit": "images/s", "n_gpus": args.gpus, "steps": used, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * args.ref_batch / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "GigaGAN unconditional 256x256 G+D step (README config), CPU oracle port of the reference",
                   "global_batch": args.ref_batch},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------- this repo's arm
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
    size, B = args.image_size, args.batch
    torch.manual_seed(0)
    gan = g.GigaGAN(generator=dict(G_CFG, image_size=size), discriminator=dict(D_CFG, image_size=size), amp=True,
                    mixed_precision_type="bf16", log_steps_every=10 ** 9, create_ema_generator_at_init=True).to(dev)
    gan.use_cuda_graphs = not args.no_graphs
    torch.manual_seed(1234 + rank)

    n_batches = args.steps + args.warmup
    dev_pool = [real_batch(s, world, rank, B, size).to(dev) for s in range(min(n_batches, 8))]

    class Pool:
        batch_size = B

        def __init__(self, items):
            self.items = items

        def __iter__(self):
            return iter(self.items)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, first_step, it):
        last = None
        for s in range(n):
            step = first_step + s
            gp_on = step % 4 == 0
            d = gan.train_discriminator_step(dl_iter=it, apply_gradient_penalty=gp_on)
            gl = gan.train_generator_step(batch_size=B, dl_iter=it)
            last = (d, gl)
        return last

    # ---- device-resident inputs: `value`
    from gigagan_pytorch_b200.trainer import cycle
    it = cycle(Pool(dev_pool))
    # one-time setup, independent of --warmup: every step variant (plain / gradient-penalty discriminator step,
    # generator step) is seen twice so that its CUDA graph is captured before anything is timed (a capture inside the
    # timed region would be a ~0.2 s host stall); step numbering continues so the penalty cadence is unchanged
    PRIME = 0 if args.no_graphs else 8
    run_steps(PRIME, 1, it)
    run_steps(args.warmup, PRIME + 1, it)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    g0 = getattr(gan, "graph_kernel_launches", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps, PRIME + args.warmup + 1, it)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0 + getattr(gan, "graph_kernel_launches", 0) - g0      # own kernels in the timed region
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API: pinned host batches, H2D inside the timed region, loss read back
    host_pool = [real_batch(100 + s, world, rank, B, size, pin=True) for s in range(min(n_batches, 8))]
    gan2_iter_holder = Pool(host_pool)
    gan.train_dl = None
    gan.set_dataloader(gan2_iter_holder)
    gan._host_steps = 1
    gan(steps=max(1, min(args.warmup, 4)))
    sync_all()
    d2h = 0
    e0.record()
    t_wall = time.perf_counter()
    losses = []
    it2 = cycle(gan2_iter_holder)
    for s in range(args.steps):
        step = s + 1
        d = gan.train_discriminator_step(dl_iter=it2, apply_gradient_penalty=step % 4 == 0)
        gl = gan.train_generator_step(batch_size=B, dl_iter=it2)
        losses.append((float(d.divergence), float(gl.divergence)))      # device -> host read of the step result
        d2h += 8
    e1.record()
    sync_all()
    ms2 = e0.elapsed_time(e1)
    t = torch.tensor([ms2], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms2 = t.item()
    e2e = world * B * args.steps / (ms2 / 1e3)
    h2d = B * 3 * size * size * 4

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv): flops / CUDA-event time over one more step
    pk = peaks()
    prof = ops.ConvProfiler()
    gan.use_cuda_graphs = False
    with prof:
        run_steps(1, 5, it)                 # a plain (non gradient-penalty) step
    torch.cuda.synchronize()
    tc = prof.summary()
    roof = {"bound": "tensor", "achieved": tc["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": tc["tflops"] / pk["tflops"] if tc["tflops"] else None, "traffic": NCU_CONV_DRAM_BYTES_PER_LAUNCH,
            "traffic_source": "ncu dram__bytes_read+write per launch, averaged over the same launches of one plain "
                              "step (profiles/r01_ncu_step_launches.txt)",
            "kernel": "convolution fprop+dgrad launches of one plain step: conv_fprop_tc_kernel (tcgen05 implicit GEMM) "
                      "+ conv_thin_tc_kernel (128^2/256^2 layers)", "launches": tc["launches"],
            "kernel_ms_per_step": tc["ms"], "peak_source": pk["src"],
            "step_useful_tflops": FLOP_PER_IMG_CYCLE * B * args.steps / (ms / 1e3) / 1e12 / 1.0}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(size)
    if rank == 0:
        print(json.dumps({
            "metric": "images/sec, 256x256 unconditional G+D training step", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"GigaGAN unconditional {size}x{size} G+D step (README config: G dim_capacity 8, "
                                   f"D dim_capacity 16, dim_max 512, 4 skip-layer-excite), gradient penalty every 4th step",
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "l2": "per-step activations (several GB) exceed the 126 MB L2; no explicit flush",
                       "cuda_graphs": bool(not args.no_graphs)},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": ms2 / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "losses_last": losses[-1]}))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(size):
    """Oracle (port of the reference algorithm) on the host cores: one plain G+D step at batch 1 (bounded sample)."""
    val, cores, used, desc = run_cpu_steps(size, 1, 0, 1, budget_s=90.0)
    return {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": desc}


if __name__ == "__main__":
    a = parse()
    if a.impl == "cpu-worker":
        cpu_worker(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
