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
    """`bench.py --impl reference` (oracle port on the host cores, bounded) at a tiny size: one JSON line with the
    contract keys; never touches CUDA"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--image-size", "64",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=280, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
