"""Static checks of the compiled sm_100a objects (cuobjdump, no GPU): the tensor-core kernels really are tcgen05 / TMA
kernels, and the issue rule of DESIGN.md section 4 holds - UTCHMMA operands come from uniform registers, i.e. no
R2UR(.BROADCAST) conversion sits in front of a tcgen05.mma (each one cost ~100 issue cycles before the fix)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gigagan_pytorch_b200", "csrc")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def sass(obj):
    path = os.path.join(CSRC, obj)
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    out = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    return [l for l in out.stdout.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]


def near(lines, mnemonic, what, window=8):
    total = hits = 0
    for i, l in enumerate(lines):
        if mnemonic in l:
            total += 1
            hits += any(what in x for x in lines[max(0, i - window):i])
    return total, hits


pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


@pytest.mark.parametrize("obj,max_r2ur_mma", [("conv_tc.o", 0), ("conv_thin_tc.o", 0), ("bmm_tc.o", 0), ("attn_tc.o", 11),
                                              ("attn_tc2.o", 60)])      # 392 UTCHMMA in 21 instantiations, 43 near an R2UR
def test_tcgen05_mma_is_issued_from_uniform_registers(obj, max_r2ur_mma):
    lines = sass(obj)
    n_mma, r2ur_mma = near(lines, "UTCHMMA", "R2UR")
    assert n_mma > 0, f"{obj}: no tcgen05.mma (UTCHMMA) in the SASS"
    assert r2ur_mma <= max_r2ur_mma, f"{obj}: {r2ur_mma} of {n_mma} UTCHMMA are fed through R2UR again (divergent issue code?)"
    assert any("UTCBAR" in l for l in lines), f"{obj}: no tcgen05.commit (UTCBAR)"
    assert any("LDTM" in l for l in lines), f"{obj}: no tcgen05.ld (LDTM)"


@pytest.mark.parametrize("obj", ["conv_tc.o", "bmm_tc.o", "attn_tc.o", "attn_tc2.o"])
def test_tma_loads_present(obj):
    lines = sass(obj)
    assert any("UTMALDG" in l for l in lines), f"{obj}: no TMA tensor load (UTMALDG)"


def test_epilogues_use_256_bit_stores():
    for obj in ("conv_tc.o", "conv_thin_tc.o", "attn_tc.o", "attn_tc2.o"):
        assert any("STG.E.ENL2.256" in l for l in sass(obj)), f"{obj}: no 256-bit global store"


def test_second_generation_attention_reads_tmem_in_whole_slabs():
    """attn_tc2.cu: a softmax thread pulls its column slab of S / dP out of TMEM with 32-column loads (one round trip per
    tile), not in 16-column chunks"""
    lines = sass("attn_tc2.o")
    assert sum("LDTM.x32" in l for l in lines) >= 20
