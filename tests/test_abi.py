"""CPU checks of the C-ABI boundary: the shared library loads and exports every symbol the public header declares
(no compute calls without a GPU); host-side constants (resampling matrices) against torch."""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as ge
    so = os.path.join(ROOT, "gigagan_pytorch_b200", "libgigagan_sm100.so")
    if not os.path.exists(so):
        ge.build()
    return so


def test_library_exports_every_declared_symbol():
    so = _ensure_built()
    from gigagan_pytorch_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 25
    L = ctypes.CDLL(so)
    for name in protos:
        assert hasattr(L, name), f"{name} declared in include/gigagan_sm100.h but not exported"
    L.gg_version.restype = ctypes.c_int
    assert L.gg_version() >= 100
    assert L.gg_has_tcgen05() in (0, 1)


def test_product_has_no_cpu_path():
    _ensure_built()
    from gigagan_pytorch_b200 import ops
    with pytest.raises(RuntimeError):
        ops.leaky_relu(torch.zeros(4))          # CPU tensor -> loud failure, never a fallback


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gigagan_pytorch_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "") or f == "__init__.py", f


def test_resample_matrices_match_torch():
    from gigagan_pytorch_b200.ops import bilinear_matrix, blur_matrix
    for n in (4, 8, 16):
        x = torch.randn(1, 1, n, n, dtype=torch.float64)
        up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        a = bilinear_matrix(n, 2 * n)
        torch.testing.assert_close(a @ x[0, 0] @ a.t(), up[0, 0])
        k1 = torch.tensor([1.0, 2.0, 1.0], dtype=torch.float64)
        k = (k1[:, None] * k1[None]) / 16
        ref = F.conv2d(F.pad(up, (1, 1, 1, 1), mode="reflect"), k[None, None])
        b = blur_matrix(2 * n)
        torch.testing.assert_close(b @ up[0, 0] @ b.t(), ref[0, 0])
    for n_in, n_out in ((64, 16), (32, 8), (256, 64)):
        x = torch.randn(1, 1, n_in, n_in, dtype=torch.float64)
        ref = F.interpolate(x, n_out, mode="bilinear")
        a = bilinear_matrix(n_in, n_out)
        torch.testing.assert_close(a @ x[0, 0] @ a.t(), ref[0, 0])


def test_seeded_init_matches_reference_layout():
    """state_dict keys/shapes of the drop-in classes equal the fixture captured from the reference."""
    import gigagan_pytorch_b200 as g
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "ka4_generator.pt"), weights_only=False)
    torch.manual_seed(0)
    G = g.Generator(**fx["cfg"])
    sd = G.state_dict()
    assert list(sd.keys()) == list(fx["sd"].keys())
    for k in sd:
        assert sd[k].shape == fx["sd"][k].shape, k
    same = [k for k in sd if "1.1.weight" not in k and "1.4.weight" not in k]      # noise weights were perturbed
    assert all(torch.equal(sd[k], fx["sd"][k]) for k in same), "seeded init differs from the reference"
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "ka5_discriminator.pt"), weights_only=False)
    torch.manual_seed(0)
    D = g.Discriminator(**fx["cfg"])
    sd = D.state_dict()
    assert list(sd.keys()) == list(fx["sd"].keys())
    assert all(torch.equal(sd[k], fx["sd"][k]) for k in sd)
