"""CPU-side checks of the Chamfer op: C-ABI validation happens before any launch, the Python
wrapper refuses CPU tensors, fscore follows chamfer/fscore.py of the reference."""
import ctypes as C

import pytest
import torch


def test_c_abi_validation_and_scratch_size():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    assert lib.gsl_chamfer_scratch_bytes(2, 1000, 500) == 8 * 2 * 1500
    assert lib.gsl_chamfer_scratch_bytes(-1, 10, 10) == 0
    assert lib.gsl_chamfer_forward(1, -5, None, 3, None, None, None, None, None, None, None) == L.GSL_EINVAL
    assert b"bad sizes" in lib.gsl_last_error()
    assert lib.gsl_chamfer_forward(1, 10, None, 10, None, None, None, None, None, None, None) == L.GSL_EINVAL
    assert b"NULL" in lib.gsl_last_error()
    assert lib.gsl_chamfer_backward(1, 10, None, 10, None, None, None, None, None, None, None, None) == L.GSL_EINVAL
    assert lib.gsl_chamfer_forward(0, 10, None, 10, None, None, None, None, None, None, None) == 0   # empty batch: nothing to do
    assert lib.gsl_chamfer_forward(70000, 1, None, 1, None, None, None, None, None, None, None) == L.GSL_EINVAL


def test_python_wrapper_refuses_cpu_tensors_and_bad_shapes():
    from gs_lidar_b200.chamfer import chamfer_3DDist
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chamfer_3DDist()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(AssertionError, match="Wrong last dimension"):
        chamfer_3DDist()(torch.zeros(1, 4, 2), torch.zeros(1, 4, 3))


def test_fscore_follows_the_reference_formula():
    from gs_lidar_b200.chamfer import fscore
    g = torch.Generator().manual_seed(0)
    d1, d2 = torch.rand(3, 50, generator=g) * 0.002, torch.rand(3, 70, generator=g) * 0.002
    d1[2], d2[2] = 1.0, 1.0                                   # nothing within the threshold: 0 / 0 -> 0
    f, p1, p2 = fscore(d1, d2)
    q1 = (d1 < 0.001).float().mean(dim=1)                     # chamfer/fscore.py:12-13
    q2 = (d2 < 0.001).float().mean(dim=1)
    want = 2 * q1 * q2 / (q1 + q2)                            # :14
    want[torch.isnan(want)] = 0                               # :15
    assert torch.equal(p1, q1) and torch.equal(p2, q2) and torch.equal(f, want) and float(f[2]) == 0.0
