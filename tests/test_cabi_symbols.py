"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/rrrmc_b200.h declares, and
fails loudly (no fallback) when there is no device. Host-side lattice logic is compared with the oracle."""
import os
import re

import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from rrrmc_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rrrmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rrrmc_[A-Za-z0-9_]+)\s*\(", src)) - {"rrrmc_hook_fn"})


def test_library_exports_every_declared_symbol():
    L = _ffi.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes signature"


def test_version_and_opts_default():
    assert b"sm_100a" in _ffi.lib().rrrmc_version()
    import ctypes as C
    o = _ffi.Opts()
    assert _ffi.lib().rrrmc_opts_default(C.byref(o)) == 0
    assert o.planes_K == 5 and o.planes_M == 4 and o.staged_thr_fact == 5.0 and np.isnan(o.staged_thr)
    assert o.schedule == _ffi.SCHED_RANDOM_SITE   # the reference order and sampling contract is the default; lattice sweeps are opt-in


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(rb.RRRMCError, match="no CPU fallback"):
        rb.Context(0)
    with pytest.raises(rb.RRRMCError):
        rb.GraphEA(4, 2)


@pytest.mark.parametrize("L,D", [(2, 1), (2, 3), (3, 2), (4, 3), (5, 3), (6, 2), (3, 4)])
def test_host_adjacency_matches_oracle(L, D):
    A = rb.gen_EA(L, D)
    assert np.array_equal(A, ffi.gen_EA(L, D))
    n = int((A > np.arange(1, len(A) + 1)[:, None]).sum())
    d = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(rb.gen_J(lambda k: d[:k], A), ffi.gen_J(A, d))


def test_config_bit_layout():
    bits = np.random.default_rng(1).integers(0, 2, (3, 130))
    c = rb.Config.from_bits(bits)
    assert c.chunks.shape == (3, 3)
    for r in range(3):
        for i in (0, 63, 64, 129):
            assert (int(c.chunks[r, i >> 6]) >> (i & 63)) & 1 == bits[r, i]   # Common.jl:15-22
    assert np.array_equal(c.s, bits.astype(bool))
    assert int(c.chunks[0, 2]) >> 2 == 0  # unused high bits stay zero (Interface.jl:26)
