#!/usr/bin/env python
"""Generates tests/golden/eo_v1.npz from the CPU oracle: extremal_opt (RRRMC.jl:468-521, EOCache DeltaE.jl:413-543) on
the DiscrGraph families — energies at the hook instants, final configuration, Emin, Cmin, itmin — together with the rank
table fτ each run used (so that the vectors do not depend on the host's `pow`). Same caveat as make_golden.py: the
reference (pure Julia) ships no golden vectors and cannot run here; these freeze the oracle's and the engine's answers.
Run:  python tests/golden/make_golden_eo.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ffi  # noqa: E402
from tests.helpers import ea_instance, random_config  # noqa: E402

CASES = {
    "EA(4,3)": lambda: ffi.Graph.ea_int(*ea_instance(4, 3, (-1, 1), 21), (-1, 1)),
    "EA(2,3)": lambda: ffi.Graph.ea_int(*ea_instance(2, 3, (-1, 1), 22), (-1, 1)),
    "EA(5,2,(-1,0,1))": lambda: ffi.Graph.ea_int(*ea_instance(5, 2, (-1, 0, 1), 23), (-1, 0, 1)),
    "EA(4,2,(-2,-1,1,2))": lambda: ffi.Graph.ea_int(*ea_instance(4, 2, (-2, -1, 1, 2), 24), (-2, -1, 1, 2)),
    "QT(12,4)": lambda: ffi.Graph.qt(12, 4, 0.73),
}
TAU, ITERS, STEP, SEED = 1.3, 800, 40, 20261018


def ftau_of(N):
    return np.cumsum(np.arange(1, N + 1, dtype=np.float64) ** -TAU)


def compute(ftaus=None):
    out = {}
    for name, mk in CASES.items():
        g = mk()
        ft = ftau_of(g.N) if ftaus is None else ftaus[name]
        s = random_config(g.N, seed=12)
        Es, Cmin, res = ffi.extremal_opt(g, ft, ITERS, s, ffi.PhiloxDraws(SEED, chain=2), step=STEP)
        out[f"{name}/ftau"] = ft
        out[f"{name}/Es"] = Es
        out[f"{name}/final"] = s
        out[f"{name}/Cmin"] = Cmin
        out[f"{name}/Emin_itmin"] = np.array([res.Emin, res.itmin])
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eo_v1.npz")
    np.savez_compressed(path, **compute())
    print("wrote", path, os.path.getsize(path), "bytes")
