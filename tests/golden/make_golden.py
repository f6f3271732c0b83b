#!/usr/bin/env python
"""Generates tests/golden/hotpath_v1.npz from the CPU oracle (oracle/rrrmc_oracle.c).

The reference (pure Julia) ships no golden vectors and cannot run in this image, so these vectors do
NOT pin the oracle to Julia output — they freeze the oracle's (and therefore the engine's) current answers on the
hot-path graph families so that later refactors of either side cannot drift silently. Inputs are regenerated
deterministically by tests.helpers (numpy PCG64 seeds); draws come from the shared Philox chain source.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ffi  # noqa: E402
from tests.helpers import ea_instance, random_config, sk_binary, sk_gauss  # noqa: E402

CASES = {
    "EA(4,2)": lambda: ffi.Graph.ea_int(*ea_instance(4, 2, (-1, 1), 1), (-1, 1)),
    "EA(2,3)": lambda: ffi.Graph.ea_int(*ea_instance(2, 3, (-1, 1), 2), (-1, 1)),
    "EA(3,3,(-1,0,1))": lambda: ffi.Graph.ea_int(*ea_instance(3, 3, (-1, 0, 1), 3), (-1, 0, 1)),
    "EANormal(3,2)": lambda: ffi.Graph.ea_f64(*ea_instance(3, 2, seed=4, gaussian=True)),
    "EANormalDiscretized(3,2,(-1,0,1))": lambda: ffi.Graph.ea_discretized(*ea_instance(3, 2, seed=14, gaussian=True), (-1, 0, 1)),
    "SK(10)": lambda: ffi.Graph.sk_bin(sk_binary(10, 5)),
    "SKNormal(10)": lambda: ffi.Graph.sk_f64(sk_gauss(10, 6)),
    "QT(12,4)": lambda: ffi.Graph.qt(12, 4, 0.73),
    "Quant(6,4,SK)": lambda: ffi.Graph.quant(6, 4, 0.5, 2.0, ffi.SK_BIN, sk_binary(6, 7)),
    "Quant(6,4,SKNormal)": lambda: ffi.Graph.quant(6, 4, 0.5, 2.0, ffi.SK_F64, sk_gauss(6, 8)),
    "Quant(6,4,Empty)": lambda: ffi.Graph.quant(6, 4, 0.5, 2.0, ffi.EMPTY),
    "QEAT(3,2,4)": lambda: (lambda AJ: ffi.Graph.quant(9, 4, 0.5, 2.0, ffi.EA_F64, AJ[1], AJ[0]))(ea_instance(3, 2, seed=15, gaussian=True)),
}
SAMPLERS = {"standardMC": ffi.standardMC, "rrrMC": ffi.rrrMC, "bklMC": ffi.bklMC}
BETA, ITERS, STEP, SEED = 1.7, 600, 50, 20261017


def compute():
    out = {}
    for name, mk in CASES.items():
        g = mk()
        s0 = random_config(g.N, seed=11)
        out[f"{name}/energy"] = np.array([g.energy(s0)])
        out[f"{name}/delta_energy"] = np.array([g.delta_energy(s0, i) for i in range(1, g.N + 1)])
        out[f"{name}/neighbors1"] = g.neighbors(1)
        if g.kind in (ffi.EA_INT, ffi.QT, ffi.QUANT, ffi.EA_DISCR):
            out[f"{name}/allDE"] = g.allDE()
        for sname, fn in SAMPLERS.items():
            gg = mk(); s = s0.copy()
            Es, res = fn(gg, BETA, ITERS, s, ffi.PhiloxDraws(SEED, chain=3), step=STEP)
            out[f"{name}/{sname}/Es"] = Es
            out[f"{name}/{sname}/final"] = s
            out[f"{name}/{sname}/accepted"] = np.array([res.accepted])
    # engine's checkerboard procedure (CPU model): L=4, D=3, R=64, K=5 full + M=4 merged planes, 3 sweeps
    L, D, R = 4, 3, 64
    A, J = ea_instance(L, D, seed=9)
    N = L ** D
    Jf = np.zeros((N, D), np.int8)
    for i in range(N):
        stride = 1
        for d in range(D):
            c = (i // stride) % L
            up = i + (((c + 1) % L) - c) * stride
            Jf[i, d] = J[i, np.flatnonzero(A[i] == up + 1)[0]]
            stride *= L
    sp = np.random.default_rng(13).integers(0, 2 ** 32, (N, R // 32), dtype=np.uint32)
    out["checkerboard/initial"] = sp.copy()
    acc = np.zeros(R, np.int64)
    ffi.checkerboard_sweeps(L, D, R, sp, Jf, ffi.thresholds_fixed64(0.9, D), 5, 77, 0, 3, acc, M=4)
    out["checkerboard/final"] = sp
    out["checkerboard/accepted"] = acc
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_v1.npz")
    np.savez_compressed(path, **compute())
    print("wrote", path, os.path.getsize(path), "bytes")
