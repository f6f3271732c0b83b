#!/usr/bin/env python
"""Stress the multi-sweep brick kernel (k_checkerboard_flow) at BASELINE size: one oracle trajectory, many device runs
from the same start, every mismatch located (site -> brick, replica -> group). usage: stress_flow.py [reps] [ladder|uniform] [nsweeps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rrrmc_b200 as rb
from oracle import ffi
from rrrmc_b200._ffi import check, lib, ptr
from tests.helpers import ea_instance
from tests.test_gpu_checkerboard import _fwd, _multispin, _ladder_tbls, _poisson_tbl

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mode = sys.argv[2] if len(sys.argv) > 2 else "ladder"
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L, D, R = 64, 3, 1024
A, J = ea_instance(L, D, seed=64)
X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
sp = _multispin(C0)
t0 = time.time()
if mode == "ladder":
    lo, hi, NW = (sys.argv[4].split(",") if len(sys.argv) > 4 else ("0.8", "1.6", "4"))
    bg = np.geomspace(float(lo), float(hi), 8); tbls = _ladder_tbls(bg); NW = int(NW); print("ladder", lo, hi, "NW", NW)
    ffi.checkerboard_sweeps_poisson_ladder(L, D, R, sp, _fwd(A, J, L, D), tbls, NW, 0xABCDEF, 5, nsw)
else:
    tbl = _poisson_tbl(float(mode.split(":")[1]) if ":" in mode else 1.0, D); NW = ffi.cb_poisson_nw(tbl); print("NW", NW)
    ffi.checkerboard_sweeps_poisson(L, D, R, sp, _fwd(A, J, L, D), tbl, NW, 0xABCDEF, 5, nsw)
print("oracle %.1fs" % (time.time() - t0), flush=True)
from tests.test_gpu_checkerboard import _from_multispin
want_cfg = _from_multispin(sp, R)
want_chunks = np.asarray(want_cfg.chunks).copy()
bad = 0
t0 = time.time()
for rep in range(reps):
    X._upload(C0)
    if mode == "ladder":
        check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(X._state, ptr(tbls), 8, NW, 0xABCDEF, 5, nsw))
    else:
        check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), NW, 0xABCDEF, 5, nsw))
    dl = X._download()
    if np.array_equal(np.asarray(dl.chunks), want_chunks):
        continue
    bad += 1
    again = X._download()
    stable = np.array_equal(np.asarray(again.chunks), np.asarray(dl.chunks))
    got = _multispin(dl)
    diff = got ^ sp
    sites = np.flatnonzero(diff.any(axis=1))
    col = [(int(s % L) + int(s // L % L) + int(s // (L * L))) & 1 for s in sites]
    c0 = [(int(s % L), int(s // L % L), int(s // (L * L))) for s, c in zip(sites, col) if c == 0]
    c1 = [(int(s % L), int(s // L % L), int(s // (L * L))) for s, c in zip(sites, col) if c == 1]
    per_group = [int(sum(bin(int(v)).count("1") for v in diff[sites][:, 4 * g:4 * g + 4].ravel())) for g in range(8)]
    print("rep %d: %d sites differ (second download identical: %s); colour-0 sites %s; colour-1 sites %d %s; differing bits per group %s" % (
        rep, len(sites), stable, c0, len(c1), c1[:6], per_group), flush=True)
print("mismatching runs: %d of %d (%s, %d sweeps) in %.1fs" % (bad, reps, mode, nsw, time.time() - t0))
