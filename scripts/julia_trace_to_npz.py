#!/usr/bin/env python
"""Converts the text dump of scripts/dump_julia_trace.jl into tests/golden/julia_trace_*.npz
(read by tests/test_oracle_pins.py::test_julia_trace_replays).   usage: julia_trace_to_npz.py <dump dir> <out.npz>"""
import os
import sys

import numpy as np


def main(d, out):
    hdr = dict(line.split(None, 1) for line in open(os.path.join(d, "header.txt")).read().splitlines() if line.strip())
    N, D = int(hdr["N"]), int(hdr["D"])

    def arr(name, conv):
        return np.array([conv(x) for x in open(os.path.join(d, name)).read().split()])
    u64 = lambda x: int(x, 16) if x.startswith("0x") else int(x)
    A = arr("A.txt", int).astype(np.int64).reshape(N, 2 * D)
    J = arr("J.txt", float).astype(np.float64).reshape(N, 2 * D)
    np.savez_compressed(out, sampler=hdr["sampler"], graph=hdr["graph"], L=int(hdr["L"]), D=D, beta=float(hdr["beta"]),
                        iters=int(hdr["iters"]), step=int(hdr["step"]), seed=int(hdr["seed"]), julia=hdr["julia"],
                        A=A, J=J, C0=arr("C0.txt", u64).astype(np.uint64), C1=arr("C1.txt", u64).astype(np.uint64),
                        kind=arr("kind.txt", u64).astype(np.uint8), ival=arr("ival.txt", int).astype(np.int64),
                        fval=arr("fval.txt", float).astype(np.float64), Es=arr("Es.txt", float).astype(np.float64))
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
