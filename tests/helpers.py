"""Shared fixtures: reference-style toy instances (test/runtests.jl:26-123) built on the oracle."""
import numpy as np

from oracle import ffi


def ea_instance(L, D, lev=(-1, 1), seed=0, gaussian=False):
    """GraphEA(L,D,LEV) / GraphEANormal(L,D): lattice from gen_EA, couplings via gen_J draw order (EA.jl:181-189,565-574)."""
    rng = np.random.default_rng(seed)
    A = ffi.gen_EA(L, D)
    nb = int((A > np.arange(1, len(A) + 1)[:, None]).sum())
    draws = rng.standard_normal(nb) if gaussian else rng.choice(np.asarray(lev, dtype=np.float64), nb)
    J = ffi.gen_J(A, draws)
    return A, (J if gaussian else J.astype(np.int64))


def sk_gauss(N, seed=0):
    """gen_J_gauss (SK.jl:170-179)."""
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((N, N)) / np.sqrt(N)
    J = np.triu(J, 1)
    return J + J.T


def sk_binary(N, seed=0):
    """gen_J (SK.jl:17-26)."""
    rng = np.random.default_rng(seed)
    J = np.triu(rng.integers(0, 2, (N, N)), 1)
    return (J + J.T).astype(np.uint8)


def random_config(N, seed=0):
    rng = np.random.default_rng(seed)
    nch = (N + 63) // 64
    ch = rng.integers(0, 2 ** 64, nch, dtype=np.uint64)
    if N % 64:
        ch[-1] &= np.uint64((1 << (N % 64)) - 1)
    return ch


def bits(ch, N):
    return np.array([(int(ch[i >> 6]) >> (i & 63)) & 1 for i in range(N)], dtype=np.int64)


def reference_graphs(seed=1):
    """The hot-path subset of test/runtests.jl's graph list (lines 46-67, 78-81)."""
    out = {}
    for (L, D) in ((2, 3), (3, 2)):
        A, J = ea_instance(L, D, (-1, 1), seed)
        out[f"EA({L},{D})"] = ffi.Graph.ea_int(A, J, (-1, 1))
        A, J = ea_instance(L, D, (-1, 0, 1), seed + 1)
        out[f"EA({L},{D},(-1,0,1))"] = ffi.Graph.ea_int(A, J, (-1, 0, 1))
        A, J = ea_instance(L, D, seed=seed + 2, gaussian=True)
        out[f"EANormal({L},{D})"] = ffi.Graph.ea_f64(A, J)
        A, cJ = ea_instance(L, D, seed=seed + 5, gaussian=True)
        out[f"EANormalDiscretized({L},{D},(-1,0,1))"] = ffi.Graph.ea_discretized(A, cJ, (-1, 0, 1))
    out["SK(10)"] = ffi.Graph.sk_bin(sk_binary(10, seed))
    out["SKNormal(10)"] = ffi.Graph.sk_f64(sk_gauss(10, seed))
    out["Quant(10,8,Empty)"] = ffi.Graph.quant(10, 8, 0.5, 2.0, ffi.EMPTY)
    out["Quant(10,8,SK)"] = ffi.Graph.quant(10, 8, 0.5, 2.0, ffi.SK_BIN, sk_binary(10, seed + 3))
    out["Quant(10,8,SKNormal)"] = ffi.Graph.quant(10, 8, 0.5, 2.0, ffi.SK_F64, sk_gauss(10, seed + 4))
    # random regular graphs, test/runtests.jl:35-44
    import rrrmc_b200 as rb   # host-side generators only (gen_RRG, gen_J_graph)
    rng = np.random.default_rng(seed + 7)
    Ar = rb.gen_RRG(10, 3, rng)
    out["RRG(10,3)"] = ffi.Graph.rrg_int(Ar, rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), Ar).astype(np.int64))
    out["RRG(10,3,(-1,0,1))"] = ffi.Graph.rrg_int(Ar, rb.gen_J_graph(lambda n: rng.choice([-1.0, 0.0, 1.0], n), Ar).astype(np.int64), (-1, 0, 1))
    out["RRGNormalDiscretized(10,3,(-1,0,1))"] = ffi.Graph.rrg_discretized(Ar, rb.gen_J_graph(lambda n: rng.standard_normal(n), Ar), (-1, 0, 1))
    out["RRGNormal(10,3)"] = ffi.Graph.ea_f64(Ar, rb.gen_J_graph(lambda n: rng.standard_normal(n), Ar))
    A, J = ea_instance(3, 2, seed=seed + 6, gaussian=True)
    out["QEAT(3,2,5)"] = ffi.Graph.quant(9, 5, 0.5, 2.0, ffi.EA_F64, J, A)   # GraphQEAT, QAliases.jl:51-81
    return out
