"""Golden vectors (tests/golden/hotpath_v1.npz, made by tests/golden/make_golden.py from the oracle):
CPU: the oracle still reproduces them; GPU: the engine reproduces them through the C ABI without the oracle in the loop."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hotpath_v1.npz"))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def test_oracle_reproduces_golden():
    now = mg.compute()
    assert set(now) == set(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(np.asarray(now[k]), GOLD[k]), k


def _engine_graph(name, R):
    import rrrmc_b200 as rb
    from tests.helpers import ea_instance, sk_binary, sk_gauss
    if name == "EA(4,2)":
        A, J = ea_instance(4, 2, (-1, 1), 1); return rb.GraphEA(4, 2, replicas=R, A=A, J=J)
    if name == "EA(2,3)":
        A, J = ea_instance(2, 3, (-1, 1), 2); return rb.GraphEA(2, 3, replicas=R, A=A, J=J)
    if name == "EA(3,3,(-1,0,1))":
        A, J = ea_instance(3, 3, (-1, 0, 1), 3); return rb.GraphEA(3, 3, (-1, 0, 1), replicas=R, A=A, J=J)
    if name == "EANormal(3,2)":
        A, J = ea_instance(3, 2, seed=4, gaussian=True); return rb.GraphEANormal(3, 2, replicas=R, A=A, J=J)
    if name == "EANormalDiscretized(3,2,(-1,0,1))":
        A, cJ = ea_instance(3, 2, seed=14, gaussian=True); return rb.GraphEANormalDiscretized(3, 2, (-1, 0, 1), replicas=R, A=A, cJ=cJ)
    if name == "SK(10)":
        return rb.GraphSK(10, replicas=R, J=sk_binary(10, 5))
    if name == "SKNormal(10)":
        return rb.GraphSKNormal(10, replicas=R, J=sk_gauss(10, 6))
    if name == "QT(12,4)":
        return rb.GraphQT(12, 4, 0.73, replicas=R)
    if name == "Quant(6,4,SK)":
        return rb.GraphQSKT(6, 4, 0.5, 2.0, replicas=R, J=sk_binary(6, 7))
    if name == "Quant(6,4,SKNormal)":
        return rb.GraphQSKNormalT(6, 4, 0.5, 2.0, replicas=R, J=sk_gauss(6, 8))
    if name == "QEAT(3,2,4)":
        A, J = ea_instance(3, 2, seed=15, gaussian=True); return rb.GraphQEAT(3, 2, 4, 0.5, 2.0, replicas=R, A=A, J=J)
    if name == "Quant(6,4,Empty)":
        return rb.GraphQ0T(6, 4, 0.5, 2.0, replicas=R)
    raise KeyError(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mg.CASES))
def test_engine_reproduces_golden(name):
    import rrrmc_b200 as rb
    from tests.helpers import random_config
    R = 4  # the golden chain is Philox chain 3 = replica 3 of the batch
    X = _engine_graph(name, R)
    s0 = random_config(X.N, seed=11)
    C0 = rb.Config(X.N, R, chunks=np.tile(s0, (R, 1)))
    assert np.atleast_1d(rb.energy(X, C0))[3] == GOLD[f"{name}/energy"][0]
    assert np.array_equal(np.asarray(rb.all_delta_energy(X, C0, 3), np.float64), GOLD[f"{name}/delta_energy"])
    assert tuple(GOLD[f"{name}/neighbors1"]) == rb.neighbors(X, 1)
    if f"{name}/allDE" in GOLD.files:
        assert np.array_equal(np.asarray(rb.allDeltaE(X), np.float64), GOLD[f"{name}/allDE"])
    for sname in mg.SAMPLERS:
        kw = {"schedule": "random"} if sname == "standardMC" else {}
        Es, Cf = getattr(rb, sname)(X, mg.BETA, mg.ITERS, step=mg.STEP, seed=mg.SEED, C0=C0, quiet=True, **kw)
        assert np.array_equal(np.asarray(Es, np.float64)[:, 3], GOLD[f"{name}/{sname}/Es"]), sname
        assert np.array_equal(Cf.chunks[3], GOLD[f"{name}/{sname}/final"]), sname


@pytest.mark.gpu
def test_engine_checkerboard_reproduces_golden():
    import rrrmc_b200 as rb
    from oracle import ffi  # thresholds helper only (host-side table)
    from rrrmc_b200._ffi import check, lib, ptr
    from tests.helpers import ea_instance
    L, D, R = 4, 3, 64
    A, J = ea_instance(L, D, seed=9)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    sp = GOLD["checkerboard/initial"]
    bits = np.unpackbits(sp.view(np.uint8).reshape(X.N, R // 8), axis=1, bitorder="little").T
    X._upload(rb.Config.from_bits(bits))
    thr = ffi.thresholds_fixed64(0.9, D)
    check(lib().rrrmc_checkerboard_sweeps(X._state, ptr(thr), D, 5, 4, 77, 0, 3))
    got = X._download()
    want = np.unpackbits(GOLD["checkerboard/final"].view(np.uint8).reshape(X.N, R // 8), axis=1, bitorder="little").T
    assert np.array_equal(got.s, want.astype(bool))


# ---- count-table acceptance procedures of the checkerboard kernels (tests/golden/checkerboard_v2.npz) ------------
GOLD_CB = np.load(os.path.join(HERE, "golden", "checkerboard_v2.npz"))
_spec_cb = importlib.util.spec_from_file_location("make_golden_cb", os.path.join(HERE, "golden", "make_golden_cb.py"))
mgcb = importlib.util.module_from_spec(_spec_cb)
_spec_cb.loader.exec_module(mgcb)


def test_oracle_reproduces_checkerboard_golden():
    now = mgcb.compute()
    assert set(now) == set(GOLD_CB.files)
    for k in GOLD_CB.files:
        assert np.array_equal(np.asarray(now[k]), GOLD_CB[k]), k


def test_library_tables_match_checkerboard_golden():
    """Host-only entry points of the C ABI (no device): the table builders reproduce the committed tables."""
    from rrrmc_b200._ffi import check, lib, ptr
    for name in mgcb.CASES:
        thr = np.ascontiguousarray(GOLD_CB[f"{name}/thr"]); want = GOLD_CB[f"{name}/tbl"]
        tbl = np.zeros(len(want), np.uint32)
        check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), mgcb.D, ptr(tbl), len(tbl)))
        assert np.array_equal(tbl, want), name
    thr = np.ascontiguousarray(GOLD_CB["sparse_b0.9/thr"]); want = GOLD_CB["sparse_b0.9/tbl"]
    tbl = np.zeros(len(want), np.uint32)
    check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), mgcb.D, ptr(tbl), len(tbl)))
    assert np.array_equal(tbl, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mgcb.CASES) + ["sparse_b0.9"])
def test_engine_checkerboard_count_procedures_reproduce_golden(name):
    """The CUDA kernels against the committed vectors, through the C ABI, without the oracle in the loop."""
    import rrrmc_b200 as rb
    from rrrmc_b200._ffi import check, lib, ptr
    from tests.helpers import ea_instance
    L, D, R = mgcb.L, mgcb.D, mgcb.R
    A, J = ea_instance(L, D, seed=21)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    sp = GOLD_CB["initial"]
    bits = np.unpackbits(sp.view(np.uint8).reshape(X.N, R // 8), axis=1, bitorder="little").T
    X._upload(rb.Config.from_bits(bits))
    tbl = np.ascontiguousarray(GOLD_CB[f"{name}/tbl"])
    if name.startswith("poisson"):
        check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), mgcb.CASES[name][1], mgcb.SEED, mgcb.SWEEP0, mgcb.NSW))
    else:
        check(lib().rrrmc_checkerboard_sweeps_sparse(X._state, ptr(tbl), len(tbl), mgcb.SEED, mgcb.SWEEP0, mgcb.NSW))
    got = X._download()
    want = np.unpackbits(GOLD_CB[f"{name}/final"].view(np.uint8).reshape(X.N, R // 8), axis=1, bitorder="little").T
    assert np.array_equal(got.s, want.astype(bool))


# ---- extremal_opt (tests/golden/eo_v1.npz, made by tests/golden/make_golden_eo.py) -----------------------------------
GOLD_EO = np.load(os.path.join(HERE, "golden", "eo_v1.npz"))
_spec_eo = importlib.util.spec_from_file_location("make_golden_eo", os.path.join(HERE, "golden", "make_golden_eo.py"))
mge = importlib.util.module_from_spec(_spec_eo)
_spec_eo.loader.exec_module(mge)


def test_oracle_reproduces_eo_golden():
    now = mge.compute({name: GOLD_EO[f"{name}/ftau"] for name in mge.CASES})
    assert set(now) == set(GOLD_EO.files)
    for k in GOLD_EO.files:
        assert np.array_equal(np.asarray(now[k]), GOLD_EO[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mge.CASES))
def test_engine_reproduces_eo_golden(name):
    import rrrmc_b200 as rb
    from tests.helpers import ea_instance, random_config
    R = 3  # the golden chain is Philox chain 2 = replica 2 of the batch
    if name == "EA(4,3)":
        A, J = ea_instance(4, 3, (-1, 1), 21); X = rb.GraphEA(4, 3, replicas=R, A=A, J=J)
    elif name == "EA(2,3)":
        A, J = ea_instance(2, 3, (-1, 1), 22); X = rb.GraphEA(2, 3, replicas=R, A=A, J=J)
    elif name == "EA(5,2,(-1,0,1))":
        A, J = ea_instance(5, 2, (-1, 0, 1), 23); X = rb.GraphEA(5, 2, (-1, 0, 1), replicas=R, A=A, J=J)
    elif name == "EA(4,2,(-2,-1,1,2))":
        A, J = ea_instance(4, 2, (-2, -1, 1, 2), 24); X = rb.GraphEA(4, 2, (-2, -1, 1, 2), replicas=R, A=A, J=J)
    else:
        X = rb.GraphQT(12, 4, 0.73, replicas=R)
    s0 = random_config(X.N, seed=12)
    C0 = rb.Config(X.N, R, chunks=np.tile(s0, (R, 1)))
    Cf, Emin, Cmin, itmin, Es = rb.extremal_opt(X, mge.TAU, mge.ITERS, step=mge.STEP, seed=mge.SEED, C0=C0, quiet=True,
                                                ftau=GOLD_EO[f"{name}/ftau"], return_Es=True)
    assert np.array_equal(np.asarray(Es, np.float64)[:, 2], GOLD_EO[f"{name}/Es"])
    assert np.array_equal(Cf.chunks[2], GOLD_EO[f"{name}/final"]) and np.array_equal(Cmin.chunks[2], GOLD_EO[f"{name}/Cmin"])
    assert float(np.atleast_1d(Emin)[2]) == GOLD_EO[f"{name}/Emin_itmin"][0] and int(np.atleast_1d(itmin)[2]) == GOLD_EO[f"{name}/Emin_itmin"][1]
