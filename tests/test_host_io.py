"""Host-side helpers that need no GPU: the reference's on-disk instance format (gen_AJ, src/graphs/EA.jl:73-118) and the
τ-EO rank table fτ = cumsum(j^-τ) (src/DeltaE.jl:443) summed like Julia's pairwise `cumsum`."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance


def test_gen_AJ_round_trip(tmp_path):
    L = 5
    A, J = ea_instance(L, 2, seed=3, gaussian=True)
    fn = tmp_path / "inst.txt"
    rb.write_AJ(fn, L, A, J, name="t")
    L2, D2, A2, J2 = rb.gen_AJ(fn)
    assert (L2, D2) == (L, 2) and np.array_equal(A2, A) and np.array_equal(J2, J)   # repr() round-trips Float64
    # the loaded instance is the same model: energies agree on the oracle
    s = np.array([0x123456789ABCDEF], np.uint64)
    assert ffi.Graph.ea_f64(A2, J2).energy(s) == ffi.Graph.ea_f64(A, J).energy(s)


def test_gen_AJ_hand_written_file(tmp_path):
    # 3x3 periodic lattice: site 1 has neighbours 2, 3, 4, 7 (gen_EA(3, 2), column-major, EA.jl:24-43)
    A = rb.gen_EA(3, 2)
    assert A[0].tolist() == [2, 3, 4, 7]
    lines = ["type: EA", "size: 3", "name: hand"]
    val = {}
    for x in range(9):
        for y in A[x]:
            if y > x + 1:
                val[(x + 1, int(y))] = 0.25 * (x + 1) - 0.5 * int(y)
                lines.append(f"{x + 1}  {int(y)}   {val[(x + 1, int(y))]}")
    fn = tmp_path / "hand.txt"
    fn.write_text("\n".join(lines) + "\n")
    L, D, A2, J = rb.gen_AJ(fn)
    assert L == 3 and D == 2 and np.array_equal(A2, A)
    for (x, y), v in val.items():
        assert J[x - 1, list(A[x - 1]).index(y)] == v and J[y - 1, list(A[y - 1]).index(x)] == v   # symmetric, slot-aligned


@pytest.mark.parametrize("corrupt", ["header", "size", "missing", "twice", "notbond", "fields"])
def test_gen_AJ_rejects_malformed(tmp_path, corrupt):
    A, J = ea_instance(4, 2, seed=1, gaussian=True)
    fn = tmp_path / "x.txt"
    rb.write_AJ(fn, 4, A, J)
    ls = fn.read_text().splitlines()
    if corrupt == "header": ls[0] = "kind: EA"
    if corrupt == "size": ls[1] = "size 4"
    if corrupt == "missing": ls.pop()
    if corrupt == "twice": ls.append(ls[5])
    if corrupt == "notbond": ls[5] = "1 11 0.5"
    if corrupt == "fields": ls[6] = "1 2"
    fn.write_text("\n".join(ls) + "\n")
    with pytest.raises(ValueError):
        rb.gen_AJ(fn)


@pytest.mark.parametrize("N", [1, 2, 127, 128, 129, 300, 5000])
def test_eo_ftau_pairwise_cumsum(N):
    tau = 1.3
    f = rb.eo_ftau(N, tau)
    v = np.arange(1, N + 1, dtype=np.float64) ** -tau
    assert f.shape == (N,) and (np.diff(f) > 0).all()
    assert np.allclose(f, np.cumsum(v), rtol=1e-13, atol=0)
    if N <= 128:   # one block: result[i] = v[1] + (v[2] + … + v[i]), the inner sum left to right (Base accumulate.jl)
        want = np.concatenate([[v[0]], v[0] + np.cumsum(v[1:])]) if N > 1 else v[:1]
        assert np.array_equal(f, want)
    import math
    assert abs(f[-1] - math.fsum(v)) <= 4 * np.spacing(f[-1])


def test_dfloat_levels_reduce_to_integer_units():
    """DFloat64 levels (src/DFloats.jl: round(x·10^5) as Int64) -> small integer levels and the gcd unit."""
    assert rb.interface.dfloat_levels((-1.5, 0.5, 1.5)) == ((-3, 1, 3), 50000)
    assert rb.interface.dfloat_levels((-1.0, 0.25)) == ((-4, 1), 25000)
    assert rb.interface.dfloat_levels((0.1, -0.3, 0.7)) == ((1, -3, 7), 10000)
    assert rb.interface.dfloat_levels((0.00001, 0.00127)) == ((1, 127), 1)
    with pytest.raises(ValueError):
        rb.interface.dfloat_levels((0.123456, 1.0))          # more than 5 decimal digits (EA.jl:133)
    with pytest.raises(NotImplementedError):
        rb.interface.dfloat_levels((0.00001, 1.0))           # would need couplings beyond int8
