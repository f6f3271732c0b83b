"""Multi-GPU host logic: replicas are independent chains on a shared read-only instance, so the replica axis is
sharded across ranks (one process per GPU) with no data-path collective. torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only for what the path really exchanges: the per-replica observables behind `hook`, and the
energies that drive parallel-tempering swaps. Swaps exchange β *labels* — every rank evaluates identical decisions
from the all-gathered energies and a shared counter-based RNG — so no spin data ever crosses NVLink.

The reference has no multi-replica or multi-process path; this module is new-engine plumbing."""
import numpy as np

ALIGN = 128  # one 128-bit multispin load serves 128 replicas: shards are multiples of it


def replica_range(rank, world, total, align=ALIGN):
    """Contiguous [lo, hi) of the `total` replicas owned by `rank`; every boundary is a multiple of `align`
    (the remainder goes to the last ranks one block at a time)."""
    if total % align:
        raise ValueError(f"total replicas {total} must be a multiple of {align}")
    blocks = total // align
    base, rem = divmod(blocks, world)
    counts = [base + (1 if r >= world - rem else 0) for r in range(world)]
    lo = sum(counts[:rank]) * align
    return lo, lo + counts[rank] * align


def _dist():
    try:
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() else None
    except Exception:
        return None


def all_gather(x_local, device=None):
    """All-gather of a per-replica observable (E, m, accepted ...): (R_local, ...) -> (R_total, ...), rank order.
    Shards may differ in size (padded to the largest for the collective). Without an initialised process group it is
    the identity."""
    x_local = np.ascontiguousarray(x_local)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return x_local
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([x_local.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(v.item()) for v in ns]
    nmax = max(ns)
    pad = np.zeros((nmax,) + x_local.shape[1:], x_local.dtype)
    pad[:x_local.shape[0]] = x_local
    t = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy()[:k] for o, k in zip(out, ns)], axis=0)


def classical_action(beta, E):
    """β·E: the weight exponent of a classical graph (GraphEA, GraphSK...)."""
    return beta * E


def quantum_action(M, Gamma):
    """β-dependent action of GraphQuant (QT.jl:163-199): E_β(s) = K(β)·e0(s) + E_cl(s)/M with
    K(β) = fourK(β)/4, fourK = round(2/β·log coth(βΓ/M), digits=8) (QT.jl:165). `terms` = (e0, E_cl_sum)."""
    def fourK(beta):
        return np.round(2.0 / beta * np.log(1.0 / np.tanh(beta * Gamma / M)), 8)

    def action(beta, terms):
        e0, ecl = terms
        return beta * (fourK(beta) / 4.0 * e0 + ecl / M)
    return action


def quant_terms(X, C):
    """(e0, E_cl) of every local replica of a GraphQuant batch for `quantum_action`: E_cl = Σ_k E_classical(slice k)
    (Renergies, QT.jl:201-211) and e0 = −Σ_{Trotter bonds} σσ′ recovered from E = fourK/4·e0 + E_cl/M (QT.jl:185-199)
    with the replica's own fourK."""
    import rrrmc_b200 as rb
    E = np.atleast_1d(np.asarray(rb.energy(X, C), np.float64))
    ecl = np.asarray(rb.Renergies(X), np.float64).reshape(X.replicas, X.M).sum(axis=1)
    fk = np.full(X.replicas, X.fourK) if getattr(X, "betas", None) is None else \
        np.round(2.0 / X.betas * np.log(1.0 / np.tanh(X.betas * X.Γ / X.M)), 8)
    return np.rint((E - ecl / X.M) * 4.0 / fk), ecl


class TemperingLadder:
    """Parallel tempering by label exchange. `order[k]` is the replica currently holding the k-th inverse
    temperature of `betas` (ascending ladder). Every rank owns an identical copy and updates it identically."""

    def __init__(self, betas, seed=0, action=classical_action):
        self.betas = np.asarray(betas, np.float64)
        self.order = np.arange(len(self.betas))
        self.seed, self.action = int(seed), action
        self.attempts = np.zeros(len(self.betas) - 1, np.int64)
        self.accepts = np.zeros(len(self.betas) - 1, np.int64)

    def beta_of_replica(self):
        b = np.empty_like(self.betas)
        b[self.order] = self.betas
        return b

    def swap(self, terms, sweep):
        """One round of neighbour swaps (pairs (k, k+1) with k ≡ sweep mod 2). `terms`: per-replica energy terms in
        replica order — an array E, or a tuple of arrays for a β-dependent action — already all-gathered.
        Acceptance min(1, exp(-ΔS)), ΔS = [S(β_k, x_b) + S(β_{k+1}, x_a)] − [S(β_k, x_a) + S(β_{k+1}, x_b)]."""
        tup = terms if isinstance(terms, tuple) else (np.asarray(terms, np.float64),)
        rng = np.random.Generator(np.random.Philox(key=self.seed, counter=[int(sweep), 0, 0, 0]))
        u = rng.random(len(self.betas))
        for k in range(int(sweep) % 2, len(self.betas) - 1, 2):
            a, b = self.order[k], self.order[k + 1]
            xa = tuple(t[a] for t in tup); xb = tuple(t[b] for t in tup)
            if not isinstance(terms, tuple):
                xa, xb = xa[0], xb[0]
            dS = (self.action(self.betas[k], xb) + self.action(self.betas[k + 1], xa)) - \
                 (self.action(self.betas[k], xa) + self.action(self.betas[k + 1], xb))
            self.attempts[k] += 1
            if dS <= 0 or u[k] < np.exp(-dS):
                self.order[k], self.order[k + 1] = b, a
                self.accepts[k] += 1
        return self.beta_of_replica()


RANK_SEED_STRIDE = 1_000_003   # seed offset per global replica index of a shard's first replica (tempered_run)


class ReplicaShard:
    """This rank's slice of a replica batch of `total` chains."""

    def __init__(self, total, rank=None, world=None, align=ALIGN):
        dist = _dist()
        self.rank = rank if rank is not None else (dist.get_rank() if dist else 0)
        self.world = world if world is not None else (dist.get_world_size() if dist else 1)
        self.total = total
        self.lo, self.hi = replica_range(self.rank, self.world, total, align)

    @property
    def count(self):
        return self.hi - self.lo

    def local(self, x_global):
        return np.asarray(x_global)[self.lo:self.hi]


def tempered_run(X, ladder, shard, rounds, iters_per_round, sampler, *, seed=1, C0=None, terms_fn=None, energy_fn=None,
                 on_device=False, **kw):
    """Runs `sampler(X, β_local, iters_per_round, ...)` on this rank's batch for `rounds` rounds with a label swap
    after each: energies are all-gathered, every rank applies the same swaps, the local β vector is refreshed.
    -> (E_history (rounds, R_total), final Config of the local batch). on_device=True keeps the configuration on the
    device between rounds (interface.ON_DEVICE): only energies and β labels cross PCIe and the batch is downloaded once
    at the end (the default round-trips it through the host three times per round: sampler result, energy(X, C), C0)."""
    C = C0
    if on_device:
        import rrrmc_b200 as rb
        if C0 is not None:
            X._upload(C0)
        else:
            from ._ffi import check, lib
            check(lib().rrrmc_state_randomize(X._ensure_state(), seed + RANK_SEED_STRIDE * shard.lo))
        C = rb.ON_DEVICE
    hist = []
    # The engine keys its counter RNG by (seed, LOCAL chain index) and draws the initial configuration from the seed
    # alone, so the shards must not share a seed: replica r of every rank would start from the same configuration and
    # consume the same stream, and the swap rule assumes independent chains. The offset is a function of the shard's
    # first GLOBAL replica, so a replica's stream does not depend on how many ranks the batch is spread over... as long
    # as the shard boundaries are the same.
    rank_seed = seed + RANK_SEED_STRIDE * shard.lo
    for rd in range(rounds):
        beta_local = shard.local(ladder.beta_of_replica())
        if hasattr(X, "set_betas"):   # GraphQuant: fourK follows the β a replica currently holds (QT.jl:165)
            X.set_betas(beta_local)
        Es, C = sampler(X, beta_local, iters_per_round, step=iters_per_round, seed=rank_seed + 7919 * rd, C0=C, quiet=True, **kw)
        # the swap weighs the configurations as they are now: the last sample of Es predates the last move (the hook
        # instant of RRRMC.jl:104 is before the move), so take energy(X, C) of the returned configuration
        if energy_fn is None:
            import rrrmc_b200 as rb
            energy_fn = rb.energy
        E_local = np.atleast_1d(np.asarray(energy_fn(X, C), np.float64)).reshape(-1)
        if terms_fn is None:
            E_all = all_gather(E_local)
            ladder.swap(E_all, rd)
            hist.append(E_all)
        else:
            t_local = terms_fn(X, C)
            t_all = tuple(all_gather(t) for t in t_local)
            ladder.swap(t_all, rd)
            hist.append(all_gather(E_local))
    return np.array(hist), (X._download() if on_device else C)


def tempered_checkerboard(X, beta_group, rounds, sweeps_per_round, *, seed=1, C0=None, read_every=0):
    """Parallel tempering of a ±J GraphEA batch on the checkerboard schedule, entirely on the device: the β ladder lies
    over the 128-replica groups (lane l of every group is one ladder, group g its rung at beta_group[g]); a round is
    `sweeps_per_round` ladder sweeps (one launch of the multi-sweep brick kernel, rrrmc_checkerboard_sweeps_poisson_ladder)
    followed by rrrmc_tempering_exchange (energies, decisions and the exchange of configurations are kernels). The host
    only launches; nothing crosses PCIe between rounds but the eight-or-so β values. Replica shards of a multi-GPU job
    hold whole ladders, so there is no collective in the loop (reductions of the observables happen after it).
    -> (exchanges accepted per neighbouring pair, summed over the 128 ladders; rounds attempted per pair)."""
    from . import _ffi
    from ._ffi import check, lib, ptr
    st = X._ensure_state()
    if C0 is not None:
        X._upload(C0)
    bg = np.ascontiguousarray(beta_group, np.float64)
    G = len(bg)
    D = X.D
    tbls = np.zeros((G, _ffi.CBP_LEN), np.uint32)
    NW = 0
    for g in range(G):
        thr = np.array([min(int(np.exp(-bg[g] * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, D + 1)], dtype=np.uint64)
        check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(tbls[g]), _ffi.CBP_LEN))
        nw = lib().rrrmc_checkerboard_poisson_nw(ptr(tbls[g]), 0.0)
        if nw == 0:
            raise ValueError(f"β={bg[g]} (group {g}) is too warm for the poisson procedure")
        NW = max(NW, nw)
    acc = np.zeros(G - 1, np.int64)
    got = np.zeros(G - 1, np.int64)
    for rd in range(rounds):
        check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(st, ptr(tbls), G, NW, seed, rd * sweeps_per_round, sweeps_per_round))
        last = rd == rounds - 1 or (read_every and (rd + 1) % read_every == 0)
        check(lib().rrrmc_tempering_exchange(st, ptr(bg), G, seed + 0x9E3779B97F4A7C15 & (2 ** 64 - 1), rd, ptr(got) if last else None))
        if last:
            acc += got
    attempts = np.array([sum(1 for rd in range(rounds) if rd % 2 == g % 2) for g in range(G - 1)], np.int64) * 128
    return acc, attempts
