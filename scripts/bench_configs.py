#!/usr/bin/env python
"""Secondary BASELINE configs (3, 4, 5): one JSON line each. The headline (config 2) is bench.py.
  C3: GraphEA 3D L=32 ±J, rrrMC and bklMC at β=3, 256 replicas              -> iterations/s and executed moves/s
  C4: GraphSKNormal N=4096, 512 replicas: tensor-core field init + lock-step Metropolis sweeps
  C5: GraphQSKT Nk=1024, M=64, Γ=0.3, rrrMC, 64 replicas
  rrg: the benchmark of the RRR paper (reference scripts/scripts.jl:23-40): GraphRRG N=10^4, K=3, ±J, β=2 with all four
       samplers (met = standardMC in the reference's random-site order, bkl, rrr, wtm), 256 replicas, and the CPU oracle
       on one host core beside each
  eo:  extremal_opt (τ = 1.3) on the same RRG instance family, moves/s beside the CPU oracle on one core
usage: python scripts/bench_configs.py [c3] [c4] [c5] [rrg] [eo] [--quick]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb

quick = "--quick" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4", "c5", "rrg"]


def emit(**kw):
    print(json.dumps(kw), flush=True)


if "c3" in which:
    L, D, R, beta = 32, 3, int(os.environ.get("C3_R", "256")), 3.0
    X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
    # equilibrate with checkerboard sweeps so that the low-T rejection-free samplers start from a typical state
    _, C = rb.standardMC(X, beta, 200 * X.N, step=200 * X.N, seed=1, quiet=True, schedule="checkerboard")
    for pick in ("reference", "rank"):      # ArraySet order on one lane (chain_ea.cu) / rank-select on a warp (chain_warp.cu)
        for name, fn, iters in (("rrrMC", rb.rrrMC, 200_000 if quick else 2_000_000), ("bklMC", rb.bklMC, 10_000_000 if quick else 200_000_000)):
            fn(X, beta, iters // 20, step=iters // 20, seed=2, C0=C, quiet=True, site_pick=pick)
            t0 = time.perf_counter()
            Es, C2 = fn(X, beta, iters, step=iters, seed=3, C0=C, quiet=True, site_pick=pick)
            dt = time.perf_counter() - t0
            info = X.last_run
            emit(config="C3", sampler=name, site_pick=pick, kernel="k_chain_warp<3>" if pick == "rank" else "k_chain_ea<6>",
                 L=L, D=D, replicas=R, beta=beta, iters_per_replica=iters,
                 iterations_per_s=R * iters / (info.device_ms * 1e-3), executed_moves_per_s=info.accepted_total / (info.device_ms * 1e-3),
                 us_per_move_per_chain=info.device_ms * 1e3 * R / max(1, info.accepted_total),
                 device_ms=info.device_ms, wall_s=dt, launches=info.launches)

if "c4" in which:
    N, R, beta = 4096, 512, 1.0
    X = rb.GraphSKNormal(N, replicas=R, rng=np.random.default_rng(2))
    C0 = rb.Config(N, R, rng=np.random.default_rng(3))
    for tc in (True, False):
        rb.sk_fields_init(X, C0, tensor_cores=tc)
        ms = min(rb.sk_fields_init(X, None, tensor_cores=tc)[2] for _ in range(3))
        flop = 2.0 * N * N * R
        emit(config="C4", op="local-field init", path="tcgen05 int8 x5 digit planes" if tc else "cuda cores, reference order (fp64)",
             N=N, replicas=R, device_ms=ms, useful_tflops=flop / (ms * 1e-3) / 1e12,
             int8_tops=(5 * flop / (ms * 1e-3) / 1e12) if tc else None)
    nsw = 1 if quick else 3
    # lock-step sweeps: the TMA-staged kernel (coupling row i+1 prefetched by a bulk copy) and the plain-load kernel.
    # algorithmic bytes per attempt = 8 (own field) + a·2·N·8 (field read+write of an accepted flip, a =
    # acceptance) + N·8/R (the coupling row shared by the lock-step replicas) = 8 + 65 536·a + 64 at N = 4096, R = 512,
    # set against the measured HBM peak. The fields live in shared memory and the rows come from L2, so this is the
    # grading convention; what actually bounds a site step is the per-site decision and two block barriers.
    import json as _json
    peak = 6539.5
    try:
        peak = float(_json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for variant, kname in (("0", "k_sk_lockstep_reg<4>"), ("2", "k_sk_lockstep_tma<4>"), ("1", "k_sk_lockstep<4>")):
        os.environ["RRRMC_SK_VARIANT"] = variant
        rb.sk_fields_init(X, C0, tensor_cores=True)
        rb.sk_metropolis_sweeps(X, beta, 2, seed=1)
        _, acc0, _ = rb.sk_metropolis_sweeps(X, beta, 0, seed=2, sweep0=2)
        # two run lengths: the slope is the cost of a sweep without the per-call work (uploads, the configuration download)
        t0 = time.perf_counter()
        rb.sk_metropolis_sweeps(X, beta, nsw, seed=2, sweep0=2)
        dt1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        E, acc, _ = rb.sk_metropolis_sweeps(X, beta, 4 * nsw, seed=2, sweep0=2 + nsw)
        dt = time.perf_counter() - t0
        a = float(acc.sum() - acc0.sum()) / (5 * nsw * N * R)
        per_sweep = (dt - dt1) / (3 * nsw)
        rate = N * R / per_sweep
        bpa = 8 + a * 2 * N * 8 + N * 8 / R
        emit(config="C4", op="lock-step Metropolis", kernel=kname, N=N, replicas=R, beta=beta, sweeps=[nsw, 4 * nsw], attempts_per_s=rate,
             us_per_site_step=per_sweep / N * 1e6, accept_rate=a, wall_s=[dt1, dt], mean_E_per_N=float(E.mean() / N),
             algorithmic_bytes_per_attempt=bpa, achieved_GBps=rate * bpa / 1e9, hbm_peak_GBps=peak, frac_of_hbm=rate * bpa / 1e9 / peak)
    os.environ.pop("RRRMC_SK_VARIANT", None)

if "c5" in which:
    Nk, M, G, beta, R = 1024, 64, 0.3, 2.0, 64
    X = rb.GraphQSKT(Nk, M, G, beta, replicas=R, rng=np.random.default_rng(4))
    iters = 20_000 if quick else 400_000
    _, C = rb.rrrMC(X, beta, iters // 10, step=iters // 10, seed=1, quiet=True)
    t0 = time.perf_counter()
    Es, C2 = rb.rrrMC(X, beta, iters, step=iters, seed=2, C0=C, quiet=True)
    dt = time.perf_counter() - t0
    info = X.last_run
    emit(config="C5", sampler="rrrMC(DoubleGraph)", Nk=Nk, M=M, Gamma=G, beta=beta, replicas=R, iters_per_replica=iters,
         iterations_per_s=R * iters / (info.device_ms * 1e-3), accepted_per_s=info.accepted_total / (info.device_ms * 1e-3),
         device_ms=info.device_ms, wall_s=dt, Qenergy_mean=float(np.mean(rb.Qenergy(X, C2))))

if "rrg" in which:
    from oracle import ffi   # CPU restatement: the baseline leg of this script only
    N, K, R, beta = 10_000, 3, int(os.environ.get("RRG_R", "256")), 2.0
    rng = np.random.default_rng(8370000274 % 2 ** 32)
    A = rb.gen_RRG(N, K, rng)
    J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), A).astype(np.int64)
    X = rb.GraphRRG(N, K, replicas=R, A=A, J=J)
    _, C = rb.standardMC(X, beta, 300 * N, step=300 * N, seed=1, quiet=True, schedule="random")   # equilibrate
    scale = 10 if quick else 1
    runs = (("standardMC", lambda it, **kw: rb.standardMC(X, beta, it, schedule="random", **kw), 20_000_000 // scale),
            ("rrrMC", lambda it, **kw: rb.rrrMC(X, beta, it, **kw), 2_000_000 // scale),
            ("bklMC", lambda it, **kw: rb.bklMC(X, beta, it, **kw), 40_000_000 // scale),
            ("rrrMC", lambda it, **kw: rb.rrrMC(X, beta, it, site_pick="rank", **kw), 2_000_000 // scale),
            ("bklMC", lambda it, **kw: rb.bklMC(X, beta, it, site_pick="rank", **kw), 40_000_000 // scale))
    for irun, (name, fn, iters) in enumerate(runs):
        fn(iters // 20, step=iters // 20, seed=2, C0=C, quiet=True)
        Es, _ = fn(iters, step=iters, seed=3, C0=C, quiet=True)
        info = X.last_run
        g = ffi.Graph.ea_int(A, J); s0 = C.chunks[0].copy()
        cpu_it = iters // 10
        t0 = time.perf_counter()
        _, res = getattr(ffi, name)(g, beta, cpu_it, s0, ffi.PhiloxDraws(3, chain=0), step=cpu_it)
        cdt = time.perf_counter() - t0
        emit(config="RRG", sampler=name, site_pick="rank (k_chain_warp)" if irun >= 3 else "reference", N=N, K=K, replicas=R, beta=beta, iters_per_replica=iters,
             iterations_per_s=R * iters / (info.device_ms * 1e-3), executed_moves_per_s=info.accepted_total / (info.device_ms * 1e-3),
             device_ms=info.device_ms, cpu_oracle_1core_iterations_per_s=cpu_it / cdt, cpu_oracle_1core_moves_per_s=res.accepted / cdt)
    samples = 40 // (4 if quick else 1)
    rb.wtmMC(X, beta, 4, step=50.0 * N, seed=2, C0=C, quiet=True)
    rb.wtmMC(X, beta, samples, step=50.0 * N, seed=3, C0=C, quiet=True)
    info = X.last_run
    g = ffi.Graph.ea_int(A, J); s0 = C.chunks[0].copy()
    t0 = time.perf_counter()
    _, res = ffi.wtmMC(g, beta, max(1, samples // 4), s0, ffi.PhiloxDraws(3, chain=0), step=50.0 * N)
    cdt = time.perf_counter() - t0
    emit(config="RRG", sampler="wtmMC", N=N, K=K, replicas=R, beta=beta, samples=samples, global_time_per_sample=50.0,
         executed_moves_per_s=info.accepted_total / (info.device_ms * 1e-3), device_ms=info.device_ms,
         cpu_oracle_1core_moves_per_s=res.accepted / cdt)

if "eo" in which:
    # extremal_opt (τ-EO, RRRMC.jl:468-521) on the RRG instance family of the paper's benchmark: moves/s, one B200 vs one host core
    from oracle import ffi   # CPU restatement: the baseline leg of this script only
    N, K, R, tau = 10_000, 3, int(os.environ.get("RRG_R", "256")), 1.3
    rng = np.random.default_rng(8370000274 % 2 ** 32)
    A = rb.gen_RRG(N, K, rng)
    J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), A).astype(np.int64)
    X = rb.GraphRRG(N, K, replicas=R, A=A, J=J)
    C0 = rb.Config(N, R, rng=np.random.default_rng(4))
    iters = 200_000 if quick else 2_000_000
    rb.extremal_opt(X, tau, iters // 20, step=iters // 20, seed=2, C0=C0, quiet=True)
    _, Emin, _, itmin = rb.extremal_opt(X, tau, iters, step=iters, seed=3, C0=C0, quiet=True)
    info = X.last_run
    g = ffi.Graph.rrg_int(A, J); s0 = C0.chunks[0].copy()
    cpu_it = iters // 4
    t0 = time.perf_counter()
    _, _, res = ffi.extremal_opt(g, rb.eo_ftau(N, tau), cpu_it, s0, ffi.PhiloxDraws(3, chain=0), step=cpu_it)
    cdt = time.perf_counter() - t0
    emit(config="RRG", sampler="extremal_opt", N=N, K=K, replicas=R, tau=tau, iters_per_replica=iters,
         moves_per_s=R * iters / (info.device_ms * 1e-3), device_ms=info.device_ms, Emin_mean_per_spin=float(np.mean(Emin)) / N,
         cpu_oracle_1core_moves_per_s=cpu_it / cdt)
