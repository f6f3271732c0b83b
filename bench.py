#!/usr/bin/env python
"""bench.py — spin-flip attempts/s, 3D EA L=64 ±J × 1024 replicas (BASELINE.json configs[1]) on N B200s.

A "step" = SWEEPS_PER_STEP checkerboard Metropolis sweeps over the whole replica batch of one GPU
(L^3 * R * SWEEPS_PER_STEP spin-flip attempts). Replicas shard across ranks (weak scaling: every rank owns its own
1024 replicas of the same instance); NCCL is used only for the barrier, the max-over-ranks time and the gather of
the per-replica energies behind `hook`.

  python bench.py [--gpus N --steps K --warmup W]          # our arm
  python bench.py --impl reference [...]                    # CPU arm: the oracle restatement of RRRMC.jl's
                                                            # standardMC on all host cores (Julia cannot run here)
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, D, R_PER_GPU = 64, 3, 1024
N_SITES = L ** D
SWEEPS_PER_STEP = int(os.environ.get("BENCH_SWEEPS", "400"))
BETA = 1.0
PLANES_K = int(os.environ.get("BENCH_K", "5"))
PLANES_M = int(os.environ.get("BENCH_M", "4"))
METHOD = os.environ.get("BENCH_METHOD", "poisson")  # acceptance procedure of the checkerboard kernel: poisson | sparse | planes
SEED = 0x5EEDEA64
METRIC = "spin-flip attempts/s, 3D EA L=64 ±J ×1024 replicas"
UNIT = "attempts/s"
# 2 bits spin RMW per attempt + 3 coupling bits per site per sweep shared by R replicas
ALG_BYTES_PER_SWEEP = 2 * N_SITES * R_PER_GPU // 8 + 3 * N_SITES // 8


def synthetic_instance(reference_arm=False):
    """3 forward bonds/site iid uniform{-1,+1}; one instance shared by all replicas and ranks.
    The reference arm builds it with the oracle's gen_EA/gen_J so that it never maps the product library; both give
    the same instance (same draw order, EA.jl:45-71; tests/test_oracle_basic.py checks the two generators agree)."""
    rng = np.random.default_rng(SEED)
    if reference_arm:
        from oracle import ffi
        A = ffi.gen_EA(L, D)
        nb = int((A > np.arange(1, len(A) + 1)[:, None]).sum())
        return A, ffi.gen_J(A, rng.choice(np.array([-1.0, 1.0]), nb)).astype(np.int64)
    import rrrmc_b200 as rb
    A = rb.gen_EA(L, D)
    J = rb.gen_J(lambda n: rng.choice(np.array([-1.0, 1.0]), n), A)
    return A, J.astype(np.int64)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's standardMC (random-site Metropolis with cached local
# fields, RRRMC.jl:81-127 + EA.jl:195-275), one replica per host thread, each with its own graph copy.
# ----------------------------------------------------------------------------------------------------
def cpu_reference_rate(A, J, iters_per_thread, nthreads):
    from oracle import ffi
    graphs = [ffi.Graph.ea_int(A, J) for _ in range(nthreads)]
    srcs = [ffi.XoshiroDraws(SEED + t) for t in range(nthreads)]
    cfgs = [s.config(N_SITES) for s in srcs]
    for g, c in zip(graphs, cfgs):
        g.energy(c)
    done = [0] * nthreads

    def work(t):
        _, res = ffi.standardMC(graphs[t], BETA, iters_per_thread, cfgs[t], srcs[t], step=iters_per_thread)
        done[t] = res.iters_done
    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    A, J = synthetic_instance(reference_arm=True)   # oracle helpers only: this arm must not load librrrmc_b200.so
    cores = os.cpu_count() or 1
    iters = 20_000_000  # per thread per step (~2 s): a bounded sample of the workload (a full step is 1.07e11 attempts per GPU)
    for _ in range(max(0, args.warmup)):
        cpu_reference_rate(A, J, 200_000, cores)
    rates, t_tot = [], 0.0
    for _ in range(max(1, args.steps)):
        r, dt = cpu_reference_rate(A, J, iters, cores)
        rates.append(r); t_tot += dt
    v = float(np.mean(rates))
    sample = f"{cores} independent replicas (1 per thread), {iters} random-site attempts each per step, L=64 3D ±J, β={BETA}"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u32 bit-sliced / int64 local fields", "data": "synthetic",
           "config": {"workload": "GraphEA 3D L=64 ±J, checkerboard Metropolis, 1024 replicas per GPU (BASELINE configs[1])",
                      "L": L, "D": D, "replicas_per_gpu": R_PER_GPU, "beta": BETA,
                      "reference_sampler": "standardMC (random-site Metropolis, RRRMC.jl:81-127) — the reference has no checkerboard schedule; "
                                           "one replica per host thread, bounded sample per step"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "reference is pure Julia (no julia binary in the image): timed the C restatement oracle/rrrmc_oracle.c"}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import rrrmc_b200 as rb
    from rrrmc_b200 import _ffi
    from rrrmc_b200._ffi import check, lib, ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = rb.Context(local_rank)
    A, J = synthetic_instance()
    X = rb.GraphEA(L, D, replicas=R_PER_GPU, A=A, J=J, ctx=ctx)
    st = X._ensure_state()
    check(lib().rrrmc_state_randomize(st, SEED + rank))
    beta = BETA
    thr = np.array([min(int(np.exp(-beta * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, D + 1)], dtype=np.uint64)

    tbl = np.zeros(_ffi.CBS_T1 + (D - 1) * _ffi.CBS_TC, np.uint32)
    check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), len(tbl)))

    ptbl = np.zeros(_ffi.CBP_LEN, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(ptbl), len(ptbl)))
    NW = lib().rrrmc_checkerboard_poisson_nw(ptr(ptbl), 0.0)   # the library's default choice (what AUTO uses)

    def step(k):
        if METHOD == "poisson":
            check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(ptbl), len(ptbl), NW, SEED + 1000 * rank, k * SWEEPS_PER_STEP, SWEEPS_PER_STEP))
        elif METHOD == "sparse":
            check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), len(tbl), SEED + 1000 * rank, k * SWEEPS_PER_STEP, SWEEPS_PER_STEP))
        else:
            check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, PLANES_K, PLANES_M, SEED + 1000 * rank, k * SWEEPS_PER_STEP, SWEEPS_PER_STEP))

    # ---- device-resident arm ("value"): inputs already in HBM, CUDA events on the launching stream ----------
    for k in range(args.warmup):
        step(k)
    ctx.flush_l2()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ms_steps = []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ctx.timer_start()
        step(args.warmup + k)
        ms_steps.append(ctx.timer_stop())
        ctx.flush_l2()  # evict L2 between timed steps (outside the event pair); the 32 MiB state would otherwise stay resident
        ctx.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - l0  # sweep kernels only (the L2-flush kernel is not counted by the library)
    clk = clocks.stop() if rank == 0 else None
    ms_total = float(sum(ms_steps))
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    attempts_per_step = N_SITES * R_PER_GPU * SWEEPS_PER_STEP
    value = world * attempts_per_step * args.steps / (ms_total * 1e-3)

    # ---- end-to-end arm: the public sampler call with HOST buffers (pinned), H2D + D2H inside the timed region --
    nch = (N_SITES + 63) // 64
    h_in = torch.empty((R_PER_GPU, nch), dtype=torch.int64).pin_memory()
    h_out = torch.empty((R_PER_GPU, nch), dtype=torch.int64).pin_memory()
    h_E = torch.empty((1, R_PER_GPU), dtype=torch.float64).pin_memory()
    rng = np.random.default_rng(SEED + 7 + rank)
    h_in.numpy().view(np.uint64)[...] = rng.integers(0, 2 ** 64, (R_PER_GPU, nch), dtype=np.uint64)
    betas = np.full(R_PER_GPU, beta)
    opts = _ffi.Opts(); check(lib().rrrmc_opts_default(C.byref(opts)))
    opts.planes_K = PLANES_K
    opts.planes_M = PLANES_M
    opts.count_accepted = 0
    opts.schedule = _ffi.SCHED_CHECKERBOARD   # opt-in: the library default is the reference order
    opts.cb_method = {"poisson": _ffi.CB_POISSON, "sparse": _ffi.CB_SPARSE, "planes": _ffi.CB_PLANES}[METHOD]
    info = _ffi.RunInfo()
    iters = SWEEPS_PER_STEP * N_SITES

    # Two pipelines per GPU: a second context (= a second stream), state and set of pinned buffers, each driven by its
    # own host thread through the same three public calls. While one pipeline's sweeps run, the other one's H2D/D2H
    # copies and transposes proceed — every step still uploads its input and downloads its result inside the timed
    # region, but the copies of step k+1 hide behind the sweeps of step k (a sampling service keeps the GPU busy this way).
    NPIPE = 2
    sweep_lock = threading.Lock()
    pipes = [{"ctx": ctx, "X": X, "st": st, "h_in": h_in, "h_out": h_out, "h_E": h_E, "info": info}]
    for _ in range(NPIPE - 1):
        c2 = rb.Context(local_rank)
        X2 = rb.GraphEA(L, D, replicas=R_PER_GPU, A=A, J=J, ctx=c2)
        hi2 = torch.empty((R_PER_GPU, nch), dtype=torch.int64).pin_memory()
        hi2.copy_(h_in)
        pipes.append({"ctx": c2, "X": X2, "st": X2._ensure_state(), "h_in": hi2,
                      "h_out": torch.empty((R_PER_GPU, nch), dtype=torch.int64).pin_memory(),
                      "h_E": torch.empty((1, R_PER_GPU), dtype=torch.float64).pin_memory(), "info": _ffi.RunInfo()})

    def e2e_step(P, k, phases=None):
        t = [time.perf_counter()]

        def mark():
            if phases is not None:
                P["ctx"].sync()
                t.append(time.perf_counter())
        check(lib().rrrmc_state_upload(P["st"], 0, R_PER_GPU, P["h_in"].data_ptr())); mark()
        with sweep_lock:   # one sampler call at a time per GPU: the sweep kernel is a persistent grid that fills the device, so
            # two of them gain nothing from being in flight together; the other pipeline's copies are what overlaps
            check(lib().rrrmc_standard_mc(P["st"], ptr(betas), iters, iters, SEED + 17 * k + 1000 * rank, C.cast(None, _ffi.HOOK), None,
                                          C.byref(opts), P["h_E"].data_ptr(), 1, C.byref(P["info"]))); mark()
        check(lib().rrrmc_state_download(P["st"], 0, R_PER_GPU, P["h_out"].data_ptr())); mark()
        if phases is not None:
            names = ["h2d_upload_transpose", "standard_mc_sweeps_energy_d2h", "download_transpose_d2h"]
            for n, a, b in zip(names, t[:-1], t[1:]):
                phases[n] = phases.get(n, 0.0) + 1e3 * (b - a)
            phases["sweep_kernels_device_ms"] = phases.get("sweep_kernels_device_ms", 0.0) + float(P["info"].device_ms)
        return P["h_E"][0].clone()

    def run_steps(ks):
        """steps ks spread over the pipelines (step k on pipeline k % NPIPE), one host thread per pipeline -> energies"""
        Es = {}
        errs = []

        def work(p):
            try:
                for k in ks[p::NPIPE]:
                    Es[k] = e2e_step(pipes[p], k)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(p,)) for p in range(NPIPE)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errs:
            raise errs[0]
        if world > 1:   # the observable reduction behind `hook`: per-replica energies of every rank, step by step
            hs = []
            for k in ks:
                out = [torch.empty(R_PER_GPU, dtype=torch.float64, device="cuda") for _ in range(world)]
                hs.append((dist.all_gather(out, Es[k].cuda(non_blocking=True), async_op=True), out))
            for w, _ in hs:
                w.wait()
        return Es
    run_steps([0, 1])
    barrier()
    t0 = time.perf_counter()
    Es_all = run_steps([2 + k for k in range(args.steps)])
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    # per-phase breakdown of ONE pipeline's step (two extra steps outside the timed region, a sync after every phase)
    phases = {}
    for k in range(2):
        e2e_step(pipes[0], 100 + k, phases)
    phases = {n: v / 2 for n, v in phases.items()}
    if world > 1:
        keys = sorted(phases)
        t = torch.tensor([phases[n] for n in keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = {n: float(v) for n, v in zip(keys, t.tolist())}
    assert all(float(e.mean()) < -1.0 * N_SITES for e in Es_all.values()), "e2e energies look wrong"
    e2e_value = world * attempts_per_step * args.steps / t_e2e

    if rank == 0:
        peak, peak_src = measured_peak()
        # the dominant kernel's launches inside the timed region (counted by the library): ONE launch of the multi-sweep
        # kernel per step for the poisson procedure on the brick path, two colour launches per sweep otherwise
        launches_per_step = max(1, int(round(launches / max(1, args.steps))))
        launch_ms = ms_total / (args.steps * launches_per_step)
        alg_bytes_per_launch = ALG_BYTES_PER_SWEEP * SWEEPS_PER_STEP / launches_per_step
        achieved = alg_bytes_per_launch / (launch_ms * 1e-3) / 1e9
        flow = launches_per_step == 1
        traffic = None
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this configuration, from the ncu
        # capture scripts/round_check.sh takes of this very command (profiles/r2_checkerboard_traffic.json names the
        # capture and the library it was taken with); null when the capture is of another kernel / launch shape
        tp = os.path.join(ROOT, "profiles", "r2_checkerboard_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("sweeps_per_launch") == (SWEEPS_PER_STEP if flow else 0.5) and tj.get("method") == METHOD and abs(tj.get("beta", -1) - beta) < 1e-9:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cores = os.cpu_count() or 1
        # the same sample as one step of the reference arm (bench.py --impl reference), after the same warm-up call, so
        # that the two CPU figures of a box agree; five of them, ~10 s
        cpu_iters = 20_000_000
        if args.no_cpu_baseline:
            cpu_iters = 1_000_000
        cpu_reference_rate(A, J, 200_000, cores)
        reps = 1 if args.no_cpu_baseline else 5
        cpu_runs = [cpu_reference_rate(A, J, cpu_iters, cores) for _ in range(reps)]
        cpu_v, cpu_dt = float(np.mean([r for r, _ in cpu_runs])), float(sum(d for _, d in cpu_runs))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 bit-sliced (multispin, 32 replicas/word)", "data": "synthetic",
            "config": {"workload": "GraphEA 3D L=64 ±J, checkerboard Metropolis, 1024 replicas per GPU (BASELINE configs[1])",
                       "L": L, "D": D, "replicas_per_gpu": R_PER_GPU, "beta": beta, "sweeps_per_step": SWEEPS_PER_STEP,
                       "rng": {"poisson": "Philox4x32-10; per task (128 replicas of a site) and hit level a Poisson count (inverse CDF, 32-bit tables) "
                                          "+ uniform positions with replacement: exact per-(site,replica) Bernoulli(exp(-βΔE)); NW=%d static position words" % NW,
                               "sparse": "Philox4x32-10; per task and ΔE class a binomial count of passing lanes (inverse CDF, 32-bit tables) + "
                                         "uniform distinct positions: exact per-(site,replica) Bernoulli(exp(-βΔE))",
                               "planes": "Philox4x32-10, exact per-(site,replica) Bernoulli via %d full + %d merged bit planes + 32-bit tail" % (PLANES_K, PLANES_M)}[METHOD],
                       "acceptance_procedure": METHOD,
                       "parallelism": f"replica-sharded x{world}", "l2": "flushed between timed steps (256 MiB write)",
                       "accepted_counters": "off in the timed loop"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": {"poisson": ("k_checkerboard_flow<%d,2,false,false>" if flow else "k_checkerboard_tma<%d,2>") % NW, "sparse": "k_checkerboard_sparse<3,true>",
                                    "planes": "k_checkerboard<3,true>"}[METHOD],
                         "algorithmic_bytes_per_launch": int(alg_bytes_per_launch), "launch_ms": launch_ms,
                         "launches_per_step": launches_per_step, "sweeps_per_launch": SWEEPS_PER_STEP / launches_per_step,
                         "note": "bound by integer instruction issue (ALU pipe, ncu), not by HBM: the 32 MiB state is L2 resident (see DESIGN.md §5)"},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{reps} x ({cores} replicas (1/thread) x {cpu_iters} random-site attempts), same instance, beta={beta}; {cpu_dt:.1f}s"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(R_PER_GPU * nch * 8),
                    "d2h_bytes_per_step": int(R_PER_GPU * nch * 8 + R_PER_GPU * 8),
                    "api": "rrrmc_state_upload + rrrmc_standard_mc + rrrmc_state_download (host pinned buffers); two pipelines per GPU "
                           "(two contexts/states, one host thread each) so that a step's copies overlap the other pipeline's sweeps",
                    "pipelines_per_gpu": NPIPE,
                    "phases_ms_max_over_ranks": phases},
            "gpu_launches": int(launches),
            "clocks": clk,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--beta", type=float, default=BETA, help="inverse temperature (BASELINE configs[1] sweeps 0.5-2.0; headline: 1.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (β-sweep runs)")
    args = ap.parse_args()
    globals()["BETA"] = args.beta
    if "BENCH_METHOD" not in os.environ and args.beta < 0.54:
        globals()["METHOD"] = "planes"   # what RRRMC_CB_AUTO selects for warm runs: the count procedures lose to bit planes below β ≈ 0.54
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
