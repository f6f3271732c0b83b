# RRRMCB200.jl — the Julia host side of the B200 engine: RRRMC.jl's calling convention on top of the C ABI of
# include/rrrmc_b200.h (librrrmc_b200.so), bound with `ccall`. No CUDA.jl, no CPU fallback.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image and the GPU box have no `julia` binary; the same symbols are
# driven from Python (rrrmc.jl_b200/interface.py) by the tests. This file is kept declarative and thin so that a
# maintainer can check it against the header line by line (INTEGRATION.md walks through it).
#
# Usage (mirrors src/RRRMC.jl:81-88, 149-157, 311-317 and src/Interface.jl:87-270):
#     using RRRMCB200
#     X = RRRMCB200.GraphEA(64, 3; replicas=1024)          # GraphEA{Int,(-1,1),6} on a replica batch
#     Es, C = standardMC(X, 1.0, 10^3 * X.N; step = 10 * X.N)
module RRRMCB200

export standardMC, rrrMC, bklMC, wtmMC, extremal_opt
export transverse_mag, Qenergy, Renergies, overlaps, GraphQT, replay, replay_wtm, tempering_exchange!

const lib = get(ENV, "RRRMC_B200_LIB", joinpath(@__DIR__, "..", "lib", "librrrmc_b200.so"))

const RRRMC_OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = Cint(0), Cint(-1), Cint(-2), Cint(-3), Cint(-4)
const EA_PM1, EA_INT, EA_F64, SK_F64, SK_BIN, QT, QUANT, EMPTY, EA_DISCR = Cint.(1:9)

last_error() = unsafe_string(ccall((:rrrmc_last_error, lib), Cstring, ()))
function check(st::Cint)
    st == RRRMC_OK && return
    msg = last_error()
    st == ERR_ARG && throw(ArgumentError(msg))          # the reference throws ArgumentError (e.g. RRRMC.jl:94,159)
    throw(ErrorException("rrrmc_b200 [$st]: $msg"))
end

# ---- context ------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rrrmc_ctx_create, lib), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, r))
        c = new(r[]); finalizer(c -> ccall((:rrrmc_ctx_destroy, lib), Cint, (Ptr{Cvoid},), c.h), c); c
    end
end
const default_ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (default_ctx[] === nothing && (default_ctx[] = Context()); default_ctx[])

# ---- Config (src/Interface.jl:21-54) for a batch: chunks[:, r] is replica r's BitVector chunk array ----------
struct Config
    N::Int
    chunks::Matrix{UInt64}      # (cld(N,64), R), column r = BitVector(s).chunks of replica r
end
Config(N::Integer, R::Integer = 1) = Config(N, zeros(UInt64, cld(N, 64), R))
Base.BitVector(C::Config, r::Integer = 1) = (b = BitVector(undef, C.N); b.chunks .= view(C.chunks, :, r); b)

# ---- graphs ---------------------------------------------------------------------------------------
abstract type AbstractGraph{ET<:Real} end          # src/Interface.jl:66
mutable struct Graph{ET} <: AbstractGraph{ET}
    h::Ptr{Cvoid}; state::Ptr{Cvoid}; N::Int; replicas::Int; kind::Cint
end
function _finish(h::Ptr{Cvoid}, ET, replicas, kind)
    n = Ref{Int64}(0); check(ccall((:rrrmc_getN, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), h, n))
    s = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_state_create, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), h, replicas, s))
    X = Graph{ET}(h, s[], n[], replicas, kind)
    finalizer(X) do x
        ccall((:rrrmc_state_destroy, lib), Cint, (Ptr{Cvoid},), x.state)
        ccall((:rrrmc_graph_destroy, lib), Cint, (Ptr{Cvoid},), x.h)
    end
    X
end
getN(X::Graph) = X.N                                  # Interface.jl:145

"GraphEA{ET,LEV,twoD}(A, J) (src/graphs/EA.jl:145-168) / GraphEANormal (EA.jl:540-552): A, J as N×2D matrices (row-major copy)."
function GraphEA(L::Integer, D::Integer, A::Matrix{Int64}, J::Matrix; replicas::Integer = 1)
    kind = eltype(J) <: AbstractFloat ? EA_F64 : (all(abs.(J) .== 1) ? EA_PM1 : EA_INT)
    Jc = eltype(J) <: AbstractFloat ? Matrix{Float64}(permutedims(J)) : Matrix{Int64}(permutedims(J))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_graph_ea_create, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                ctx().h, L, D, kind, permutedims(A), Jc, r))
    _finish(r[], kind == EA_F64 ? Float64 : Int, replicas, kind)
end
"gen_AJ(fname) (src/graphs/EA.jl:73-118): `type:`/`size: L`/`name:` header, then `x y Jxy` per bond of the L×L lattice -> (L, D, A, J) as N×2D matrices."
function gen_AJ(fname::AbstractString)
    D = 2
    open(fname) do f
        startswith(strip(readline(f)), "type:") || throw(ArgumentError("$fname: first line must start with type:"))
        ls = split(readline(f)); (length(ls) == 2 && ls[1] == "size:") || throw(ArgumentError("$fname: second line must be `size: L`"))
        L = parse(Int, ls[2])
        startswith(strip(readline(f)), "name:") || throw(ArgumentError("$fname: third line must start with name:"))
        N = L^D; At = zeros(Int64, 2D, N)
        check(ccall((:rrrmc_gen_ea_adjacency, lib), Cint, (Cint, Cint, Ptr{Int64}), L, D, At))
        A = permutedims(At); J = fill(NaN, N, 2D)
        for l in eachline(f)
            ls = split(l); length(ls) == 3 || throw(ArgumentError("$fname: expected `x y Jxy`, got: $l"))
            x, y, Jxy = parse(Int, ls[1]), parse(Int, ls[2]), parse(Float64, ls[3])
            for (a, b) in ((x, y), (y, x))
                k = findfirst(==(b), view(A, a, :))
                k === nothing && throw(ArgumentError("$fname: $a and $b are not neighbours"))
                isnan(J[a, k]) || throw(ArgumentError("$fname: bond $x-$y given twice"))
                J[a, k] = Jxy
            end
        end
        any(isnan, J) && throw(ArgumentError("$fname: bonds missing"))
        L, D, A, J
    end
end
"GraphEANormal(fname) (src/graphs/EA.jl:576-580)"
GraphEANormal(fname::AbstractString; replicas::Integer = 1) = ((L, D, A, J) = gen_AJ(fname); GraphEA(L, D, A, J; replicas = replicas))
"GraphEANormalDiscretized{Int,LEV,twoD} (src/graphs/EA.jl:311-360) from the continuous couplings cJ (N×2D, slot-aligned with A)."
function GraphEANormalDiscretized(L::Integer, D::Integer, LEV::NTuple{K,Int}, A::Matrix{Int64}, cJ::Matrix{Float64}; replicas::Integer = 1) where {K}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    lev = collect(Int64, LEV)
    check(ccall((:rrrmc_graph_ea_discretized_create, lib), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
                ctx().h, L, D, permutedims(A), Matrix{Float64}(permutedims(cJ)), lev, length(lev), r))
    _finish(r[], Float64, replicas, EA_DISCR)
end
"GraphRRG{Int,LEV,K}(A, J) / GraphRRGNormal (src/graphs/RRG.jl:112-137): explicit K-regular adjacency, A and J as N×K matrices."
function GraphRRG(A::Matrix{Int64}, J::Matrix; replicas::Integer = 1)
    N, K = size(A)
    kind = eltype(J) <: AbstractFloat ? EA_F64 : (all(abs.(J) .== 1) ? EA_PM1 : EA_INT)
    Jc = eltype(J) <: AbstractFloat ? Matrix{Float64}(permutedims(J)) : Matrix{Int64}(permutedims(J))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_graph_rrg_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Int64}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                ctx().h, N, K, kind, permutedims(A), Jc, r))
    _finish(r[], kind == EA_F64 ? Float64 : Int, replicas, kind)
end
"GraphRRGNormalDiscretized{Int,LEV,K} (src/graphs/RRG.jl:274-330) from the continuous couplings cJ (N×K, slot-aligned with A)."
function GraphRRGNormalDiscretized(LEV::NTuple{M,Int}, A::Matrix{Int64}, cJ::Matrix{Float64}; replicas::Integer = 1) where {M}
    N, K = size(A)
    r = Ref{Ptr{Cvoid}}(C_NULL); lev = collect(Int64, LEV)
    check(ccall((:rrrmc_graph_rrg_discretized_create, lib), Cint,
                (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
                ctx().h, N, K, permutedims(A), Matrix{Float64}(permutedims(cJ)), lev, length(lev), r))
    _finish(r[], Float64, replicas, EA_DISCR)
end
"GraphQEAT (src/QAliases.jl:51-81): GraphQuant over GraphEANormal{2D}; A, J as N×2D matrices (reference layout)."
function GraphQEAT(L::Integer, D::Integer, M::Integer, Γ::Float64, β::Float64, A::Matrix{Int64}, J::Matrix{Float64}; replicas::Integer = 1)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_graph_quant_ea_create, lib), Cint,
                (Ptr{Cvoid}, Cint, Cint, Int64, Cdouble, Cdouble, Ptr{Int64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx().h, L, D, M, Γ, β, permutedims(A), Matrix{Float64}(permutedims(J)), r))
    _finish(r[], Float64, replicas, QUANT)
end
"GraphEA(L, D) with ±1 couplings drawn here (src/graphs/EA.jl:181-191)."
function GraphEA(L::Integer, D::Integer; replicas::Integer = 1)
    N = L^D
    A = Matrix{Int64}(undef, 2D, N)
    check(ccall((:rrrmc_gen_ea_adjacency, lib), Cint, (Cint, Cint, Ptr{Int64}), L, D, A))
    A = permutedims(A); J = zeros(Int64, N, 2D)
    for x = 1:N, k = 1:2D                                  # gen_J, EA.jl:45-71
        y = A[x, k]; x < y || continue
        J[x, k] = rand((-1, 1)); l = findfirst(==(0), view(J, y, :)); J[y, l] = J[x, k]
    end
    GraphEA(L, D, A, J; replicas = replicas)
end
"GraphSKNormal(N) / GraphSK(N) with explicit couplings (src/graphs/SK.jl:181-199 / :28-49)."
function GraphSK(J::Matrix; replicas::Integer = 1)
    N = size(J, 1); kind = eltype(J) <: AbstractFloat ? SK_F64 : SK_BIN
    Jc = kind == SK_F64 ? Matrix{Float64}(J) : Matrix{UInt8}(J)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_graph_sk_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx().h, N, kind, Jc, r))
    _finish(r[], Float64, replicas, kind)
end
"GraphQuant(Nk, M, Γ, β, inner, J) (src/graphs/QT.jl:163-170): inner = SK_BIN (GraphQSKT), SK_F64 (GraphQSKNormalT), EMPTY (GraphQ0T)."
function GraphQuant(Nk::Integer, M::Integer, Γ::Real, β::Real, inner::Cint, J = nothing; replicas::Integer = 1)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    Jp = J === nothing ? C_NULL : pointer(J)
    GC.@preserve J check(ccall((:rrrmc_graph_quant_create, lib), Cint,
        (Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx().h, Nk, M, Γ, β, inner, Jp, r))
    _finish(r[], Float64, replicas, QUANT)
end

# ---- Interface (src/Interface.jl:87-270) on the batch ------------------------------------------------
upload!(X::Graph, C::Config) = (C.N == X.N || throw(ArgumentError("Invalid C0, wrong N, expected $(X.N), given: $(C.N)"));
    check(ccall((:rrrmc_state_upload, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{UInt64}), X.state, 0, X.replicas, C.chunks)))
function download(X::Graph)
    C = Config(X.N, X.replicas)
    check(ccall((:rrrmc_state_download, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{UInt64}), X.state, 0, X.replicas, C.chunks)); C
end
function energy(X::Graph, C::Config)                    # Interface.jl:105 — also resets the caches
    upload!(X, C); E = zeros(X.replicas)
    check(ccall((:rrrmc_energy, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), X.state, E)); E
end
function delta_energy(X::Graph, C::Config, move::Integer) # Interface.jl:130
    upload!(X, C); dE = zeros(X.replicas)
    check(ccall((:rrrmc_delta_energy, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), X.state, move, dE)); dE
end
function delta_energy_residual(X::Graph, C::Config, move::Integer) # Interface.jl:254-261
    upload!(X, C); dE = zeros(X.replicas)
    check(ccall((:rrrmc_delta_energy_residual, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), X.state, move, dE)); dE
end
function neighbors(X::Graph, i::Integer)                # Interface.jl:158
    m = Ref{Int64}(0); check(ccall((:rrrmc_max_neighbors, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), X.h, m))
    out = zeros(Int64, max(m[], 1)); n = Ref{Cint}(0)
    check(ccall((:rrrmc_neighbors, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ref{Cint}), X.h, i, out, n)); out[1:n[]]
end
function allΔE(X::Graph)                                # Interface.jl:200-201
    out = zeros(64); n = Ref{Cint}(0)
    check(ccall((:rrrmc_allDE, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cint}), X.h, out, n)); Tuple(out[1:n[]])
end
"spinflip!(X, C, move) = flip + update_cache! (Interface.jl:89-92) on every replica of the batch."
function spinflip!(X::Graph, C::Config, move::Integer)
    upload!(X, C)
    check(ccall((:rrrmc_spinflip, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{UInt32}), X.state, move, C_NULL))
    C.chunks .= download(X).chunks; C
end
update_cache!(X::Graph, C::Config, move::Integer) = nothing  # the device caches are rebuilt lazily from the configuration

# ---- samplers (src/RRRMC.jl:81-127, 149-290, 311-359) ---------------------------------------------------
struct Opts
    schedule::Cint; planes_K::Cint; count_accepted::Cint; staged_thr::Cdouble; staged_thr_fact::Cdouble
    planes_M::Cint; cb_method::Cint; site_pick::Cint; reserved::NTuple{5,Cint}   # cb_method: 0 auto, 1 planes, 2 sparse, 3 poisson; site_pick: 0 reference (ArraySet order), 1 rank (rrrmc_b200.h)
end
struct RunInfo
    nsamples::Int64; iters_done::Int64; launches::Int64; device_ms::Cfloat; accepted_total::Int64
end
# hook(it, X, C, accepted, E)::Bool of RRRMC.jl:61-64 arrives through a C callback; `user` points at a mutable box
# that carries the closure (a mutable struct has a stable address under GC.@preserve; a Ref of a tuple would hand the
# callback a RefValue, not the tuple). An exception inside the hook must not unwind through the C frames of the
# library: it is parked in the box, the run is stopped (hook result false) and the exception is rethrown by _run.
mutable struct _HookBox
    hook::Any
    X::Any
    err::Any
end
const _default_hook = (x...) -> true
function _hook_tramp(user::Ptr{Cvoid}, it::Int64, E::Ptr{Cdouble}, acc::Ptr{Int64}, R::Int64)::Cint
    box = unsafe_pointer_to_objref(user)::_HookBox
    try
        ok = box.hook(it, box.X, download(box.X), unsafe_wrap(Array, acc, R), unsafe_wrap(Array, E, R))
        return Cint(ok ? 1 : 0)
    catch e
        box.err = e
        return Cint(0)
    end
end
function _run(sym::Symbol, X::Graph, β, iters::Integer; seed = 167432777111, step::Integer = 1, hook = _default_hook,
              C0::Union{Config,Nothing} = nothing, quiet::Bool = false, staged_thr::Real = NaN, staged_thr_fact::Real = 5.0,
              schedule::Integer = 1)
    isfinite(β) || throw(ArgumentError("β must be finite, given: $β"))                    # RRRMC.jl:159
    C0 === nothing ? check(ccall((:rrrmc_state_randomize, lib), Cint, (Ptr{Cvoid}, UInt64), X.state, seed > 0 ? seed : rand(UInt64))) :
                     upload!(X, C0)
    o = Ref{Opts}(); check(ccall((:rrrmc_opts_default, lib), Cint, (Ref{Opts},), o))
    o[] = Opts(schedule, o[].planes_K, o[].count_accepted, staged_thr, staged_thr_fact, o[].planes_M, o[].cb_method, o[].reserved)
    cap = min(10^8, iters ÷ step)                                                          # RRRMC.jl:90
    Es = zeros(X.replicas, max(cap, 1)); info = Ref{RunInfo}()
    betas = fill(Float64(β), X.replicas)
    # the default hook is not passed at all: no callback, no per-sample download of the whole batch
    box = _HookBox(hook, X, nothing)
    cb = hook === _default_hook ? C_NULL : @cfunction(_hook_tramp, Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Int64}, Int64))
    GC.@preserve box check(ccall((sym, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Opts}, Ptr{Cdouble}, Int64, Ref{RunInfo}),
        X.state, betas, iters, step, seed > 0 ? seed : 0, cb, pointer_from_objref(box), o, Es, cap, info))
    box.err === nothing || throw(box.err)
    quiet || (println("samples = ", info[].nsamples); println("iters = ", info[].iters_done))
    Es[:, 1:info[].nsamples], download(X)
end
"standardMC(X, β, iters; seed, step, hook, C0, quiet) (src/RRRMC.jl:81-127); schedule=1 (default) the reference's rand(1:N) order, 0 checkerboard lattice sweeps (opt-in)."
standardMC(X::Graph, β::Real, iters::Integer; schedule::Integer = (X.kind == EA_PM1 ? 0 : 1), kw...) =
    _run(:rrrmc_standard_mc, X, β, iters; schedule = schedule, kw...)
"rrrMC(X, β, iters; seed, step, hook, C0, staged_thr, staged_thr_fact, quiet) (src/RRRMC.jl:149-290)"
rrrMC(X::Graph, β::Real, iters::Integer; kw...) = _run(:rrrmc_rrr_mc, X, β, iters; kw...)
"bklMC(X, β, iters; seed, step, hook, C0, quiet) (src/RRRMC.jl:311-359)"
bklMC(X::Graph, β::Real, iters::Integer; kw...) = _run(:rrrmc_bkl_mc, X, β, iters; kw...)
"wtmMC(X, β, samples; seed, step::Float64, hook, C0, quiet) (src/RRRMC.jl:376-430); hook(t, X, C, num_moves, E) gets the global time."
function wtmMC(X::Graph, β::Real, samples::Integer; seed = 167432777111, step::Float64 = 1.0, hook = (x...) -> true,
               C0::Union{Config,Nothing} = nothing, quiet::Bool = false)
    C0 === nothing ? check(ccall((:rrrmc_state_randomize, lib), Cint, (Ptr{Cvoid}, UInt64), X.state, seed > 0 ? seed : rand(UInt64))) :
                     upload!(X, C0)
    cap = min(10^8, samples)
    Es = zeros(X.replicas, max(cap, 1)); info = Ref{RunInfo}()
    betas = fill(Float64(β), X.replicas)
    timed = (k, X_, C, acc, E) -> hook(k * step / X.N, X_, C, acc, E)      # sample index -> global time (RRRMC.jl:405)
    box = _HookBox(timed, X, nothing); cb = @cfunction(_hook_tramp, Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Int64}, Int64))
    GC.@preserve box check(ccall((:rrrmc_wtm_mc, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Cdouble, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{RunInfo}),
        X.state, betas, samples, step, seed > 0 ? seed : 0, cb, pointer_from_objref(box), Es, cap, info))
    box.err === nothing || throw(box.err)
    quiet || (println("samples = ", info[].nsamples); println("num_moves = ", info[].iters_done))
    Es[:, 1:info[].nsamples], download(X)
end

"set_betas!(X, betas): a GraphQuant batch as a β ladder — replica r at betas[r] with its own fourK(β) (QT.jl:165); returns the fourK vector."
function set_betas!(X::Graph, betas::Vector{Float64})
    fk = zeros(X.replicas)
    check(ccall((:rrrmc_state_set_quant_betas, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), X.state, betas, fk)); fk
end
# hook(it, X, C, E, Emin)::Bool of RRRMC.jl:499
function _eo_hook_tramp(user::Ptr{Cvoid}, it::Int64, E::Ptr{Cdouble}, Emin::Ptr{Cdouble}, R::Int64)::Cint
    box = unsafe_pointer_to_objref(user)::_HookBox
    try
        return Cint(box.hook(it, box.X, download(box.X), unsafe_wrap(Array, E, R), unsafe_wrap(Array, Emin, R)) ? 1 : 0)
    catch e
        box.err = e
        return Cint(0)
    end
end
"extremal_opt(X, τ, iters; seed, step, hook, C0, quiet) (src/RRRMC.jl:468-521) -> (C, Emin, Cmin, itmin), one entry per replica; DiscrGraph models only."
function extremal_opt(X::Graph, τ::Real, iters::Integer; seed = 167432777111, step::Integer = 1, hook = (x...) -> true,
                      C0::Union{Config,Nothing} = nothing, quiet::Bool = false)
    C0 === nothing ? check(ccall((:rrrmc_state_randomize, lib), Cint, (Ptr{Cvoid}, UInt64), X.state, seed > 0 ? seed : rand(UInt64))) :
                     upload!(X, C0)
    fτ = cumsum([j^(-Float64(τ)) for j = 1:X.N])                           # DeltaE.jl:443: Julia's own `^` and pairwise cumsum
    Emin = zeros(X.replicas); itmin = zeros(Int64, X.replicas); Cmin = Config(X.N, X.replicas); info = Ref{RunInfo}()
    box = _HookBox(hook, X, nothing); cb = @cfunction(_eo_hook_tramp, Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int64))
    GC.@preserve box check(ccall((:rrrmc_extremal_opt, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Int64, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}, Ptr{UInt64},
         Ptr{Cdouble}, Int64, Ref{RunInfo}),
        X.state, fτ, 0, iters, step, seed > 0 ? seed : 0, cb, pointer_from_objref(box), Emin, itmin, Cmin.chunks, C_NULL, 0, info))
    box.err === nothing || throw(box.err)
    quiet || (println("iters = ", info[].iters_done); println("min [it = $itmin] = $Emin"))
    download(X), Emin, Cmin, itmin
end

# ---- GraphQT and the observables of the quantum graphs (QT.jl:46-54, 113-121, 201-268) ------------------------
function GraphQT(N::Integer, M::Integer, fourK::Float64; replicas::Integer = 1)     # GraphQT{fourK}(N, M), QT.jl:46-54
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrrmc_graph_qt_create, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble, Ref{Ptr{Cvoid}}), ctx().h, N, M, fourK, r))
    _finish(r[], Float64, replicas, QT)
end
function transverse_mag(X::Graph, C::Config, β::Float64)                             # QT.jl:113-121
    upload!(X, C); out = zeros(X.replicas)
    check(ccall((:rrrmc_transverse_mag, lib), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), X.state, β, out)); out
end
function Qenergy(X::Graph, C::Config)                                                # QT.jl:253-268
    upload!(X, C); out = zeros(X.replicas)
    check(ccall((:rrrmc_Qenergy, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), X.state, out)); out
end
function Renergies(X::Graph, M::Integer)                                             # QT.jl:201-211 (after energy(X, C))
    out = zeros(M, X.replicas)
    check(ccall((:rrrmc_Renergies, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), X.state, out)); out
end
function overlaps(X::Graph, M::Integer)                                              # QT.jl:213-251
    out = zeros(M ÷ 2, X.replicas)
    check(ccall((:rrrmc_overlaps, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), X.state, out)); out
end

# ---- replay mode (SURVEY Appendix B): one chain fed the typed draw stream a run of RRRMC.jl consumed ---------------
# kind[k] = 0 for a rand(1:n) value (ival[k]), 1 for a rand() value (fval[k]); scripts/dump_julia_trace.jl writes them
function replay(X::Graph, C0::Config, sampler::Symbol, β::Float64, iters::Integer, kind::Vector{UInt8}, ival::Vector{Int64},
                fval::Vector{Float64}; step::Integer = 1, replica::Integer = 0)
    upload!(X, C0)
    o = Ref{Opts}(); check(ccall((:rrrmc_opts_default, lib), Cint, (Ref{Opts},), o))
    Es = zeros(max(iters ÷ step, 1)); info = Ref{RunInfo}()
    code = Dict(:standardMC => 0, :rrrMC => 1, :bklMC => 2)[sampler]
    check(ccall((:rrrmc_replay, lib), Cint,
        (Ptr{Cvoid}, Int64, Cint, Cdouble, Int64, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Cdouble}, Int64, Ref{Opts}, Ptr{Cdouble}, Int64, Ref{RunInfo}),
        X.state, replica, code, β, iters, step, kind, ival, fval, length(kind), o, Es, length(Es), info))
    Es[1:info[].nsamples], download(X)
end
function replay_wtm(X::Graph, C0::Config, β::Float64, samples::Integer, kind::Vector{UInt8}, ival::Vector{Int64}, fval::Vector{Float64};
                    step::Float64 = 1.0, replica::Integer = 0)
    upload!(X, C0)
    Es = zeros(max(samples, 1)); info = Ref{RunInfo}()
    check(ccall((:rrrmc_replay_wtm, lib), Cint,
        (Ptr{Cvoid}, Int64, Cdouble, Int64, Cdouble, Ptr{UInt8}, Ptr{Int64}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Ref{RunInfo}),
        X.state, replica, β, samples, step, kind, ival, fval, length(kind), Es, length(Es), info))
    Es[1:info[].nsamples], download(X)
end

# ---- parallel tempering on the checkerboard schedule (new engine; the reference runs one standardMC per β) -------
# standardMC(X, βs, ...) with opts.schedule = checkerboard takes a β vector that is constant inside each group of 128
# replicas; tempering_exchange! then swaps the configurations of neighbouring groups on the device.
function tempering_exchange!(X::Graph, β_group::Vector{Float64}, seed::Integer, round::Integer; read::Bool = false)
    acc = zeros(Int64, max(length(β_group) - 1, 1))
    check(ccall((:rrrmc_tempering_exchange, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, UInt64, UInt64, Ptr{Int64}),
                X.state, β_group, length(β_group), seed, round, read ? pointer(acc) : C_NULL))
    acc
end

end # module
