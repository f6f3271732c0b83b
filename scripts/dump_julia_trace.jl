# Dumps a replay trace from REAL RRRMC.jl (SURVEY.md Appendix B) — run this OFF-BOX, where Julia and the reference are:
#
#   julia --project=/path/to/RRRMC.jl scripts/dump_julia_trace.jl standardMC EA 4 3 1.3 2000 100 7 out_dir
#   python scripts/julia_trace_to_npz.py out_dir tests/golden/julia_trace_standardMC_EA_4_3.npz
#
# args: sampler (standardMC|rrrMC|bklMC), graph (EA = GraphEA ±J | EANormal), L, D, β, iters, step, seed, output directory.
# The image this repository is built in has no Julia, so THIS SCRIPT HAS NEVER BEEN RUN; it is committed so that anyone
# with Julia can pin the oracle against the real package: tests/test_oracle_pins.py::test_julia_trace_replays replays
# every tests/golden/julia_trace_*.npz through the oracle and (on a GPU) tests/test_gpu_chain.py through rrrmc_replay.
#
# How: the reference samplers draw from the default RNG (`rand()`, `rand(1:n)`: RRRMC.jl:39,43,113,192,202,
# DeltaE.jl:143,148,323, ArraySets.jl:83, DynamicSamplers.jl:154). A logging RNG that forwards to a seeded inner
# generator and records every TYPED draw (kind 1 = Float64 in [0,1), kind 0 = integer in 1:n) is installed as the
# default RNG for the duration of the run; the reference code itself is not modified. `seed=0` is passed to the sampler
# so that it does not reseed (RRRMC.jl:89).
using Random
using RRRMC

struct LogRNG <: Random.AbstractRNG
    inner::Random.AbstractRNG
    kind::Vector{UInt8}
    ival::Vector{Int64}
    fval::Vector{Float64}
end
LogRNG(inner) = LogRNG(inner, UInt8[], Int64[], Float64[])

function Random.rand(r::LogRNG, ::Type{Float64})
    x = rand(r.inner, Float64)
    push!(r.kind, 0x01); push!(r.ival, 0); push!(r.fval, x)
    return x
end
Random.rand(r::LogRNG) = rand(r, Float64)
function Random.rand(r::LogRNG, rg::UnitRange{Int})
    first(rg) == 1 || error("the hot path only draws rand(1:n); got $rg")
    k = rand(r.inner, rg)
    push!(r.kind, 0x00); push!(r.ival, k); push!(r.fval, 0.0)
    return k
end
# anything else would be a draw the trace format does not know: fail loudly instead of logging garbage
Random.rand(r::LogRNG, sp::Random.Sampler) = error("unexpected draw through sampler $(typeof(sp))")

function main(args)
    length(args) == 9 || error("usage: dump_julia_trace.jl sampler graph L D beta iters step seed outdir")
    sampler, gname = args[1], args[2]
    L, D = parse(Int, args[3]), parse(Int, args[4])
    β, iters, step, seed = parse(Float64, args[5]), parse(Int, args[6]), parse(Int, args[7]), parse(Int, args[8])
    outdir = args[9]
    mkpath(outdir)

    Random.seed!(seed)
    X = gname == "EA" ? RRRMC.GraphEA(L, D) : gname == "EANormal" ? RRRMC.GraphEANormal(L, D) : error("graph must be EA or EANormal")
    N = RRRMC.getN(X)
    C0 = RRRMC.Config(N)
    chunks0 = copy(C0.s.chunks)

    logger = LogRNG(Random.Xoshiro(seed + 1))
    # install the logger as the default RNG (method overwrite: every `rand()` of the package now lands in `logger`)
    @eval Random.default_rng() = $logger
    @eval Random.default_rng(::Int) = $logger
    f = sampler == "standardMC" ? RRRMC.standardMC : sampler == "rrrMC" ? RRRMC.rrrMC : sampler == "bklMC" ? RRRMC.bklMC : error("unknown sampler")
    Es, C1 = Base.invokelatest(f, X, β, iters; seed = 0, step = step, C0 = C0, quiet = true)

    open(joinpath(outdir, "header.txt"), "w") do io
        println(io, "sampler ", sampler); println(io, "graph ", gname == "EA" ? "GraphEA" : "GraphEANormal")
        println(io, "L ", L); println(io, "D ", D); println(io, "N ", N); println(io, "beta ", repr(β))
        println(io, "iters ", iters); println(io, "step ", step); println(io, "seed ", seed); println(io, "julia ", VERSION)
    end
    writearr(name, v) = open(io -> foreach(x -> println(io, repr(x)), v), joinpath(outdir, name), "w")
    writearr("A.txt", vcat([collect(a) for a in X.A]...))            # N rows of 2D neighbours, row-major
    writearr("J.txt", vcat([collect(j) for j in X.J]...))
    writearr("C0.txt", chunks0); writearr("C1.txt", C1.s.chunks)
    writearr("kind.txt", logger.kind); writearr("ival.txt", logger.ival); writearr("fval.txt", logger.fval)
    writearr("Es.txt", Es)
    println("wrote ", length(logger.kind), " draws, ", length(Es), " samples to ", outdir)
end

main(ARGS)
