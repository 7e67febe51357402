# GRMCuda.jl — Julia `ccall` shim over libgrm_cuda (include/grm_cuda.h).
#
# Gives the B200 backend the reference's exact signatures (GenomicBreedingCore.jl src/grm/calc.jl):
#     grmsimple(genomes::Genomes; max_iter::Int64 = 1_000, verbose::Bool = false)::GRM                        # :115
#     grmploidyaware(genomes::Genomes; ploidy::Int64 = 2, max_iter::Int64 = 1_000, verbose::Bool = false)::GRM # :189
#     inflatediagonals!(X::Matrix{Float64}; max_iter::Int64 = 1_000, verbose::Bool = false)::Nothing          # :41
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image or on the GPU boxes.  The Python
# ctypes binding (genomicbreedingcore.jl_b200/_lib.py) makes the same calls with the same argument order and is what
# the tests drive.  julia/selftest.jl runs the reference and this shim side by side where Julia is available.
module GRMCuda

using GenomicBreedingCore: Genomes, GRM, checkdims
import LinearAlgebra
import StatsBase

const LIB = get(ENV, "GRM_CUDA_LIB", joinpath(@__DIR__, "..", "genomicbreedingcore.jl_b200", "libgrm_cuda.so"))

# mirrors of grm_cuda_options / grm_cuda_info (include/grm_cuda.h, ABI 0.2.0: 80 and 120 bytes)
Base.@kwdef mutable struct Options
    struct_size::Int32 = 80
    path::Int32 = 0
    denominator::Int32 = 0
    max_iter::Int32 = 1000
    skip_repair::Int32 = 0
    input_location::Int32 = 0
    output_location::Int32 = 0
    rank::Int32 = 0
    world::Int32 = 1
    missing_policy::Int32 = 0
    chunk_loci::Int64 = 0
    maf::Float64 = 0.0
    max_locus_sparsity::Float64 = 1.0
    distributed::Int32 = 0
    input_slab::Int32 = 0
    output_mode::Int32 = 0
    reserved0::Int32 = 0
end

Base.@kwdef mutable struct Info
    struct_size::Int32 = 120
    path::Int32 = 0
    denominator::Int32 = 0
    repair_iters::Int32 = 0
    eps_total::Float64 = 0.0
    scale::Float64 = 0.0
    tiles::Int64 = 0
    ms_h2d::Float32 = 0
    ms_pack::Float32 = 0
    ms_rowdot::Float32 = 0
    ms_syrk::Float32 = 0
    ms_finalize::Float32 = 0
    ms_repair::Float32 = 0
    ms_d2h::Float32 = 0
    ms_total::Float32 = 0
    kernel_launches::Int32 = 0
    loci_kept::Int32 = 0
    ms_exchange::Float32 = 0
    ms_gather::Float32 = 0
    ms_host_busy::Float32 = 0
    ingest::Int32 = 0
    host_threads::Int32 = 0
    world::Int32 = 1
    exchange_waves::Int32 = 0
    syrk_window::Int32 = 0
    repair_lu::Int32 = 0
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function context()::Ptr{Cvoid}
    if CTX[] == C_NULL
        dev = parse(Int32, get(ENV, "GRM_CUDA_DEVICE", "0"))
        st = ccall((:grm_cuda_ctx_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), dev, CTX)
        st == 0 || error("libgrm_cuda: " * unsafe_string(ccall((:grm_cuda_last_error, LIB), Cstring, ())))
    end
    CTX[]
end

lasterror() = unsafe_string(ccall((:grm_cuda_last_error, LIB), Cstring, ()))

# status -> the exception type the reference raises at the corresponding place
function throwstatus(st::Int32)
    msg = lasterror()
    if st == 2
        # the reference fails with a MethodError(convert, (Float64, missing)) inside a TaskFailedException (calc.jl:127)
        throw(MethodError(convert, (Float64, missing)))
    elseif st in (1, 3, 4, 5, 6, 9, 10)
        throw(ArgumentError(msg))                                                # 4: "The input matrix is not symmetric." (calc.jl:44)
    elseif st == 11
        throw(ErrorException(msg))                                               # filter.jl:363-365
    elseif st == 13
        throw(ErrorException("libgrm_cuda communicator: $msg"))
    elseif st == 12
        throw(LinearAlgebra.PosDefException(1))                                  # what cholesky / MvNormal throw
    else
        throw(ErrorException("libgrm_cuda status $st: $msg"))
    end
end

# ---- genomes.allele_frequencies as Julia stores it ----------------------------------------------------------------
# `Matrix{Union{Float64,Missing}}` (src/all_structs.jl:66) is an isbits-Union array: n*p Float64 payloads followed, in the
# same allocation, by n*p selector ("type tag") bytes.  The library reads both in place (grm_cuda_*_union): no
# conversion pass, no copy, no pinning on this side.  Returns (payload pointer, selector pointer or C_NULL, selector
# value of a Float64 cell).
function unionptrs(A::Matrix{Union{Float64,Missing}})
    T = eltype(A)
    tag = Int32(findfirst(==(Float64), Base.uniontypes(T)) - 1)      # component index of Float64 in the union
    payload = Ptr{Float64}(pointer(A))
    sel = @static if VERSION >= v"1.11"
        mem = A.ref.mem                                                # Memory{T}: data, then the selector bytes
        off = (UInt(pointer(A)) - UInt(pointer(mem))) ÷ sizeof(Float64)
        ccall(:jl_genericmemory_typetagdata, Ptr{UInt8}, (Any,), mem) + off
    else
        ccall(:jl_array_typetagdata, Ptr{UInt8}, (Any,), A)
    end
    (payload, sel, tag)
end
unionptrs(A::Matrix{Float64}) = (pointer(A), Ptr{UInt8}(C_NULL), Int32(0))   # NaN then stands for `missing`

# anything else (views, other element types) is materialised once; the two concrete layouts above are zero-copy
hostmatrix(A::Matrix{Union{Float64,Missing}}) = A
hostmatrix(A::Matrix{Float64}) = A
hostmatrix(A::AbstractMatrix) = convert(Matrix{Union{Float64,Missing}}, A)

# dense Float64 copy (missing -> NaN) for the entry points that take plain double* (distances)
dense(A::Matrix{Float64}) = A
dense(A::AbstractMatrix) = Float64[ismissing(x) ? NaN : Float64(x) for x in A]

# ---- multi-GPU: one Julia process per GPU (Distributed.jl / MPI.jl), the exchange lives inside the library -----------
# On ONE process:  id = GRMCuda.uniqueid();  hand the 128 bytes to the others (remotecall, MPI.Bcast!, a file);  on
# EVERY process:  GRMCuda.initcomm!(id, rank, world)  (0-based rank).  Afterwards grmsimple / grmploidyaware with
# `distributed = true` are collective calls: each process passes the same Genomes (or, with `slab = true`, only its own
# loci range  L*rank+1 : min(p, L*(rank+1)),  L = cld(p, world))  and every process receives the complete GRM.
function uniqueid()::Vector{UInt8}
    id = Vector{UInt8}(undef, 128)
    st = ccall((:grm_cuda_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
    st == 0 || throwstatus(st)
    id
end

function initcomm!(id::Vector{UInt8}, rank::Integer, world::Integer)
    length(id) == 128 || throw(ArgumentError("the communicator id is 128 bytes"))
    st = ccall((:grm_cuda_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), context(), id, Int32(rank), Int32(world))
    st == 0 || throwstatus(st)
    nothing
end

function rungrm(genomes::Genomes, centred::Bool, ploidy::Int64, max_iter::Int64, verbose::Bool;
                distributed::Bool = false, slab::Bool = false, total_loci::Int64 = 0)::GRM
    if !checkdims(genomes)
        throw(ArgumentError("The Genomes struct is corrupted ☹."))          # calc.jl:118-120, :191-193
    end
    A = hostmatrix(genomes.allele_frequencies)
    n, p = size(A)
    ptotal = (distributed && slab) ? total_loci : p
    grm = GRM(genomes.entries, genomes.loci_alleles, Matrix{Float64}(undef, n, n))   # calc.jl:123, :196
    G = grm.genomic_relationship_matrix
    opts = Options(max_iter = Int32(max_iter), distributed = Int32(distributed), input_slab = Int32(distributed && slab))
    info = Info()
    payload, sel, tag = unionptrs(A)
    st = GC.@preserve A G begin
        if centred
            ccall((:grm_cuda_ploidyaware_union, LIB), Int32,
                  (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Int32, Int64, Int64, Int64, Int32, Ptr{Float64}, Int64, Ref{Options}, Ref{Info}),
                  context(), payload, sel, tag, n, ptotal, n, Int32(ploidy), G, n, opts, info)
        else
            ccall((:grm_cuda_simple_union, LIB), Int32,
                  (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Int32, Int64, Int64, Int64, Ptr{Float64}, Int64, Ref{Options}, Ref{Info}),
                  context(), payload, sel, tag, n, ptotal, n, G, n, opts, info)
        end
    end
    st == 0 || throwstatus(st)
    if verbose                                                                    # calc.jl:67-69
        println("Performed $(info.repair_iters) iterations of diagonal inflation to ensure matrix invertibility (ϵ_total = $(info.eps_total)).")
    end
    if !centred && !checkdims(grm)
        throw(ErrorException("Error computing a simple GRM."))                   # calc.jl:135-137
    end
    grm
end

# the reference's signatures; the keyword-only extras (`distributed`, `slab`, `total_loci`) select the multi-GPU call
grmsimple(genomes::Genomes; max_iter::Int64 = 1_000, verbose::Bool = false, kw...)::GRM =
    rungrm(genomes, false, 0, max_iter, verbose; kw...)

grmploidyaware(genomes::Genomes; ploidy::Int64 = 2, max_iter::Int64 = 1_000, verbose::Bool = false, kw...)::GRM =
    rungrm(genomes, true, ploidy, max_iter, verbose; kw...)

function inflatediagonals!(X::Matrix{Float64}; max_iter::Int64 = 1_000, verbose::Bool = false)::Nothing
    if size(X, 1) != size(X, 2)
        # calc.jl:43 runs issymmetric before the squareness check, so this is the message a caller sees
        throw(ArgumentError("The input matrix is not symmetric."))
    end
    n = size(X, 1)
    iters = Ref{Int32}(0)
    total = Ref{Float64}(0.0)
    st = GC.@preserve X ccall((:grm_cuda_inflate_diagonals, LIB), Int32,
                              (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, Int32, Ref{Int32}, Ref{Float64}),
                              context(), X, n, n, Int32(max_iter), Int32(0), iters, total)
    st == 0 || throwstatus(st)
    if verbose
        println("Performed $(iters[]) iterations of diagonal inflation to ensure matrix invertibility (ϵ_total = $(total[])).")
    end
    nothing
end

# ---- consumer side (SURVEY.md §8 f3): MvNormal(μ, Σ) factorises Σ (simulate_effects.jl:479-481) -------------------
# cholesky(Σ).U computed on the GPU; `lastfactor(n)` hands out the factor the repair of the last GRM call ended with, so
# `MvNormal(μ, PDMat(Σ, Cholesky(lastfactor(n), 'U', 0)))` skips the second O(n³) factorisation.
function cholesky_upper(X::Matrix{Float64})::Matrix{Float64}
    n = size(X, 1)
    size(X, 2) == n || throw(DimensionMismatch("matrix is not square"))
    U = Matrix{Float64}(undef, n, n)
    st = GC.@preserve X U ccall((:grm_cuda_cholesky, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, Ptr{Float64}, Int64, Int32), context(), X, n, n, 0, U, n, 0)
    st == 0 || throwstatus(st)
    U
end

function lastfactor(n::Integer)::Matrix{Float64}
    U = Matrix{Float64}(undef, n, n)
    st = GC.@preserve U ccall((:grm_cuda_last_factor, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Int32),
        context(), n, U, n, 0)
    st == 0 || throwstatus(st)
    U
end

# ---- distances(genomes) (SURVEY.md §8 f4; src/genomes/genomes.jl:454-654) -------------------------------------------
# Same keyword arguments and return value as the reference; the pair loops run in grm_cuda_distances.
function distances(genomes::Genomes;
        distance_metrics::Vector{String} = ["euclidean", "correlation", "mad", "rmsd", "χ²"],
        idx_loci_alleles::Union{Nothing,Vector{Int64}} = nothing, include_loci_alleles::Bool = true,
        include_entries::Bool = true, include_counts::Bool = true,
        verbose::Bool = false)::Tuple{Vector{String},Vector{String},Dict{String,Matrix{Float64}}}
    checkdims(genomes) || throw(ArgumentError("The genomes struct is corrupted ☹."))
    recognised = ["euclidean", "correlation", "mad", "rmsd", "χ²"]
    metrics = unique(distance_metrics)
    length(metrics) >= 1 || throw(ArgumentError("Please supply at least 1 distance metric. Choose from:\n\t‣ " * join(recognised, "\n\t‣ ")))
    bad = filter(m -> !(m in recognised), metrics)
    isempty(bad) || throw(ArgumentError("Unrecognised metric/s:\n\t‣ " * join(bad, "\n\t‣ ") * "\nPlease choose from:\n\t‣ " * join(recognised, "\n\t‣ ")))
    A = dense(genomes.allele_frequencies)
    n, p = size(A)
    idx = if isnothing(idx_loci_alleles)
        sort(StatsBase.sample(1:p, 100, replace = false))
    else
        (minimum(idx_loci_alleles) < 1 || maximum(idx_loci_alleles) > p) &&
            throw(ArgumentError("We accept `idx_loci_alleles` from 1 to `n_loci_alleles` of `genomes`."))
        unique(sort(idx_loci_alleles))
    end
    t = length(idx)
    idx0 = Int64.(idx .- 1)
    slot = Dict("euclidean" => 1, "correlation" => 2, "mad" => 3, "rmsd" => 4, "χ²" => 5)
    out = Dict{String,Matrix{Float64}}()
    for (name, dim, r, on) in (("loci_alleles", 0, t, include_loci_alleles && t > 1), ("entries", 1, n, include_entries && n > 1))
        on || continue
        mats = Vector{Union{Nothing,Matrix{Float64}}}(nothing, 6)
        for m in metrics
            mats[slot[m]] = Matrix{Float64}(undef, r, r)
        end
        include_counts && (mats[6] = Matrix{Float64}(undef, r, r))
        ptr(x) = isnothing(x) ? Ptr{Float64}(C_NULL) : pointer(x)
        st = GC.@preserve A idx0 mats ccall((:grm_cuda_distances, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int32, Ptr{Int64}, Int64, Int32, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
            context(), A, n, p, n, 0, idx0, t, dim, ptr(mats[1]), ptr(mats[2]), ptr(mats[3]), ptr(mats[4]), ptr(mats[5]),
            ptr(mats[6]), r, 0)
        st == 0 || throwstatus(st)
        for m in metrics
            out[string(name, "|", m)] = mats[slot[m]]
        end
        include_counts && (out[string(name, "|counts")] = mats[6])
    end
    isempty(out) && throw(ErrorException("Genomes struct is too sparse. No distance matrix was calculated."))
    (genomes.loci_alleles[idx], genomes.entries, out)
end

end # module
