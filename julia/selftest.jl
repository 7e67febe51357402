# selftest.jl — run where Julia 1.12 + GenomicBreedingCore.jl + a B200 are available:
#     julia --threads=auto julia/selftest.jl
# Compares the reference's own functions with the libgrm_cuda shim on identical simulategenomes inputs, and dumps
# the inputs/outputs in the raw format tests/golden/README.md describes, so they can be committed as
# reference-generated golden vectors (which would turn "parity unpinned" into "pinned").
using GenomicBreedingCore, LinearAlgebra
include(joinpath(@__DIR__, "GRMCuda.jl"))

function dump(path, M::Matrix{Float64})
    open(path, "w") do io
        write(io, Int64(size(M, 1)), Int64(size(M, 2)))
        write(io, M)                       # column-major little-endian Float64
    end
end

relerr(A, B) = maximum(abs.(A .- B)) / maximum(abs.(B))

for (n, l) in ((100, 1_000), (123, 5_000), (100, 10_000))
    genomes = simulategenomes(n = n, l = l, verbose = false)
    A = Float64.(genomes.allele_frequencies)
    dosage = deepcopy(genomes)
    dosage.allele_frequencies = round.(A .* 4) ./ 4          # tetraploid dosages: the int8 path
    for (name, g) in (("cont", genomes), ("tet", dosage))
        ref_s = grmsimple(g).genomic_relationship_matrix
        ref_p = grmploidyaware(g, ploidy = 4).genomic_relationship_matrix
        new_s = GRMCuda.grmsimple(g).genomic_relationship_matrix
        new_p = GRMCuda.grmploidyaware(g, ploidy = 4).genomic_relationship_matrix
        println("n=$n l=$l $name  simple rel.err = ", relerr(new_s, ref_s), "  ploidyaware rel.err = ", relerr(new_p, ref_p),
                "  symmetric = ", issymmetric(new_s) && issymmetric(new_p), "  det>eps = ", det(new_s) > eps(Float64))
        dump("golden_$(name)_$(n)x$(l)_A.bin", Float64.(g.allele_frequencies))
        dump("golden_$(name)_$(n)x$(l)_grmsimple.bin", ref_s)
        dump("golden_$(name)_$(n)x$(l)_grmploidyaware4.bin", ref_p)
    end
end
