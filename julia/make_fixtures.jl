# Pins the oracle to the REAL reference.  Runs the unmodified CellScopes.jl functions of the hot path on the matrix
# that tests/golden/make_ref_inputs.py wrote and dumps every stage's output as raw little-endian binary + a manifest:
#
#   python tests/golden/make_ref_inputs.py                # writes tests/golden/ref_in/ (matrix.mtx, meta.json)
#   julia --project=<env with CellScopes.jl> julia/make_fixtures.jl [tests/golden/ref_in] [tests/golden/ref_out]
#
# tests/test_reference_fixtures.py consumes tests/golden/ref_out/ when it exists (oracle on CPU, CUDA path on the GPU)
# and is skipped with a clear message when it does not.  Nothing here runs in the build container: it has no Julia
# (DESIGN.md section 2); the script only uses the reference's public functions and the standard library.
import CellScopes as cs
using SparseArrays, LinearAlgebra, Statistics, MatrixMarket, Printf

indir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "ref_in")
outdir = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden", "ref_out")
mkpath(outdir)

counts = SparseMatrixCSC{Int64,Int64}(MatrixMarket.mmread(joinpath(indir, "matrix.mtx")))     # genes x cells
G, N = size(counts)
genes = ["g$(i-1)" for i in 1:G]          # 0-based names: the Python side maps them back to row ids
cells = ["c$(i-1)" for i in 1:N]
meta = Dict{String,Any}("n_genes" => G, "n_cells" => N, "julia" => string(VERSION))

function dump(name, a::AbstractArray{Float64})
    open(joinpath(outdir, name * ".f64"), "w") do io; write(io, Array{Float64}(a)); end
    meta[name] = collect(size(a))
end
function dump(name, a::AbstractArray{<:Integer})
    open(joinpath(outdir, name * ".i64"), "w") do io; write(io, Array{Int64}(a)); end
    meta[name] = collect(size(a))
end

obj = cs.scRNAObject(cs.RawCountObject(counts, cells, genes))            # subset_matrix (utils.jl:313-327)
kept = [parse(Int, g[2:end]) for g in obj.rawCount.gene_name]
dump("kept_genes", kept)                                                 # 0-based ids of the genes the constructor kept
dump("kept_cells", [parse(Int, c[2:end]) for c in obj.rawCount.cell_name])

obj = cs.normalize_object(obj)                                           # processing.jl:40-48
nm = obj.normCount.count_mtx
dump("norm_colptr", nm.colptr .- 1); dump("norm_rowval", nm.rowval .- 1); dump("norm_nzval", nm.nzval)

obj = cs.find_variable_genes(obj)                                        # processing.jl:157-162 (nFeatures 2000, span 0.3)
vst = obj.varGene.vst_data
dump("vst_gene", [parse(Int, g[2:end]) for g in vst.gene])               # rank order, 0-based ORIGINAL ids
dump("vst_mean", vst.mean); dump("vst_variance", vst.variance)
dump("vst_variance_expected", vst.variance_expected); dump("vst_variance_standardized", vst.variance_standardized)
dump("var_gene", [parse(Int, g[2:end]) for g in obj.varGene.var_gene])

obj = cs.scale_object(obj; features = obj.varGene.var_gene)              # processing.jl:88-93
dump("scale_gene", [parse(Int, g[2:end]) for g in obj.scaleCount.gene_name])
dump("scaled", Matrix{Float64}(obj.scaleCount.count_mtx))                # H x N column-major == [N][H] row-major

obj = cs.run_pca(obj; maxoutdim = 50)                                    # processing.jl:164-193
emb = obj.dimReduction.pca.cell_embedding
dump("pca_embedding", permutedims(emb))                                  # k x N column-major == [N][k] row-major
# explained variance of the exact SVD, for the principalvars / tvar check (the default branch stores none)
Z = Matrix{Float64}(obj.scaleCount.count_mtx); Zc = Z .- mean(Z, dims = 2)
sv = svdvals(Zc)
dump("principalvars", (sv .^ 2) ./ (N - 1))

# the adjacency the reference hands to Leiden, built by ITS loops from a given neighbour table: the exact cosine kNN of
# the reference's embedding (NN-descent is approximate and unseeded, so the table itself cannot be a fixture)
using Distances
k = 20
En = emb ./ sqrt.(sum(abs2, emb, dims = 2))
D = max.(1 .- En * En', 0.0)
for i in 1:size(D, 1); D[i, i] = Inf; end
ind = Matrix{Int64}(undef, k, size(D, 1))
for i in 1:size(D, 1)
    ord = sortperm(collect(zip(D[i, :], 1:size(D, 2))))[1:k]            # ascending distance, ties by lower index
    ind[:, i] = ord
end
dump("knn_exact_indices", ind .- 1)                                      # k x N column-major == [N][k] row-major
n = size(ind, 2)
rows = Int64[]; cols = Int64[]; vals = Int64[]
for i in 1:n, j in 1:size(ind, 1)                                        # processing.jl:221-231 verbatim semantics
    push!(rows, ind[j, i]); push!(cols, i); push!(vals, 1)
    push!(rows, i); push!(cols, ind[j, i]); push!(vals, 1)
end
adj = sparse(rows, cols, vals, n, n)
dump("adj_colptr", adj.colptr .- 1); dump("adj_rowval", adj.rowval .- 1); dump("adj_nzval", adj.nzval)

open(joinpath(outdir, "manifest.json"), "w") do io
    print(io, "{")
    first = true
    for (k_, v) in meta
        first || print(io, ","); first = false
        print(io, "\"", k_, "\":", v isa AbstractString ? "\"$v\"" : (v isa AbstractVector ? "[" * join(v, ",") * "]" : v))
    end
    print(io, "}")
end
println("fixtures written to ", outdir)
