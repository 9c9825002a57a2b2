# Runs the reference's own CPU path and the GPU shim on the same synthetic counts and reports the deviations.
#   julia --project=<env with CellScopes> julia/compare_with_reference.jl [n_genes=3000] [n_cells=4000]
# Needs Julia 1.10, CellScopes.jl with its dependencies, and a B200 for the shim.  (Not runnable in the build
# container: no Julia there.  See DESIGN.md section 2.)
import CellScopes as cs
include(joinpath(@__DIR__, "CellScopesCUDA.jl"))
import .CellScopesCUDA as gpu
using SparseArrays, Random, LinearAlgebra, Statistics

G = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 3000
N = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 4000
Random.seed!(1234)
# module-structured Poisson counts (same spirit as cellscopes.jl_b200/synth.py; the exact generator does not matter here)
n_mod = 30
base = exp.(randn(G) .* 1.5 .- 3.0)
modu = rand(1:n_mod, G)
load = 0.5 .+ rand(G)
h = randn(N, n_mod)
lib = exp.(randn(N) .* 0.4)
I = Int64[]; J = Int64[]; V = Int64[]
for c in 1:N, g in 1:G
    lam = base[g] * lib[c] * exp(load[g] * h[c, modu[g]] - load[g]^2 / 2)
    x = lam < 30 ? count(_ -> true, Iterators.takewhile(t -> t < lam, cumsum(-log.(rand(200))))) : round(Int, lam + sqrt(lam) * randn())
    if x > 0
        push!(I, g); push!(J, c); push!(V, x)
    end
end
counts = sparse(I, J, V, G, N)
genes = ["g$i" for i in 1:G]; cells = ["c$i" for i in 1:N]
mk() = cs.scRNAObject(cs.RawCountObject(copy(counts), copy(cells), copy(genes)))

ref = mk(); dev = mk()
ref = cs.normalize_object(ref);            dev = gpu.normalize_object(dev)
relerr(a, b) = maximum(abs.(a .- b) ./ max.(abs.(b), 1e-300))
println("normalize   max rel err  ", relerr(nonzeros(dev.normCount.count_mtx), nonzeros(ref.normCount.count_mtx)))
ref = cs.find_variable_genes(ref);         dev = gpu.find_variable_genes(dev)
println("HVG overlap (of 2000)    ", length(intersect(ref.varGene.var_gene, dev.varGene.var_gene)))
feats = ref.varGene.var_gene               # same rows on both sides from here on
ref = cs.scale_object(ref; features=feats); dev = gpu.scale_object(dev; features=feats)
zr = Matrix(ref.scaleCount.count_mtx); zd = Matrix(dev.scaleCount.count_mtx)
println("scale       max abs err  ", maximum(abs.(zr .- zd)))
ref.varGene.var_gene = feats; dev.varGene.var_gene = feats
ref = cs.run_pca(ref; maxoutdim=30);       dev = gpu.run_pca(dev; maxoutdim=30)
er = ref.dimReduction.pca.cell_embedding;  ed = dev.dimReduction.pca.cell_embedding
qr_r = Matrix(qr(er).Q); qr_d = Matrix(qr(ed).Q)
println("PCA  cos of largest principal angle  ", minimum(svdvals(qr_r' * qr_d)))
for j in 1:5
    s = sign(dot(er[:, j], ed[:, j]))
    println("PCA  component $j  max |diff| / rms  ", maximum(abs.(er[:, j] .- s .* ed[:, j])) / sqrt(mean(er[:, j] .^ 2)))
end
# exact kNN of the shim against the reference's NN-descent on the SAME embedding (recall; NN-descent is approximate)
dev.dimReduction.pca.cell_embedding = er
ind_gpu, _, _ = gpu.knn(dev; n_neighbors=20, metric="cosine")
using NearestNeighborDescent, Distances
graph = nndescent([collect(r) for r in eachrow(er)], 20, CosineDist())
ind_ref, _ = knn_matrices(graph)
recall = mean(length(intersect(ind_ref[:, i], ind_gpu[:, i])) / 20 for i in 1:N)
println("kNN  recall of NN-descent against the exact GPU lists  ", recall)
