# CellScopesCUDA.jl -- Julia binding of libcellscopes_cuda.so (include/cellscopes_cuda.h).
#
# The reference (HaojiaWu/CellScopes.jl) has no FFI: its boundary for the preprocessing-to-graph path is
# multiple dispatch on
#     normalize_object      src/scrna/processing.jl:15,33,40
#     find_variable_genes   src/scrna/processing.jl:133,157
#     scale_object          src/scrna/processing.jl:65,78,88
#     run_pca               src/scrna/processing.jl:164
#     run_clustering        src/scrna/processing.jl:309 (graph part :195-232, :263-278)
# This module provides the same five functions with the same keywords, defaults, mutate-and-return behaviour
# and output fields, computed on a B200 through `ccall`.  Two ways to use it:
#
#     import CellScopes as cs
#     include("CellScopesCUDA.jl"); import .CellScopesCUDA as gpu
#     obj = gpu.normalize_object(obj); obj = gpu.find_variable_genes(obj); ...      # explicit
#     gpu.enable!()           # or: re-point cs.normalize_object & co. at the GPU path for the object methods
#
# There is no CPU fallback in here: if the library or a B200 is missing every call throws.
# NOTE: Julia is not available in the build container, so this file is reviewed but not executed by the
# repo's tests; every entry point it binds is exercised through the identical C ABI from Python
# (cellscopes.jl_b200/_capi.py).  julia/compare_with_reference.jl checks it against the reference's own
# functions on a machine that has Julia.
module CellScopesCUDA

using Libdl
using SparseArrays
using DataFrames
using Random
import CellScopes
const cs = CellScopes

# ------------------------------------------------------------------------------------------------
# library loading and error plumbing
# ------------------------------------------------------------------------------------------------
const _lib = Ref{Ptr{Cvoid}}(C_NULL)
const _sym = Dict{Symbol,Ptr{Cvoid}}()

function libpath()
    get(ENV, "CELLSCOPES_CUDA_LIB",
        joinpath(@__DIR__, "..", "cellscopes.jl_b200", "lib", "libcellscopes_cuda.so"))
end

function sym(name::Symbol)
    if _lib[] == C_NULL
        _lib[] = Libdl.dlopen(libpath())       # throws if the library is missing: no CPU path
        v = ccall(Libdl.dlsym(_lib[], :csc_abi_version), Int32, ())
        v == 2 || error("libcellscopes_cuda ABI version $v, this shim binds version 2")
    end
    get!(_sym, name) do
        Libdl.dlsym(_lib[], name)
    end
end

const DT_F32, DT_F64, DT_I32, DT_I64 = Int32(0), Int32(1), Int32(2), Int32(3)
dtcode(::Type{Float32}) = DT_F32
dtcode(::Type{Float64}) = DT_F64
dtcode(::Type{Int32}) = DT_I32
dtcode(::Type{Int64}) = DT_I64

mutable struct Context
    h::Ptr{Cvoid}
    device::Int
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall(sym(:csc_create), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, r)
        if rc != 0
            msg = unsafe_string(ccall(sym(:csc_last_error), Cstring, (Ptr{Cvoid},), C_NULL))
            error("libcellscopes_cuda: $msg")            # like the reference's error(...) calls
        end
        ctx = new(r[], device)
        finalizer(c -> (c.h != C_NULL && ccall(sym(:csc_destroy), Int32, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
        ctx
    end
end

function check(ctx::Context, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall(sym(:csc_last_error), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == -1 ? throw(ArgumentError(msg)) : error("libcellscopes_cuda error $rc: $msg")
end

# opaque device handle, freed by the matching csc_*_free
mutable struct Handle
    h::Ptr{Cvoid}
    free::Symbol
    ctx::Context                      # keeps the context alive
    function Handle(h, free, ctx)
        x = new(h, free, ctx)
        finalizer(v -> (v.h != C_NULL && ccall(sym(v.free), Int32, (Ptr{Cvoid},), v.h); v.h = C_NULL), x)
        x
    end
end

const _default_ctx = Ref{Union{Context,Nothing}}(nothing)
default_context() = (_default_ctx[] === nothing && (_default_ctx[] = Context(0)); _default_ctx[])

# ------------------------------------------------------------------------------------------------
# thin wrappers, one per C entry point
# ------------------------------------------------------------------------------------------------
function vec_alloc(ctx, n::Integer)                       # zero-filled fp64
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(sym(:csc_vec_alloc), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Ptr{Cvoid}}), ctx.h, n, DT_F64, r))
    Handle(r[], :csc_vec_free, ctx)
end

function vec_download(ctx, v::Handle, ::Type{T}, n::Integer) where {T}
    out = Vector{T}(undef, n)
    GC.@preserve out check(ctx, ccall(sym(:csc_vec_download), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), v.h, pointer(out), n))
    out
end

"rawCount.count_mtx (genes x cells, SparseMatrixCSC{<:Real,Int64}) -> device CSC shard; 1-based indices go over as they are"
function upload(ctx, m::SparseMatrixCSC{Tv,Int64}) where {Tv<:Union{Int64,Float64,Float32,Int32}}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    cp, rv, nz = m.colptr, m.rowval, m.nzval
    GC.@preserve cp rv nz check(ctx, ccall(sym(:csc_sparse_upload), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}),
        ctx.h, size(m, 1), size(m, 2), pointer(cp), pointer(rv), DT_I64, pointer(nz), dtcode(Tv), Int32(1), r))
    Handle(r[], :csc_sparse_free, ctx)
end
# other element types and matrix kinds (Adjoint from read_merfish, dense loom matrices: SURVEY A.1) are materialised first
upload(ctx, m::SparseMatrixCSC) = upload(ctx, SparseMatrixCSC{Float64,Int64}(m))
upload(ctx, m::AbstractMatrix{<:Real}) = upload(ctx, SparseMatrixCSC{Float64,Int64}(sparse(m)))

"values of a device matrix that shares the pattern of `pattern` -> SparseMatrixCSC{Float64,Int64}"
function download_like(ctx, h::Handle, pattern::SparseMatrixCSC)
    nz = Vector{Float64}(undef, nnz(pattern))
    GC.@preserve nz check(ctx, ccall(sym(:csc_sparse_download), Int32,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int32),
        h.h, 0, size(pattern, 2), C_NULL, C_NULL, pointer(nz), Int32(1)))
    SparseMatrixCSC{Float64,Int64}(size(pattern, 1), size(pattern, 2), copy(pattern.colptr), copy(pattern.rowval), nz)
end

norm_code(m::AbstractString) = m == "logarithm" ? Int32(0) : m == "none" ? Int32(1) :
    throw(ArgumentError("Unknown normalization method: $m"))              # processing.jl:28

function dev_normalize(ctx, raw::Handle, n_genes; scale_factor, norm_method, pseudocount, nan_to_zero::Bool)
    gstats = vec_alloc(ctx, 2 * n_genes)        # sum x, sum x^2 per gene   (find_variable_genes)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(sym(:csc_normalize), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Float64, Int32, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
        ctx.h, raw.h, Float64(scale_factor), norm_code(norm_method), Float64(pseudocount), Int32(nan_to_zero), gstats.h, r))
    Handle(r[], :csc_sparse_free, ctx), gstats
end

function dev_hvg(ctx, gstats::Handle, n_genes, n_cells, nFeatures, span)
    mean = Vector{Float64}(undef, n_genes); var = similar(mean); vexp = similar(mean); vstd = similar(mean)
    order = Vector{Int64}(undef, n_genes)
    GC.@preserve mean var vexp vstd order check(ctx, ccall(sym(:csc_hvg_finish), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
        ctx.h, gstats.h, n_genes, n_cells, nFeatures, span, pointer(mean), pointer(var), pointer(vexp), pointer(vstd),
        pointer(order)))
    mean, var, vexp, vstd, order .+ 1
end

# ------------------------------------------------------------------------------------------------
# device-resident state of one object (so consecutive stages do not round-trip through the host)
# ------------------------------------------------------------------------------------------------
mutable struct Session
    ctx::Context
    raw::Union{Handle,Nothing}
    raw_id::UInt
    gstats::Union{Handle,Nothing}
    norm::Union{Handle,Nothing}
    scaled::Union{Handle,Nothing}       # dense, cell-major [N x H]
    emb::Union{Handle,Nothing}          # dense [N x k]
end
const _sessions = WeakKeyDict{Any,Session}()
function session(obj; ctx=nothing)
    get!(_sessions, obj) do
        Session(ctx === nothing ? default_context() : ctx, nothing, UInt(0), nothing, nothing, nothing, nothing)
    end
end
function ensure_raw!(s::Session, obj)
    m = obj.rawCount.count_mtx
    if s.raw === nothing || s.raw_id != objectid(m)
        s.raw = upload(s.ctx, m)
        s.raw_id = objectid(m)
        s.gstats = nothing; s.norm = nothing
    end
    s.raw
end

# ------------------------------------------------------------------------------------------------
# a1  normalize_object (processing.jl:15-48)
# ------------------------------------------------------------------------------------------------
function normalize_object(mtx::AbstractMatrix{<:Real}; scale_factor=10000, norm_method="logarithm", pseudocount=1, ctx=default_context())
    m = mtx isa SparseMatrixCSC ? mtx : sparse(mtx)
    raw = upload(ctx, m)
    nh, _, _ = dev_normalize(ctx, raw, size(m, 1); scale_factor, norm_method, pseudocount, nan_to_zero=false)
    out = download_like(ctx, nh, m)
    mtx isa SparseMatrixCSC ? out : Matrix(out)       # dense in, dense out (atac_processing.jl:87-89)
end

function normalize_object(ct_obj::cs.RawCountObject; scale_factor=10000, norm_method="logarithm", pseudocount=1, ctx=default_context())
    norm_count = normalize_object(ct_obj.count_mtx; scale_factor, norm_method, pseudocount, ctx)
    cs.NormCountObject(norm_count, ct_obj.cell_name, ct_obj.gene_name, scale_factor, norm_method, pseudocount)
end

function normalize_object(sc_obj::cs.get_object_group("All"); scale_factor=10000, norm_method="logarithm", pseudocount=1, ctx=nothing)
    s = session(sc_obj; ctx)
    raw = ensure_raw!(s, sc_obj)
    m = sc_obj.rawCount.count_mtx
    # processing.jl:44-46: imaging objects replace NaN (zero-sum cells) by 0
    imaging = isa(sc_obj, Union{cs.MerfishObject,cs.CartanaObject,cs.XeniumObject})
    s.norm, s.gstats = dev_normalize(s.ctx, raw, size(m, 1); scale_factor, norm_method, pseudocount, nan_to_zero=imaging)
    pattern = m isa SparseMatrixCSC ? m : sparse(m)
    sc_obj.normCount = cs.NormCountObject(download_like(s.ctx, s.norm, pattern), sc_obj.rawCount.cell_name,
                                          sc_obj.rawCount.gene_name, scale_factor, norm_method, pseudocount)
    sc_obj
end

# ------------------------------------------------------------------------------------------------
# a2  find_variable_genes (processing.jl:133-162)
# ------------------------------------------------------------------------------------------------
function _vst(genes, mean, var, vexp, vstd, order, nFeatures)
    vst = DataFrame(mean=mean[order], variance=var[order], variance_expected=vexp[order],
                    variance_standardized=vstd[order], gene=genes[order])          # sorted descending, stable (:152)
    vst, genes[order[1:min(nFeatures, length(genes))]]
end

function find_variable_genes(ct::cs.RawCountObject; nFeatures::Int64=2000, span::Float64=0.3, ctx=default_context())
    raw = upload(ctx, ct.count_mtx)
    g, n = size(ct.count_mtx)
    gstats = vec_alloc(ctx, 2g)
    check(ctx, ccall(sym(:csc_gene_stats_partial), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx.h, raw.h, gstats.h))
    _vst(ct.gene_name, dev_hvg(ctx, gstats, g, n, nFeatures, span)..., nFeatures)
end

function find_variable_genes(sc_obj::cs.get_object_group("All"); nFeatures::Int64=2000, span::Float64=0.3, ctx=nothing)
    s = session(sc_obj; ctx)
    raw = ensure_raw!(s, sc_obj)
    g, n = size(sc_obj.rawCount.count_mtx)
    if s.gstats === nothing                       # normalize_object already took the moments in its pass
        s.gstats = vec_alloc(s.ctx, 2g)
        check(s.ctx, ccall(sym(:csc_gene_stats_partial), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), s.ctx.h, raw.h, s.gstats.h))
    end
    vst, feats = _vst(sc_obj.rawCount.gene_name, dev_hvg(s.ctx, s.gstats, g, n, nFeatures, span)..., nFeatures)
    sc_obj.varGene = cs.VariableGeneObject(feats, vst)
    sc_obj
end

# ------------------------------------------------------------------------------------------------
# a3  scale_object (processing.jl:65-93)
# ------------------------------------------------------------------------------------------------
"rows `fidx` (1-based gene ids, any order; nothing = all) of the normalised device matrix -> dense device matrix + host copy"
function dev_scale(ctx, norm::Handle, n_genes, n_cells, fidx; scale_max=10.0, do_scale=true, do_center=true, want_host=true)
    rows = fidx === nothing ? collect(1:n_genes) : sort(unique(fidx))        # subset_count keeps original order (utils.jl:72-76)
    H = length(rows)
    f0 = fidx === nothing ? Int64[] : Int64.(fidx .- 1)
    fptr = fidx === nothing ? Ptr{Int64}(C_NULL) : pointer(f0)
    part = vec_alloc(ctx, 2H)
    GC.@preserve f0 begin
        check(ctx, ccall(sym(:csc_scale_stats_partial), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cvoid}),
                         ctx.h, norm.h, fptr, length(f0), part.h))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx, ccall(sym(:csc_scale_fill), Int32,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cvoid}, Int64, Float64, Int32, Int32, Ptr{Ptr{Cvoid}}),
            ctx.h, norm.h, fptr, length(f0), part.h, n_cells, Float64(scale_max), Int32(do_scale), Int32(do_center), r))
        dense = Handle(r[], :csc_dense_free, ctx)
        host = nothing
        if want_host
            # device layout is cell-major [N x H] row-major == Julia's column-major H x N: one straight copy
            host = Matrix{Float64}(undef, H, n_cells)
            GC.@preserve host check(ctx, ccall(sym(:csc_dense_download), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64),
                                               dense.h, pointer(host), DT_F64, 0, n_cells))
        end
        return dense, host, rows
    end
end

function scale_object(count_mtx::AbstractMatrix{<:Real}; scale_max=10.0, do_scale::Bool=true, do_center::Bool=true, ctx=default_context())
    m = count_mtx isa SparseMatrixCSC ? count_mtx : sparse(count_mtx)
    norm = upload(ctx, m)
    _, host, _ = dev_scale(ctx, norm, nothing, size(m, 1), size(m, 2), nothing; scale_max, do_scale, do_center)
    host
end

function scale_object(sc_obj::cs.get_object_group("All"); features::Union{Vector{String},Nothing}=nothing, scale_max=10.0,
                      do_scale::Bool=true, do_center::Bool=true, ctx=nothing, kwargs...)
    s = session(sc_obj; ctx)
    nc = sc_obj.normCount
    if s.norm === nothing
        s.norm = upload(s.ctx, nc.count_mtx)
    end
    fidx = features === nothing ? nothing : Int64.(filter(!isnothing, indexin(features, nc.gene_name)))
    g, n = size(nc.count_mtx)
    # processing.jl:78-83 forwards only kwargs... to the matrix method: the effective values are always the
    # defaults (10.0, true, true) while the user's values are recorded in the object.  Mirrored, not fixed.
    eff = (scale_max=get(kwargs, :scale_max, 10.0), do_scale=get(kwargs, :do_scale, true), do_center=get(kwargs, :do_center, true))
    s.scaled, host, rows = dev_scale(s.ctx, s.norm, g, n, fidx; eff...)
    sc_obj.scaleCount = cs.ScaleCountObject(host, nc.cell_name, nc.gene_name[rows], do_scale, do_center, scale_max)
    sc_obj
end

# ------------------------------------------------------------------------------------------------
# a4  run_pca (processing.jl:164-193)
# ------------------------------------------------------------------------------------------------

"""
    scale_gram_partial!(ctx, norm, features0, part, n_cells, colsum, gram; scale_max=10.0, do_scale=true, do_center=true)

`scale_object` fused with the covariance accumulation of `run_pca` for callers that do not keep `scaleCount`: adds the
column sums and `Z'Z` of this process's cells into `colsum` / `gram` (device fp64 vectors, ready for the all-reduce)
without a dense scaled matrix in device memory.  `features0` are 0-based gene ids.
"""
function scale_gram_partial!(ctx, norm, features0::Vector{Int64}, part, n_cells::Integer, colsum, gram;
                             scale_max::Real=10.0, do_scale::Bool=true, do_center::Bool=true)
    GC.@preserve features0 check(ctx, ccall(sym(:csc_scale_gram_partial), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cvoid}, Int64, Float64, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
        ctx.h, norm.h, pointer(features0), length(features0), part.h, n_cells, Float64(scale_max), Int32(do_scale),
        Int32(do_center), colsum.h, gram.h, C_NULL))          # C_NULL: the fp16 operands are not kept (csc_pca_transform_split)
    return nothing
end

function run_pca(sc_obj::cs.get_object_group("All"); method=:svd, pratio=1, maxoutdim=10, package="MLJ", ctx=nothing)
    s = session(sc_obj; ctx)
    features = sc_obj.varGene.var_gene
    sgenes = sc_obj.scaleCount.gene_name
    n = length(sc_obj.scaleCount.cell_name)
    if length(sgenes) == length(sc_obj.rawCount.gene_name) && length(features) != length(sgenes)
        # processing.jl:166-167: scaleCount still has all genes -> take the variable genes (rows are scaled per gene,
        # so re-scaling the subset from normCount gives the same rows)
        s.norm === nothing && (s.norm = upload(s.ctx, sc_obj.normCount.count_mtx))
        fidx = Int64.(indexin(features, sc_obj.normCount.gene_name))
        dense, _, _ = dev_scale(s.ctx, s.norm, length(sc_obj.normCount.gene_name), n, fidx; want_host=false)
    elseif s.scaled !== nothing
        dense = s.scaled
    else
        z = Matrix{Float64}(sc_obj.scaleCount.count_mtx)                       # H x N column-major == [N x H] row-major
        r = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve z check(s.ctx, ccall(sym(:csc_dense_upload), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Ptr{Cvoid}}),
                                          s.ctx.h, pointer(z), DT_F64, size(z, 2), size(z, 1), r))
        dense = Handle(r[], :csc_dense_free, s.ctx)
    end
    H = Ref{Int64}(0)
    check(s.ctx, ccall(sym(:csc_dense_info), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Cvoid}}), dense.h, C_NULL, H, C_NULL))
    colsum = vec_alloc(s.ctx, H[]); gram = vec_alloc(s.ctx, H[] * H[])
    check(s.ctx, ccall(sym(:csc_pca_partial), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), s.ctx.h, dense.h, colsum.h, gram.h))
    pr = Ref{Ptr{Cvoid}}(C_NULL)
    check(s.ctx, ccall(sym(:csc_pca_solve), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int32, Float64, Int32, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Float64}),
        s.ctx.h, colsum.h, gram.h, H[], n, Int32(maxoutdim), 1e-9, Int32(500), pr, C_NULL, C_NULL))
    model = Handle(pr[], :csc_pca_free, s.ctx)
    er = Ref{Ptr{Cvoid}}(C_NULL)
    check(s.ctx, ccall(sym(:csc_pca_transform), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), s.ctx.h, dense.h, model.h, er))
    s.emb = Handle(er[], :csc_dense_free, s.ctx)
    k = Ref{Int32}(0)
    check(s.ctx, ccall(sym(:csc_pca_info), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       model.h, C_NULL, k, C_NULL, C_NULL, C_NULL, C_NULL))
    pv = Vector{Float64}(undef, k[]); tv = Ref{Float64}(0.0)
    GC.@preserve pv check(s.ctx, ccall(sym(:csc_pca_info), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                       model.h, C_NULL, C_NULL, pointer(pv), tv, C_NULL, C_NULL))
    kxn = Matrix{Float64}(undef, k[], n)                                        # [N x k] row-major == k x N column-major
    GC.@preserve kxn check(s.ctx, ccall(sym(:csc_dense_download), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64),
                                        s.emb.h, pointer(kxn), DT_F64, 0, n))
    proj = permutedims(kxn)                                                     # N x k Float64 (utils.jl:108-117, int_objects.jl:153)
    # "MLJ" branch: no model, no explained variance (:180-181).  The other branch stores percent_var (:186); its
    # loadings-as-embedding quirk (:185) is not replicated.
    percent_var = package == "MLJ" ? nothing : pv ./ tv[] .* 100
    sc_obj.dimReduction = cs.ReductionObject(cs.PCAObject(proj, nothing, percent_var, "PC", method, pratio, maxoutdim), nothing, nothing)
    sc_obj
end

# ------------------------------------------------------------------------------------------------
# a5/a6  run_clustering (processing.jl:195-318): exact kNN + adjacency on the GPU, then the reference's own Leiden
# ------------------------------------------------------------------------------------------------
metric_code(m) = occursin("cosine", lowercase(string(m))) ? Int32(0) : occursin("euclid", lowercase(string(m))) ? Int32(1) :
    throw(ArgumentError("unsupported metric $m (CosineDist, Euclidean)"))

"exact k nearest neighbours of every cell in PCA space: (indices k x n 1-based, distances k x n), self excluded"
function knn(sc_obj; n_neighbors=30, metric="cosine", ctx=nothing)
    s = session(sc_obj; ctx)
    emb = sc_obj.dimReduction.pca.cell_embedding
    n, d = size(emb)
    if s.emb === nothing
        kxn = permutedims(Matrix{Float64}(emb))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve kxn check(s.ctx, ccall(sym(:csc_dense_upload), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Ptr{Cvoid}}),
                                            s.ctx.h, pointer(kxn), DT_F64, n, d, r))
        s.emb = Handle(r[], :csc_dense_free, s.ctx)
    end
    ir = Ref{Ptr{Cvoid}}(C_NULL); dr = Ref{Ptr{Cvoid}}(C_NULL)
    check(s.ctx, ccall(sym(:csc_knn), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
        s.ctx.h, s.emb.h, s.emb.h, 0, Int32(0), Int32(d), Int32(n_neighbors), metric_code(metric), ir, dr))
    idx = Handle(ir[], :csc_vec_free, s.ctx); dist = Handle(dr[], :csc_vec_free, s.ctx)
    # [n x k] row-major == k x n column-major, the layout of knn_matrices (processing.jl:201)
    # Int64 / 1-based / Float64 on the device before the copy (csc_vec_download_as), straight into the Julia arrays
    indices = Matrix{Int64}(undef, n_neighbors, n); dists = Matrix{Float64}(undef, n_neighbors, n)
    GC.@preserve indices check(s.ctx, ccall(sym(:csc_vec_download_as), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int64),
                                            idx.h, pointer(indices), n * n_neighbors, DT_I64, 1))
    GC.@preserve dists check(s.ctx, ccall(sym(:csc_vec_download_as), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int64),
                                          dist.h, pointer(dists), n * n_neighbors, DT_F64, 0))
    indices, dists, idx
end

const GRAPH_BINARY, GRAPH_SUM, GRAPH_SNN = Int32(0), Int32(1), Int32(2)
function adjacency(s::Session, idx::Handle, n, k, mode; snn_prune=0.0)
    gr = Ref{Ptr{Cvoid}}(C_NULL)
    check(s.ctx, ccall(sym(:csc_graph_build), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int64, Int64, Int32, Float64, Ptr{Ptr{Cvoid}}),
                       s.ctx.h, idx.h, n, Int32(k), 0, n, mode, snn_prune, gr))
    g = Handle(gr[], :csc_graph_free, s.ctx)
    nz = Ref{Int64}(0)
    check(s.ctx, ccall(sym(:csc_graph_info), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), g.h, C_NULL, C_NULL, nz))
    cp = Vector{Int64}(undef, n + 1); rv = Vector{Int64}(undef, nz[])
    if mode == GRAPH_SNN
        v = Vector{Float64}(undef, nz[])
        GC.@preserve cp rv v check(s.ctx, ccall(sym(:csc_graph_download), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Int32, Int32),
                                                g.h, pointer(cp), pointer(rv), pointer(v), DT_F64, Int32(1)))
        return SparseMatrixCSC{Float64,Int64}(n, n, cp, rv, v), g
    end
    vi = Vector{Int64}(undef, nz[])            # the reference's adjacency is SparseMatrixCSC{Int64,Int64}
    GC.@preserve cp rv vi check(s.ctx, ccall(sym(:csc_graph_download), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Int32, Int32),
                                             g.h, pointer(cp), pointer(rv), pointer(vi), DT_I64, Int32(1)))
    SparseMatrixCSC{Int64,Int64}(n, n, cp, rv, vi), g
end

# f1  Leiden.leiden(adj_mat, resolution = res) (processing.jl:233-234, 279-280) on the device graph: csc_leiden
"communities of the device graph `g` (all n columns on this device): (labels 1-based Vector{Int}, quality, levels)"
function leiden_gpu(ctx::Context, g::Handle, n::Integer; resolution=0.06, max_levels=0, max_sweeps=0)
    lab = Vector{Int32}(undef, n)
    nc = Ref{Int32}(0); lv = Ref{Int32}(0); q = Ref{Float64}(0.0)
    GC.@preserve lab check(ctx, ccall(sym(:csc_leiden), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
        ctx.h, g.h, Float64(resolution), Int32(max_levels), Int32(max_sweeps), pointer(lab), nc, q, lv))
    Int.(lab) .+ 1, q[], Int(lv[]), Int(nc[])
end

function run_clustering(sc_obj::cs.get_object_group("All"); n_neighbors=30, metric=cs.CosineDist(), res=0.06, seed_use=1234,
                        graph=nothing, leiden::Symbol=:reference, ctx=nothing)
    s = session(sc_obj; ctx)
    n = size(sc_obj.rawCount.count_mtx, 2)
    indices, _, idx = knn(sc_obj; n_neighbors, metric, ctx)
    # processing.jl:310-317: n < 10000 -> dense binary union (:268-278); otherwise COO-summed sparse, mutual pairs carry 2 (:218-232)
    mode = graph === nothing ? (n < 10000 ? GRAPH_BINARY : GRAPH_SUM) : graph
    adj, g = adjacency(s, idx, n, n_neighbors, mode)
    adj_mat = (mode == GRAPH_BINARY && n < 10000) ? Matrix{Int64}(adj) : adj
    cluster = Vector{String}(undef, n)
    if leiden == :gpu
        # csc_leiden: the same CPM quality as Leiden.jl, parallel and deterministic (590 s -> 0.3 s at 1M cells); communities
        # are numbered in order of first appearance, `result` mimics the named tuple Leiden.leiden returns
        lab, quality, _, nc = leiden_gpu(s.ctx, g, n; resolution=res)
        partition = [Int[] for _ in 1:nc]
        for (i, c) in enumerate(lab)
            push!(partition[c], i)
        end
        result = (quality=quality, partition=partition)
        cluster .= string.(lab)
    elseif leiden == :reference
        Random.seed!(seed_use)
        result = cs.Leiden.leiden(adj_mat, resolution=res)                      # the reference's own community detection
        for (i, members) in enumerate(result.partition)
            cluster[members] .= string(i)
        end
    else
        throw(ArgumentError("leiden must be :reference or :gpu"))
    end
    df = DataFrame(cluster=cluster, cell_id=sc_obj.rawCount.cell_name)           # already in colnames(sc_obj) order (:241)
    sc_obj.clustData = cs.ClusteringObject(df, string(metric), adj_mat, result, res)
    sc_obj.metaData.cluster = df.cluster
    if isa(sc_obj, cs.get_object_group("Imaging"))
        sc_obj.spmetaData.cell.cluster = df.cluster                               # (:246-259; polygon remapping left to the reference)
    end
    if sc_obj.dimReduction.umap !== nothing
        sc_obj.dimReduction.umap.knn_data = indices                               # what run_clustering_atlas reads (:196-197)
    end
    sc_obj
end

# ------------------------------------------------------------------------------------------------
# f2-f4: thin bindings of the entry points behind run_harmony, find_markers / find_all_markers, read_10x, merge_matrices and
# run_tf_idf.  One wrapper per C function, Julia arrays in and out; the bookkeeping around them (cluster labels -> group
# codes, the Mann-Whitney normal approximation from the rank sums, Bonferroni, gene / barcode names, the HarmonyObject
# fields) is host code that cellscopes.jl_b200/{processing,integration,ingest}.py spell out line by line against the
# reference (processing.jl:350-414, int_objects.jl:115-236, fileio.jl:1-37, int_utils.jl:1-23, atac_processing.jl:1-13).
# ------------------------------------------------------------------------------------------------
"csc_kmeans: Lloyd's k-means (fp64) from the rows `init_idx` (1-based) of x (n x d); unit_rows = Harmony's Z_cos transform first"
function kmeans_gpu(ctx::Context, x::Matrix{Float64}, K::Integer, init_idx::Vector{Int}; max_iter=100, tol=1e-4, unit_rows=false)
    n, d = size(x)
    xt = permutedims(x)                                   # cell-major memory = the reference's d x N data_mat
    idx0 = Int64.(init_idx .- 1)
    cen = Matrix{Float64}(undef, d, K); lab = Vector{Int32}(undef, n)
    obj = Ref{Float64}(0.0); it = Ref{Int32}(0)
    GC.@preserve xt idx0 cen lab check(ctx, ccall(sym(:csc_kmeans), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Int64}, Int32, Float64, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
        ctx.h, pointer(xt), n, Int32(d), Int32(K), pointer(idx0), Int32(max_iter), Float64(tol), Int32(unit_rows), pointer(cen),
        pointer(lab), obj, it))
    permutedims(cen), Int.(lab) .+ 1, obj[], Int(it[])    # centres K x d
end

"csc_harmony: Z_corr (n x d) of the embedding z (n x d); batch = 0-based level codes; y0 = K x d k-means centres of the unit rows"
function harmony_gpu(ctx::Context, z::Matrix{Float64}, batch::Vector{Int32}, n_batch::Integer, y0::Matrix{Float64},
                     theta::Vector{Float64}, lamb::Vector{Float64}, sigma::Vector{Float64}; block_size=0.05,
                     max_iter_harmony=10, max_iter_kmeans=20, eps_kmeans=1e-5, eps_harmony=1e-4, seed=123)
    n, d = size(z); K = size(y0, 1)
    zt = permutedims(z); y0t = permutedims(y0)
    out = Matrix{Float64}(undef, d, n)
    obj = zeros(Float64, max_iter_harmony + 2); rounds = zeros(Int32, max_iter_harmony + 1)
    n_obj = Ref{Int32}(length(obj)); n_rounds = Ref{Int32}(length(rounds))
    GC.@preserve zt batch y0t theta lamb sigma out obj rounds check(ctx, ccall(sym(:csc_harmony), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Float64, Int32, Int32, Float64, Float64, UInt64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        ctx.h, pointer(zt), n, Int32(d), pointer(batch), Int32(n_batch), pointer(y0t), Int32(K), pointer(theta), pointer(lamb),
        pointer(sigma), Float64(block_size), Int32(max_iter_harmony), Int32(max_iter_kmeans), Float64(eps_kmeans),
        Float64(eps_harmony), UInt64(seed), pointer(out), pointer(obj), n_obj, pointer(rounds), n_rounds, C_NULL))
    permutedims(out), obj[1:n_obj[]], Int.(rounds[1:n_rounds[]])
end

"csc_marker_stats: per-(gene, group) tables of the Mann-Whitney tests for EVERY group at once; labels 0-based, -1 = not taking part"
function marker_stats(ctx::Context, norm::Handle, n_genes::Integer, labels::Vector{Int32}, n_groups::Integer; expr_cutoff=0.0)
    tab() = Matrix{Float64}(undef, n_groups, n_genes)      # C row-major [n_genes x n_groups] = Julia n_groups x n_genes
    rank_sum, sum_expm1, n_expr, nnzs, sum_x = tab(), tab(), tab(), tab(), tab()
    tie_adj, n_neg, n_zero = (Vector{Float64}(undef, n_genes) for _ in 1:3)
    n_incl = Ref{Int64}(0)
    GC.@preserve labels rank_sum sum_expm1 n_expr nnzs sum_x tie_adj n_neg n_zero check(ctx, ccall(sym(:csc_marker_stats), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
        ctx.h, norm.h, pointer(labels), Int32(n_groups), Float64(expr_cutoff), pointer(rank_sum), pointer(sum_expm1), pointer(n_expr),
        pointer(nnzs), pointer(sum_x), pointer(tie_adj), pointer(n_neg), pointer(n_zero), n_incl))
    (rank_sum=rank_sum, sum_expm1=sum_expm1, n_expr=n_expr, nnz=nnzs, sum_x=sum_x, tie_adj=tie_adj, n_neg=n_neg,
     n_zero_stored=n_zero, n_included=n_incl[])
end

"csc_read_mtx: matrix.mtx / matrix.mtx.gz of read_10x (fileio.jl:7,16) -> device CSC (genes x cells)"
function read_mtx(ctx::Context, path::AbstractString)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(sym(:csc_read_mtx), Int32, (Ptr{Cvoid}, Cstring, Ptr{Ptr{Cvoid}}), ctx.h, path, r))
    Handle(r[], :csc_sparse_free, ctx)
end

"csc_sparse_merge: merge_matrices (int_utils.jl:1-23); row_maps[i][r] = 1-based row of the merged matrix for row r of matrix i"
function merge_gpu(ctx::Context, mats::Vector{Handle}, row_maps::Vector{Vector{Int}}, n_rows_merged::Integer)
    maps0 = [Int64.(m .- 1) for m in row_maps]
    hp = [m.h for m in mats]; mp = [pointer(m) for m in maps0]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve maps0 hp mp check(ctx, ccall(sym(:csc_sparse_merge), Int32,
        (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Int64}}, Int32, Int64, Ptr{Ptr{Cvoid}}),
        ctx.h, pointer(hp), pointer(mp), Int32(length(mats)), n_rows_merged, r))
    Handle(r[], :csc_sparse_free, ctx)
end

"csc_tf_idf: run_tf_idf (atac_processing.jl:1-13) on the stored pattern; gene_stats = the table csc_gene_stats_partial filled"
function tf_idf_gpu(ctx::Context, raw::Handle, gene_stats::Handle, n_cells_total::Integer; scale_factor=10000)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall(sym(:csc_tf_idf), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float64, Ptr{Ptr{Cvoid}}),
                     ctx.h, raw.h, gene_stats.h, n_cells_total, Float64(scale_factor), r))
    Handle(r[], :csc_sparse_free, ctx)
end

# ------------------------------------------------------------------------------------------------
# drop-in: re-point the reference's object-level methods at this module
# ------------------------------------------------------------------------------------------------
"Overwrite CellScopes' object-level methods of the five hot-path functions with the GPU path."
function enable!()
    # the methods are defined inside CellScopes and call back into THIS module, wherever it was loaded (the module object
    # itself is spliced into the expression: no assumption that it lives in Main)
    shim = @__MODULE__
    @eval CellScopes begin
        normalize_object(sc_obj::get_object_group("All"); kw...) = $shim.normalize_object(sc_obj; kw...)
        find_variable_genes(sc_obj::get_object_group("All"); kw...) = $shim.find_variable_genes(sc_obj; kw...)
        scale_object(sc_obj::get_object_group("All"); kw...) = $shim.scale_object(sc_obj; kw...)
        run_pca(sc_obj::get_object_group("All"); kw...) = $shim.run_pca(sc_obj; kw...)
        run_clustering(sc_obj::get_object_group("All"); kw...) = $shim.run_clustering(sc_obj; kw...)
    end
    nothing
end

end # module
