# GeoStatsModelsB200.jl -- thin Julia shim that routes the interpolation hot path of
# GeoStatsModels.jl (v0.13.10) to libgsm_b200.so through `ccall`.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image.  The file is kept
# literal and small so a maintainer can check it against include/gsm_b200.h line by line; the Python
# package `geostatsmodels.jl_b200/api.py` implements the same orchestration and is what the tests run.
#
# What it does: adds more-specific methods of `GeoStatsModels.fitpredictneigh` / `fitpredictfull`
# (src/models.jl:131-234) for the signatures the GPU path claims -- Kriging (Simple / Ordinary / Universal
# built by `(deg, dim)`, i.e. stock `Kriging(γ, deg, dim)`), IDW and NN with Euclidean distance, univariate
# Float64 (or Union{Missing,Float64}) Cartesian PointSet data, CartesianGrid or PointSet targets, point support or
# (Kriging on a CartesianGrid) volume support: `point=false` hands the library the points upstream integrates over.
# Every other signature keeps dispatching to the untouched generic methods (that is dispatch, not a CPU
# fallback inside the library).  `fitpredict` itself (models.jl:102-129) is unchanged, so `_pointsupport`,
# `_scale` and the final `georef` run exactly as in the reference and the shim receives ALREADY SCALED
# model, data and domain.
#
# Devices: GSM_B200_DEVICES="0,1,2,3" (default "0").  More than one device goes through gsm_multi_* (samples
# replicated by one NCCL broadcast inside the library, target slabs, one host thread per device).
module GeoStatsModelsB200

using GeoStatsModels, GeoStatsFunctions, Meshes, GeoTables, Tables, Distributions, Distances
using LinearAlgebra: I
using Unitful: ustrip
import GeoStatsModels: fitpredictneigh, fitpredictfull, SimpleKriging, OrdinaryKriging, UniversalKriging, IDW, NN

const LIB = get(ENV, "GSM_B200_LIB", "libgsm_b200.so")
const MAXK = 64            # GSM_MAX_NEIGHBORS
const GSM_ERR_UNSUPPORTED = 4

# ---- POD mirrors of include/gsm_b200.h (ABI version 2) -----------------------------------------
struct GsmStructure
  kind::Int32; _pad::Int32; sill::Float64; nugget::Float64; range::Float64; metric::NTuple{9,Float64}
end
struct GsmVariogram
  kind::Int32; nstruct::Int32; sill::Float64; nugget::Float64; range::Float64; structs::NTuple{4,GsmStructure}
end
struct GsmModel; kind::Int32; ndrift::Int32; mean::Float64; exps::NTuple{30,Int32}; end
struct GsmTargets
  kind::Int32; dim::Int32
  origin::NTuple{3,Float64}; spacing::NTuple{3,Float64}; dims::NTuple{3,Int64}
  points::Ptr{Float64}; npoints::Int64; first::Int64; count::Int64
end
struct GsmNeighborOpts; maxneighbors::Int32; minneighbors::Int32; radius::Float64; end
struct GsmSummary; missing::Int64; singular::Int64; nonpositive_var::Int64; total::Int64; end

const VARIO_KIND = Dict(GaussianVariogram => 0, SphericalVariogram => 1, ExponentialVariogram => 2, CubicVariogram => 3,
                        PentasphericalVariogram => 4, SineHoleVariogram => 5, CircularVariogram => 6, NuggetEffect => 7)
const NOSTRUCT = GsmStructure(0, 0, 0.0, 0.0, 1.0, ntuple(_ -> 0.0, 9))

# ---- handles: one context, or one multi-GPU handle, per process ----------------------------------
mutable struct Handle
  ptr::Ptr{Cvoid}
  multi::Bool
  function Handle(devices::Vector{Int32})
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    multi = length(devices) > 1
    rc = multi ? ccall((:gsm_multi_create, LIB), Cint, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devices, length(devices), ref) :
                 ccall((:gsm_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), devices[1], ref)
    rc == 0 || error(unsafe_string(multi ? ccall((:gsm_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL) :
                                           ccall((:gsm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = new(ref[], multi)
    finalizer(x -> x.multi ? ccall((:gsm_multi_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr) :
                             ccall((:gsm_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
  end
end
const HANDLE = Ref{Union{Nothing,Handle}}(nothing)
handle() = (HANDLE[] === nothing && (HANDLE[] = Handle(parse.(Int32, split(get(ENV, "GSM_B200_DEVICES", "0"), ",")))); HANDLE[])
lasterror(h) = unsafe_string(h.multi ? ccall((:gsm_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr) :
                                       ccall((:gsm_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))
struct Unsupported <: Exception end                      # the library does not claim this call: take the generic method
check(h, rc) = rc == 0 ? nothing : (rc == GSM_ERR_UNSUPPORTED ? throw(Unsupported()) : throw(ErrorException(lasterror(h))))

# page-locked output arrays (gsm_host_alloc): results DMA straight into the memory the table columns wrap
function pinned(::Type{T}, n::Integer) where {T}
  ref = Ref{Ptr{Cvoid}}(C_NULL)
  ccall((:gsm_host_alloc, LIB), Cint, (Ref{Ptr{Cvoid}}, Int64), ref, max(n, 1) * sizeof(T)) == 0 || error("gsm_host_alloc failed")
  a = unsafe_wrap(Array, Ptr{T}(ref[]), n; own=false)
  finalizer(x -> ccall((:gsm_host_free, LIB), Cint, (Ptr{Cvoid},), pointer(x)), a)
end

# ---- descriptors ---------------------------------------------------------------------------------
# one structure: family, sill, nugget and metric ball (isotropic: range; anisotropic: diag(1 ./ radii) * R', range 1)
function structure(γ, c=1.0)
  k = get(VARIO_KIND, typeof(γ).name.wrapper, -1)
  k < 0 && throw(Unsupported())
  ball = GeoStatsFunctions.metricball(γ)
  r = ustrip.(radii(ball))
  if length(unique(r)) == 1
    GsmStructure(k, 0, c * ustrip(sill(γ)), c * ustrip(nugget(γ)), r[1], (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0))
  else
    d = length(r); R = Matrix(rotation(ball)); M = zeros(3, 3)
    M[1:d, 1:d] = (1 ./ r) .* R'                               # row i scaled by 1 / r_i
    GsmStructure(k, 0, c * ustrip(sill(γ)), c * ustrip(nugget(γ)), 1.0, Tuple(M'))   # row-major
  end
end
function variogram(γ)
  nvariables(γ) == 1 || throw(Unsupported())
  if γ isa GeoStatsFunctions.CompositeFunction                 # sum of scaled structures
    cs, fs = GeoStatsFunctions.structures(γ)[2:3]              # (nugget, coefficients, functions)
    cn = GeoStatsFunctions.structures(γ)[1]
    ss = GsmStructure[]
    ustrip(first(cn)) > 0 && push!(ss, GsmStructure(7, 0, ustrip(first(cn)), ustrip(first(cn)), 1.0, Tuple(Matrix(1.0I, 3, 3))))
    for (c, f) in zip(cs, fs); push!(ss, structure(f, ustrip(first(c)))); end
    length(ss) <= 4 || throw(Unsupported())
    return GsmVariogram(0, length(ss), 0.0, 0.0, 1.0, ntuple(i -> i <= length(ss) ? ss[i] : NOSTRUCT, 4))
  end
  s = structure(γ)
  iso = s.metric == (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
  iso ? GsmVariogram(s.kind, 0, s.sill, s.nugget, s.range, ntuple(_ -> NOSTRUCT, 4)) :
        GsmVariogram(0, 1, 0.0, 0.0, 1.0, (s, NOSTRUCT, NOSTRUCT, NOSTRUCT))
end

# UniversalKriging(fun, deg, dim) = UniversalKriging(fun, monomials(deg, dim)) (universal.jl:50): every drift is the
# closure `p -> prod(x(p) .^ n)` of universal.jl:60-62, which captures its exponent vector `n` as a field.  Recover the
# table from there (stock `Kriging(γ, deg, dim)` is then accelerated); anything else is an arbitrary closure
# (universal.jl:26) that cannot cross a C ABI.
function drift_exponents(drifts, dim)
  es = Vector{Vector{Int}}()
  for f in drifts
    (parentmodule(typeof(f)) === GeoStatsModels && hasfield(typeof(f), :n)) || return nothing
    n = collect(Int, getfield(f, :n))
    (length(n) == dim && all(>=(0), n)) || return nothing
    push!(es, n)
  end
  1 <= length(es) <= 10 ? es : nothing
end
exps_tuple(es, dim) = ntuple(i -> (n = (i - 1) ÷ 3 + 1; d = (i - 1) % 3 + 1; n <= length(es) && d <= dim ? Int32(es[n][d]) : Int32(0)), 30)
gsmmodel(m::SimpleKriging, dim) = GsmModel(0, 0, first(m.mean), ntuple(_ -> Int32(0), 30))
gsmmodel(m::OrdinaryKriging, dim) = GsmModel(1, 0, 0.0, ntuple(_ -> Int32(0), 30))
function gsmmodel(m::UniversalKriging, dim)
  es = drift_exponents(m.drifts, dim)
  es === nothing && throw(Unsupported())
  GsmModel(2, length(es), 0.0, exps_tuple(es, dim))
end

coordmatrix(pset::PointSet) = reduce(hcat, [collect(ustrip.(to(p))) for p in pset])  # dim x n, column-major == AoS

function targets(dom::CartesianGrid)
  dim = embeddim(dom)
  o = ustrip.(to(minimum(dom))); s = ustrip.(spacing(dom)); d = size(dom)
  pad(t, z) = ntuple(i -> i <= dim ? t[i] : z, 3)
  GsmTargets(0, dim, pad(o, 0.0), pad(s, 0.0), pad(d, 1), C_NULL, 0, 0, -1), nothing
end
function targets(dom::PointSet)
  P = coordmatrix(dom)
  GsmTargets(1, size(P, 1), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0, 0, 0), pointer(P), size(P, 2), 0, -1), P
end

column(dat) = Tables.getcolumn(Tables.columns(values(dat)), 1)
claimed(dat, dom, point, distance) =
  distance isa Euclidean && domain(dat) isa PointSet && dom isa Union{CartesianGrid,PointSet} &&
  length(Tables.columnnames(Tables.columns(values(dat)))) == 1 &&
  nonmissingtype(eltype(column(dat))) === Float64

# Change of support (point=false, models.jl:114-115,136,205): the points GeoStatsFunctions integrates over for one
# cell, relative to its centroid.  Every cell of a CartesianGrid is congruent, so one table serves the whole domain;
# the elements of a PointSet are points anyway, and IDW / NN only look at centroids (idw.jl:82, nn.jl).  The sampling
# rule stays upstream's own (`GeoStatsFunctions._sample`, an internal name [RECALLED]); without it the call is not claimed.
function blocktable(γ, dom, point)
  (point || !(dom isa CartesianGrid)) && return nothing
  isdefined(GeoStatsFunctions, :_sample) || throw(Unsupported())
  g = dom[1]
  c = collect(ustrip.(to(centroid(g))))
  B = reduce(hcat, [collect(ustrip.(to(p))) .- c for p in GeoStatsFunctions._sample(γ, g)])   # dim x nq == AoS
  size(B, 2) <= 512 || throw(Unsupported())                  # GSM_MAX_BLOCK_POINTS
  B
end
contexts(h) = h.multi ? [ccall((:gsm_multi_context, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), h.ptr, i - 1)
                         for i in 1:ccall((:gsm_multi_device_count, LIB), Cint, (Ptr{Cvoid},), h.ptr)] : [h.ptr]
function setblock!(h, B)
  for c in contexts(h)
    rc = B === nothing ? ccall((:gsm_set_block_support, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), c, C_NULL, 0, 0) :
                         ccall((:gsm_set_block_support, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), c, B, size(B, 1), size(B, 2))
    check(h, rc)
  end
end

# neighbour limits exactly like models.jl:144-150; the library takes up to MAXK neighbours
clampmax(maxneighbors, nobs) = (maxneighbors > nobs || maxneighbors < 1) ? nobs : maxneighbors

function setsamples!(h, dat)
  X = coordmatrix(domain(dat))
  z = Float64[ismissing(v) ? NaN : v for v in column(dat)]     # `missing` crosses the ABI as NaN
  f = h.multi ? :gsm_multi_set_samples : :gsm_set_samples
  GC.@preserve X z check(h, ccall((f, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Float64}),
                                   h.ptr, X, size(X, 1), size(X, 2), z))
end
function summary(h)
  s = Ref(GsmSummary(0, 0, 0, -1))
  ccall((h.multi ? :gsm_multi_get_summary : :gsm_get_summary, LIB), Cint, (Ptr{Cvoid}, Ref{GsmSummary}), h.ptr, s)
  s[]
end
generic_neigh(args...) = invoke(fitpredictneigh, Tuple{GeoStatsModel,Any,Any,Any,Any,Any,Any,Any,Any}, args...)
generic_full(args...) = invoke(fitpredictfull, Tuple{GeoStatsModel,Any,Any,Any,Any}, args...)

# status byte -> table column: 1 = `missing` (models.jl:178-181); 2 = factorisation failed: the reference's
# check=false Bunch-Kaufman (krig.jl:152) returns non-finite values there, the library NaN
function columnof(mean, vari, status, prob, s::GsmSummary)
  if s.total == length(mean) && s.missing == 0            # the common case: no scan of the status column at all
    if prob
      s.nonpositive_var > 0 && throw(DomainError(-1.0, "non-positive Kriging variance"))   # Normal(μ, σ) at krig.jl:228
      return s.singular == 0 ? reinterpret_normal(mean, vari) : [st == 2 ? Normal(NaN, NaN) : Normal(mean[i], vari[i]) for (i, st) in enumerate(status)]
    end
    return mean
  end
  prob ? [st == 1 ? missing : Normal(mean[i], vari[i]) for (i, st) in enumerate(status)] :
         [st == 1 ? missing : mean[i] for (i, st) in enumerate(status)]
end
reinterpret_normal(mean, vari) = Normal.(mean, vari)       # Normal.(mean, var), models.jl:298-299

# ---- fitpredictneigh (models.jl:131-198) -------------------------------------------------------
function fitpredictneigh(model::Union{SimpleKriging,OrdinaryKriging,UniversalKriging}, dat, dom, point, prob,
                         minneighbors, maxneighbors, neighborhood, distance)
  args = (model, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  (claimed(dat, dom, point, distance) && (neighborhood === nothing || neighborhood isa MetricBall) &&
   clampmax(maxneighbors, nrow(dat)) <= MAXK) || return generic_neigh(args...)
  try
    h = handle()
    v = Ref(variogram(model.fun)); m = Ref(gsmmodel(model, embeddim(dom)))
    B = blocktable(model.fun, dom, point)
    setsamples!(h, dat); setblock!(h, B)
    M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
    mean = pinned(Float64, M); vari = prob ? pinned(Float64, M) : Float64[]; status = pinned(UInt8, M)
    tg, keep = targets(dom)
    o = Ref(GsmNeighborOpts(maxneighbors, minneighbors, neighborhood === nothing ? 0.0 : ustrip(radius(neighborhood))))
    GC.@preserve keep mean vari status begin
      if h.multi
        check(h, ccall((:gsm_multi_local_krige, LIB), Cint,
          (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{GsmTargets}, Ref{GsmNeighborOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
          h.ptr, v, m, Ref(tg), o, mean, prob ? pointer(vari) : C_NULL, status))
      else
        check(h, ccall((:gsm_local_krige, LIB), Cint,
          (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{GsmTargets}, Ref{GsmNeighborOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int32}),
          h.ptr, v, m, Ref(tg), o, mean, prob ? pointer(vari) : C_NULL, status, C_NULL))
      end
    end
    B === nothing || setblock!(h, nothing)
    col = columnof(mean, vari, status, prob, summary(h))
    return (; var => col) |> Tables.materializer(values(dat))
  catch e
    e isa Unsupported || rethrow()
    return generic_neigh(args...)
  end
end

# ---- fitpredictfull (models.jl:200-234) --------------------------------------------------------
function fitpredictfull(model::Union{SimpleKriging,OrdinaryKriging,UniversalKriging}, dat, dom, point, prob)
  args = (model, dat, dom, point, prob)
  (claimed(dat, dom, point, Euclidean()) && !any(ismissing, column(dat))) || return generic_full(args...)
  try
    h = handle()
    v = Ref(variogram(model.fun)); m = Ref(gsmmodel(model, embeddim(dom)))
    B = blocktable(model.fun, dom, point)
    setsamples!(h, dat); setblock!(h, B)
    M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
    mean = pinned(Float64, M); vari = prob ? pinned(Float64, M) : Float64[]
    tg, keep = targets(dom)
    if h.multi
      ok = Ref{Int32}(0)
      GC.@preserve keep mean vari check(h, ccall((:gsm_multi_global_krige, LIB), Cint,
        (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{GsmTargets}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
        h.ptr, v, m, Ref(tg), mean, prob ? pointer(vari) : C_NULL, ok))
    else
      fit = Ref{Ptr{Cvoid}}(C_NULL)
      check(h, ccall((:gsm_global_krige_fit, LIB), Cint, (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{Ptr{Cvoid}}), h.ptr, v, m, fit))
      try
        GC.@preserve keep mean vari check(h, ccall((:gsm_global_krige_predict, LIB), Cint,
          (Ptr{Cvoid}, Ref{GsmTargets}, Ptr{Float64}, Ptr{Float64}), fit[], Ref(tg), mean, prob ? pointer(vari) : C_NULL))
      finally
        ccall((:gsm_global_krige_free, LIB), Cint, (Ptr{Cvoid},), fit[])
      end
    end
    B === nothing || setblock!(h, nothing)
    col = prob ? Normal.(mean, vari) : mean
    return (; var => col) |> Tables.materializer(values(dat))
  catch e
    e isa Unsupported || rethrow()
    return generic_full(args...)
  end
end

# ---- IDW (idw.jl:43-89) and NN (nn.jl:47-54) -----------------------------------------------------
function idwcall(model::IDW, dat, dom, opts, prob)
  h = handle(); setsamples!(h, dat)
  M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
  mean = pinned(Float64, M); status = pinned(UInt8, M)
  tg, keep = targets(dom)
  GC.@preserve keep mean status check(h, ccall((h.multi ? :gsm_multi_idw : :gsm_idw, LIB), Cint,
    (Ptr{Cvoid}, Float64, Ref{GsmTargets}, Ptr{GsmNeighborOpts}, Ptr{Float64}, Ptr{UInt8}),
    h.ptr, Float64(model.exponent), Ref(tg), opts === nothing ? C_NULL : Ref(opts), mean, status))
  # a `missing` among the weighted values makes the sum `missing` (idw.jl:59-73): NaN on the device
  gone(i, s) = s == 1 || isnan(mean[i])
  col = (any(==(0x01), status) || any(ismissing, column(dat))) ?
        [gone(i, s) ? missing : (prob ? Dirac(mean[i]) : mean[i]) for (i, s) in enumerate(status)] :
        (prob ? Dirac.(mean) : mean)                         # predictprob = Dirac(predict) (idw.jl:57)
  (; var => col) |> Tables.materializer(values(dat))
end
function fitpredictfull(model::IDW{<:Real,<:Euclidean}, dat, dom, point, prob)
  claimed(dat, dom, point, model.distance) || return generic_full(model, dat, dom, point, prob)
  idwcall(model, dat, dom, nothing, prob)
end
function fitpredictneigh(model::IDW{<:Real,<:Euclidean}, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  (claimed(dat, dom, point, distance) && clampmax(maxneighbors, nrow(dat)) <= MAXK &&
   (neighborhood === nothing || neighborhood isa MetricBall)) ||
    return generic_neigh(model, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  idwcall(model, dat, dom, GsmNeighborOpts(maxneighbors, minneighbors, neighborhood === nothing ? 0.0 : ustrip(radius(neighborhood))), prob)
end
function fitpredictneigh(model::NN{<:Euclidean}, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  h = handle()
  plain = !any(ismissing, column(dat)) && neighborhood === nothing
  (claimed(dat, dom, point, distance) && !h.multi && (plain || (clampmax(maxneighbors, nrow(dat)) <= MAXK &&
   (neighborhood === nothing || neighborhood isa MetricBall)))) ||
    return generic_neigh(model, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  setsamples!(h, dat)
  M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
  val = pinned(Float64, M)
  tg, keep = targets(dom)
  if !plain   # missings or a ball: nearest of the k nearest that has a value, `missing` when none (gsm_nn_local)
    status = pinned(UInt8, M)
    o = Ref(GsmNeighborOpts(maxneighbors, minneighbors, neighborhood === nothing ? 0.0 : ustrip(radius(neighborhood))))
    GC.@preserve keep val status check(h, ccall((:gsm_nn_local, LIB), Cint,
      (Ptr{Cvoid}, Ref{GsmTargets}, Ref{GsmNeighborOpts}, Ptr{Float64}, Ptr{UInt8}), h.ptr, Ref(tg), o, val, status))
    col = [s == 1 ? missing : (prob ? Dirac(val[i]) : val[i]) for (i, s) in enumerate(status)]
    return (; var => col) |> Tables.materializer(values(dat))
  end
  GC.@preserve keep val check(h, ccall((:gsm_nn, LIB), Cint, (Ptr{Cvoid}, Ref{GsmTargets}, Ptr{Float64}), h.ptr, Ref(tg), val))
  (; var => (prob ? Dirac.(val) : val)) |> Tables.materializer(values(dat))
end

end # module
