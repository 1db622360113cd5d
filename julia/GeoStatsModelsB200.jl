# GeoStatsModelsB200.jl -- thin Julia shim that routes the interpolation hot path of
# GeoStatsModels.jl (v0.13.10) to libgsm_b200.so through `ccall`.
#
# NOT EXECUTED in this repository's CI: Julia is not installed in the build image.  The file is kept
# literal and small so a maintainer can check it against include/gsm_b200.h line by line; the Python
# package `geostatsmodels.jl_b200/api.py` implements the same orchestration and is what the tests run.
#
# What it does: adds more-specific methods of `GeoStatsModels.fitpredictneigh` / `fitpredictfull`
# (src/models.jl:131-234) for the signatures the GPU path claims -- Kriging (Simple / Ordinary /
# Universal built by (deg, dim)) and IDW with Euclidean distance, univariate Float64 Cartesian PointSet
# data without missings, CartesianGrid or PointSet targets, point support.  Every other signature keeps
# dispatching to the untouched generic methods.  `fitpredict` itself (models.jl:102-129) is unchanged,
# so `_pointsupport`, `_scale` and the final `georef` run exactly as in the reference and the shim
# receives ALREADY SCALED model, data and domain.
module GeoStatsModelsB200

using GeoStatsModels, GeoStatsFunctions, Meshes, GeoTables, Tables, Distributions, Distances
import GeoStatsModels: fitpredictneigh, fitpredictfull, SimpleKriging, OrdinaryKriging, UniversalKriging, IDW

const LIB = get(ENV, "GSM_B200_LIB", "libgsm_b200.so")

# ---- POD mirrors of include/gsm_b200.h ---------------------------------------------------------
struct GsmVariogram; kind::Int32; _pad::Int32; sill::Float64; nugget::Float64; range::Float64; end
struct GsmModel; kind::Int32; ndrift::Int32; mean::Float64; exps::NTuple{30,Int32}; end
struct GsmTargets
  kind::Int32; dim::Int32
  origin::NTuple{3,Float64}; spacing::NTuple{3,Float64}; dims::NTuple{3,Int64}
  points::Ptr{Float64}; npoints::Int64; first::Int64; count::Int64
end
struct GsmNeighborOpts; maxneighbors::Int32; minneighbors::Int32; radius::Float64; end

const VARIO_KIND = Dict(GaussianVariogram => 0, SphericalVariogram => 1, ExponentialVariogram => 2)

mutable struct Context
  ptr::Ptr{Cvoid}
  function Context(device::Integer=0)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:gsm_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref)
    rc == 0 || error(unsafe_string(ccall((:gsm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = new(ref[])
    finalizer(c -> ccall((:gsm_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
  end
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
context() = (CTX[] === nothing && (CTX[] = Context(parse(Int, get(ENV, "GSM_B200_DEVICE", "0")))); CTX[])

check(ctx, rc) = rc == 0 || throw(ErrorException(unsafe_string(ccall((:gsm_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr))))

# ---- descriptors ---------------------------------------------------------------------------------
supported(γ) = γ isa Union{GaussianVariogram,SphericalVariogram,ExponentialVariogram} && isisotropic(γ)
variogram(γ) = GsmVariogram(VARIO_KIND[typeof(γ).name.wrapper], 0, ustrip(sill(γ)), ustrip(nugget(γ)), ustrip(range(γ)))

exps_tuple(es, dim) = ntuple(i -> (n = (i - 1) ÷ 3 + 1; d = (i - 1) % 3 + 1; n <= length(es) && d <= dim ? Int32(es[n][d]) : Int32(0)), 30)
gsmmodel(m::SimpleKriging, dim) = GsmModel(0, 0, first(m.mean), ntuple(_ -> Int32(0), 30))
gsmmodel(m::OrdinaryKriging, dim) = GsmModel(1, 0, 0.0, ntuple(_ -> Int32(0), 30))
# UniversalKriging(fun, deg, dim) stores closures; the shim re-derives the exponent table of
# universal.jl:68-77 from (deg, dim), which the caller records in `DRIFT_TABLE` when constructing.
const DRIFT_TABLE = IdDict{Any,Vector{Vector{Int}}}()
function UniversalKrigingB200(fun, deg::Int, dim::Int)
  m = UniversalKriging(fun, deg, dim)
  DRIFT_TABLE[m.drifts] = collect(collect.(GeoStatsModels.exponents(deg, dim)))
  m
end
gsmmodel(m::UniversalKriging, dim) = (es = DRIFT_TABLE[m.drifts]; GsmModel(2, length(es), 0.0, exps_tuple(es, dim)))

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

claimed(model, dat, dom, point, distance) =
  point && distance isa Euclidean && domain(dat) isa PointSet && dom isa Union{CartesianGrid,PointSet} &&
  length(Tables.columnnames(Tables.columns(values(dat)))) == 1 &&
  eltype(Tables.getcolumn(Tables.columns(values(dat)), 1)) === Float64

function setsamples!(ctx, dat)
  X = coordmatrix(domain(dat))
  z = Tables.getcolumn(Tables.columns(values(dat)), 1)
  GC.@preserve X z check(ctx, ccall((:gsm_set_samples, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Float64}),
                                     ctx.ptr, X, size(X, 1), size(X, 2), z))
end

# ---- fitpredictneigh (models.jl:131-198) -------------------------------------------------------
function fitpredictneigh(model::Union{SimpleKriging,OrdinaryKriging,UniversalKriging}, dat, dom, point, prob,
                         minneighbors, maxneighbors, neighborhood, distance)
  if !(claimed(model, dat, dom, point, distance) && supported(model.fun) &&
       (neighborhood === nothing || neighborhood isa MetricBall) &&
       (!(model isa UniversalKriging) || haskey(DRIFT_TABLE, model.drifts)) && min(maxneighbors, nrow(dat)) <= 32)
    return invoke(fitpredictneigh, Tuple{GeoStatsModel,Any,Any,Any,Any,Any,Any,Any,Any}, model, dat, dom, point, prob,
                  minneighbors, maxneighbors, neighborhood, distance)
  end
  ctx = context(); setsamples!(ctx, dat)
  M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
  mean = Vector{Float64}(undef, M); vari = prob ? Vector{Float64}(undef, M) : Float64[]; status = Vector{UInt8}(undef, M)
  tg, keep = targets(dom)
  v = Ref(variogram(model.fun)); m = Ref(gsmmodel(model, embeddim(dom)))
  o = Ref(GsmNeighborOpts(maxneighbors, minneighbors, neighborhood === nothing ? 0.0 : ustrip(radius(neighborhood))))
  GC.@preserve keep mean vari status check(ctx, ccall((:gsm_local_krige, LIB), Cint,
    (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{GsmTargets}, Ref{GsmNeighborOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int32}),
    ctx.ptr, v, m, Ref(tg), o, mean, prob ? pointer(vari) : C_NULL, status, C_NULL))
  col = prob ? [s == 1 ? missing : Normal(mean[i], vari[i]) for (i, s) in enumerate(status)] :  # Normal.(mean, var), models.jl:298-299
               [s == 1 ? missing : mean[i] for (i, s) in enumerate(status)]
  (; var => map(identity, col)) |> Tables.materializer(values(dat))
end

# ---- fitpredictfull (models.jl:200-234) --------------------------------------------------------
function fitpredictfull(model::Union{SimpleKriging,OrdinaryKriging,UniversalKriging}, dat, dom, point, prob)
  if !(claimed(model, dat, dom, point, Euclidean()) && supported(model.fun) &&
       (!(model isa UniversalKriging) || haskey(DRIFT_TABLE, model.drifts)))
    return invoke(fitpredictfull, Tuple{GeoStatsModel,Any,Any,Any,Any}, model, dat, dom, point, prob)
  end
  ctx = context(); setsamples!(ctx, dat)
  M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
  mean = Vector{Float64}(undef, M); vari = prob ? Vector{Float64}(undef, M) : Float64[]
  tg, keep = targets(dom)
  fit = Ref{Ptr{Cvoid}}(C_NULL)
  check(ctx, ccall((:gsm_global_krige_fit, LIB), Cint, (Ptr{Cvoid}, Ref{GsmVariogram}, Ref{GsmModel}, Ref{Ptr{Cvoid}}),
                   ctx.ptr, Ref(variogram(model.fun)), Ref(gsmmodel(model, embeddim(dom))), fit))
  try
    GC.@preserve keep mean vari check(ctx, ccall((:gsm_global_krige_predict, LIB), Cint,
      (Ptr{Cvoid}, Ref{GsmTargets}, Ptr{Float64}, Ptr{Float64}), fit[], Ref(tg), mean, prob ? pointer(vari) : C_NULL))
  finally
    ccall((:gsm_global_krige_free, LIB), Cint, (Ptr{Cvoid},), fit[])
  end
  col = prob ? Normal.(mean, vari) : mean
  (; var => col) |> Tables.materializer(values(dat))
end

# ---- IDW (idw.jl:43-89) ---------------------------------------------------------------------------
function idwcall(model::IDW, dat, dom, opts)
  ctx = context(); setsamples!(ctx, dat)
  M = nelements(dom); var = first(Tables.columnnames(Tables.columns(values(dat))))
  mean = Vector{Float64}(undef, M); status = Vector{UInt8}(undef, M)
  tg, keep = targets(dom)
  GC.@preserve keep mean status check(ctx, ccall((:gsm_idw, LIB), Cint,
    (Ptr{Cvoid}, Float64, Ref{GsmTargets}, Ptr{GsmNeighborOpts}, Ptr{Float64}, Ptr{UInt8}),
    ctx.ptr, Float64(model.exponent), Ref(tg), opts === nothing ? C_NULL : Ref(opts), mean, status))
  (; var => map(identity, [s == 1 ? missing : mean[i] for (i, s) in enumerate(status)])) |> Tables.materializer(values(dat))
end
function fitpredictfull(model::IDW{<:Real,<:Euclidean}, dat, dom, point, prob)
  (claimed(model, dat, dom, point, model.distance) && !prob) ||
    return invoke(fitpredictfull, Tuple{GeoStatsModel,Any,Any,Any,Any}, model, dat, dom, point, prob)
  idwcall(model, dat, dom, nothing)
end
function fitpredictneigh(model::IDW{<:Real,<:Euclidean}, dat, dom, point, prob, minneighbors, maxneighbors, neighborhood, distance)
  (claimed(model, dat, dom, point, distance) && !prob && min(maxneighbors, nrow(dat)) <= 64 &&
   (neighborhood === nothing || neighborhood isa MetricBall)) ||
    return invoke(fitpredictneigh, Tuple{GeoStatsModel,Any,Any,Any,Any,Any,Any,Any,Any}, model, dat, dom, point, prob,
                  minneighbors, maxneighbors, neighborhood, distance)
  idwcall(model, dat, dom, GsmNeighborOpts(maxneighbors, minneighbors, neighborhood === nothing ? 0.0 : ustrip(radius(neighborhood))))
end

end # module
