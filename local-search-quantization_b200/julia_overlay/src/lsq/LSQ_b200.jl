# Overlay for src/lsq/LSQ.jl — train_lsq (reference LSQ.jl:10-88) as ONE ccall: the whole alternation
# update_codebooks <-> encoding_icm runs with X, the codes, the unaries and the pair tables resident
# on the GPU (the reference re-sends X to its workers on every encoding_icm call, encode_icm.jl:165-172),
# followed by the norm codebook of LSQ.jl:68-84.  Same name, same signature, same return tuple.
# Include AFTER the reference's src/lsq/LSQ.jl so that this method replaces the original one.
include("../lsq_b200.jl")

function train_lsq{T <: AbstractFloat}(
  X::Matrix{T},         # d-by-n matrix of data points to train on.
  m::Integer,           # number of codebooks
  h::Integer,           # number of entries per codebook
  R::Matrix{T},         # init rotation
  B::Matrix{Int16},     # init codes
  C::Vector{Matrix{T}}, # init codebooks (overwritten before use in the reference too, LSQ.jl:34)
  niter::Integer,       # number of optimization iterations
  ilsiter::Integer,     # number of ILS iterations to use during encoding
  icmiter::Integer,     # number of iterations in local search
  randord::Bool,        # whether to use random order
  npert::Integer,       # The number of codes to perturb
  V::Bool=false)        # whether to print progress

  d, n    = size( X )
  Xf      = convert( Matrix{Cfloat}, X )
  Rf      = convert( Matrix{Cfloat}, R )
  Bout    = copy( B )                      # in: init codes, out: final codes
  K       = Array{Cfloat,3}( d, h, m )
  cbnorms = zeros( Cfloat, h )
  B_norms = zeros( Int16, n )
  obj     = zeros( Cfloat, max(niter, 1) )

  lsq_check( ccall((:lsq_train_lsq, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Cint, Cint, Ptr{Cfloat}, Ptr{Int16}, Ptr{Cfloat},
     Cint, Cint, Cint, Cint, Cint, UInt64, Ptr{Cfloat}, Ptr{Int16}, Ptr{Cfloat}, Cint),
    Xf, d, n, m, h, Rf, Bout, K,
    niter, ilsiter, icmiter, randord, npert, LSQ_B200_SEED[1], cbnorms, B_norms, obj, V) )
  # the call consumed (niter+1)*ilsiter ILS iterations of the schedule
  LSQ_B200_COUNTER[1] += UInt32( (niter + 1) * ilsiter )

  new_C = Vector{Matrix{Float32}}( m )     # K2vec (utils.jl:72-87)
  for i = 1:m
    new_C[i] = K[:, :, i]
  end

  # same shapes as the reference's return values: cbnorms 1-by-h (dbnormsq.centers), B_norms 1-by-n
  return new_C, Bout, reshape( cbnorms, 1, h ), reshape( B_norms, 1, n ), obj[1:niter]
end
