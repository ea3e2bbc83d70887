# Overlay for src/encodings/encode_icm.jl — same name, same signature, same return type.
# encoding_icm (reference encode_icm.jl:131-189) becomes one ccall; encode_icm_fully! (4-127) is gone:
# perturbation, the ICM sweeps, both veccost passes and the keep-if-better step all run on the GPU.
include("../lsq_b200.jl")

# Encode a full dataset: ONE ILS iteration
function encoding_icm{T <: AbstractFloat}(
  X::Matrix{T},         # d-by-n matrix. Data to encode
  oldB::Matrix{Int16},  # m-by-n matrix. Previous encoding
  C::Vector{Matrix{T}}, # m-long vector with d-by-h codebooks
  niter::Integer,       # number of ICM iterations
  randord::Bool,        # whether to use random order
  npert::Integer,       # the number of codes to perturb
  V::Bool=false)        # whether to print progress

  d, n = size( X )
  m    = length( C )
  _, h = size( C[1] )

  Xf = convert( Matrix{Cfloat}, X )
  Cf = lsq_pack_codebooks( C )
  B  = Matrix{Int16}( m, n )

  ils_iter = LSQ_B200_COUNTER[1]
  LSQ_B200_COUNTER[1] += 1

  lsq_check( ccall((:lsq_encoding_icm, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Int16}, Ptr{Int16}, Ptr{Cfloat},
     Cint, Cint, Cint, Cint, Cint, UInt64, UInt32, UInt64, Cint),
    Xf, d, n, oldB, B, Cf,
    m, h, niter, randord, npert, LSQ_B200_SEED[1], ils_iter, 0, V) )

  return B
end
