# Overlay for src/encodings/encode_icm_cuda.jl — no CUDAdrv / CuArrays / PTX module any more.
# encode_icm_cuda (reference encode_icm_cuda.jl:253-296) and encode_icm_cuda_single (22-234) become one
# ccall: every ILS iteration runs device-side in a single kernel launch per memory chunk.
include("../utils.jl")
include("../read/read_datasets.jl")
include("../initializations.jl")
include("../lsq_b200.jl")

"Encodes a database with ILS on the GPU"
function encode_icm_cuda(
  RX::Matrix{Float32},         # in. The data to encode
  B::Matrix{Int16},            # in. Initial list of codes
  C::Vector{Matrix{Float32}},  # in. Codebooks
  ilsiters::Vector{Int64},     # in. ILS iterations to record Bs and obj function. Its max is the total number of iterations
  icmiter::Integer,            # in. Number of ICM iterations
  npert::Integer,              # in. Number of entries to perturb
  randord::Bool,               # in. Whether to randomize the order in which nodes are visited in ILS
  nsplits::Integer=2,          # in. Number of splits of the data (bounds device memory only)
  V::Bool=false)

  d, n = size( RX )
  m    = length( C )
  _, h = size( C[1] )
  nr   = length( ilsiters )

  Cf    = lsq_pack_codebooks( C )
  Bsbuf = Array{Int16,3}( m, n, nr )
  objs  = zeros( Float32, nr )

  lsq_check( ccall((:lsq_encode_icm_cuda, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Int16}, Ptr{Cfloat}, Cint, Cint,
     Ptr{Int64}, Cint, Cint, Cint, Cint, Cint, UInt64, UInt64,
     Ptr{Int16}, Ptr{Cfloat}, Cint),
    RX, d, n, B, Cf, m, h,
    ilsiters, nr, icmiter, npert, randord, nsplits, LSQ_B200_SEED[1], 0,
    Bsbuf, objs, V) )

  Bs = Vector{Matrix{Int16}}( nr )
  for i = 1:nr
    Bs[i] = Bsbuf[:, :, i]
  end
  return Bs, objs
end

encode_icm_cuda_single(RX, B, C, ilsiters, icmiter, npert, randord, V=false) =
  encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, 1, V)
