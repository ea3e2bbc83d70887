# Overlay for src/encodings/encode_chain.jl — encoding_viterbi (reference encode_chain.jl:95-127) as one
# ccall; encode_viterbi! (1-92) is gone: unaries, the (m-1) pair tables, the forward min-sum pass and the
# backward trace all run on the GPU.  Same name, same signature, same return type (m-by-n Matrix{Int16}).
# Append `include("encode_chain_b200.jl")` at the end of src/encodings/encode_chain.jl.
include("../lsq_b200.jl")

function encoding_viterbi(
  X::Matrix{Float32},         # d-by-n matrix. Data to encode
  C::Vector{Matrix{Float32}}, # m-long vector with d-by-h codebooks
  V::Bool=false)              # whether to print progress

  d, n = size( X )
  m    = length( C )
  _, h = size( C[1] )
  B    = Matrix{Int16}( m, n )

  lsq_check( ccall((:lsq_encoding_viterbi, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Cfloat}, Cint, Cint, Ptr{Int16}, Cint),
    X, d, n, lsq_pack_codebooks( C ), m, h, B, V) )

  return B
end
