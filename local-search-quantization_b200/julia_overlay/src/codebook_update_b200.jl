# Overlay for src/codebook_update.jl — update_codebooks (reference codebook_update.jl:52-86) as one
# ccall.  The chain / generic variants of the reference file (88-169, ChainQ initialiser) are outside
# the replaced path: keep the reference's definitions for those by including the original file first.
include("utils.jl")
include("lsq_b200.jl")

function update_codebooks(
  X::Matrix{Float32}, # d-by-n matrix to update codebooks on.
  B::Matrix{Int16},   # m-by-n matrix. X encoded.
  h::Integer,         # number of entries per codebook.
  V::Bool=false,      # whether to print progress
  codebook_upd_method::AbstractString="lsqr")   # lsqr or lsmr (both map to the same min-norm LS solve)

  if !(codebook_upd_method in ["lsmr", "lsqr"]); error("Codebook update method unknown"); end

  d, n = size( X )
  m, _ = size( B )
  K    = Array{Cfloat,3}( d, h, m )

  lsq_check( ccall((:lsq_update_codebooks, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Int16}, Cint, Cint, Ptr{Cfloat}, Cstring, Cint),
    X, d, n, B, m, h, K, codebook_upd_method, V) )

  # K2vec (utils.jl:72-87): back to an m-long vector of d-by-h codebooks
  new_C = Vector{Matrix{Float32}}( m )
  for i = 1:m
    new_C[i] = K[:, :, i]
  end
  return new_C
end
