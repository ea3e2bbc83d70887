# Overlay for the helpers of src/utils.jl that sit on the hot path: same names, same signatures, same
# return types as the reference, each one ccall into liblsq_b200.so.  Optional — the reference's Julia
# definitions keep working unchanged on top of the overlaid encoders; include this file AFTER
# src/utils.jl (e.g. at its end: include("utils_b200.jl")) to move them to the GPU too.
# Only the Float32 methods are replaced (the library computes in fp32 like the reference GPU path).
include("lsq_b200.jl")

# reconstruct (utils.jl:203-223) -> d-by-n matrix
function reconstruct(B::Matrix{Int16}, C::Vector{Matrix{Float32}})
  m, n = size( B )
  d, h = size( C[1] )
  CB   = Matrix{Cfloat}( d, n )
  lsq_check( ccall((:lsq_reconstruct, LSQ_B200_LIB), Cint,
    (Ptr{Int16}, Int64, Ptr{Cfloat}, Cint, Cint, Cint, Ptr{Cfloat}),
    B, n, lsq_pack_codebooks( C ), d, m, h, CB) )
  return CB
end

# veccost (utils.jl:225-254) -> n-long vector of per-vector squared errors
function veccost(X::Matrix{Float32}, B::Matrix{Int16}, C::Vector{Matrix{Float32}})
  d, n = size( X )
  m, _ = size( B )
  _, h = size( C[1] )
  cost = zeros( Cfloat, n )
  lsq_check( ccall((:lsq_veccost, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Int16}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}),
    X, d, n, B, lsq_pack_codebooks( C ), m, h, cost) )
  return cost
end

# qerror (utils.jl:257-285) -> mean squared error
function qerror(X::Matrix{Float32}, B::Matrix{Int16}, C::Vector{Matrix{Float32}})
  d, n = size( X )
  m, _ = size( B )
  _, h = size( C[1] )
  out  = Cfloat[ 0 ]
  lsq_check( ccall((:lsq_qerror, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Int16}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}),
    X, d, n, B, lsq_pack_codebooks( C ), m, h, out) )
  return out[1]
end

# quantize_norms (utils.jl:6-31) -> n-long Vector{Int16}, 1-based index into cbnorms
function quantize_norms(B::Matrix{Int16}, C::Vector{Matrix{Float32}}, cbnorms::Vector{Float32})
  m, n = size( B )
  d, h = size( C[1] )
  dbnormsB = Vector{Int16}( n )
  lsq_check( ccall((:lsq_quantize_norms, LSQ_B200_LIB), Cint,
    (Ptr{Int16}, Int64, Ptr{Cfloat}, Cint, Cint, Cint, Ptr{Cfloat}, Cint, Ptr{Int16}),
    B, n, lsq_pack_codebooks( C ), d, m, h, cbnorms, length( cbnorms ), dbnormsB) )
  return dbnormsB
end

# get_unaries (utils.jl:94-122) -> m-long vector of h-by-n matrices
function get_unaries(X::Matrix{Float32}, C::Vector{Matrix{Float32}}, V::Bool=false)
  d, n = size( X )
  m    = length( C )
  _, h = size( C[1] )
  U    = Array{Cfloat,3}( h, n, m )
  lsq_check( ccall((:lsq_get_unaries, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Int64, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}),
    X, d, n, lsq_pack_codebooks( C ), m, h, U) )
  return Matrix{Float32}[ U[:, :, i] for i = 1:m ]
end

# get_binaries (utils.jl:125-144) -> (binaries, cbi): ncbi h-by-h tables and the 2-by-ncbi index of pairs
function get_binaries(C::Vector{Matrix{Float32}})
  m    = length( C )
  d, h = size( C[1] )
  ncbi = div( m * (m - 1), 2 )
  G    = Array{Cfloat,3}( h, h, max(ncbi, 1) )
  cbi  = zeros( Int32, 2, max(ncbi, 1) )
  lsq_check( ccall((:lsq_get_binaries, LSQ_B200_LIB), Cint,
    (Ptr{Cfloat}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Int32}),
    lsq_pack_codebooks( C ), d, m, h, G, cbi) )
  binaries = Matrix{Float32}[ G[:, :, i] for i = 1:ncbi ]
  return binaries, cbi[:, 1:ncbi]
end

# splitarray (utils.jl:152-177) stays in Julia: it is index arithmetic (lsq_splitarray is the same rule,
# exported for non-Julia hosts).
