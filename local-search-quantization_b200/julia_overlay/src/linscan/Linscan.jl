# Overlay for src/linscan/Linscan.jl.  The library exports the reference's two C symbols with
# byte-identical signatures (linscan_aqd_query, linscan_aqd.cpp:107-113; linscan_aqd_query_extra_byte,
# linscan_aqd_pairwise_byte.cpp:97-104), so the ONLY change against the reference file is the library
# path in the two ccalls.  (Alternatively leave Linscan.jl untouched and symlink
# src/linscan/cpp/linscan_aqd.so and linscan_aqd_pairwise_byte.so to liblsq_b200.so.)
using Distances
include("../lsq_b200.jl")

# Linear scan using PQ codebooks no rotation
function linscan_pq(
  B::Matrix{UInt8},           # m-by-n. The database, encoded
  X::Matrix{Cfloat},         # d-by-nq. The queries.
  C::Vector{Matrix{Cfloat}}, # The cluster centers
  b:: Int,                    # Number of bits per code -- log2(h) * m
  k:: Int = 10000)            # Number of knn results to return

  m, n  = size( B )
  d, nq = size( X )

  dists = zeros( Cfloat, k, nq )
  res   = zeros(  Cuint, k, nq )

  ccall((:linscan_aqd_query, LSQ_B200_LIB), Void,
    (Ptr{Cfloat}, Ptr{Cuint}, Ptr{Cuchar}, Ptr{Cfloat},
    Ptr{Cfloat}, Cint, Cuint, Cint, Cint, Cint, Cint, Cint),
    dists, res, B, cat(3,C...), X, Cint(n), Cuint(nq),
    Cint(b), Cint(k), Cint(m), Cint(d), Cint(d/m) )

  return dists, (res.+=1)
end

# Linear scan using OPQ
function linscan_opq(
  B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int,
  R::Matrix{Cfloat}, k::Int = 10000)
  RX = R'*X
  return linscan_pq( B, RX, C, b, k )
end

# Linear scan using LSQ, with dbnorms not encoded.
function linscan_lsq(
  B::Matrix{UInt8},           # m-by-n. The database, encoded
  X::Matrix{Cfloat},         # d-by-nq. The queries.
  C::Vector{Matrix{Cfloat}}, # The cluster centers
  dbnorms::Vector{Cfloat},   # n-long. Database norms
  R::Matrix{Cfloat},         # Rotation matrix
  k::Int = 10000)             # Number of knn results to return

  RX = R' * X;

  m, n  = size( B );
  d, nq = size( RX );
  _, h  = size( C[1] );

  dists = zeros( Cfloat, k, nq );
  res   = zeros(  Cint,  k, nq  );

  ccall((:linscan_aqd_query_extra_byte, LSQ_B200_LIB), Void,
    (Ptr{Cfloat}, Ptr{Cint},
    Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
    Cuint, Cint, Cint, Cint, Cint, Cint),
    dists, res,
    B, RX, hcat(C...), dbnorms,
    Cint(nq), Cint(n), Cint(m), Cint(h), Cint(d), Cint(k) );

  return dists, res
end

# eval_recall (Linscan.jl:76-117): the rank search over the k-by-nq id matrix runs on the GPU
# (lsq_eval_recall); the printed r@N lines and the returned recall_at_i vector are the reference's.
function eval_recall{T <: Integer}(
  ids_gnd::Vector{T},
  ids_predicted::Matrix{T},
  k::Integer)

  nquery = size( ids_predicted, 2 );
  assert( nquery == length( ids_gnd) );
  ld = size( ids_predicted, 1 );

  recall_at_i = zeros( Cdouble, k );
  lsq_check( ccall((:lsq_eval_recall, LSQ_B200_LIB), Cint,
    (Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Ptr{Cdouble}),
    convert(Vector{Int32}, ids_gnd), convert(Matrix{Int32}, ids_predicted), nquery, ld, k, recall_at_i) )

  for i = [1 2 5 10 20 50 100 200 500 1000 2000 5000 10000]
    if i <= k
      println("r@$(i) = $(recall_at_i[i] * 100)");
    end
  end
  return recall_at_i
end
