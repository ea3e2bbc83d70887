# linscan_b200.jl — search-side bindings of liblsq_b200.so.
#
# The reference needs NO Julia change for the scan: the library exports the two C symbols its Linscan.jl
# already binds (linscan_aqd_query / linscan_aqd_query_extra_byte) with the same signatures, so pointing
# src/linscan/cpp/linscan_aqd.so and linscan_aqd_pairwise_byte.so at liblsq_b200.so (symlink) is enough.
# This file is the alternative for callers that want an error channel: it rebinds the three public entry
# points of the scan to the status-returning twins (lsq_linscan_pq / lsq_linscan_lsq) and moves the recall
# bookkeeping to the GPU.  Include it after the reference's Linscan.jl; method signatures are unchanged.
include("../lsq_b200.jl")

# one scan call: `kind` picks the LUT flavour; outputs are allocated here and fully overwritten by the library
function b200_scan(kind::Symbol, codes::Matrix{UInt8}, queries::Matrix{Cfloat}, books::Array{Cfloat},
                   norms, nbits::Int, knn::Int)
  m, n   = size( codes )
  d, nq  = size( queries )
  dists  = Matrix{Cfloat}( knn, nq )
  if kind == :pq
    ids  = Matrix{Cuint}( knn, nq )
    lsq_check( ccall((:lsq_linscan_pq, LSQ_B200_LIB), Cint,
      (Ptr{Cfloat}, Ptr{Cuint}, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cuint, Cint, Cint, Cint, Cint, Cint),
      dists, ids, codes, books, queries, n, nq, nbits, knn, m, d, div(d, m)) )
    return dists, ids .+ one(Cuint)      # C side is 0-based for PQ (linscan_aqd.cpp:88)
  else
    ids  = Matrix{Cint}( knn, nq )
    h    = div( size(books, 2), m )
    lsq_check( ccall((:lsq_linscan_lsq, LSQ_B200_LIB), Cint,
      (Ptr{Cfloat}, Ptr{Cint}, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Cint, Cint, Cint),
      dists, ids, codes, queries, books, norms, nq, n, m, h, d, knn) )
    return dists, ids                    # already 1-based (linscan_aqd_pairwise_byte.cpp:75)
  end
end

linscan_pq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int, k::Int=10000) =
  b200_scan( :pq, B, X, cat(3, C...), C_NULL, b, k )

linscan_opq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int, R::Matrix{Cfloat}, k::Int=10000) =
  b200_scan( :pq, B, R' * X, cat(3, C...), C_NULL, b, k )

linscan_lsq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, dbnorms::Vector{Cfloat},
            R::Matrix{Cfloat}, k::Int=10000) =
  b200_scan( :lsq, B, R' * X, hcat(C...), dbnorms, 0, k )

# recall@1..k from the ranked id lists (one column per query); the rank search runs on the GPU
function eval_recall{T <: Integer}(ids_gnd::Vector{T}, ids_predicted::Matrix{T}, k::Integer)
  ld, nquery = size( ids_predicted )
  nquery == length( ids_gnd ) || error("one ground-truth id per query expected")
  curve = zeros( Cdouble, k )
  lsq_check( ccall((:lsq_eval_recall, LSQ_B200_LIB), Cint,
    (Ptr{Int32}, Ptr{Int32}, Cint, Cint, Cint, Ptr{Cdouble}),
    convert(Vector{Int32}, ids_gnd), convert(Matrix{Int32}, ids_predicted), nquery, ld, k, curve) )
  for i in (1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000)
    i <= k && println("r@$(i) = $(100 * curve[i])")
  end
  return curve
end
