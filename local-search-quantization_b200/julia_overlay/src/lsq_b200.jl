# lsq_b200.jl — shared ccall plumbing of the B200 overlay.  Argument marshalling only: every byte of
# compute happens inside liblsq_b200.so (C ABI: include/lsq_b200.h).  Written in the reference's own
# Julia 0.6 dialect; all entry points return Cint, so no Void/Cvoid appears and the same file also
# parses on Julia >= 0.7 once `cat(3, C...)` is spelled `cat(C...; dims=3)`.

if !isdefined(:LSQ_B200_LIB)
  # cwd-relative like the reference's own .so/.ptx paths (Linscan.jl:19,63; encode_icm_cuda.jl:64)
  const LSQ_B200_LIB = get(ENV, "LSQ_B200_LIB", "local-search-quantization_b200/liblsq_b200.so")
  # schedule state: the reference draws perturbations from Julia's global RNG on every call; the
  # library uses Philox(seed, ils_iter, global vector index), so successive calls advance ils_iter.
  const LSQ_B200_SEED    = UInt64[ parse(UInt64, get(ENV, "LSQ_B200_SEED", "0")) ]
  const LSQ_B200_COUNTER = UInt32[ 0 ]
end

function lsq_check(rc::Integer)
  if rc != 0
    msg = unsafe_string( ccall((:lsq_last_error, LSQ_B200_LIB), Cstring, ()) )
    error("lsq_b200 error $rc: $msg")
  end
  return nothing
end

# Replaces CudaUtilsModule.init / finit (cudaUtilsModule.jl:37-43)
lsq_init(gpuid::Integer=0) = lsq_check( ccall((:lsq_init, LSQ_B200_LIB), Cint, (Cint,), gpuid) )
lsq_finit()                = lsq_check( ccall((:lsq_finalize, LSQ_B200_LIB), Cint, ()) )
# Several GPUs of the box (the reference is hard-wired to device 0, encode_icm_cuda.jl:59-64): after this call
# encoding_icm / encode_icm_cuda / update_codebooks / train_lsq / linscan_* split their input over the listed
# devices inside the library; results do not depend on the device set.  An empty list binds every visible GPU.
# Without any Julia change the same is selected by the environment: LSQ_B200_DEVICES=all (or 0,1,2,...).
function lsq_init_devices(gpuids::Vector{Int}=Int[])
  devs = convert(Vector{Cint}, gpuids)
  lsq_check( ccall((:lsq_init_devices, LSQ_B200_LIB), Cint, (Ptr{Cint}, Cint), devs, length(devs)) )
  return Int( ccall((:lsq_num_bound_devices, LSQ_B200_LIB), Cint, ()) )
end

# d-by-h-by-m stack of the codebooks: exactly what Linscan.jl:22 already builds for the PQ scan
lsq_pack_codebooks{T <: AbstractFloat}(C::Vector{Matrix{T}}) = convert(Array{Cfloat,3}, cat(3, C...))
