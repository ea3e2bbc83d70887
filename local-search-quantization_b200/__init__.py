"""local-search-quantization_b200 — B200-native LSQ hot path behind the reference's API.

The product is `liblsq_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/lsq_b200.h).
This package is the host-side mirror of the reference's Julia interface for that path — same function
names, argument meaning and error behaviour as src/encodings/encode_icm.jl, encode_icm_cuda.jl,
src/codebook_update.jl, src/linscan/Linscan.jl and the utils they use — implemented as thin ctypes
calls.  Julia itself is not installable in this image; `julia_overlay/` holds the equivalent `ccall`
shim.

There is NO CPU fallback: importing works anywhere (so the symbol table can be checked), but every
compute call raises LsqError when the library or a GPU is missing.

Array conventions (numpy, C-order) are byte-identical to the Julia column-major arrays:
  X: (n, d) float32  == Julia d-by-n;   B: (n, m) int16 1-based == Julia m-by-n Matrix{Int16};
  C: (m, h, d) float32 == cat(3, C...) of m d-by-h codebooks.
Because the package directory name contains '-', import it with importlib or via the `lsq_b200`
alias module at the repository root.
"""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401
