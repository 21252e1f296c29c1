"""lol_b200 -- B200 (sm_100a) back end for Lol's cyclotomic `Tensor` hot path.

The product is the C-ABI shared library `libctensor_b200.so` (sources in
`lol_b200/csrc`, interface in `include/lol_b200.h`).  This package only holds

  * `build`   -- compiles the library in-tree with nvcc for sm_100a,
  * `capi`    -- a ctypes binding of the C ABI (no torch types cross it),
  * `tensor`  -- a host-side mirror of the reference's `Tensor` class methods
                 for this path (same names and argument meaning as
                 lol/Crypto/Lol/Cyclotomic/Tensor.hs:86-193) over torch CUDA
                 tensors, which are used for device memory and streams only,
  * `extension` -- the same for the two-index methods (twace / embed / coeffs over O_m'/O_m),
  * `symmshe` -- the same for the SymmSHE steps between the CRTs, `shard` -- batch partitioning over GPUs.

There is no CPU implementation anywhere in this package: importing works
without a GPU, calling an operator without one raises.
"""
from .build import build_library, library_path  # noqa: F401

__all__ = ["build_library", "library_path"]
