"""cosmopp_b200 -- B200-native generator of the pixel-space CMB covariance (Cosmo++ CMatrixGenerator -> CMatrix).

The product is the CUDA library lib/libcosmopp_b200.so behind the C ABI of include/cmg.h and the C++
drop-in classes of include/c_matrix.hpp / include/c_matrix_generator.hpp.  This Python package is the
test/bench harness over that ABI (ctypes) and uses PyTorch only for device memory, streams and
torch.distributed plumbing.  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .capi import CmgError, Context, library, library_path, build_library  # noqa: F401
from .generator import CMatrix, CMatrixGenerator  # noqa: F401
from .partition import column_partition, tqu_shard_sizes  # noqa: F401
