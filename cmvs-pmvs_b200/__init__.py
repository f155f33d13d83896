"""B200-native PMVS patch-optimisation path: CUDA kernels behind a C ABI (csrc/, include/pmvs_b200.h),
a ctypes binding for tests and benchmarks (binding.py) and deterministic synthetic scenes (synth.py).

The directory name carries a hyphen (repository convention), so it is imported through
`__graft_entry__.load_package()` under the module name `cmvs_pmvs_b200`.
"""
from . import synth  # noqa: F401
from .binding import LIB_PATH, SYMBOLS, PmvsB200, PmvsError, load_library  # noqa: F401
