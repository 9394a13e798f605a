"""miro-gpu: B200-native ray-casting core for the Miro ray tracer.

The product is the native library libmiro_gpu.so (CUDA kernels for sm_100a + C++ host layer);
this package is the ctypes plumbing the tests and the benchmark use to reach its C ABI.
"""
from . import capi  # noqa: F401
from .scene import MiroScene, MiroError, RAY_DTYPE, RAY32_DTYPE, HIT_DTYPE, make_rays, pack_rays  # noqa: F401
