"""hikari-b200: Blackwell-native back end for bevy-hikari's per-frame path-tracing passes.

The product is libhikari_b200.so (CUDA sm_100a + C++ host mirror) behind the C ABI of include/hikari_b200.h; this
package is the ctypes face of it.  Importing the package does not load the library; the first call does, and fails
loudly if the library has not been built — there is no CPU / PyTorch fallback."""
from . import layout  # noqa: F401

__all__ = ["layout", "camera", "scenes", "plugin"]
