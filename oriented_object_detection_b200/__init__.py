"""geomap_b200: B200-native (sm_100a) tiled-detection hot path of Abolfazlmsl/Oriented-Object-Detection.

Layout: ``csrc/`` CUDA kernels + C ABI (``include/geomap_b200.h``), ``_lib`` ctypes binding,
``ops`` device-level API on torch tensors, ``detect`` the host-side mirror of the reference's
``Detect_OBB.py`` functions, ``sharding`` the multi-GPU row-band path, ``synth`` synthetic maps
and OBB sets.  Importing the package loads the CUDA library and raises if it was not built.
"""
from . import _lib  # noqa: F401  (raises ImportError when libgeomap_b200.so is missing)

__version__ = "0.1.0"
