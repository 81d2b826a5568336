"""repsurf_b200 — B200-native (sm_100a) implementation of the RepSurf-U hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C-ABI (include/repsurf_b200.h) -> librepsurf_b200.so
  _native.py       ctypes loader (fails loudly when the library is missing: there is NO CPU fallback)
  cls/pointops.py  mirror of the reference's classification/modules/pointops/functions/pointops.py
  cls/modules.py   UmbrellaSurfaceConstructor / SurfaceAbstractionCD (dense layout)
  seg/pointops.py  mirror of segmentation/modules/pointops/functions/pointops.py
  seg/modules.py   UmbrellaSurfaceConstructor / SurfaceAbstractionCD / SurfaceFeaturePropagationCD (packed layout)
"""
__version__ = "0.1.0"
