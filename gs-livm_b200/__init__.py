"""gs-livm_b200: B200-native (sm_100a) differentiable Gaussian-splatting rasterizer, drop-in for GS-LIVM's
`src/cuda_rasterizer` hot path. Imported as `gs_livm_b200` (see ../gs_livm_b200.py).

Layout
  csrc/          hand-written CUDA kernels, the C ABI (include/gsr.h) and the C++ operator shim
  _lib.py        ctypes binding of libgsr_b200.so (fails loudly when the library is missing)
  rasterizer.py  Python mirror of the reference's LibTorch operator surface
                 (GaussianRasterizationSettings / GaussianRasterizer / _RasterizeGaussians)
  engine.py      allocation-free forward / backward driver over the C ABI (persistent arenas, sync-free forward,
                 CUDA-graph capture, flat gradient buffer)
  scene.py       synthetic NTU-VIRAL-shaped scenes and GS-LIVM camera conventions
  window.py      view-sharded keyframe window (multi-GPU): the factored gradient exchange over peer memory
                 (FactoredExchange), the dense exchange of round 1 as fallback (GradExchange)
  train.py       fused image loss (L1 + DSSIM) and fused Adam over flat buffers
  vgpr.py        VGPR + variance-centred Gaussian initialisation of a LiDAR sweep on the GPU
  mapping.py     the mapping thread's loop: insertion + optimisation iterations over the keyframe window
"""
__version__ = "0.1.0"
