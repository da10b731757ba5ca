"""hdn_b200 -- B200-native (sm_100a) implementation of HDN's per-frame homography hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI declared in include/hdn_b200.h
  build.py     nvcc recipe -> hdn_b200/libhdn_b200.so (in-tree)
  _lib.py      ctypes binding (raises if the library is missing: there is no CPU fallback)
  ops.py       drop-ins with the reference's operator names and signatures
  engine.py    batched corr+warp+DLT chain (CUDA-graph replay, pinned-host pipeline)
  shard.py     one-process-per-GPU sharding of independent pairs / sequences
  compat/      `hdn.*` / `homo_estimator.*` packages mirroring the reference API
"""
__version__ = "0.1.0"
