"""riders_b200 - B200-native (sm_100a) implementation of the RIDERS radar-to-dense-depth hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/riders_b200.h) -> libriders_b200.so
  _lib.py, ops.py  ctypes binding and tensor-level wrappers (torch supplies memory + streams only)
  rcnet/           host mirror of the reference's RCNet/{networks,linear_attention,rcnet_model,rcnet_main}.py
  modules/         host mirror of the reference's modules/{estimator,interpolator}.py
  dropin.py        installs the mirrors under the reference's module names
  synth.py         seeded synthetic inputs of SURVEY 8(d)
  dist.py          frame sharding + the final metric all-reduce
"""
__version__ = "0.1.0"
