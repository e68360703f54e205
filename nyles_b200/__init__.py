"""nyles_b200: B200-native implementation of the Nyles LES time step.

Host layer = the reference's Python API (Nyles(param), model_les.LES, timescheme, and the
operator modules vorticity / vortex_force / kinenergy / bernoulli / tracer / projection /
mgfordriver) on CUDA tensors; compute = libnyles_b200.so (hand-written sm_100a kernels, C ABI
in include/nyles_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"
