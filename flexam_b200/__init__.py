"""flexam_b200 — B200-native (sm_100a) implementation of the FlexAM Wan2.2-Fun-5B denoising step.

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI (``include/flexam_b200.h``); ``lib``/``ops`` the
ctypes binding; ``model`` the host-side mirror of the reference transformer interface.
"""
__all__ = ["lib", "ops"]
