"""gemmgan_b200 — B200-native (sm_100a) implementation of GeMM-GAN's WGAN-GP training step.

PyTorch is used for device memory, streams and torch.distributed; the compute is in
libgemmgan_sm100a.so (hand-written CUDA, C ABI in include/gemmgan.h).
"""
__version__ = "0.1.0"
