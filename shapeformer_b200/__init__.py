"""shapeformer_b200 — B200-native (sm_100a) implementation of ShapeFormer's data-parallel hot path:
the KV-cached autoregressive (pos, val) sampler and the VQDIF implicit decoder, behind the reference's model API.

Python here is host glue only (device memory, streams, RNG, torch.distributed); the arithmetic lives in
lib/libsfb200.so (hand-written CUDA, C-ABI in include/sfb200.h).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
