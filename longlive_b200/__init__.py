"""longlive_b200 — B200-native (sm_100a) implementation of LongLive's frame-level autoregressive
denoising hot path (CausalWanModel step + rolling KV cache) behind the reference's own pipeline
entry points.  Host code is Python/PyTorch (device memory, streams); all math runs in the
hand-written CUDA kernels of libllb200.so through the C ABI in include/llb200.h.
"""
__version__ = "0.1.0"
