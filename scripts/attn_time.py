"""Development aid: decode-attention kernel time at the benchmarked shape (64 rows in groups of 4, 16 heads, L_cond 256), CUDA
events around back-to-back launches over 4 distinct caches (1.6 GB > L2).   python scripts/attn_time.py [pos ...]"""
import sys
sys.path.insert(0, '.')
import torch
from shapeformer_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
B, H, G, shared, max_len = 64, 16, 4, 256, 769
d = H * 64
caches = [(torch.randn(B, H, max_len, 64, device=dev), torch.randn(B, H, max_len, 64, device=dev)) for _ in range(4)]
qkv = torch.randn(B, 3 * d, device=dev)
out = torch.empty(B, d, device=dev); part = torch.empty(B * H * 3 * 66, device=dev); cnt = torch.zeros(B * H, dtype=torch.int32, device=dev)
def run(kc, vc, pos):
    _lib.check(lib.sfb200_attn_decode_grouped(_lib.ptr(qkv), _lib.ptr(kc), _lib.ptr(vc), _lib.ptr(out), _lib.ptr(part), _lib.ptr(cnt),
                                              B, H, max_len, pos, G, shared, _lib.stream_ptr()), "attn")
for pos in [int(a) for a in sys.argv[1:]] or [256, 511, 767]:
    for kc, vc in caches: run(kc, vc, pos)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                kc, vc = caches[i % 4]
                run(kc, vc, pos)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    byts = ((B // G) * shared + B * (pos - shared)) * 2 * 1024 * 4 + B * 4 * 1024 * 4
    print(f"pos {pos}: {us:.1f} us per launch, {byts / 1e6:.1f} MB -> {byts / us / 1e3:.0f} GB/s")
