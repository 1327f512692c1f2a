"""Development aid: time of the large-M GEMM kernels at the decode shapes (CUDA graph of back-to-back launches).
    python scripts/gemm_time.py [M ...]"""
import sys
sys.path.insert(0, '.')
import torch
from shapeformer_b200 import _lib, ops
lib = _lib.load()
dev = torch.device('cuda:0')
shapes = [(1024, 1024), (3072, 1024), (4096, 1024), (1024, 4096)]
part = torch.empty(lib.sfb200_big_partial_floats(), device=dev); cnt = torch.zeros(2048, dtype=torch.int32, device=dev)
for M in [int(a) for a in sys.argv[1:]] or [256, 512, 4096]:
    tot = {"big": 0.0, "tc": 0.0}
    for N, K in shapes:
        x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.03
        xl, Wl = ops.split_lo(x), ops.split_lo(W)
        y = torch.empty(M, N, device=dev)
        def big():
            _lib.check(lib.sfb200_linear_big(_lib.ptr(x), _lib.ptr(xl), _lib.ptr(W), _lib.ptr(Wl), None, None, _lib.ptr(y), None, M, N, K, 0,
                                             _lib.ptr(part), _lib.ptr(cnt), _lib.stream_ptr()), "big")
        def tc():
            _lib.check(lib.sfb200_linear_tc(_lib.ptr(x), _lib.ptr(W), None, None, _lib.ptr(y), M, N, K, 0, _lib.stream_ptr()), "tc")
        for name, fn in (("big", big), ("tc", tc)):
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); n = 20
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for _ in range(n): fn()
            torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            tot[name] += us
            print(f"M={M} N={N} K={K} {name}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s (algorithmic; x3 executed)")
    print(f"M={M}: per block big {tot['big']:.1f} us, tc {tot['tc']:.1f} us")
