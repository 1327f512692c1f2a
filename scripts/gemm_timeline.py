"""GPU development probe: per-launch period of back-to-back decode GEMMs (CUDA events) and the in-kernel phase timeline of
sfb200_linear_tc_ps (CTA (0,0), %globaltimer)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from shapeformer_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda")
names = ["entry", "setup done", "dep wait done", "first x tile", "first MMA", "last MMA commit", "acc drained", "cluster sync 1", "exit"]
for (M, N, K) in [(64, 1024, 1024), (64, 3072, 1024), (64, 4096, 1024), (64, 1024, 4096)]:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.03
    wt = torch.empty(lib.sfb200_tc_pretiled_floats(N, K), device=dev)
    lib.sfb200_tc_pretile(_lib.ptr(W), _lib.ptr(wt), N, K, _lib.stream_ptr())
    y = torch.empty(M, N, device=dev)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    def run(n, kind):
        for _ in range(n):
            if kind == "ps":
                lib.sfb200_linear_tc_ps(_lib.ptr(x), _lib.ptr(wt), None, None, _lib.ptr(y), M, N, K, 0, _lib.stream_ptr())
            elif kind == "tc":
                lib.sfb200_linear_tc(_lib.ptr(x), _lib.ptr(W), None, None, _lib.ptr(y), M, N, K, 0, _lib.stream_ptr())
            else:
                lib.sfb200_linear(_lib.ptr(x), _lib.ptr(W), None, None, _lib.ptr(y), M, N, K, 0, _lib.stream_ptr())
    for kind in ("ps", "tc", "ffma"):
        run(5, kind); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(3, kind)
            with torch.cuda.graph(g, stream=s):
                run(50, kind)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"M={M} N={N} K={K} {kind:4s}: {e0.elapsed_time(e1) / 50 * 1e3:7.2f} us per launch (graph of 50 back-to-back, PDL)")
    lib.sfb200_debug_ps_timeline(_lib.ptr(buf))
    run(3, "ps"); torch.cuda.synchronize()
    t = buf.cpu().tolist()
    lib.sfb200_debug_ps_timeline(None)
    print("   timeline (us from entry): " + ", ".join(f"{n} {(t[i] - t[0]) / 1e3:.2f}" for i, n in enumerate(names) if t[i]))
