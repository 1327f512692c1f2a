"""GPU experiment: fp32 cuDNN conv3d vs 3xTF32 (three TF32 tensor-core convs on hi/lo splits) for the decoder's conv stack.
Prints time and max error per layer shape (B shapes)."""
import sys, time
import torch
import torch.nn.functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")


def split(t):
    bits = t.view(torch.int32)
    hi = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    lo = t - hi
    lo = ((lo.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, lo


def conv3x(x, w, cl):
    xh, xl = split(x)
    wh, wl = split(w)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=True):
        return F.conv3d(xl, wh, None, padding=1) + F.conv3d(xh, wl, None, padding=1) + F.conv3d(xh, wh, None, padding=1)


def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


shapes = [(128, 128, 16), (128, 256, 8), (256, 512, 4), (768, 256, 8), (384, 128, 16), (128, 64, 32), (64, 64, 32), (64, 32, 64), (32, 32, 64)]
for ci, co, r in shapes:
    x = torch.randn(B, ci, r, r, r, device=dev)
    w = torch.randn(co, ci, 3, 3, 3, device=dev) / (ci * 27) ** 0.5
    for cl in (False, True):
        xx = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
        ww = w.contiguous(memory_format=torch.channels_last_3d) if cl else w
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref = F.conv3d(xx, ww, None, padding=1)
            t32 = t(lambda: F.conv3d(xx, ww, None, padding=1))
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=True):
            o1 = F.conv3d(xx, ww, None, padding=1)
            t1 = t(lambda: F.conv3d(xx, ww, None, padding=1))
        o3 = conv3x(xx, ww, cl)
        t3 = t(lambda: conv3x(xx, ww, cl))
        ref64 = F.conv3d(x[:1].double(), w.double(), None, padding=1)
        e32 = (ref[:1].double() - ref64).abs().max().item()
        e1 = (o1[:1].double() - ref64).abs().max().item()
        e3 = (o3[:1].double() - ref64).abs().max().item()
        print(f"ci={ci:4d} co={co:4d} r={r:3d} cl={int(cl)} | fp32 {t32:8.2f} ms err {e32:.1e} | tf32 {t1:8.2f} ms err {e1:.1e} | 3xtf32 {t3:8.2f} ms err {e3:.1e}", flush=True)
