"""Development aid: time of the conv prologue (32 shapes) on the tcgen05 kernels vs the cuDNN path, per layer for the former.
    python scripts/conv_time.py [B]"""
import sys, time
sys.path.insert(0, ".")
import torch
from shapeformer_b200 import decoder, synth
dev = torch.device("cuda:0")
sd = synth.vqdif_state_dict(seed=6)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
code = synth.code_grids(B, seed=1).to(dev)
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for pro in ("tc", "cudnn"):
    dec = decoder.ImplicitDecoder(sd, dev, prologue=pro)
    t = ev(lambda: dec.feature_grid_from_codes(code))
    print(f"{pro}: {t:.2f} ms per {B} shapes = {t / B:.3f} ms per shape ({96.39e9 * B / t / 1e9:.1f} algorithmic TFLOP/s)")
# per-layer times of the tc path
dec = decoder.ImplicitDecoder(sd, dev, prologue="tc")
ct = dec.conv_tc
orig_conv, orig_prep = ct._conv, ct._prep
rows = []
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        rows.append((name, a[2] if name == "conv" else "", (time.perf_counter() - t0) * 1e3)); return r
    return w
ct._conv, ct._prep = timed("conv", orig_conv), timed("prep", orig_prep)
dec.feature_grid_from_codes(code); rows.clear(); dec.feature_grid_from_codes(code)
for n, k, ms in rows: print(f"  {n:5s} {str(k)[-45:]:45s} {ms:8.3f} ms")
print(f"  sum conv {sum(r[2] for r in rows if r[0] == 'conv'):.2f} ms, prep {sum(r[2] for r in rows if r[0] == 'prep'):.2f} ms")
