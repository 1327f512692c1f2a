"""Development aid: time the fused implicit-decoder point kernel on the benchmarked size (64 shapes x 64^3 points)."""
import sys
sys.path.insert(0, ".")
import torch
from shapeformer_b200 import decoder, synth
dev = torch.device("cuda:0")
sd = synth.vqdif_state_dict(seed=6)
dec = decoder.ImplicitDecoder(sd, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.randn(B, 64, 64, 64, 32, device=dev)
X = synth.make_grid(64)[None].to(dev)
for impl in (0, 1):
    dec.decode_points(g, X, impl=impl); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dec.decode_points(g, X, impl=impl); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 30976.0 * B * 64 ** 3
    print(f"impl {impl}: {best:.3f} ms for {B} x 64^3 points = {fl / best / 1e9:.1f} algorithmic TFLOP/s ({3 * fl / best / 1e9:.1f} executed)")
a = dec.decode_points(g[:4], X, impl=0); b = dec.decode_points(g[:4], X, impl=1)
print("tc vs ffma max |d| =", (a - b).abs().max().item())
