"""Development aid: per-stage timeline of the GEMM-chain kernel (CTA 0) averaged over eager AR steps of the benchmarked batch.
    python scripts/chain_timeline.py [rows] [steps]"""
import ctypes, sys
sys.path.insert(0, '.')
import torch
from shapeformer_b200 import _lib, ar, synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda:0')
lib = _lib.load()
cfg = synth.SHIPPED_GPT
sd = synth.gpt_state_dict(cfg, seed=314, peaky=False)
s = ar.ARSampler(ar.pack_gpt_weights(sd, cfg, dev), cfg, (4096, 4096), max_rows=rows, max_cond=256, max_steps=steps, keep_history=False)
c = synth.cond_indices(max(rows // 4, 1), 256, seed=1).repeat_interleave(min(4, rows), 0)[:rows]
kw = dict(top_k=50, top_p=0.0, best_in_first=True, mask_invalid=False, mask_invalid_completion=False, stop_early=False)
s.sample(c, steps, use_graph=False, **kw)
buf = torch.zeros(128, dtype=torch.int64, device=dev)
_lib.check(lib.sfb200_debug_chain_timeline(_lib.ptr(buf)))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.sample(c, steps, use_graph=False, **kw); e1.record(); torch.cuda.synchronize()
_lib.check(lib.sfb200_debug_chain_timeline(None))
t = buf.cpu().tolist()
t = [v / 1.965 if i < 64 else v for i, v in enumerate(t)]    # SM cycles -> ns at 1965 MHz
names = ["wait_before barrier", "row statistics", "prefetch issue", "chunk loop", "partial store", "post-GEMM barrier", "reduction"]
print(f"eager pass: {e0.elapsed_time(e1):.1f} ms for {steps} steps ({e0.elapsed_time(e1) / steps * 1e3:.0f} us/step)")
tot = 0.0
for p in range(5):
    for st in range(7):
        n = t[64 + p * 8 + st]
        if n:
            print(f"phase {p} {names[st]:22s} avg {t[p * 8 + st] / n / 1e3:7.2f} us  (n={n})")
            tot += t[p * 8 + st]
for slot, nm in [(48, "X: cp.async wait + LDS"), (49, "X: wait x stage free"), (50, "X: LN + split + STS + arrive"), (51, "X: wait D full (drain)"), (40, "W: wait weights (TMA)"), (41, "W: LDS + split"), (42, "W: wait A stage free"), (43, "W: tmem st + arrive"),
                 (44, "MMA: wait D free"), (46, "MMA: wait x tile"), (45, "MMA: wait A operand"), (47, "MMA: issue + commit")]:
    n = t[64 + slot]
    if n:
        print(f"{nm:26s} avg {t[slot] / n / 1e3:7.3f} us per chunk (n={n}), {t[slot] / steps / 1e3:7.1f} us per step")
print(f"entry->dependency wait avg {t[62] / max(t[64 + 62], 1) / 1e3:.2f} us; sum of stages per step = {tot / steps / 1e3:.0f} us")
