"""Development aid: AR step time (CUDA graph replay) of the benchmarked batch for the current SFB200_* environment.
    SFB200_CHAIN=0 SFB200_ATTN_GROUPED=0 python scripts/step_time.py [rows] [steps]"""
import os, sys
sys.path.insert(0, '.')
import torch
from shapeformer_b200 import _lib, ar, synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device('cuda:0')
cfg = synth.SHIPPED_GPT
sd = synth.gpt_state_dict(cfg, seed=314, peaky=False)
s = ar.ARSampler(ar.pack_gpt_weights(sd, cfg, dev), cfg, (4096, 4096), max_rows=rows, max_cond=256, max_steps=steps, keep_history=False)
c = synth.cond_indices(max(rows // 4, 1), 256, seed=1).repeat_interleave(min(4, rows), 0)[:rows]
kw = dict(top_k=50, top_p=0.0, best_in_first=True, mask_invalid=False, mask_invalid_completion=False, stop_early=False)
s.sample(c, steps, use_graph=True, **kw)
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample(c, steps, use_graph=True, **kw); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"CHAIN={os.environ.get('SFB200_CHAIN','1')} GROUPED={os.environ.get('SFB200_ATTN_GROUPED','1')} rows={rows}: "
      f"{best:.1f} ms for prefill + {steps} steps -> <= {best / steps * 1e3:.0f} us/step")
