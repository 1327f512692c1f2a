"""Generate the committed golden fixtures from the REFERENCE's own modules (run in the build container only):

    python tests/golden/make_golden.py

Outputs (small .pt files next to this script):
  sampler_*.pt : reference ShapeFormer.sample_indices on the tiny synthetic CondTupleGPT — tokens, and for every step the
                 masked logit of the sampled token, the number of finite logits and the log-sum-exp (history summary)
  decoder.pt   : reference LocalDecoder/Quantizer decode_index logits for 2048 query points of one synthetic code grid
Inputs are re-derived from seeds (shapeformer_b200/synth.py); the Exp(1) noise the reference's torch.multinomial consumed
is reproduced by seeding the default CPU generator identically (oracle.sf_oracle.TorchNoise).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from shapeformer_b200 import synth  # noqa: E402
from tests import refutil  # noqa: E402

SAMPLER_CASES = {
    # name: (masks(mask_invalid, mask_invalid_completion), top_k, top_p, temperature, best_in_first, B, L_c, steps, seeds)
    "demo": ((True, True), 100, 0.4, 1.0, True, 4, 12, 16, (5, 2, 11)),
    "topk50_nomask": ((False, False), 50, 0.0, 1.0, True, 4, 10, 20, (6, 3, 12)),
    "greedy_nomask": ((False, False), 1, 0.001, 1.0, False, 2, 16, 24, (7, 4, 13)),
    "temp_topp": ((True, False), 0, 0.9, 0.7, False, 3, 8, 12, (8, 5, 14)),
}


def history_summary(x, hist):
    out = []
    for i, h in enumerate(hist):
        tok = x[..., i]
        out.append(dict(at_token=torch.gather(h, 2, tok[..., None])[..., 0],
                        n_finite=torch.isfinite(h).sum(-1),
                        lse=torch.logsumexp(h, -1)))
    return out


def main():
    cfg = synth.TINY_GPT
    for name, (masks, top_k, top_p, T, bif, B, Lc, steps, (wseed, cseed, rseed)) in SAMPLER_CASES.items():
        sd = synth.gpt_state_dict(cfg, seed=wseed, peaky=True)
        sf = refutil.ref_shapeformer(cfg, sd, mask_invalid=masks[0], mask_invalid_completion=masks[1])
        c = synth.cond_indices(B, Lc, seed=cseed, shared=True)
        torch.manual_seed(rseed)
        x, hist = sf.sample_indices(c_indices=c, z_indices=c[:, :0], max_steps=steps, best_in_first=bif, top_k=top_k,
                                    top_p=top_p, temperature=T)
        torch.save(dict(case=name, masks=masks, top_k=top_k, top_p=top_p, temperature=T, best_in_first=bif, B=B, L_c=Lc,
                        steps=steps, seeds=(wseed, cseed, rseed), tokens=x, hist=history_summary(x, hist)),
                   os.path.join(HERE, f"sampler_{name}.pt"))
        print(name, tuple(x.shape), x[0, :4].tolist())
    sd = synth.vqdif_state_dict(seed=4)
    dec, q = refutil.ref_vqdif_decoder(sd)
    code = synth.code_grids(1, seed=3)
    g = torch.Generator().manual_seed(0)
    Xtg = torch.rand(1, 2048, 3, generator=g) * 2 - 1
    logits = refutil.ref_decode_index(dec, q, code, Xtg)
    torch.save(dict(wseed=4, code_seed=3, pts_seed=0, n=2048, logits=logits[0, :, 0]), os.path.join(HERE, "decoder.pt"))
    print("decoder", float(logits.min()), float(logits.max()))


if __name__ == "__main__":
    main()
