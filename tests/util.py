import glob
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def noise_from_seed(seed, steps, B, V):
    """The (steps,4,B,V) Exp(1) tensor the reference's four torch.multinomial calls per step consume after
    torch.manual_seed(seed) on CPU (pos-sample, pos-best, val-sample, val-best)."""
    torch.manual_seed(seed)
    return torch.stack([torch.stack([torch.empty(B, V).exponential_(1.0) for _ in range(4)]) for _ in range(steps)])


def sampler_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN, "sampler_*.pt")))


def check_history_summary(x, hist, gold_hist, tol=5e-5):
    for i in range(2):
        h = hist[i].float().cpu()
        g = gold_hist[i]
        at = torch.gather(h, 2, x[..., i].cpu()[..., None])[..., 0]
        assert torch.equal(torch.isfinite(h).sum(-1), g["n_finite"])
        assert (at - g["at_token"]).abs().max() < tol
        assert (torch.logsumexp(h, -1) - g["lse"]).abs().max() < tol
