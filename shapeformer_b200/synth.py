"""Seeded synthetic weights and inputs (no checkpoints or datasets are reachable offline — SURVEY.md fact 9, §8d).

State dicts use the reference's key names so the same tensors feed the reference modules, the CPU
oracle and the B200 classes.
"""
import math

import torch

SHIPPED_GPT = dict(n_embd=1024, n_head=16, n_layers=(20, 4), block_size=812, vocab_sizes=(4097, 4097),
                   extra_vocab_sizes=(4097,))
TINY_GPT = dict(n_embd=128, n_head=2, n_layers=(2, 1), block_size=64, vocab_sizes=(4097, 4097),
                extra_vocab_sizes=(4097,))


def gpt_state_dict(cfg=SHIPPED_GPT, seed=314, peaky=True, dtype=torch.float32):
    """CondTupleGPT weights.  peaky=False: the reference init (transformer/mingpt.py:248-255: N(0, 0.02), LN 1/0,
    zero positional tables).  peaky=True: N(0, 0.2) on heads + embeddings and N(0, 0.02) positional tables so that logits
    spread over a few units and greedy margins are far above fp32 rounding."""
    g = torch.Generator().manual_seed(seed)
    d, V, Ve, bs = cfg["n_embd"], cfg["vocab_sizes"], cfg["extra_vocab_sizes"], cfg["block_size"]
    es = 0.2 if peaky else 0.02

    def n(*shape, std=0.02):
        return torch.randn(*shape, generator=g, dtype=dtype) * std

    sd = {}
    sd["pos_emb"] = n(1, bs, d) if peaky else torch.zeros(1, bs, d)
    sd["cond_pos_emb"] = n(1, bs, d) if peaky else torch.zeros(1, bs, d)
    for i, v in enumerate(V):
        sd[f"tok_embs.{i}.weight"] = n(v, d, std=es)
    for i, v in enumerate(Ve):
        sd[f"extra_tok_embs.{i}.weight"] = n(v, d, std=es)
    for gi, nl in enumerate(cfg["n_layers"]):
        for l in range(nl):
            p = f"blocks.{gi}.{l}."
            for ln in ("ln1", "ln2"):
                sd[p + ln + ".weight"] = torch.ones(d) + (n(d, std=0.05) if peaky else 0)
                sd[p + ln + ".bias"] = n(d, std=0.02) if peaky else torch.zeros(d)
            for nm in ("key", "query", "value", "proj"):
                sd[p + f"attn.{nm}.weight"] = n(d, d)
                sd[p + f"attn.{nm}.bias"] = n(d, std=0.01) if peaky else torch.zeros(d)
            sd[p + "mlp.0.weight"] = n(4 * d, d)
            sd[p + "mlp.0.bias"] = n(4 * d, std=0.01) if peaky else torch.zeros(4 * d)
            sd[p + "mlp.2.weight"] = n(d, 4 * d)
            sd[p + "mlp.2.bias"] = n(d, std=0.01) if peaky else torch.zeros(d)
        sd[f"heads.{gi}.0.weight"] = torch.ones(d)
        sd[f"heads.{gi}.0.bias"] = torch.zeros(d)
        sd[f"heads.{gi}.1.weight"] = n(V[gi], d, std=es)
    return sd


def _conv(g, co, ci, k):
    b = 1.0 / math.sqrt(ci * k ** 3)
    return (torch.rand(co, ci, k, k, k, generator=g) * 2 - 1) * b


def _lin(g, co, ci):
    b = 1.0 / math.sqrt(ci)
    return (torch.rand(co, ci, generator=g) * 2 - 1) * b, (torch.rand(co, generator=g) * 2 - 1) * b


def vqdif_state_dict(seed=314, vq_dim=128, n_codes=4096, hidden=32):
    """VQDIF decoder + codebook weights (keys as in VQDIF.state_dict(): decoder.*, quantizer.embedding.weight).
    torch-default-like uniform fan-in init; fc_1.weight (zero in the reference init, vqdif/layers.py:37) is drawn
    N(0, 0.1) so the residual branches are exercised; GroupNorm affine is perturbed around (1, 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {"quantizer.embedding.weight": torch.randn(n_codes, vq_dim, generator=g)}
    f = vq_dim

    def gn(pre, c):
        sd[pre + "groupnorm.weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        sd[pre + "groupnorm.bias"] = 0.1 * torch.randn(c, generator=g)

    def gcr(pre, ci, co):
        gn(pre, ci)
        sd[pre + "conv.weight"] = _conv(g, co, ci, 3)

    u = "decoder.unet3d."
    enc = [(f, f, f), (f, f, 2 * f), (2 * f, 2 * f, 4 * f)]
    for i, (ci, cm, co) in enumerate(enc):
        gcr(f"{u}encoders.{i}.basic_module.SingleConv1.", ci, cm)
        gcr(f"{u}encoders.{i}.basic_module.SingleConv2.", cm, co)
    dec = [(4 * f + 2 * f, 2 * f), (2 * f + f, f)]
    for i, (ci, co) in enumerate(dec):
        gcr(f"{u}decoders.{i}.basic_module.SingleConv1.", ci, co)
        gcr(f"{u}decoders.{i}.basic_module.SingleConv2.", co, co)
    sd[u + "final_conv.weight"] = _conv(g, f, f, 1)
    sd[u + "final_conv.bias"] = (torch.rand(f, generator=g) * 2 - 1) / math.sqrt(f)
    up = "decoder.upsampler."
    ch = [f, f // 2, f // 4]
    for s in range(2):
        for j, (ci, co) in enumerate(((ch[s], ch[s + 1]), (ch[s + 1], ch[s + 1]))):
            pre = f"{up}blocks.{3 * s + 1 + j}."
            sd[pre + "conv.weight"] = _conv(g, co, ci, 3)
            sd[pre + "groupnorm.weight"] = 1 + 0.1 * torch.randn(co, generator=g)
            sd[pre + "groupnorm.bias"] = 0.1 * torch.randn(co, generator=g)
    d = "decoder."
    sd[d + "fc_p.weight"], sd[d + "fc_p.bias"] = _lin(g, hidden, 3)
    for i in range(5):
        sd[f"{d}fc_c.{i}.weight"], sd[f"{d}fc_c.{i}.bias"] = _lin(g, hidden, hidden)
        sd[f"{d}blocks.{i}.fc_0.weight"], sd[f"{d}blocks.{i}.fc_0.bias"] = _lin(g, hidden, hidden)
        _, sd[f"{d}blocks.{i}.fc_1.bias"] = _lin(g, hidden, hidden)
        sd[f"{d}blocks.{i}.fc_1.weight"] = 0.1 * torch.randn(hidden, hidden, generator=g)
    sd[d + "fc_out.weight"], sd[d + "fc_out.bias"] = _lin(g, 1, hidden)
    sd.update(encoder_state_dict(seed + 1, hidden, vq_dim))
    return sd


def encoder_state_dict(seed=315, hidden=32, vq_dim=128, n_blocks=5):
    """LocalPoolPointnet weights (keys as in VQDIF.state_dict(): encoder.*; shipped shapenet_res16 shapes: hidden = c_dim = 32,
    Downsampler 32 -> 64 -> 128 in two k2s2 + k1 'crg' pairs).  Drawn from their own generator so that the decoder / codebook
    tensors of vqdif_state_dict keep their round-1 values.  fc_1.weight (zero in the reference init) is drawn N(0, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    e, sd = "encoder.", {}
    sd[e + "fc_pos.weight"], sd[e + "fc_pos.bias"] = _lin(g, 2 * hidden, 3)
    for i in range(n_blocks):
        sd[f"{e}blocks.{i}.fc_0.weight"], sd[f"{e}blocks.{i}.fc_0.bias"] = _lin(g, hidden, 2 * hidden)
        _, sd[f"{e}blocks.{i}.fc_1.bias"] = _lin(g, hidden, hidden)
        sd[f"{e}blocks.{i}.fc_1.weight"] = 0.1 * torch.randn(hidden, hidden, generator=g)
        sd[f"{e}blocks.{i}.shortcut.weight"] = _lin(g, hidden, 2 * hidden)[0]
    sd[e + "fc_c.weight"], sd[e + "fc_c.bias"] = _lin(g, hidden, hidden)
    ch = [hidden, 2 * hidden, 4 * hidden]
    assert ch[-1] == vq_dim
    for s in range(2):
        for j, (ci, co, k) in enumerate(((ch[s], ch[s + 1], 2), (ch[s + 1], ch[s + 1], 1))):
            pre = f"{e}downsampler.blocks.{2 * s + j}."
            sd[pre + "conv.weight"] = _conv(g, co, ci, k)
            sd[pre + "groupnorm.weight"] = 1 + 0.1 * torch.randn(co, generator=g)
            sd[pre + "groupnorm.bias"] = 0.1 * torch.randn(co, generator=g)
    return sd


def partial_cloud(B, T=4096, seed=0):
    """Synthetic partial scans (§8d cfg 5): T points on a random half of a noisy sphere shell inside [-0.9, 0.9]^3."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(B):
        v = torch.randn(T, 3, generator=g)
        v = v / v.norm(dim=1, keepdim=True)
        n = torch.randn(3, generator=g)
        n = n / n.norm()
        v = torch.where((v @ n)[:, None] < 0, v - 2 * (v @ n)[:, None] * n, v)      # keep the half facing n
        r = 0.45 + 0.3 * torch.rand(1, generator=g) + 0.02 * torch.randn(T, 1, generator=g)
        ctr = (torch.rand(3, generator=g) - 0.5) * 0.2
        out.append((v * r + ctr).clamp(-0.9, 0.9))
    return torch.stack(out)


def cond_indices(B, L_c, seed=0, n_pos=4096, n_val=4096, end_tokens=(4096, 4096), shared=False):
    """Conditioning tuples: L_c-1 sorted distinct positions with random codes + the (end, end) terminator (§8d cfg 2).
    shared=True repeats one conditioning for all rows (the reference's sample_n expansion, shapeformer.py:229)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(1 if shared else B):
        pos = torch.randperm(n_pos, generator=g)[:L_c - 1].sort()[0]
        val = torch.randint(0, n_val, (L_c - 1,), generator=g)
        t = torch.stack([pos, val], 1)
        rows.append(torch.cat([t, torch.tensor([list(end_tokens)])], 0))
    c = torch.stack(rows).long()
    return c.expand(B, -1, -1).contiguous() if shared else c


def code_grids(B, seed=0, res=16, n_codes=4096, occupied=(200, 500)):
    """Code grids with a mode-filled background and a few hundred random non-empty cells (§8d cfg 3)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(B, res, res, res, dtype=torch.int64)
    for b in range(B):
        mode = int(torch.randint(0, n_codes, (1,), generator=g))
        grid = torch.full((res ** 3,), mode, dtype=torch.int64)
        k = int(torch.randint(occupied[0], occupied[1], (1,), generator=g))
        cells = torch.randperm(res ** 3, generator=g)[:k]
        grid[cells] = torch.randint(0, n_codes, (k,), generator=g)
        out[b] = grid.view(res, res, res)
    return out


def make_grid(res=64, lo=-1.0, hi=1.0, dtype=torch.float32):
    """xgutils.nputil.makeGrid([lo]*3,[hi]*3,[res]*3, indexing="ij") flattened (nputil.py:618-654): x slowest, z fastest."""
    import numpy as np
    ax = np.linspace(lo, hi, res)
    grid = np.stack(np.meshgrid(ax, ax, ax, sparse=False, indexing="ij"), axis=-1).reshape(-1, 3)
    return torch.from_numpy(grid).to(dtype)
