"""CPU restatement (torch fp32, no CUDA) of ShapeFormer's data-parallel hot path.

*** TEST INFRASTRUCTURE — NOT PRODUCT CODE ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module, and only as the checker / the CPU arm.  ``shapeformer_b200/`` never imports it; the product path raises if the
CUDA library is missing.

Parity pin: the reference has no tests or golden vectors for this path, so this restatement is pinned
against the reference's OWN modules executed in the build container through ``oracle/ref_shim.py``:
``tests/test_oracle_vs_reference.py`` (runs whenever ``/root/reference`` is present) and the committed fixtures under
``tests/golden/`` produced by ``tests/golden/make_golden.py`` from the reference modules.  The arithmetic underneath
(softmax, layer_norm, gelu, conv3d, group_norm, sort, cumsum ...) is PyTorch ATen in both, the reference's own
third-party dependency (pinned torch 1.7.0 in environment.yml:24; 2.11 here).

All state is passed as plain ``dict[str, Tensor]`` with the reference's state_dict key names.
Every function cites the reference lines it restates (paths relative to /root/reference/shapeformer/models/).
"""
import math

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# ----------------------------------------------------------------------------------------------------------------------
# Representer: extra indices + sampling mask
# ----------------------------------------------------------------------------------------------------------------------
def next_cond_pos(c_pos, z_pos, end_token):
    """First conditioning position strictly greater than each generated position; end stays end.

    Restates shapeformer/representers.py:432-442 (get_next_cond).  c_pos (B,Lc) sorted ascending incl. the end token,
    z_pos (B,Lz) -> (B,Lz) int64.
    """
    if z_pos.shape[1] == 0:
        return z_pos.clone()
    idx = torch.searchsorted(c_pos.contiguous(), z_pos.contiguous(), right=True)
    is_end = z_pos == end_token
    idx = torch.where(is_end, torch.full_like(idx, c_pos.shape[1] - 1), idx)
    out = torch.gather(c_pos, 1, idx)
    return torch.where(is_end, torch.full_like(out, end_token), out)


def extra_indices(c_idx, z_idx, end_token):
    """AR_N.get_extra_indices — shapeformer/representers.py:187-196.  (B,Lc,2),(B,Lz,2) -> (B,Lc+Lz,1)."""
    c_pos = c_idx[..., 0]
    z_extra = next_cond_pos(c_pos, z_idx[..., 0], end_token)
    return torch.cat([c_pos.clone(), z_extra], 1)[..., None]


def sampling_masker(logits, idx, L_cond, step_j, tuple_i, end_tokens, mask_invalid=True, mask_invalid_completion=False):
    """ShapeRepresenter.sampling_masker — shapeformer/representers.py:120-155.

    logits (B,V) fp32; idx (B,L+1,2) whose row -1 is the tuple being sampled and row -2 the newest complete tuple.
    """
    out = logits.clone()
    B, V = out.shape
    last = idx[:, -2, 0]
    if tuple_i == 1:
        ended = idx[:, -1, 0] == end_tokens[0]
        out[ended, :] = NEG_INF
        out[ended, end_tokens[1]] = 1.0
        return out
    v = torch.arange(V, dtype=idx.dtype)[None, :]
    if mask_invalid and step_j > 0:
        bad = v <= last[:, None]
        bad[:, end_tokens[0]] = False
        out[bad] = NEG_INF
    if mask_invalid_completion:
        cond = idx[:, :L_cond, 0]
        cond_plus = torch.cat([cond, torch.full((B, 1), end_tokens[0] + 1, dtype=idx.dtype)], 1).contiguous()
        nxt_i = torch.searchsorted(cond_plus, last[:, None].contiguous(), right=True)
        nxt = torch.gather(cond_plus, 1, nxt_i)
        out[v > nxt] = NEG_INF
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Sampling: temperature / top-k / top-p filter, then multinomial as argmax(p / Exp(1))
# ----------------------------------------------------------------------------------------------------------------------
def filter_logits_row(row, top_k, top_p, temperature):
    """filter_sampling_logits — shapeformer/common.py:260-285 (one row, returns a new tensor)."""
    row = row / temperature
    k = min(int(top_k), row.shape[-1])
    if k > 0:
        kth = torch.topk(row, k)[0][-1]
        row = torch.where(row < kth, torch.full_like(row, NEG_INF), row)
    if top_p > 0.0:
        srt, order = torch.sort(row, descending=True)
        cum = torch.cumsum(F.softmax(srt, dim=-1), dim=-1)
        drop = cum > top_p
        drop = torch.cat([torch.zeros(1, dtype=torch.bool), drop[:-1]])
        row = row.clone()
        row[order[drop]] = NEG_INF
    return row


def sample_rows(logits, noise, top_k, top_p, temperature):
    """sample_logits — shapeformer/common.py:288-299 with torch.multinomial(n=1) written as argmax(p / q).

    ``noise`` (B,V) is the Exp(1) draw multinomial makes internally (SURVEY.md fact 4 [probe]); passing it in makes the
    draw an explicit input.  Returns int64 (B,).
    """
    filt = torch.stack([filter_logits_row(logits[b], top_k, top_p, temperature) for b in range(logits.shape[0])])
    probs = F.softmax(filt, dim=-1)
    return torch.argmax(probs / noise, dim=-1)


# ----------------------------------------------------------------------------------------------------------------------
# CondTupleGPT
# ----------------------------------------------------------------------------------------------------------------------
class GPTSpec:
    """Static shape of a CondTupleGPT (transformer/mingpt.py:187-244)."""

    def __init__(self, n_embd=1024, n_head=16, n_layers=(20, 4), block_size=812, vocab_sizes=(4097, 4097),
                 extra_vocab_sizes=(4097,)):
        self.n_embd, self.n_head, self.n_layers, self.block_size = n_embd, n_head, tuple(n_layers), block_size
        self.vocab_sizes, self.extra_vocab_sizes = tuple(vocab_sizes), tuple(extra_vocab_sizes)


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def gpt_embed(sd, idx, extra, L_cond):
    """get_embeddings — transformer/mingpt.py:256-286.  idx (B,L,2), extra (B,L,1) -> (B,L,d)."""
    L = idx.shape[1]
    x = sd["tok_embs.0.weight"][idx[..., 0]] + sd["tok_embs.1.weight"][idx[..., 1]]
    x = x + sd["extra_tok_embs.0.weight"][extra[..., 0]]
    pos = torch.cat([sd["cond_pos_emb"][:, :L_cond], sd["pos_emb"][:, :L - L_cond]], 1)
    return x + pos


def gpt_block(sd, pre, x, n_head, kv_prefix=None):
    """Block.forward + CausalSelfAttention.forward — transformer/mingpt.py:74-91,108-111.

    With ``kv_prefix=(K,V)`` of shape (B,H,T0,hd) the new positions attend to prefix + themselves causally (the
    KV-cached evaluation, mathematically identical — SURVEY.md fact 2).  Returns (x, (K,V)) with the updated cache.
    """
    B, T, C = x.shape
    hd = C // n_head
    h = _ln(x, sd[pre + "ln1.weight"], sd[pre + "ln1.bias"])
    q = F.linear(h, sd[pre + "attn.query.weight"], sd[pre + "attn.query.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    k = F.linear(h, sd[pre + "attn.key.weight"], sd[pre + "attn.key.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    v = F.linear(h, sd[pre + "attn.value.weight"], sd[pre + "attn.value.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    T0 = 0
    if kv_prefix is not None:
        T0 = kv_prefix[0].shape[2]
        k = torch.cat([kv_prefix[0], k], 2)
        v = torch.cat([kv_prefix[1], v], 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
    qi = torch.arange(T)[:, None] + T0
    ki = torch.arange(T0 + T)[None, :]
    att = att.masked_fill(ki > qi, NEG_INF)
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    x = x + F.linear(y, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    h = _ln(x, sd[pre + "ln2.weight"], sd[pre + "ln2.bias"])
    h = F.gelu(F.linear(h, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"]))
    x = x + F.linear(h, sd[pre + "mlp.2.weight"], sd[pre + "mlp.2.bias"])
    return x, (k, v)


def gpt_head(sd, g, x):
    """heads[g] = LayerNorm + Linear(no bias) — transformer/mingpt.py:222-231 (head_hidden_layers=0)."""
    return F.linear(_ln(x, sd[f"heads.{g}.0.weight"], sd[f"heads.{g}.0.bias"]), sd[f"heads.{g}.1.weight"])


def gpt_group(sd, spec, g, x):
    for l in range(spec.n_layers[g]):
        x, _ = gpt_block(sd, f"blocks.{g}.{l}.", x, spec.n_head)
    return x


def gpt_forward(sd, spec, idx, extra, L_cond, target_idx):
    """CondTupleGPT.forward / compute_logits — transformer/mingpt.py:287-296,311-319.  Returns [logits0, logits1]."""
    x = gpt_embed(sd, idx, extra, L_cond)
    out = []
    for g in range(len(spec.n_layers)):
        x = gpt_group(sd, spec, g, x)
        out.append(gpt_head(sd, g, x))
        x = x + sd[f"tok_embs.{g}.weight"][target_idx[..., g]]
    return out


class GPTCachedStepper:
    """KV-cached evaluation of sample_next_tuple (transformer/mingpt.py:297-310) — one new position per call.

    ``group0(idx_new, extra_new, L_cond)`` consumes the newest tuples (B,T,2) and returns logits0 of the last position;
    ``group1(target_pos)`` adds tok_embs[0](target) to the stored group-0 output of the positions not yet pushed through
    blocks[1] and returns logits1 of the last position.  Mathematically identical to recomputing the whole prefix.
    """

    def __init__(self, sd, spec):
        self.sd, self.spec = sd, spec
        self.kv = [[None] * n for n in spec.n_layers]
        self.len0 = 0          # positions pushed through blocks[0]
        self.pending = None    # group-0 outputs awaiting blocks[1]  (B,T,d)

    def group0(self, idx_new, extra_new, L_cond):
        sd, spec = self.sd, self.spec
        B, T, _ = idx_new.shape
        p0 = self.len0
        x = sd["tok_embs.0.weight"][idx_new[..., 0]] + sd["tok_embs.1.weight"][idx_new[..., 1]]
        x = x + sd["extra_tok_embs.0.weight"][extra_new[..., 0]]
        rows = []
        for t in range(p0, p0 + T):
            rows.append(sd["cond_pos_emb"][0, t] if t < L_cond else sd["pos_emb"][0, t - L_cond])
        x = x + torch.stack(rows)[None]
        for l in range(spec.n_layers[0]):
            x, self.kv[0][l] = gpt_block(sd, f"blocks.0.{l}.", x, spec.n_head, self.kv[0][l])
        self.len0 += T
        self.pending = x if self.pending is None else torch.cat([self.pending, x], 1)
        return gpt_head(sd, 0, x[:, -1:, :])[:, 0]

    def group1(self, target_pos):
        """target_pos (B,T) = the pos element of the NEXT tuple for each pending position."""
        sd, spec = self.sd, self.spec
        x = self.pending + sd["tok_embs.0.weight"][target_pos]
        self.pending = None
        for l in range(spec.n_layers[1]):
            x, self.kv[1][l] = gpt_block(sd, f"blocks.1.{l}.", x, spec.n_head, self.kv[1][l])
        return gpt_head(sd, 1, x[:, -1:, :])[:, 0]


# ----------------------------------------------------------------------------------------------------------------------
# ShapeFormer.sample_indices
# ----------------------------------------------------------------------------------------------------------------------
class TorchNoise:
    """Exp(1) noise drawn the way torch.multinomial(n=1) draws it: empty_like(p).exponential_(1) from the default
    generator (ATen Distributions multinomial fast path)."""

    def __init__(self, generator=None):
        self.generator = generator

    def __call__(self, B, V):
        return torch.empty(B, V, dtype=torch.float32).exponential_(1.0, generator=self.generator)


class ListNoise:
    """Replays a pre-drawn (n_draws, B, V) tensor, in call order."""

    def __init__(self, noise):
        self.noise, self.i = noise, 0

    def __call__(self, B, V):
        q = self.noise[self.i]
        self.i += 1
        assert q.shape == (B, V)
        return q


def sample_indices(sd, spec, c_indices, z_indices, max_steps, end_tokens=(4096, 4096), best_in_first=False, top_k=100,
                   top_p=0.8, temperature=1.0, mask_invalid=True, mask_invalid_completion=False, noise=None,
                   cached=True):
    """ShapeFormer.sample_indices — shapeformer/shapeformer.py:54-123, for the AR_N representer.

    ``mask_invalid`` / ``mask_invalid_completion`` are the REPRESENTER's attributes (the reference ignores the same-named
    kwargs of sample_indices — SURVEY.md App. C-2).  ``cached=False`` re-runs the full forward every step exactly like
    the reference (used as the faithful CPU baseline); ``cached=True`` uses GPTCachedStepper (fast checker).
    Four noise draws of shape (B,V) are consumed per step in the reference's order: pos-sample, pos-best, val-sample,
    val-best.  Returns (x (B,steps,2) int64, [hist0, hist1] each (B,steps,V) fp32).
    """
    noise = noise or TorchNoise()
    B, L_c, tuple_n = c_indices.shape
    assert tuple_n == 2 and z_indices.shape[1] == 0, "oracle covers the shipped call: empty z prefix, (pos,val) tuples"
    assert L_c + max_steps < spec.block_size, "overflow crop (shapeformer.py:73-76, buggy) is out of scope"
    V0, V1 = spec.vocab_sizes
    sampled = torch.zeros(B, L_c + max_steps, 2, dtype=torch.int64)
    sampled[:, :L_c] = c_indices
    hist = [[], []]
    L = L_c
    stepper = GPTCachedStepper(sd, spec) if cached else None
    for j in range(max_steps):
        c, z = sampled[:, :L_c], sampled[:, L_c:L]
        extra = extra_indices(c, z, end_tokens[0])
        if cached:
            lo = stepper.len0
            logits = stepper.group0(sampled[:, lo:L], extra[:, lo:L], L_c)
        else:
            x = gpt_embed(sd, sampled[:, :L], extra, L_c)
            x = gpt_group(sd, spec, 0, x)
            logits = gpt_head(sd, 0, x)[:, -1]
        # --- tuple element 0: position
        logits = sampling_masker(logits, sampled[:, :L + 1], L_c, j, 0, end_tokens, mask_invalid, mask_invalid_completion)
        hist[0].append(logits)
        new = sample_rows(logits, noise(B, V0), top_k, top_p, temperature)
        best = sample_rows(logits, noise(B, V0), 1, 0.001, temperature)
        if best_in_first:
            new[0] = best[0]
        sampled[:, L, 0] = new
        # --- tuple element 1: value
        if cached:
            n_pending = stepper.pending.shape[1]
            logits = stepper.group1(sampled[:, L + 1 - n_pending:L + 1, 0])
        else:
            x = x + sd["tok_embs.0.weight"][sampled[:, 1:L + 1, 0]]
            x = gpt_group(sd, spec, 1, x)
            logits = gpt_head(sd, 1, x)[:, -1]
        logits = sampling_masker(logits, sampled[:, :L + 1], L_c, j, 1, end_tokens, mask_invalid, mask_invalid_completion)
        hist[1].append(logits)
        new = sample_rows(logits, noise(B, V1), top_k, top_p, temperature)
        best = sample_rows(logits, noise(B, V1), 1, 0.001, temperature)
        if best_in_first:
            new[0] = best[0]
        sampled[:, L, 1] = new
        L += 1
        ended = (sampled[:, L - 1, 0] == end_tokens[0]) | (sampled[:, L - 1, 1] == end_tokens[1])
        if bool(ended.all()):
            break
    hist = [torch.stack(h, 1) for h in hist]
    return sampled[:, L_c:L], hist


# ----------------------------------------------------------------------------------------------------------------------
# VQDIF decoder
# ----------------------------------------------------------------------------------------------------------------------
def get_code(sd, code_ind):
    """Quantizer.get_code — vqdif/quantizer.py:19-30.  (B,R,R,R) int64 -> (B,C,R,R,R)."""
    return sd["quantizer.embedding.weight"][code_ind].permute(0, 4, 1, 2, 3).contiguous()


def _gcr(sd, pre, x):
    """SingleConv order 'gcr' — vqdif/unet3d.py:18-60,74-92: GroupNorm(8) -> Conv3d(3,pad 1,no bias) -> ReLU."""
    x = F.group_norm(x, 8, sd[pre + "groupnorm.weight"], sd[pre + "groupnorm.bias"], 1e-5)
    return F.relu(F.conv3d(x, sd[pre + "conv.weight"], None, padding=1))


def _crg(sd, pre, x):
    """ConvLayer order 'crg' — vqdif/updown.py:79-99: Conv3d -> ReLU -> GroupNorm(8)."""
    x = F.relu(F.conv3d(x, sd[pre + "conv.weight"], None, padding=1))
    return F.group_norm(x, 8, sd[pre + "groupnorm.weight"], sd[pre + "groupnorm.bias"], 1e-5)


def unet3d(sd, x, pre="decoder.unet3d.", num_levels=3):
    """Abstract3DUNet.forward — vqdif/unet3d.py:449-474 with DoubleConv encoders/decoders (:103-144, :222-300)."""
    feats = []
    for i in range(num_levels):
        if i > 0:
            x = F.max_pool3d(x, 2)
        x = _gcr(sd, f"{pre}encoders.{i}.basic_module.SingleConv1.", x)
        x = _gcr(sd, f"{pre}encoders.{i}.basic_module.SingleConv2.", x)
        feats.insert(0, x)
    for i, skip in enumerate(feats[1:]):
        x = F.interpolate(x, size=skip.shape[2:], mode="nearest")
        x = torch.cat([skip, x], 1)
        x = _gcr(sd, f"{pre}decoders.{i}.basic_module.SingleConv1.", x)
        x = _gcr(sd, f"{pre}decoders.{i}.basic_module.SingleConv2.", x)
    return F.conv3d(x, sd[pre + "final_conv.weight"], sd[pre + "final_conv.bias"])


def upsampler(sd, x, pre="decoder.upsampler.", steps=2):
    """Upsampler.forward — vqdif/updown.py:119-132: per step nearest x2 then two 'crg' convs."""
    for s in range(steps):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = _crg(sd, f"{pre}blocks.{3 * s + 1}.", x)
        x = _crg(sd, f"{pre}blocks.{3 * s + 2}.", x)
    return x


def normalize_3d(p, padding=0.1):
    """normalize_3d_coordinate — vqdif/common.py:260-276 (the data-dependent ifs are equivalent to elementwise selects)."""
    p = p / (1 + padding + 10e-4) + 0.5
    p = torch.where(p >= 1, torch.full_like(p, 1 - 10e-4), p)
    p = torch.where(p < 0, torch.zeros_like(p), p)
    return p


def grid_feature(p, grid):
    """LocalDecoder.sample_grid_feature — vqdif/dec.py:62-68 written out as the explicit 8-corner blend.

    p (B,N,3) in [-0.5,0.5], grid (B,C,D,H,W) -> (B,N,C).  Equals F.grid_sample(bilinear, border, align_corners=True)
    (checked in tests/test_oracle_vs_reference.py); component 0 of p indexes W (last dim), 1 -> H, 2 -> D.
    """
    B, C, D, H, W = grid.shape
    pn = normalize_3d(p).float()
    vg = 2.0 * pn - 1.0
    out = F.grid_sample(grid, vg[:, :, None, None], padding_mode="border", align_corners=True, mode="bilinear")
    return out[..., 0, 0].transpose(1, 2)


def grid_feature_explicit(p, grid):
    """Same as grid_feature with the arithmetic spelled out the way the CUDA kernel does it (fp32)."""
    B, C, D, H, W = grid.shape
    pn = normalize_3d(p).float()
    vg = 2.0 * pn - 1.0
    outs = []
    for b in range(B):
        fx = ((vg[b, :, 0] + 1) / 2) * (W - 1)
        fy = ((vg[b, :, 1] + 1) / 2) * (H - 1)
        fz = ((vg[b, :, 2] + 1) / 2) * (D - 1)
        fx = fx.clamp(0, W - 1); fy = fy.clamp(0, H - 1); fz = fz.clamp(0, D - 1)
        x0 = fx.floor(); y0 = fy.floor(); z0 = fz.floor()
        tx, ty, tz = fx - x0, fy - y0, fz - z0
        x0 = x0.long(); y0 = y0.long(); z0 = z0.long()
        x1 = (x0 + 1).clamp(max=W - 1); y1 = (y0 + 1).clamp(max=H - 1); z1 = (z0 + 1).clamp(max=D - 1)
        g = grid[b]
        acc = 0
        for zz, wz in ((z0, 1 - tz), (z1, tz)):
            for yy, wy in ((y0, 1 - ty), (y1, ty)):
                for xx, wx in ((x0, 1 - tx), (x1, tx)):
                    acc = acc + g[:, zz, yy, xx] * (wx * wy * wz)[None]
        outs.append(acc.t())
    return torch.stack(outs)


def decoder_mlp(sd, p, c, pre="decoder.", n_blocks=5):
    """LocalDecoder.forward MLP — vqdif/dec.py:86-97 + ResnetBlockFC.forward vqdif/layers.py:39-48.  -> (B,N,1)."""
    net = F.linear(p.float(), sd[pre + "fc_p.weight"], sd[pre + "fc_p.bias"])
    for i in range(n_blocks):
        net = net + F.linear(c, sd[f"{pre}fc_c.{i}.weight"], sd[f"{pre}fc_c.{i}.bias"])
        h = F.linear(F.relu(net), sd[f"{pre}blocks.{i}.fc_0.weight"], sd[f"{pre}blocks.{i}.fc_0.bias"])
        net = net + F.linear(F.relu(h), sd[f"{pre}blocks.{i}.fc_1.weight"], sd[f"{pre}blocks.{i}.fc_1.bias"])
    return F.linear(F.relu(net), sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"])


def feature_grid(sd, code_ind):
    """get_code -> UNet3D -> Upsampler: the per-shape prologue of decode_index (vqdif/dec.py:75-83)."""
    return upsampler(sd, unet3d(sd, get_code(sd, code_ind)))


def decode_points(sd, grid, Xtg):
    """Per-point half of LocalDecoder.forward given the 32x64^3 feature grid; Xtg in [-1,1] (vqdif/vqdif.py:71)."""
    p = Xtg / 2.0
    return decoder_mlp(sd, p, grid_feature(p, grid))


def decode_index(sd, code_ind, Xtg):
    """VQDIF.decode_index — vqdif/vqdif.py:60-76.  -> {"logits": (B,N,1)}."""
    return {"logits": decode_points(sd, feature_grid(sd, code_ind), Xtg)}


# ----------------------------------------------------------------------------------------------------------------------
# token glue
# ----------------------------------------------------------------------------------------------------------------------
def tokens_to_dense(tokens, empty_index, res=16, end_tokens=(4096, 4096)):
    """filter_end_tokens + batch_sparse2dense for one row — shapeformer/common.py:50-55,171-189 (caller
    shapeformer/shapeformer.py:342-351).  tokens (L,2) int64 -> (res,res,res) int64; later duplicates win (index_put)."""
    dense = torch.full((res ** 3,), int(empty_index), dtype=torch.int64)
    keep = (tokens[:, 0] != end_tokens[0]) & (tokens[:, 1] != end_tokens[1])
    t = tokens[keep]
    for i in range(t.shape[0]):
        dense[int(t[i, 0])] = int(t[i, 1])
    return dense.view(res, res, res)


# ----------------------------------------------------------------------------------------------------------------------
# VQDIF encoder + quantiser + token packing  (the step before the hot path)
# ----------------------------------------------------------------------------------------------------------------------
def _resnet_fc(sd, pre, x):
    """ResnetBlockFC.forward — vqdif/layers.py:39-48 (shortcut when size_in != size_out)."""
    net = F.linear(F.relu(x), sd[pre + "fc_0.weight"], sd[pre + "fc_0.bias"])
    dx = F.linear(F.relu(net), sd[pre + "fc_1.weight"], sd[pre + "fc_1.bias"])
    xs = F.linear(x, sd[pre + "shortcut.weight"]) if (pre + "shortcut.weight") in sd else x
    return xs + dx


def encoder_forward(sd, p, reso=64, pre="encoder."):
    """LocalPoolPointnet.forward + generate_grid_features — vqdif/enc.py:66-140 (plane_type 'grid', scatter 'max',
    c2i_order 'original', downsampler with 2 steps).  p (B,T,3) in [-0.5,0.5] -> (grid_feat (B,128,16,16,16), mask
    (B,16,16,16) bool).  torch_scatter is restated with Tensor.scatter_reduce (max: empty cells 0, only gathered where
    points exist; mean: include_self=False)."""
    B, T, _ = p.shape
    p_nor = normalize_3d(p.clone())
    xi = (p_nor * reso).long()
    index = xi[..., 0] + reso * (xi[..., 1] + reso * xi[..., 2])                 # (B,T)  coordinate2index, 'original'
    net = F.linear(p, sd[pre + "fc_pos.weight"], sd[pre + "fc_pos.bias"])
    net = _resnet_fc(sd, pre + "blocks.0.", net)
    n_blocks = 1
    while (pre + f"blocks.{n_blocks}.fc_0.weight") in sd:
        n_blocks += 1
    for i in range(1, n_blocks):
        C = net.shape[2]
        idx = index[:, :, None].expand(-1, -1, C)
        pooled = torch.zeros(B, reso ** 3, C).scatter_reduce(1, idx, net, reduce="amax", include_self=False)
        pooled = torch.gather(pooled, 1, idx)                                    # pool_local, enc.py:96-114
        net = _resnet_fc(sd, pre + f"blocks.{i}.", torch.cat([net, pooled], 2))
    c = F.linear(net, sd[pre + "fc_c.weight"], sd[pre + "fc_c.bias"])             # (B,T,c_dim)
    C = c.shape[2]
    idx = index[:, :, None].expand(-1, -1, C)
    fea = torch.zeros(B, reso ** 3, C).scatter_reduce(1, idx, c, reduce="mean", include_self=False)
    fea = fea.permute(0, 2, 1).reshape(B, C, reso, reso, reso)                    # enc.py:72-74
    i = 0
    while (pre + f"downsampler.blocks.{i}.conv.weight") in sd:                    # Downsampler 'crg', updown.py:98-113
        dp = pre + f"downsampler.blocks.{i}."
        k = sd[dp + "conv.weight"].shape[-1]
        fea = F.relu(F.conv3d(fea, sd[dp + "conv.weight"], None, stride=k, padding=0))
        fea = F.group_norm(fea, 8, sd[dp + "groupnorm.weight"], sd[dp + "groupnorm.bias"], 1e-5)
        i += 1
    R = fea.shape[-1]
    mi = (p_nor * R).long()                                                       # enc.py:84-91
    mask = torch.zeros(B, R, R, R, dtype=torch.bool)
    b = torch.arange(B)[:, None].expand(-1, T)
    mask[b, mi[..., 2], mi[..., 1], mi[..., 0]] = True
    return fea, mask


def quantize(sd, grid_feat, key="quantizer.embedding.weight"):
    """Quantizer.forward (eval) — vqdif/quantizer.py:31-53: nearest code by |x|^2 - 2 x.w + |w|^2, first index on ties.
    Returns (quant_ind (B,R,R,R) int64, distances (B*R^3, n_codes))."""
    B, C = grid_feat.shape[:2]
    flat = grid_feat.permute(0, 2, 3, 4, 1).contiguous().view(-1, C)
    w = sd[key]
    dist = (flat ** 2).sum(1, keepdim=True) - 2 * torch.mm(flat, w.t()) + (w.t() ** 2).sum(0, keepdim=True)
    ind = torch.max(-dist, dim=1)[1]
    return ind.view(B, *grid_feat.shape[2:]), dist


def mode_smallest(x):
    """pth_get_mode — shapeformer/common.py:20-23 (= torch.mode: the smallest of the most frequent values)."""
    vals, counts = torch.unique(x.reshape(-1), return_counts=True)
    return vals[torch.argmax(counts)]


def quantize_cloud(sd, cloud):
    """VQDIF.quantize_cloud — vqdif/vqdif.py:36-58: cloud (B,T,3) in [-1,1] -> (quant_ind with the batch-wide mode in
    unoccupied cells, mode, raw quant_ind, mask)."""
    fea, mask = encoder_forward(sd, cloud / 2.0)
    raw, _ = quantize(sd, fea)
    mode = mode_smallest(raw)
    out = torch.zeros_like(raw) + mode
    out[mask] = raw[mask]
    return out, mode, raw, mask


def batch_dense2sparse(indices, max_length=None, end_tokens=(4096, 4096)):
    """batch_dense2sparse + unpack_sparse — shapeformer/common.py:84-122,152-169: (B,R,R,R) -> ((B,L,2) int64 padded with
    end tokens, L = longest row + 1, cropped to max_length with a forced final end tuple; mode)."""
    B = indices.shape[0]
    flat = indices.reshape(B, -1)
    mode = torch.mode(indices.reshape(-1))[0]
    rows = [torch.stack([(flat[b] != mode).nonzero()[:, 0], flat[b][flat[b] != mode]], 1) for b in range(B)]
    L = max(r.shape[0] for r in rows) + 1
    out = torch.tensor(list(end_tokens), dtype=torch.int64).repeat(B, L, 1)
    for b, r in enumerate(rows):
        out[b, :r.shape[0]] = r
    if max_length is not None and L > max_length:
        out = out[:, :max_length].clone()
        out[:, max_length - 1] = torch.tensor(list(end_tokens))
    return out, mode


def get_indices(sd, Xct, max_length=406, end_tokens=(4096, 4096)):
    """AR_N.encode_cloud + get_indices for inference (Xbd None) — shapeformer/representers.py:68-103: partial cloud (B,T,3)
    -> (c_indices (B,L_c,2), empty_index (mode))."""
    quant_ind, _, _, _ = quantize_cloud(sd, Xct)
    c_indices, mode = batch_dense2sparse(quant_ind, max_length, end_tokens)
    return c_indices, mode
