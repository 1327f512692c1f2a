"""GPU parity AT THE CONFIGURATIONS bench.py TIMES, and through the drop-in classes a user of the reference would call.

  * cfg 4/5 batch: the shipped 325M transformer, 64 rows = 16 shapes x sample_n 4, L_cond 256, top_k 50, masks off, CUDA graph
    replay in 32-step chunks, grouped-prefix attention — tokens bit-exact against the KV-cached oracle over a long run, with a
    first-divergence diagnostic (step, row, top-1 / top-2 margin of the oracle's logits) on mismatch (SURVEY.md §7 hard parts);
  * cfg 2: one row, L_cond 256, 512 greedy steps (GEMV / split-KV path);
  * the YAML-built ShapeFormer / VQDIF classes: reference-keyed checkpoint -> load_state_dict -> .sample() -> tokens + CPU
    history against the fixture generated from the reference; decode_index / decode_sample_indices (fp64 grid) vs the golden.
"""
import os

import numpy as np
import pytest
import torch

from oracle import sf_oracle as O
from shapeformer_b200 import ar, synth
from shapeformer_b200.xgutils import nputil, optutil, sysutil
from tests import util

pytestmark = pytest.mark.gpu
END = (4096, 4096)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def divergence_report(x, ox, oh):
    """First (step, row, tuple element) where the token streams differ + the oracle's top-1 / top-2 logit margin there."""
    n = min(x.shape[1], ox.shape[1])
    diff = (x[:, :n] != ox[:, :n])
    if x.shape[1] != ox.shape[1] and not diff.any():
        return f"same tokens for {n} steps but lengths differ: {x.shape[1]} vs oracle {ox.shape[1]}"
    steps = diff.any(-1).any(0).nonzero()
    j = int(steps[0])
    rows = diff[:, j].any(-1).nonzero()[:, 0].tolist()
    b = rows[0]
    i = 0 if diff[b, j, 0] else 1
    top = torch.topk(oh[i][b, j], 2).values
    return (f"first divergence at step {j} (of {n}), rows {rows}, tuple element {i}: got {x[b, j].tolist()} oracle "
            f"{ox[b, j].tolist()}; oracle top-1/top-2 logits {top[0]:.6f}/{top[1]:.6f} (margin {top[0] - top[1]:.2e}); "
            f"{int(diff.any(-1).sum())} of {diff.shape[0] * n} tuples differ")


def run_pair(cuda, cfg, sd, c, steps, top_k, top_p, best_in_first, noise, **kw):
    B, Lc = c.shape[:2]
    ox, oh = O.sample_indices(sd, O.GPTSpec(**cfg), c, c[:, :0], steps, END, best_in_first, top_k, top_p, 1.0, False, False,
                              noise=O.ListNoise(noise.reshape(-1, B, 4097)), cached=True)
    s = ar.ARSampler(ar.pack_gpt_weights(sd, cfg, cuda), cfg, END, max_rows=B, max_cond=Lc, max_steps=steps,
                     keep_history=True, **kw)
    x, hist = s.sample(c, steps, top_k=top_k, top_p=top_p, best_in_first=best_in_first, mask_invalid=False,
                       mask_invalid_completion=False, noise=noise, use_graph=True, stop_early=False)
    x = x.cpu()
    assert ox.shape[1] == steps      # masks off: nothing ends early
    worst = 0.0
    for a, b in zip(hist, oh):
        worst = max(worst, (a.cpu() - b).abs().max().item())
    assert torch.equal(x, ox), divergence_report(x, ox, oh)
    return worst


def test_bench_batch_shipped_model_long_run(cuda):
    """The batch bench.py times (rows, grouping, L_cond, top-k, graph chunks of 32), 256 of its 512 steps (the cached CPU
    oracle for all 512 would not fit the GPU test budget; positions 256..511 are what the first 256 steps visit)."""
    cfg = synth.SHIPPED_GPT
    sd = synth.gpt_state_dict(cfg, seed=314, peaky=True)
    B, n, Lc, steps = 64, 4, 256, 256
    c = synth.cond_indices(B // n, Lc, seed=1000).repeat_interleave(n, 0)
    noise = util.noise_from_seed(17, steps, B, 4097)
    worst = run_pair(cuda, cfg, sd, c, steps, 50, 0.0, True, noise)
    print(f"bench-config batch: tokens bit-exact over {steps} steps x {B} rows; max |dlogit| = {worst:.2e}")
    assert worst < 2e-4, worst


def test_large_batch_decode_path(cuda):
    """Decode batches of more than 64 rows (the default bench batch is 512) leave the persistent chain kernel for the TMA-fed
    large-M GEMM (csrc/tc_big.cu) + LayerNorm / split kernels, and — from 1024 (head, group) pairs on — the one-CTA-per-(head,
    group) form of the grouped attention: 256 rows = 64 shapes x sample_n 4 of the shipped model, tokens bit-exact against the
    KV-cached oracle."""
    cfg = synth.SHIPPED_GPT
    sd = synth.gpt_state_dict(cfg, seed=314, peaky=True)
    B, n, Lc, steps = 256, 4, 64, 32
    c = synth.cond_indices(B // n, Lc, seed=2000).repeat_interleave(n, 0)
    noise = util.noise_from_seed(19, steps, B, 4097)
    worst = run_pair(cuda, cfg, sd, c, steps, 50, 0.0, True, noise)
    print(f"large-batch decode path: tokens bit-exact over {steps} steps x {B} rows; max |dlogit| = {worst:.2e}")
    assert worst < 2e-4, worst


def test_cfg2_single_row_greedy_512(cuda):
    """BASELINE cfg 2: B = 1, L_cond 256, 512 greedy steps (top_k 1, top_p 0.001), fixed-length mode."""
    cfg = synth.SHIPPED_GPT
    sd = synth.gpt_state_dict(cfg, seed=314, peaky=True)
    c = synth.cond_indices(1, 256, seed=3)
    steps = 512
    noise = util.noise_from_seed(23, steps, 1, 4097)
    worst = run_pair(cuda, cfg, sd, c, steps, 1, 0.001, False, noise)
    print(f"cfg 2: tokens bit-exact over {steps} greedy steps; max |dlogit| = {worst:.2e}")
    assert worst < 2e-4, worst


# ----------------------------------------------------------------------------------------------------------------------
def build_models(cuda, cfg, masks):
    """ShapeFormer + VQDIF built the way the trainer does: YAML -> instantiate_from_opt, with the transformer shrunk to the
    fixture's tiny configuration."""
    opt = optutil.load_option(os.path.join(ROOT, "configs", "b200", "shapeformer_b200.yaml"))["pl_model_opt"]
    kw = opt["kwargs"]
    for k in ("vocab_sizes", "extra_vocab_sizes", "block_size"):
        kw[k] = list(cfg[k]) if isinstance(cfg[k], tuple) else cfg[k]
        kw["transformer_opt"]["kwargs"][k] = kw[k]
    kw["transformer_opt"]["kwargs"].update(n_layers=list(cfg["n_layers"]), n_head=cfg["n_head"], n_embd=cfg["n_embd"])
    kw["representer_opt"]["kwargs"].update(block_size=cfg["block_size"], mask_invalid=masks[0],
                                           mask_invalid_completion=masks[1])
    return sysutil.instantiate_from_opt(opt)


def reference_checkpoint(cfg, wseed, vseed):
    """A state dict with the reference's key set for a full ShapeFormer checkpoint: transformer.* (incl. the attn.mask
    buffers) and the frozen VQDIF under representer.vqvae_model.* (decoder, quantiser with EMA buffers)."""
    ckpt = {"transformer." + k: v for k, v in synth.gpt_state_dict(cfg, seed=wseed, peaky=True).items()}
    bs = cfg["block_size"]
    for g, nl in enumerate(cfg["n_layers"]):
        for l in range(nl):
            ckpt[f"transformer.blocks.{g}.{l}.attn.mask"] = torch.tril(torch.ones(bs, bs)).view(1, 1, bs, bs)
    vsd = synth.vqdif_state_dict(seed=vseed)
    ckpt.update({"representer.vqvae_model." + k: v for k, v in vsd.items()})
    ckpt["representer.vqvae_model.quantizer.N"] = torch.zeros(4096)
    ckpt["representer.vqvae_model.quantizer.z_avg"] = vsd["quantizer.embedding.weight"].clone()
    return ckpt, vsd


@pytest.mark.parametrize("case", ["demo", "topk50_nomask"])
def test_dropin_shapeformer_sample_reproduces_reference_fixture(cuda, case):
    """ShapeFormer.sample through the YAML-built class with a checkpoint loaded AFTER a first (warm-up) sample with other
    weights: tokens bit-exact vs the reference fixture, CPU history matches, and returned tensors are fresh (a second call
    does not overwrite them)."""
    g = torch.load(os.path.join(util.GOLDEN, f"sampler_{case}.pt"))
    cfg = synth.TINY_GPT
    wseed, cseed, rseed = g["seeds"]
    model = build_models(cuda, cfg, g["masks"]).to(cuda)
    c = synth.cond_indices(g["B"], g["L_c"], seed=cseed, shared=True)
    noise = util.noise_from_seed(rseed, g["steps"], g["B"], 4097)
    kw = dict(z_indices=c[:, :0], max_steps=g["steps"], temperature=g["temperature"], sample=True,
              best_in_first=g["best_in_first"], top_k=g["top_k"], top_p=g["top_p"], noise=noise)
    model.sample(c_indices=c.to(cuda), **kw)                       # warm-up with the constructor's random weights
    ckpt, _ = reference_checkpoint(cfg, wseed, 4)
    model.load_state_dict(ckpt, strict=True)                       # nested load must invalidate the packed weights
    out_x, x, hist = model.sample(c_indices=c.to(cuda), **kw)
    assert out_x.device.type == "cuda" and x.dtype == torch.int64 and hist[0].device.type == "cpu"
    assert x.shape == g["tokens"].shape and torch.equal(x.cpu(), g["tokens"]) and torch.equal(out_x, x)
    util.check_history_summary(x.cpu(), hist, g["hist"])
    assert hist[0].data_ptr() != hist[1].data_ptr() and not torch.equal(hist[0], hist[1])
    # outputs are fresh tensors: another call with different noise leaves them untouched
    keep_x, keep_h = x.clone(), [h.clone() for h in hist]
    model.sample(c_indices=c.to(cuda), **dict(kw, noise=util.noise_from_seed(rseed + 1, g["steps"], g["B"], 4097)))
    assert torch.equal(x, keep_x) and all(torch.equal(a, b) for a, b in zip(hist, keep_h))
    # the reference's ranking (compute_log_probs, shapeformer.py:407-418) on the returned history
    from shapeformer_b200.models.shapeformer.shapeformer import compute_log_probs
    lp = compute_log_probs(x.cpu().numpy(), [h.numpy() for h in hist])
    for i in range(2):
        want = g["hist"][i]["at_token"].double() - g["hist"][i]["lse"].double()
        assert np.abs(lp[..., i] - want.numpy()).max() < 1e-4


def test_dropin_vqdif_decode_index_and_decode_sample_indices(cuda):
    """VQDIF.decode_index through the nn.Module (checkpoint loaded through the parent ShapeFormer), and
    decode_sample_indices with the reference's fp64 makeGrid query points (shapeformer.py:382-391)."""
    from shapeformer_b200.models.shapeformer.shapeformer import decode_sample_indices
    g = torch.load(util.GOLDEN + "/decoder.pt")
    cfg = synth.TINY_GPT
    model = build_models(cuda, cfg, (True, True)).to(cuda)
    ckpt, vsd = reference_checkpoint(cfg, 5, g["wseed"])
    model.load_state_dict(ckpt, strict=True)
    vq = model.representer.vqvae_model
    assert vq.device.type == "cuda"
    code = synth.code_grids(1, seed=g["code_seed"])
    Xtg = torch.rand(1, g["n"], 3, generator=torch.Generator().manual_seed(g["pts_seed"])) * 2 - 1
    out = vq.decode_index(code.to(cuda), Xtg.to(cuda))["logits"]
    assert out.shape == (1, g["n"], 1)
    assert (torch.sigmoid(out[0, :, 0].cpu()) - torch.sigmoid(g["logits"])).abs().max() < 1e-4
    assert (out[0, :, 0].cpu() - g["logits"]).abs().max() < 2e-4
    # decode_sample_indices: numpy (16,16,16) code grid + numpy fp64 (N,3) grid -> numpy (N,) occupancy
    grid = nputil.makeGrid([-1, -1, -1], [1, 1, 1], [20, 20, 20], indexing="ij")
    assert grid.dtype == np.float64
    occ = decode_sample_indices(vq, grid, code[0].numpy())
    assert isinstance(occ, np.ndarray) and occ.shape == (8000,) and occ.dtype == np.float32
    ref = torch.sigmoid(O.decode_index(vsd, code, torch.from_numpy(grid)[None].float())["logits"])[0, :, 0].numpy()
    assert np.abs(occ - ref).max() < 1e-4
