"""GPU parity of the VQDIF encoder + quantiser + token packing against the CPU oracle (which is pinned
against the reference's own LocalPoolPointnet / Quantizer / batch_dense2sparse in tests/test_oracle_vs_reference.py).

Index outputs are compared exactly wherever the oracle's decision is not a numerical near-tie: a code index is an argmin over
4096 distances of magnitude ~250 computed from features that went through four GroupNorms over mostly-empty grids (fp32 noise
amplified to ~1e-4, the same size as the oracle-vs-reference difference), so a cell whose two best distances differ by less
than 5e-2 may legitimately resolve either way; all other cells, the occupancy mask, the mode and the token packing are exact."""
import os

import pytest
import torch

from oracle import sf_oracle as O
from shapeformer_b200 import encoder, synth
from shapeformer_b200.xgutils import optutil, sysutil

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("B,T,seed", [(1, 2048, 0), (3, 5000, 1), (2, 16384, 2)])
def test_encoder_features_mask_and_codes(cuda, B, T, seed):
    sd = synth.vqdif_state_dict(seed=4)
    cloud = synth.partial_cloud(B, T, seed=seed)
    cloud[0, :4] = torch.tensor([[-1., -1, -1], [1, 1, 1], [0.999, -0.999, 0.5], [1.2, -1.3, 0.]])    # borders / out of range
    ofeat, omask = O.encoder_forward(sd, cloud / 2.0)
    oraw, dist = O.quantize(sd, ofeat)
    enc = encoder.PointEncoder(sd, cuda)
    raw, mask, feat = enc.encode_quant(cloud, return_feat=True)
    assert torch.equal(mask.cpu(), omask)
    err = (feat.cpu() - ofeat).abs().max().item()
    print(f"encoder features: max |d| = {err:.2e} (B={B}, T={T})")
    assert err < 1e-3
    top2 = torch.topk(-dist, 2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).view(B, 16, 16, 16)
    diff = raw.cpu() != oraw
    print(f"code indices: {int(diff.sum())} of {diff.numel()} differ; smallest oracle margin among them "
          f"{margin[diff].min().item() if diff.any() else float('nan'):.3g}; occupied cells {int(omask.sum())}")
    assert not (diff & (margin > 5e-2)).any()
    assert diff.float().mean() < 2e-3
    # determinism: the scatter-mean uses fixed-point atomics, the scatter-max integer atomics
    raw2, mask2, feat2 = enc.encode_quant(cloud, return_feat=True)
    assert torch.equal(raw, raw2) and torch.equal(feat, feat2)


def test_quantize_cloud_and_tokens_exact_given_indices(cuda):
    """The integer part in isolation: mode fill, empty index and (pos, val) packing on the GPU's own raw indices are exactly the
    oracle's (torch.unique / torch.mode / nonzero semantics, incl. the max_length crop with the forced final end tuple)."""
    sd = synth.vqdif_state_dict(seed=4)
    enc = encoder.PointEncoder(sd, cuda)
    cloud = synth.partial_cloud(3, 6000, seed=3)
    for max_length in (406, 64):
        out = enc.quantize_cloud(cloud, max_length=max_length)
        raw, mask = out["raw_ind"].cpu(), out["mask"].cpu()
        mode = O.mode_smallest(raw)
        dense = torch.zeros_like(raw) + mode
        dense[mask] = raw[mask]
        assert int(out["mode"]) == int(mode) and torch.equal(out["quant_ind"].cpu(), dense)
        toks, empty = O.batch_dense2sparse(dense, max_length)
        assert int(out["empty_index"]) == int(empty)
        assert torch.equal(out["c_indices"].cpu(), toks)


def test_representer_get_indices_through_the_yaml_built_model(cuda):
    """AR_N.get_indices (representers.py:80-103) through the drop-in classes: cloud -> c_indices usable by ShapeFormer.sample."""
    opt = optutil.load_option(os.path.join(ROOT, "configs", "b200", "shapeformer_b200.yaml"))["pl_model_opt"]
    cfg = synth.TINY_GPT
    kw = opt["kwargs"]
    kw["block_size"] = kw["transformer_opt"]["kwargs"]["block_size"] = kw["representer_opt"]["kwargs"]["block_size"] = 812
    kw["transformer_opt"]["kwargs"].update(n_layers=list(cfg["n_layers"]), n_head=cfg["n_head"], n_embd=cfg["n_embd"])
    model = sysutil.instantiate_from_opt(opt)
    sd = synth.vqdif_state_dict(seed=4)
    model.representer.vqvae_model.load_state_dict(sd, strict=False)
    model.to(cuda)
    cloud = synth.partial_cloud(1, 4096, seed=5).to(cuda)
    c, z, extra, others = model.representer.get_indices(cloud)
    oc, oempty = O.get_indices(sd, cloud.cpu())
    assert c.shape[2] == 2 and z.shape[1] == 0 and extra.shape[:2] == c.shape[:2]
    assert torch.equal(c[0, -1].cpu(), torch.tensor([4096, 4096])) and int(others["empty_index"]) == int(oempty)
    assert c.shape == oc.shape and (c.cpu() != oc).float().mean() < 0.02      # positions identical, codes up to near-ties
    assert torch.equal(c.cpu()[..., 0], oc[..., 0])
    out_x, x, hist = model.sample(c_indices=c.expand(2, -1, -1).contiguous(), z_indices=c[:, :0].expand(2, -1, -1), max_steps=6,
                                  temperature=1.0, sample=True, best_in_first=True, top_k=50, top_p=0.0)
    assert x.shape[0] == 2 and x.shape[2] == 2
