"""CPU: the oracle reproduces the committed golden fixtures generated from the reference (tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import sf_oracle as O
from shapeformer_b200 import synth
from tests import util


@pytest.mark.parametrize("path", util.sampler_goldens())
@pytest.mark.parametrize("cached", [True, False])
def test_oracle_sampler_golden(path, cached):
    g = torch.load(path)
    cfg = synth.TINY_GPT
    wseed, cseed, rseed = g["seeds"]
    sd = synth.gpt_state_dict(cfg, seed=wseed, peaky=True)
    c = synth.cond_indices(g["B"], g["L_c"], seed=cseed, shared=True)
    noise = O.ListNoise(util.noise_from_seed(rseed, g["steps"], g["B"], 4097).reshape(-1, g["B"], 4097))
    x, hist = O.sample_indices(sd, O.GPTSpec(**cfg), c, c[:, :0], g["steps"], (4096, 4096), g["best_in_first"], g["top_k"],
                               g["top_p"], g["temperature"], g["masks"][0], g["masks"][1], noise=noise, cached=cached)
    assert torch.equal(x, g["tokens"])
    util.check_history_summary(x, hist, g["hist"])


def test_oracle_decoder_golden():
    g = torch.load(util.GOLDEN + "/decoder.pt")
    sd = synth.vqdif_state_dict(seed=g["wseed"])
    code = synth.code_grids(1, seed=g["code_seed"])
    gen = torch.Generator().manual_seed(g["pts_seed"])
    Xtg = torch.rand(1, g["n"], 3, generator=gen) * 2 - 1
    out = O.decode_index(sd, code, Xtg)["logits"][0, :, 0]
    assert (out - g["logits"]).abs().max() < 1e-5
