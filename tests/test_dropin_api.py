"""CPU: the drop-in boundary — YAML `class:` plugin loading, constructor/state_dict compatibility with the reference, and the
loud failure of the product path without a GPU (no CPU fallback)."""
import os

import pytest
import torch

from shapeformer_b200 import _lib, decoder, synth
from shapeformer_b200.xgutils import optutil, sysutil
from tests import refutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_yaml_plugin_builds_the_b200_classes():
    opt = optutil.load_option(os.path.join(ROOT, "configs", "b200", "shapeformer_b200.yaml"))
    model = sysutil.instantiate_from_opt(opt["pl_model_opt"])
    assert type(model).__module__ == "shapeformer_b200.models.shapeformer.shapeformer"
    assert model.transformer.get_block_size() == 812 and not model.transformer.training
    assert model.representer.mask_invalid_completion is True and model.representer.max_length == 406
    n = sum(p.numel() for p in model.transformer.parameters())
    assert abs(n - 324.95e6) < 0.05e6
    vq = model.representer.vqvae_model
    assert type(vq).__name__ == "VQDIF" and sum(p.numel() for p in vq.decoder.parameters()) > 16e6
    opt2 = optutil.load_option(os.path.join(ROOT, "configs", "b200", "vqdif_b200.yaml"))
    assert type(sysutil.instantiate_from_opt(opt2["pl_model_opt"])).__name__ == "VQDIF"


def test_inherit_from_merges_recursively(tmp_path):
    (tmp_path / "base.yaml").write_text("a: {x: 1, y: {z: 2}}\nb: 3\n")
    (tmp_path / "child.yaml").write_text("inherit_from: base.yaml\na: {y: {w: 5}}\n")
    assert optutil.load_option(str(tmp_path / "child.yaml")) == {"a": {"x": 1, "y": {"z": 2, "w": 5}}, "b": 3}


@pytest.mark.skipif(not refutil.have_reference(), reason="/root/reference not present")
def test_state_dict_keys_match_the_reference():
    from shapeformer_b200.models.shapeformer.transformer.mingpt import CondTupleGPT
    from shapeformer_b200.models.vqdif.vqdif import VQDIF
    cfg = synth.TINY_GPT
    sd = synth.gpt_state_dict(cfg)
    ref = refutil.ref_gpt(cfg, sd)
    mine = CondTupleGPT(vocab_sizes=cfg["vocab_sizes"], extra_vocab_sizes=cfg["extra_vocab_sizes"],
                        block_size=cfg["block_size"], tuple_n=2, n_layers=cfg["n_layers"], n_head=cfg["n_head"],
                        n_embd=cfg["n_embd"])
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items() if not k.endswith("attn.mask")}
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == want
    mine.load_state_dict(ref.state_dict())                # reference checkpoints (with mask buffers) load strictly
    vsd = synth.vqdif_state_dict()
    dec, q = refutil.ref_vqdif_decoder(vsd)
    opt = optutil.load_option(os.path.join(ROOT, "configs", "b200", "vqdif_b200.yaml"))
    vq = sysutil.instantiate_from_opt(opt["pl_model_opt"])
    ref_keys = {"decoder." + k: tuple(v.shape) for k, v in dec.state_dict().items()}
    ref_keys.update({"quantizer." + k: tuple(v.shape) for k, v in q.state_dict().items()})
    ref_keys.update({"encoder." + k: tuple(v.shape) for k, v in refutil.ref_vqdif_encoder(vsd).state_dict().items()})
    assert {k: tuple(v.shape) for k, v in vq.state_dict().items()} == ref_keys


def test_product_path_fails_loudly_without_gpu():
    from shapeformer_b200.models.shapeformer.transformer.mingpt import CondTupleGPT
    cfg = synth.TINY_GPT
    gpt = CondTupleGPT(vocab_sizes=cfg["vocab_sizes"], extra_vocab_sizes=cfg["extra_vocab_sizes"], block_size=64, tuple_n=2,
                       n_layers=cfg["n_layers"], n_head=cfg["n_head"], n_embd=cfg["n_embd"])
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        gpt.packed_weights()
    with pytest.raises(_lib.Sfb200Error):
        decoder.ImplicitDecoder(synth.vqdif_state_dict(), "cpu")
    with pytest.raises(NotImplementedError):
        CondTupleGPT(vocab_sizes=(10, 10), extra_vocab_sizes=(10,), block_size=8, tuple_n=2, n_layers=(1, 1), n_head=3, n_embd=64)


def test_tf32_split_and_fp32_conv_prologue_on_cpu():
    """split_tf32: hi + lo reproduces the value to 2^-21 relative and both parts are TF32-representable; the fp32 mode of the
    conv prologue is the oracle's op sequence."""
    from oracle import sf_oracle as O
    t = torch.randn(10000) * torch.logspace(-3, 3, 10000)
    hi, lo = decoder.split_tf32(t)
    assert ((hi.view(torch.int32) & 0x1FFF) == 0).all() and ((lo.view(torch.int32) & 0x1FFF) == 0).all()
    assert ((hi + lo - t).abs() <= t.abs() * 2.0 ** -21).all()
    sd = synth.vqdif_state_dict(seed=6)
    x = O.get_code(sd, synth.code_grids(1, seed=1))
    assert torch.equal(decoder.conv_prologue(sd, x, up_mode="fp32"), O.upsampler(sd, O.unet3d(sd, x)))


def test_full_checkpoint_loads_through_the_parent_and_invalidates_caches():
    """A reference ShapeFormer checkpoint (transformer.* incl. attn.mask buffers, representer.vqvae_model.{encoder,decoder,
    quantizer}.*) loads STRICTLY through the parent module, and the nested load resets the packed-weight / sampler / decoder
    engine caches (nn.Module recursion bypasses the children's load_state_dict, so these are hooks)."""
    from shapeformer_b200.models.shapeformer.shapeformer import ShapeFormer
    cfg = synth.TINY_GPT
    pre = "shapeformer_b200.models."
    vq_opt = {"class": pre + "vqdif.vqdif.VQDIF", "kwargs": dict(
        decoder_opt={"class": pre + "vqdif.dec.LocalDecoder",
                     "kwargs": dict(sample_mode="bilinear", hidden_size=32, c_dim=32, unet3d=True,
                                    unet3d_kwargs=dict(num_levels=3, f_maps=128, in_channels=128, out_channels=128),
                                    upsampler=True, upsampler_kwargs=dict(in_channels=128, upsampler_steps=2))},
        quantizer_opt={"class": pre + "vqdif.quantizer.Quantizer", "kwargs": dict(vocab_size=4096, n_embd=128)})}
    model = ShapeFormer(tuple_n=2, block_size=cfg["block_size"], end_tokens=[4096, 4096], vocab_sizes=list(cfg["vocab_sizes"]),
                        extra_vocab_sizes=list(cfg["extra_vocab_sizes"]),
                        transformer_opt={"class": pre + "shapeformer.transformer.mingpt.CondTupleGPT",
                                         "kwargs": dict(tuple_n=2, vocab_sizes=cfg["vocab_sizes"],
                                                        extra_vocab_sizes=cfg["extra_vocab_sizes"], n_layers=cfg["n_layers"],
                                                        block_size=cfg["block_size"], n_head=cfg["n_head"],
                                                        n_embd=cfg["n_embd"])},
                        representer_opt={"class": pre + "shapeformer.representers.AR_N",
                                         "kwargs": dict(block_size=cfg["block_size"], end_tokens=[4096, 4096],
                                                        vqvae_opt=vq_opt)})
    ckpt = {"transformer." + k: v for k, v in synth.gpt_state_dict(cfg, seed=5).items()}
    bs = cfg["block_size"]
    for g, nl in enumerate(cfg["n_layers"]):
        for l in range(nl):
            ckpt[f"transformer.blocks.{g}.{l}.attn.mask"] = torch.tril(torch.ones(bs, bs)).view(1, 1, bs, bs)
    vsd = synth.vqdif_state_dict(seed=5)
    ckpt.update({"representer.vqvae_model." + k: v for k, v in vsd.items()})
    vq = model.representer.vqvae_model
    for k, v in vq.quantizer.state_dict().items():      # EMA buffers of the quantiser
        ckpt.setdefault("representer.vqvae_model.quantizer." + k, v.clone())
    if vq.encoder is None:   # keys of the reference's encoder (not built): accepted and dropped
        ckpt["representer.vqvae_model.encoder.fc_pos.weight"] = torch.zeros(64, 3)
    model.transformer._packed, model.transformer._samplers, vq._engine = object(), {"stale": 1}, object()
    model.load_state_dict(ckpt, strict=True)
    assert model.transformer._packed is None and model.transformer._samplers == {} and vq._engine is None
    assert torch.equal(model.transformer.tok_embs[0].weight, ckpt["transformer.tok_embs.0.weight"])
    assert torch.equal(vq.decoder.fc_out.weight, vsd["decoder.fc_out.weight"])


def test_subpixel_weight_packing_matches_upsample_then_conv():
    """decoder.pack_subpixel_weights (host logic of the sub-pixel convolutions, csrc/conv_tc.cu taps = 8): 8 phases x 8 summed taps on
    the low-resolution input == nearest x2 upsampling followed by the 3x3x3 convolution with zero padding (updown.py:119-132)."""
    import torch.nn.functional as F
    from shapeformer_b200.decoder import pack_subpixel_weights
    g = torch.Generator().manual_seed(0)
    w = torch.randn(5, 3, 3, 3, 3, generator=g)
    x = torch.randn(2, 3, 4, 4, 4, generator=g, dtype=torch.float64)
    ref = F.conv3d(F.interpolate(x, scale_factor=2, mode="nearest"), w.double(), padding=1)
    wp = pack_subpixel_weights(w).double().reshape(2, 2, 2, 2, 2, 2, 5, 3)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for pz in range(2):
        for py in range(2):
            for px in range(2):
                acc = 0
                for tz in range(2):
                    for ty in range(2):
                        for tx in range(2):
                            dz, dy, dx = pz - 1 + tz, py - 1 + ty, px - 1 + tx
                            sl = xp[:, :, 1 + dz:5 + dz, 1 + dy:5 + dy, 1 + dx:5 + dx]
                            acc = acc + torch.einsum("bizyx,oi->bozyx", sl, wp[pz, py, px, tz, ty, tx])
                out[:, :, pz::2, py::2, px::2] = acc
    assert (out - ref).abs().max() < 1e-5


def test_conv_weight_packing_is_the_implicit_gemm_of_conv3d():
    """decoder.pack_conv_weights: out[voxel, co] = sum_{tap, ci} in[voxel + offset(tap), ci] * wp[tap * Cout + co, ci] with
    tap = (dz*3 + dy)*3 + dx and zero padding — the contraction conv3d_tc_kernel runs — equals F.conv3d(padding=1)."""
    import torch.nn.functional as F
    from shapeformer_b200.decoder import pack_conv_weights
    g = torch.Generator().manual_seed(1)
    w = torch.randn(6, 4, 3, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, 4, 5, 5, 5, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, padding=1)
    wp = pack_conv_weights(w).reshape(27, 6, 4)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for tap in range(27):
        dz, dy, dx = tap // 9 - 1, (tap // 3) % 3 - 1, tap % 3 - 1
        sl = xp[:, :, 1 + dz:6 + dz, 1 + dy:6 + dy, 1 + dx:6 + dx]
        out += torch.einsum("bizyx,oi->bozyx", sl, wp[tap])
    assert (out - ref).abs().max() < 1e-12
    w1 = torch.randn(6, 4, 1, 1, 1, generator=g, dtype=torch.float64)
    assert torch.equal(pack_conv_weights(w1), w1.reshape(6, 4))
