"""Builders for the reference's OWN modules (through oracle/ref_shim.py) with synthetic weights.

TEST / BENCH INFRASTRUCTURE ONLY: used by tests/ (oracle pinning, fixture generation) and by bench.py's CPU legs.  Usable
where the reference tree is importable (``ref_shim.available()``): /root/reference in the build container, oracle/_ref on
the GPU box."""
import torch

from . import ref_shim


def have_reference():
    return ref_shim.available()


def ref_gpt(cfg, sd):
    m = ref_shim.reference_modules()
    gpt = m["mingpt"].CondTupleGPT(vocab_sizes=cfg["vocab_sizes"], extra_vocab_sizes=cfg["extra_vocab_sizes"],
                                   block_size=cfg["block_size"], tuple_n=2, n_layers=cfg["n_layers"],
                                   n_head=cfg["n_head"], n_embd=cfg["n_embd"])
    missing, unexpected = gpt.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("attn.mask") for k in missing), (missing, unexpected)
    return gpt.eval()


def ref_shapeformer(cfg, sd, mask_invalid=True, mask_invalid_completion=True, end_tokens=(4096, 4096)):
    """The reference ShapeFormer (LightningModule stub) with an AR_N representer whose frozen VQDIF is not needed."""
    m = ref_shim.reference_modules()
    rep_mod = m["representers"]
    rep_mod.Representer.init_trained_model_from_ckpt = lambda self, config: None
    pre = "shapeformer.models.shapeformer."
    sf = m["shapeformer"].ShapeFormer(
        tuple_n=2, block_size=cfg["block_size"], end_tokens=list(end_tokens), vocab_sizes=list(cfg["vocab_sizes"]),
        extra_vocab_sizes=list(cfg["extra_vocab_sizes"]), voxel_res=16,
        transformer_opt={"class": pre + "transformer.mingpt.CondTupleGPT",
                         "kwargs": dict(tuple_n=2, vocab_sizes=list(cfg["vocab_sizes"]),
                                        extra_vocab_sizes=list(cfg["extra_vocab_sizes"]), n_layers=list(cfg["n_layers"]),
                                        block_size=cfg["block_size"], n_head=cfg["n_head"], n_embd=cfg["n_embd"])},
        representer_opt={"class": pre + "representers.AR_N",
                         "kwargs": dict(voxel_res=16, uncond=False, no_val_ind=False, block_size=cfg["block_size"],
                                        end_tokens=list(end_tokens), mask_invalid=mask_invalid,
                                        mask_invalid_completion=mask_invalid_completion, vqvae_opt={})})
    sf.transformer.load_state_dict(sd, strict=False)
    return sf.eval()


def ref_vqdif_decoder(sd):
    """Reference LocalDecoder + Quantizer loaded from a synthetic VQDIF state dict."""
    m = ref_shim.reference_modules()
    dec = m["dec"].LocalDecoder(sample_mode="bilinear", hidden_size=32, c_dim=32, unet3d=True,
                                unet3d_kwargs=dict(num_levels=3, f_maps=128, in_channels=128, out_channels=128),
                                upsampler=True, upsampler_kwargs=dict(in_channels=128, upsampler_steps=2))
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    q = m["quantizer"].Quantizer(vocab_size=4096, n_embd=128)
    q.embedding.weight.data.copy_(sd["quantizer.embedding.weight"])
    return dec.eval(), q.eval()


def ref_vqdif_encoder(sd):
    """Reference LocalPoolPointnet (shipped shapenet_res16 kwargs) loaded from a synthetic VQDIF state dict; its torch_scatter
    calls run through the scatter_reduce stand-in of oracle/ref_shim.py."""
    ref_shim.reference_modules()
    import importlib
    enc_mod = importlib.import_module("shapeformer.models.vqdif.enc")
    enc = enc_mod.LocalPoolPointnet(hidden_dim=32, plane_type="grid", grid_resolution=64, c_dim=32, downsampler=True,
                                    downsampler_kwargs=dict(in_channels=32, downsample_steps=2))
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    return enc.eval()


def ref_decode_index(dec, q, code_ind, Xtg):
    """VQDIF.decode_index restated with the reference's own modules (vqdif/vqdif.py:60-76)."""
    with torch.no_grad():
        return dec(Xtg / 2.0, q.get_code(code_ind))
