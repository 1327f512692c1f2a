"""VQDIF with the reference's constructor and decode API (reference shapeformer/models/vqdif/vqdif.py:20-76).

    pl_model_opt.class: shapeformer_b200.models.vqdif.vqdif.VQDIF
    decoder_opt.class:  shapeformer_b200.models.vqdif.dec.LocalDecoder
    quantizer_opt.class: shapeformer_b200.models.vqdif.quantizer.Quantizer
"""
import torch
import torch.nn as nn

from ...decoder import ImplicitDecoder
from ...xgutils import sysutil


class VQDIF(nn.Module):
    def __init__(self, Xct_as_Xbd=False, encoder_opt=None, decoder_opt=None, quantizer_opt=None, vq_beta=1.,
                 optim_opt=None, ckpt_path=None, opt=None):
        super().__init__()
        self.encoder = None   # LocalPoolPointnet encoder: 'next' row §8f-1
        self.decoder = sysutil.instantiate_from_opt(decoder_opt)
        self.quantizer = sysutil.instantiate_from_opt(quantizer_opt) if quantizer_opt is not None else None
        self.requires_grad_(False)
        self.eval()
        self._engine = None

    @property
    def device(self):
        return self.decoder.fc_out.weight.device

    def load_state_dict(self, state_dict, strict=True):
        own = {k: v for k, v in state_dict.items() if not k.startswith("encoder.")}
        out = super().load_state_dict(own, strict=strict)
        self._engine = None
        return out

    def engine(self):
        if self._engine is None or self._engine.device != self.device:
            self._engine = ImplicitDecoder(self.state_dict(), self.device)
        return self._engine

    def encode(self, Xbd, **kwargs):
        raise NotImplementedError("VQDIF encoder is the 'next' row §8f-1")

    quantize_cloud = encode_quant = encode

    @torch.no_grad()
    def decode(self, grid_feat, Xtg=None, **kwargs):
        """vqdif.py:60-72: grid_feat (B,128,16,16,16), Xtg (B,N,3) in [-1,1] -> {"logits": (B,N,1)}."""
        return self.engine().decode(grid_feat.to(self.device).float().contiguous(), Xtg)

    @torch.no_grad()
    def decode_index(self, code_ind, Xtg):
        """vqdif.py:74-76."""
        return self.engine().decode_index(code_ind, Xtg)
