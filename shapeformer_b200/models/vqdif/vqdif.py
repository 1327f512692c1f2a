"""VQDIF with the reference's constructor and decode API (reference shapeformer/models/vqdif/vqdif.py:20-76).

    pl_model_opt.class: shapeformer_b200.models.vqdif.vqdif.VQDIF
    decoder_opt.class:  shapeformer_b200.models.vqdif.dec.LocalDecoder
    quantizer_opt.class: shapeformer_b200.models.vqdif.quantizer.Quantizer
"""
import torch
import torch.nn as nn

from ...decoder import ImplicitDecoder
from ...encoder import PointEncoder
from ...xgutils import sysutil


class VQDIF(nn.Module):
    def __init__(self, Xct_as_Xbd=False, encoder_opt=None, decoder_opt=None, quantizer_opt=None, vq_beta=1.,
                 optim_opt=None, ckpt_path=None, opt=None):
        super().__init__()
        self.encoder = sysutil.instantiate_from_opt(encoder_opt) if encoder_opt else None
        self.decoder = sysutil.instantiate_from_opt(decoder_opt)
        self.quantizer = sysutil.instantiate_from_opt(quantizer_opt) if quantizer_opt is not None else None
        self.requires_grad_(False)
        self.eval()
        self._engine = None
        self._point_encoder = None
        # nested loads (a parent ShapeFormer checkpoint carries representer.vqvae_model.encoder.*) go through
        # _load_from_state_dict, so the key filter and the engine invalidation are hooks, not a load_state_dict override
        self._register_load_state_dict_pre_hook(self._drop_unbuilt)
        self.register_load_state_dict_post_hook(self._invalidate)

    def _drop_unbuilt(self, state_dict, prefix, *args):
        if self.encoder is None:
            for k in [k for k in state_dict if k.startswith(prefix + "encoder.")]:
                del state_dict[k]

    @staticmethod
    def _invalidate(module, incompatible_keys):
        module._engine = None
        module._point_encoder = None

    @property
    def device(self):
        return self.decoder.fc_out.weight.device

    def engine(self):
        if self._engine is None or self._engine.device != self.device:
            self._engine = ImplicitDecoder(self.state_dict(), self.device)
        return self._engine

    def point_encoder(self):
        if self.encoder is None or self.quantizer is None:
            raise NotImplementedError("this VQDIF was built without encoder_opt / quantizer_opt")
        if self._point_encoder is None or self._point_encoder.device != self.device:
            self._point_encoder = PointEncoder(self.state_dict(), self.device)
        return self._point_encoder

    @torch.no_grad()
    def encode(self, Xbd, **kwargs):
        """vqdif.py:36-38: Xbd (B,T,3) in [-1,1] -> (grid_feat (B,128,16,16,16), grid_mask (B,16,16,16) bool)."""
        _, mask, feat = self.point_encoder().encode_quant(Xbd, return_feat=True)
        return feat, mask

    @torch.no_grad()
    def encode_quant(self, Xbd, **kwargs):
        """vqdif.py:40-48 (eval): quant_feat = the selected code vectors, quant_ind, grid_mask (quant_diff is a training loss)."""
        raw, mask = self.point_encoder().encode_quant(Xbd)
        return dict(quant_feat=self.engine().get_code(raw), quant_ind=raw, quant_diff=None, grid_mask=mask)

    @torch.no_grad()
    def quantize_cloud(self, cloud):
        """vqdif.py:50-58: -> (quant_ind with the batch mode in unoccupied cells, mode, encoded)."""
        out = self.point_encoder().quantize_cloud(cloud)
        encoded = dict(quant_feat=None, quant_ind=out["raw_ind"], quant_diff=None, grid_mask=out["mask"])
        return out["quant_ind"], out["mode"], encoded

    @torch.no_grad()
    def decode(self, grid_feat, Xtg=None, **kwargs):
        """vqdif.py:60-72: grid_feat (B,128,16,16,16), Xtg (B,N,3) in [-1,1] -> {"logits": (B,N,1)}."""
        return self.engine().decode(grid_feat.to(self.device).float().contiguous(), Xtg)

    @torch.no_grad()
    def decode_index(self, code_ind, Xtg):
        """vqdif.py:74-76."""
        return self.engine().decode_index(code_ind, Xtg)
