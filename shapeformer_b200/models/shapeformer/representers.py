"""AR_N representer (reference shapeformer/models/shapeformer/representers.py:53-155,187-196,432-442).

On the sampling path the extra index and the logit mask are computed inside the CUDA kernels (ar_embed_kernel,
ar_sample_kernel); the torch methods here keep the reference's method surface for callers outside the hot loop and are
small host-side index glue, not the product path.  `encode_cloud` / `get_indices` need the VQDIF encoder, which is the
"next" row §8f-1 and not built yet.
"""
import torch
import torch.nn as nn

from ...xgutils import sysutil


def get_next_cond(c_pos_indices, z_pos_indices, end_token):
    """representers.py:432-442."""
    if z_pos_indices.shape[1] == 0:
        return z_pos_indices.clone()
    idx = torch.searchsorted(c_pos_indices.contiguous(), z_pos_indices.contiguous(), right=True)
    ended = z_pos_indices == end_token
    idx = torch.where(ended, torch.full_like(idx, c_pos_indices.shape[1] - 1), idx)
    out = torch.gather(c_pos_indices, 1, idx)
    return torch.where(ended, torch.full_like(out, end_token), out)


class ShapeRepresenter(nn.Module):
    def __init__(self, voxel_res=16, end_tokens=None, input_end_tokens=None, block_size=None, uncond=False,
                 no_val_ind=False, vqvae_opt=None, cloud_shrinkage=1., random_cind_masking=False, mask_invalid=True,
                 mask_invalid_completion=False):
        super().__init__()
        self.voxel_res, self.end_tokens = voxel_res, end_tokens
        self.input_end_tokens = end_tokens if input_end_tokens is None else input_end_tokens
        self.block_size, self.max_length = block_size, block_size // 2
        self.uncond, self.no_val_ind, self.cloud_shrinkage = uncond, no_val_ind, cloud_shrinkage
        self.random_cind_masking = random_cind_masking
        self.mask_invalid, self.mask_invalid_completion = mask_invalid, mask_invalid_completion
        self.vqvae_model = self.init_trained_model_from_ckpt(vqvae_opt)

    def init_trained_model_from_ckpt(self, config):
        """representers.py:34-48: instantiate the frozen VQDIF named by `class` and load `ckpt_path` if given."""
        if not config or not config.get("class"):
            return None
        model = sysutil.load_object(config["class"])(**config.get("kwargs", {}))
        if config.get("ckpt_path"):
            ckpt = torch.load(config["ckpt_path"], map_location="cpu")
            model.load_state_dict(ckpt.get("state_dict", ckpt), strict=False)
        return model.eval().requires_grad_(False)

    @torch.no_grad()
    def encode_cloud(self, cloud):
        """representers.py:68-78: cloud (B,T,3) -> (quant_feat, quant_ind, mode, sparse_unpacked (B,L,2))."""
        vq = self.vqvae_model
        if vq is None:
            raise NotImplementedError("representer was built without a VQDIF (vqvae_opt)")
        out = vq.point_encoder().quantize_cloud(cloud * self.cloud_shrinkage, max_length=self.max_length,
                                                end_tokens=self.input_end_tokens)
        sparse = out["c_indices"]
        if self.no_val_ind:
            sparse[:, :, -1] *= 0
        return None, out["quant_ind"], out["empty_index"], sparse

    @torch.no_grad()
    def get_indices(self, Xct, Xbd=None, stage="test", **kwargs):
        """representers.py:80-103 for inference (stage != 'train': no random conditioning masking)."""
        if stage == "train" and self.random_cind_masking:
            raise NotImplementedError("training-time random conditioning masking is outside the B200 hot path")
        _, _, mode1, c_indices = self.encode_cloud(Xct)
        z_indices = c_indices[:, :0, :] if Xbd is None else self.encode_cloud(Xbd)[3]
        if self.uncond:
            B, _, tn = c_indices.shape
            c_indices = torch.tensor(list(self.input_end_tokens), dtype=torch.int64, device=c_indices.device).repeat(B, 1, 1)
        others = dict(empty_index=mode1, origin_c_indices=c_indices, origin_z_indices=z_indices)
        extra = self.get_extra_indices(c_indices, z_indices)
        c_indices, z_indices = self.convert_input_indices(c_indices, z_indices)
        return c_indices, z_indices, extra, others

    def get_extra_indices(self, c_indices, z_indices):
        cz = torch.cat([c_indices, z_indices], 1)
        return torch.zeros(cz.shape[0], cz.shape[1], 1, dtype=cz.dtype, device=cz.device)

    def convert_input_indices(self, c_indices, z_indices):
        return c_indices, z_indices

    def convert_output_indices(self, indices):
        return indices

    def sampling_masker(self, logits, idx, extra_idx=None, L_cond=None, step_j=None, tuple_i=None):
        """representers.py:120-155 (torch restatement for API compatibility; the sampler uses the fused kernel)."""
        out = logits.clone()
        end = self.end_tokens
        if tuple_i == 1:
            m = idx[:, -1, 0] == end[0]
            out[m, :] = float("-inf")
            out[m, end[1]] = 1.0
            return out
        last = idx[:, -2, 0]
        v = torch.arange(out.shape[-1], device=idx.device, dtype=idx.dtype)[None]
        if self.mask_invalid and step_j > 0:
            bad = v <= last[:, None]
            bad[:, end[0]] = False
            out[bad] = float("-inf")
        if self.mask_invalid_completion:
            cond = torch.cat([idx[:, :L_cond, 0], torch.full((idx.shape[0], 1), end[0] + 1, dtype=idx.dtype,
                                                             device=idx.device)], 1).contiguous()
            nxt = torch.gather(cond, 1, torch.searchsorted(cond, last[:, None].contiguous(), right=True))
            out[v > nxt] = float("-inf")
        return out


class AR(ShapeRepresenter):
    pass


class AR_N(ShapeRepresenter):
    def get_extra_indices(self, c_indices, z_indices):
        z_extra = get_next_cond(c_indices[..., 0], z_indices[..., 0], self.end_tokens[0])
        return torch.cat([c_indices[..., 0].clone(), z_extra], 1)[..., None]
