"""Host side of the VQDIF encoder path: partial point cloud -> code grid -> (pos, val) conditioning tuples.

    VQDIF.encode / encode_quant / quantize_cloud   reference shapeformer/models/vqdif/vqdif.py:36-58
    AR_N.encode_cloud / get_indices                reference shapeformer/models/shapeformer/representers.py:68-103

Every arithmetic step runs in libsfb200 (csrc/enc_kernels.cu); PyTorch owns the buffers.  No CPU fallback."""
import ctypes

import torch

from . import _lib

CHUNK = 8      # clouds per library call (about 150 MB of workspace each)


class PointEncoder:
    """`sd`: VQDIF state dict (encoder.*, quantizer.embedding.weight)."""

    def __init__(self, sd, device, prefix="encoder.", codebook_key="quantizer.embedding.weight"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.Sfb200Error("PointEncoder needs a CUDA device (no CPU fallback)")
        g = lambda k: sd[prefix + k].detach().to(self.device, torch.float32).contiguous()
        if tuple(sd[prefix + "fc_pos.weight"].shape) != (64, 3) or tuple(sd[prefix + "fc_c.weight"].shape) != (32, 32):
            raise _lib.Sfb200Error("only the shipped LocalPoolPointnet (hidden = c_dim = 32, 64^3 grid, 2 downsample steps)")
        self.keep = []      # device tensors the struct points into
        w = _lib.EncWeights()

        def put(name, t):
            self.keep.append(t)
            setattr(w, name, ctypes.cast(ctypes.c_void_p(t.data_ptr()), _lib.c_f32p))

        def put_i(name, i, t):
            self.keep.append(t)
            getattr(w, name)[i] = ctypes.cast(ctypes.c_void_p(t.data_ptr()), _lib.c_f32p)

        put("fc_pos_w", g("fc_pos.weight")); put("fc_pos_b", g("fc_pos.bias"))
        for i in range(5):
            put_i("fc0_w", i, g(f"blocks.{i}.fc_0.weight")); put_i("fc0_b", i, g(f"blocks.{i}.fc_0.bias"))
            put_i("fc1_w", i, g(f"blocks.{i}.fc_1.weight")); put_i("fc1_b", i, g(f"blocks.{i}.fc_1.bias"))
            put_i("sc_w", i, g(f"blocks.{i}.shortcut.weight"))
        put("fcc_w", g("fc_c.weight")); put("fcc_b", g("fc_c.bias"))
        want = [(64, 32, 2), (64, 64, 1), (128, 64, 2), (128, 128, 1)]
        for i, (co, ci, k) in enumerate(want):
            cw = g(f"downsampler.blocks.{i}.conv.weight")
            if tuple(cw.shape) != (co, ci, k, k, k):
                raise _lib.Sfb200Error(f"unexpected Downsampler conv {i} shape {tuple(cw.shape)}")
            put_i("ds_wT", i, cw.permute(2, 3, 4, 1, 0).contiguous().view(k ** 3 * ci, co))
            put_i("ds_gn_w", i, g(f"downsampler.blocks.{i}.groupnorm.weight"))
            put_i("ds_gn_b", i, g(f"downsampler.blocks.{i}.groupnorm.bias"))
        cb = sd[codebook_key].detach().to(self.device, torch.float32).contiguous()
        if cb.shape[1] != 128:
            raise _lib.Sfb200Error("codebook dimension must be 128")
        put("codebook", cb)
        w.n_codes = cb.shape[0]
        self.w, self.n_codes = w, cb.shape[0]
        self._ws = None

    def _workspace(self, B, T):
        need = self.lib.sfb200_encoder_workspace_bytes(B, T, self.n_codes)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def encode_quant(self, Xbd, return_feat=False):
        """VQDIF.encode + Quantizer.forward: Xbd (B, T, 3) in [-1, 1] -> raw_ind (B,16,16,16) int64, mask (B,16,16,16) bool
        (and grid_feat (B,128,16,16,16) when return_feat)."""
        with torch.cuda.device(self.device):
            x = Xbd.to(self.device, torch.float32).contiguous()
            B, T, _ = x.shape
            raw = torch.empty(B, 16, 16, 16, dtype=torch.int64, device=self.device)
            mask = torch.empty(B, 4096, dtype=torch.uint8, device=self.device)
            feat = torch.empty(B, 128, 16, 16, 16, dtype=torch.float32, device=self.device) if return_feat else None
            for b0 in range(0, B, CHUNK):
                n = min(CHUNK, B - b0)
                ws = self._workspace(n, T)
                _lib.check(self.lib.sfb200_encode_cloud(ctypes.byref(self.w), _lib.ptr(x[b0:b0 + n]), n, T, _lib.ptr(ws),
                                                        _lib.ptr(raw[b0:b0 + n]), _lib.ptr(mask[b0:b0 + n]),
                                                        _lib.ptr(feat[b0:b0 + n]) if return_feat else None, _lib.stream_ptr()),
                           "sfb200_encode_cloud")
            mask = mask.view(B, 16, 16, 16).bool()
            return (raw, mask, feat) if return_feat else (raw, mask)

    def quantize_cloud(self, cloud, max_length=406, end_tokens=(4096, 4096)):
        """VQDIF.quantize_cloud + batch_dense2sparse: -> dict(quant_ind (B,16,16,16) with the batch mode in empty cells, mode,
        raw_ind, mask, c_indices (B, L, 2) end-padded tuples, empty_index)."""
        with torch.cuda.device(self.device):
            raw, mask = self.encode_quant(cloud)
            B = raw.shape[0]
            dense = torch.empty_like(raw)
            tokens = torch.empty(B, max_length, 2, dtype=torch.int64, device=self.device)
            lengths = torch.empty(B, dtype=torch.int32, device=self.device)
            modes = torch.empty(2, dtype=torch.int64, device=self.device)
            hist = torch.empty(self.n_codes, dtype=torch.int32, device=self.device)
            m8 = mask.view(B, 4096).to(torch.uint8).contiguous()
            _lib.check(self.lib.sfb200_dense_to_tokens(_lib.ptr(raw), _lib.ptr(m8), B, 4096, self.n_codes, max_length,
                                                       int(end_tokens[0]), int(end_tokens[1]), _lib.ptr(hist), _lib.ptr(dense),
                                                       _lib.ptr(tokens), _lib.ptr(lengths), _lib.ptr(modes), _lib.stream_ptr()),
                       "sfb200_dense_to_tokens")
            longest = int(lengths.max())                  # one small D2H sync per batch: the tuple count is data dependent
            L = longest + 1
            if L > max_length:                            # unpack_sparse's crop: the last tuple is forced to the end tokens
                L = max_length
                tokens[:, L - 1, 0], tokens[:, L - 1, 1] = int(end_tokens[0]), int(end_tokens[1])
            return dict(quant_ind=dense, mode=modes[0], raw_ind=raw, mask=mask, c_indices=tokens[:, :L].contiguous(),
                        empty_index=modes[1])
