"""ShapeFormer with the reference's constructor and sampling API (reference
shapeformer/models/shapeformer/shapeformer.py:16-130,382-391), the AR loop running in libsfb200.

    pl_model_opt.class:   shapeformer_b200.models.shapeformer.shapeformer.ShapeFormer
    transformer_opt.class: shapeformer_b200.models.shapeformer.transformer.mingpt.CondTupleGPT
    representer_opt.class: shapeformer_b200.models.shapeformer.representers.AR_N
"""
import numpy as np
import torch
import torch.nn as nn

from ...xgutils import sysutil


class ShapeFormer(nn.Module):
    def __init__(self, tuple_n=None, block_size=None, end_tokens=None, vocab_sizes=None, extra_vocab_sizes=None,
                 voxel_res=16, transformer_opt=None, representer_opt=None, optim_opt=None):
        super().__init__()
        self.tuple_n, self.block_size, self.end_tokens = tuple_n, block_size, end_tokens
        self.vocab_sizes, self.extra_vocab_sizes, self.voxel_res = vocab_sizes, extra_vocab_sizes, voxel_res
        self.transformer = sysutil.instantiate_from_opt(transformer_opt)
        self.representer = sysutil.instantiate_from_opt(representer_opt)
        assert "TupleGPT" in transformer_opt["class"]
        self.history_device = "cpu"   # the reference returns the logits history on the CPU (shapeformer.py:94)
        # False (default): sample()/sample_indices() return FRESH tensors like the reference.  True: `x` and the history are
        # views of the sampler's persistent buffers (device token buffer / pinned host staging) — valid only until the next
        # sample() call of this model; opt-in for callers that consume them immediately (saves a 1 GB host copy at 64 rows).
        self.zero_copy_outputs = False
        self.eval()

    @property
    def device(self):
        return self.transformer.pos_emb.device

    @torch.no_grad()
    def sample_indices(self, c_indices, z_indices, max_steps, sample=False, best_in_first=False, top_k=100, top_p=.8,
                       temperature=1.0, mask_invalid=True, mask_invalid_completion=False, callback=lambda k: None,
                       noise=None, generator=None):
        """Same contract as the reference: returns (x (B, steps, 2) int64 on the model's device, [hist0, hist1] fp32
        (B, steps, V) on the CPU).  Like the reference (App. C-2) the `mask_invalid*`, `sample` and `callback` arguments are
        ignored: masking is governed by the representer's attributes.  `noise` / `generator` are extensions: explicit
        Exp(1) draws (parity tests) or a torch.Generator for the device RNG."""
        assert not self.transformer.training
        if z_indices.shape[1] != 0:
            raise NotImplementedError("a non-empty generated prefix (z_indices) is not on the B200 path yet")
        B, L_c, tuple_n = c_indices.shape
        rep = self.representer
        keep = self.history_device is not None
        s = self.transformer.sampler(B, L_c, max_steps, self.end_tokens, keep_history=keep)
        with torch.cuda.device(self.device):
            x, hist = self._run_sampler(s, c_indices, max_steps, top_k, top_p, temperature, best_in_first, rep, noise,
                                        generator)
        if hist is None:
            hist = [None] * tuple_n
        elif self.history_device == "cpu":
            with torch.cuda.device(self.device):
                hist = [self._to_host(h, i) for i, h in enumerate(hist)]
            if not self.zero_copy_outputs:
                hist = [h.clone() for h in hist]
        elif not self.zero_copy_outputs:
            hist = [h.clone() for h in hist]
        return (x if self.zero_copy_outputs else x.clone()), hist

    @staticmethod
    def _run_sampler(s, c_indices, max_steps, top_k, top_p, temperature, best_in_first, rep, noise, generator):
        return s.sample(c_indices, max_steps, top_k=top_k, top_p=top_p, temperature=temperature,
                           best_in_first=best_in_first, mask_invalid=getattr(rep, "mask_invalid", True),
                           mask_invalid_completion=getattr(rep, "mask_invalid_completion", False), noise=noise,
                           generator=generator)

    _pinned = {}

    def _to_host(self, t, slot):
        """One D2H copy of the (B, steps, V) history slab of tuple element `slot` into a reusable pinned staging buffer
        (the reference pays a synchronising .cpu() per sub-step)."""
        key = (slot, tuple(t.shape), t.dtype)
        buf = ShapeFormer._pinned.get(key)
        if buf is None:
            for k in [k for k in ShapeFormer._pinned if k[0] == slot]:
                del ShapeFormer._pinned[k]          # one staging buffer per tuple element
            buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            ShapeFormer._pinned[key] = buf
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return buf

    @torch.no_grad()
    def sample(self, **sampling_kwargs):
        x, logits_history = self.sample_indices(**sampling_kwargs)
        return self.representer.convert_output_indices(x), x, logits_history


def decode_sample_indices(decoder, Xtg, voxel_vqind):
    """shapeformer.py:382-391: (16,16,16) numpy code grid + (N,3) numpy query points -> (N,) numpy occupancy."""
    Xtg_t = torch.from_numpy(np.asarray(Xtg)[None, ...]).to(decoder.device)
    ind = torch.from_numpy(np.asarray(voxel_vqind))[None, ...].long().to(decoder.device)
    with torch.no_grad():
        logits = decoder.decode_index(ind, Xtg=Xtg_t)["logits"]
        return torch.sigmoid(logits)[0, ..., 0].cpu().numpy()


def compute_log_probs(samples, logits_history):
    """shapeformer.py:407-418: log-softmax of the stored logits gathered at the sampled tokens -> (S, L, tuple_n)."""
    out = np.zeros(samples.shape)
    for ti in range(samples.shape[-1]):
        h = torch.as_tensor(np.asarray(logits_history[ti])).double()
        lp = torch.log_softmax(h, -1)
        out[..., ti] = torch.gather(lp, 2, torch.as_tensor(np.asarray(samples[..., ti]))[..., None])[..., 0].numpy()
    return out
