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


class _HistoryStreamer:
    """The reference returns the masked logits of every step on the CPU (shapeformer.py:94,118): 16.8 MB per row, 8.6 GB for a
    512-row batch.  Instead of one big device->host copy + host copy after the loop, every chunk of steps travels while the NEXT
    chunk is being computed: a side stream copies the chunk's slab into a pinned staging ring as soon as the chunk has finished
    on the GPU, and the host thread — idle while the GPU works — moves the previous chunk from the ring into the fresh result
    tensors.  Only the last chunk's transfer is exposed."""
    _ring = {}          # (B, chunk, V tuple) -> [pinned (2 slots) per tuple element], reused across calls (never returned)
    _side = {}          # device index -> copy stream

    def __init__(self, sampler, B, max_steps):
        self.s, self.B = sampler, B
        self.V = list(sampler.spec["vocab_sizes"])
        self.dev = sampler.device
        self.views = sampler.history_views(B)
        self.out = [torch.empty(B, min(int(max_steps), sampler.max_steps), v, dtype=torch.float32) for v in self.V]     # fresh, pageable
        key = (B, sampler.chunk_steps, tuple(self.V))
        ring = _HistoryStreamer._ring.get(key)
        if ring is None:
            _HistoryStreamer._ring.clear()
            ring = [[torch.empty(B, sampler.chunk_steps, v, dtype=torch.float32).pin_memory() for _ in range(2)] for v in self.V]
            _HistoryStreamer._ring[key] = ring
        self.ring = ring
        idx = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        if idx not in _HistoryStreamer._side:
            _HistoryStreamer._side[idx] = torch.cuda.Stream(device=self.dev)
        self.side = _HistoryStreamer._side[idx]
        self.pending = []     # (first_step, n, slot, event)
        self.k = 0

    def _drain(self, upto):
        while len(self.pending) > upto:
            first, n, slot, ev = self.pending.pop(0)
            ev.synchronize()
            for i in range(len(self.V)):
                self.out[i][:, first:first + n].copy_(self.ring[i][slot][:, :n])

    def on_chunk(self, first, n):
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())          # fires when this chunk's steps have run
        self._drain(1)                                     # the slot about to be reused must have been moved out
        slot = self.k % 2
        self.k += 1
        with torch.cuda.stream(self.side):
            self.side.wait_event(done)
            for i in range(len(self.V)):
                self.ring[i][slot][:, :n].copy_(self.views[i][:, first:first + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.pending.append((first, n, slot, ev))
        self._drain(1)                                     # move the PREVIOUS chunk while the GPU runs this one

    def finish(self, steps):
        self._drain(0)
        return [o[:, :steps] for o in self.out]


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
        stream_hist = keep and self.history_device == "cpu" and not self.zero_copy_outputs
        with torch.cuda.device(self.device):
            streamer = _HistoryStreamer(s, B, max_steps) if stream_hist else None
            x, hist = self._run_sampler(s, c_indices, max_steps, top_k, top_p, temperature, best_in_first, rep, noise,
                                        generator, streamer.on_chunk if streamer else None)
            if streamer:
                hist = streamer.finish(x.shape[1])
        if hist is None:
            hist = [None] * tuple_n
        elif streamer:
            pass
        elif self.history_device == "cpu":
            with torch.cuda.device(self.device):
                hist = [self._to_host(h, i) for i, h in enumerate(hist)]
        elif not self.zero_copy_outputs:
            hist = [h.clone() for h in hist]
        return (x if self.zero_copy_outputs else x.clone()), hist

    @staticmethod
    def _run_sampler(s, c_indices, max_steps, top_k, top_p, temperature, best_in_first, rep, noise, generator, on_chunk=None):
        return s.sample(c_indices, max_steps, top_k=top_k, top_p=top_p, temperature=temperature,
                           best_in_first=best_in_first, mask_invalid=getattr(rep, "mask_invalid", True),
                           mask_invalid_completion=getattr(rep, "mask_invalid_completion", False), noise=noise,
                           generator=generator, on_chunk=on_chunk)

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
