"""Host side of the KV-cached autoregressive sampler: owns the device buffers (packed weights, KV cache, workspace, token
buffer, logits history) as torch tensors and drives libsfb200's sfb200_ar_* entry points.

Replaces the body of ShapeFormer.sample_indices (reference shapeformer/models/shapeformer/shapeformer.py:54-123).
PyTorch is used for memory, streams and RNG only; every arithmetic step runs in the CUDA library.
"""
import ctypes
import os

import torch

from . import _lib


def _cfg_struct(spec, end_tokens, max_rows, max_len, max_steps, prefill_rows, max_cond, keep_history):
    c = _lib.ArConfig()
    c.n_embd, c.n_head = spec["n_embd"], spec["n_head"]
    c.n_layers[0], c.n_layers[1] = spec["n_layers"]
    c.block_size = spec["block_size"]
    c.vocab[0], c.vocab[1] = spec["vocab_sizes"]
    c.extra_vocab = spec["extra_vocab_sizes"][0]
    c.end_tokens[0], c.end_tokens[1] = int(end_tokens[0]), int(end_tokens[1])
    c.max_rows, c.max_len, c.max_steps = max_rows, max_len, max_steps
    c.prefill_rows, c.max_cond, c.keep_history = prefill_rows, max_cond, int(keep_history)
    return c


# reference state_dict key -> (tensor id, how) for one transformer block
_BLOCK_KEYS = [
    ("ln1.weight", _lib.W_LN1_W), ("ln1.bias", _lib.W_LN1_B), ("attn.proj.weight", _lib.W_PROJ_W),
    ("attn.proj.bias", _lib.W_PROJ_B), ("ln2.weight", _lib.W_LN2_W), ("ln2.bias", _lib.W_LN2_B),
    ("mlp.0.weight", _lib.W_FC1_W), ("mlp.0.bias", _lib.W_FC1_B), ("mlp.2.weight", _lib.W_FC2_W),
    ("mlp.2.bias", _lib.W_FC2_B),
]


def pack_gpt_weights(sd, spec, device):
    """Copy a CondTupleGPT state_dict (keys as SURVEY.md App. A-4, no prefix) into the library's packed fp32 layout
    (sfb200_ar_weight_offset).  query/key/value are concatenated into one (3d, d) matrix."""
    lib = _lib.load()
    cfg = _cfg_struct(spec, (0, 0), 1, 2, 1, 1, 1, 0)
    total = lib.sfb200_ar_weight_floats(ctypes.byref(cfg))
    if total <= 0:
        raise _lib.Sfb200Error("unsupported CondTupleGPT shape (need n_embd == 64 * n_head, vocab <= 8192)")
    blob = torch.empty(total, dtype=torch.float32, device=device)

    def put(t, tid, g=0, l=0):
        off = lib.sfb200_ar_weight_offset(ctypes.byref(cfg), tid, g, l)
        assert off >= 0
        t = t.detach().to(device=device, dtype=torch.float32).reshape(-1)
        blob[off:off + t.numel()].copy_(t)

    put(sd["pos_emb"], _lib.W_POS_EMB)
    put(sd["cond_pos_emb"], _lib.W_COND_POS_EMB)
    put(sd["tok_embs.0.weight"], _lib.W_TOK_EMB0)
    put(sd["tok_embs.1.weight"], _lib.W_TOK_EMB1)
    put(sd["extra_tok_embs.0.weight"], _lib.W_EXTRA_EMB)
    for g, nl in enumerate(spec["n_layers"]):
        put(sd[f"heads.{g}.0.weight"], _lib.W_HEAD_LN_W, g)
        put(sd[f"heads.{g}.0.bias"], _lib.W_HEAD_LN_B, g)
        put(sd[f"heads.{g}.1.weight"], _lib.W_HEAD_W, g)
        for l in range(nl):
            p = f"blocks.{g}.{l}."
            for key, tid in _BLOCK_KEYS:
                put(sd[p + key], tid, g, l)
            put(torch.cat([sd[p + "attn.query.weight"], sd[p + "attn.key.weight"], sd[p + "attn.value.weight"]], 0),
                _lib.W_QKV_W, g, l)
            put(torch.cat([sd[p + "attn.query.bias"], sd[p + "attn.key.bias"], sd[p + "attn.value.bias"]], 0),
                _lib.W_QKV_B, g, l)
    return blob


class ARSampler:
    """KV-cached sampler for up to `max_rows` rows.  Buffers are allocated once and reused across batches."""

    def __init__(self, weights_blob, spec, end_tokens=(4096, 4096), max_rows=1, max_cond=406, max_steps=512,
                 keep_history=True, prefill_tokens=8192, chunk_steps=32, device=None, pretile=True):
        self.lib = _lib.load()
        self.spec, self.end_tokens = dict(spec), tuple(int(e) for e in end_tokens)
        self.device = device or weights_blob.device
        self.max_rows, self.max_cond, self.max_steps = max_rows, max_cond, max_steps
        # the last token / K,V / pos_emb index touched is L_cond + max_steps - 1.  The reference's call site uses
        # max_steps = 512 with block_size 812 and L_cond up to 406: the capacity is clamped to block_size and sample()
        # raises only if a batch actually reaches the limit before every row has ended (the reference's overflow crop,
        # shapeformer.py:73-76, is buggy — SURVEY.md App. C-3 — and is not reproduced).
        self.max_len = min(max_cond + max_steps, spec["block_size"])
        if max_cond >= self.max_len:
            raise _lib.Sfb200Error(f"L_cond {max_cond} leaves no room to generate within block_size {spec['block_size']}")
        self.keep_history = bool(keep_history)
        self.chunk_steps = int(chunk_steps)
        prefill_rows = max(1, min(max_rows, prefill_tokens // max(1, max_cond)))
        self.cfg = _cfg_struct(spec, self.end_tokens, max_rows, self.max_len, max_steps, prefill_rows, max_cond,
                               self.keep_history)
        self.Vmax = max(spec["vocab_sizes"])
        self._pretile = bool(pretile)
        dev = self.device
        self.weights = weights_blob
        with torch.cuda.device(dev):
            self._allocate(dev)

    def _allocate(self, dev):
        self.kv = torch.empty(self.lib.sfb200_ar_kv_bytes(ctypes.byref(self.cfg)), dtype=torch.uint8, device=dev)
        self.ws = torch.empty(self.lib.sfb200_ar_workspace_bytes(ctypes.byref(self.cfg)), dtype=torch.uint8, device=dev)
        self.tokens = torch.zeros(self.max_rows, self.max_len, 2, dtype=torch.int64, device=dev)
        nh = self.lib.sfb200_ar_history_floats(ctypes.byref(self.cfg))
        self.hist = torch.empty(max(nh, 1), dtype=torch.float32, device=dev)
        self._noise = {}
        self.status_host = torch.empty(8, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None
        h = ctypes.c_void_p()
        _lib.check(self.lib.sfb200_ar_create(ctypes.byref(self.cfg), _lib.ptr(self.weights), _lib.ptr(self.kv),
                                             _lib.ptr(self.ws), _lib.ptr(self.tokens),
                                             _lib.ptr(self.hist) if self.keep_history else None, ctypes.byref(h)),
                   "sfb200_ar_create")
        self.handle = h
        self._status_ptr = self.lib.sfb200_ar_status_ptr(self.handle)
        off = (self.lib.sfb200_ar_logprob_ptr(self.handle) - self.ws.data_ptr()) // 4
        self.logp = self.ws.view(torch.float32)[off:off + self.max_rows * self.max_steps * 2].view(
            self.max_rows, self.max_steps, 2)
        self.last_log_prob = None
        # decode steps of <= 64 rows run the persistent GEMM-chain kernel straight from the fp32 blob (csrc/ar_chain.cu).  Only
        # with SFB200_CHAIN=0 (per-GEMM kernels) do batches of 9..64 rows use pre-split TF32 weight tiles (2x the weight bytes)
        self.pretiled = None
        chain = os.environ.get("SFB200_CHAIN", "1")[:1] != "0"
        if self._pretile and not chain and 9 <= self.max_rows <= 64:
            n = self.lib.sfb200_ar_pretiled_floats(ctypes.byref(self.cfg))
            self.pretiled = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.check(self.lib.sfb200_ar_set_pretiled(self.handle, _lib.ptr(self.pretiled), _lib.stream_ptr()),
                       "sfb200_ar_set_pretiled")

        # GEMMs over more than 64 rows (prefill of >= 1 shape, decode batches of > 64 rows) run the TMA-fed kernel of
        # csrc/tc_big.cu, which reads the low parts of the TF32 operand split from memory: one extra copy of the weight blob,
        # shared by every sampler built on the same packed weights (SFB200_BIG=0 keeps the in-kernel-split GEMM)
        self.weights_lo = None
        if os.environ.get("SFB200_BIG", "1")[:1] != "0":
            lo = getattr(self.weights, "_sfb200_lo", None)
            if lo is None:
                lo = torch.empty_like(self.weights)
                try:
                    self.weights._sfb200_lo = lo
                except Exception:
                    pass
            self.weights_lo = lo
            _lib.check(self.lib.sfb200_ar_set_lo_weights(self.handle, _lib.ptr(lo), _lib.stream_ptr()), "sfb200_ar_set_lo_weights")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.sfb200_ar_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------------
    def _read_status(self):
        """(steps_done, first_all_ended_step) — one 32-byte D2H copy + sync per chunk (the reference syncs every
        sub-step).  The library's state words live at the start of the workspace (sfb200_ar_status_ptr)."""
        assert self._status_ptr == self.ws.data_ptr()
        self.status_host.copy_(self.ws[:32].view(torch.int32), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(self.status_host[0]), int(self.status_host[1])

    def history_views(self, B):
        """Device views (B, max_steps, V) of the two logits-history slabs (rows beyond the steps run so far are stale)."""
        V = self.spec["vocab_sizes"]
        n0 = self.max_rows * self.max_steps * V[0]
        h0 = self.hist[:n0].view(self.max_rows, self.max_steps, V[0])[:B]
        h1 = self.hist[n0:n0 + self.max_rows * self.max_steps * V[1]].view(self.max_rows, self.max_steps, V[1])[:B]
        return [h0, h1]

    def _noise_buf(self, B):
        buf = self._noise.get(B)
        if buf is None:
            buf = torch.empty(self.chunk_steps, 4, B, self.Vmax, dtype=torch.float32, device=self.device)
            self._noise[B] = buf
        return buf

    def _draw_noise(self, buf, n_steps, B, generator):
        """Exp(1) draws in the reference's order: per step pos-sample, pos-best, val-sample, val-best, each what
        torch.multinomial(n=1) draws internally on this device: empty_like(probs).exponential_(1) (common.py:296)."""
        V = self.spec["vocab_sizes"]
        for s in range(n_steps):
            for d in range(4):
                v = V[d // 2]
                if v == self.Vmax:
                    buf[s, d].exponential_(1.0, generator=generator)   # contiguous (B, V): same draw as empty_like(p)
                else:
                    q = torch.empty(B, v, dtype=torch.float32, device=self.device).exponential_(1.0, generator=generator)
                    buf[s, d, :, :v].copy_(q)

    def sample(self, *args, **kwargs):
        """See _sample; runs with the sampler's device current (the library launches on the current device's stream)."""
        with torch.cuda.device(self.device):
            return self._sample(*args, **kwargs)

    def _sample(self, c_indices, max_steps, top_k=100, top_p=0.8, temperature=1.0, best_in_first=False,
               mask_invalid=True, mask_invalid_completion=False, noise=None, generator=None, use_graph=True,
               stop_early=True, share_prefix=True, on_chunk=None):
        """Run the AR loop.  c_indices (B, L_c, 2) int64 (any device).  noise: optional (>= max_steps, 4, B, Vmax)
        tensor of Exp(1) draws to use instead of the device RNG (parity tests).  Returns (x (B, steps, 2) int64 on the
        device, [hist0, hist1] device views (B, steps, V) or None).  on_chunk(first_step, n_steps): called right after a chunk of
        steps has been enqueued (before the host waits for it) — the hook through which ShapeFormer.sample streams the logits
        history to the host while the next chunk runs."""
        B, L_c, tn = c_indices.shape
        if tn != 2:
            raise _lib.Sfb200Error("tuple_n must be 2 (pos, val)")
        if B > self.max_rows or L_c > self.max_cond or max_steps > self.max_steps:
            raise _lib.Sfb200Error(f"batch (B={B}, L_c={L_c}, steps={max_steps}) exceeds the sampler's capacity "
                                   f"({self.max_rows}, {self.max_cond}, {self.max_steps})")
        steps_cap = min(max_steps, self.max_len - L_c)   # steps that fit the context window (block_size)
        if steps_cap < 1:
            raise _lib.Sfb200Error(f"L_cond {L_c} leaves no room to generate within block_size {self.spec['block_size']}")
        V = self.spec["vocab_sizes"]
        if int(c_indices[..., 0].max()) >= V[0] or int(c_indices[..., 1].max()) >= V[1] or int(c_indices.min()) < 0:
            raise _lib.Sfb200Error("conditioning indices out of vocabulary range")
        self.tokens[:B].zero_()
        self.tokens[:B, :L_c].copy_(c_indices.to(self.device, non_blocking=True))
        sp = _lib.ArSampling(int(top_k), float(top_p), float(temperature), int(bool(best_in_first)),
                             int(bool(mask_invalid)), int(bool(mask_invalid_completion)))
        stream = _lib.stream_ptr()
        # rows with identical conditioning (the reference's sample_n expansion) share one prefill
        row_src, seen = (ctypes.c_int32 * B)(), {}
        if share_prefix:
            c_host = c_indices.detach().cpu().contiguous()
            for b in range(B):
                row_src[b] = seen.setdefault(c_host[b].numpy().tobytes(), b)
        else:
            for b in range(B):
                row_src[b] = b
        _lib.check(self.lib.sfb200_ar_begin_shared(self.handle, B, L_c, ctypes.byref(sp),
                                                   ctypes.cast(row_src, ctypes.c_void_p), stream), "sfb200_ar_begin_shared")
        done, steps, ended = 0, steps_cap, -1
        while done < steps_cap:
            n = min(self.chunk_steps, steps_cap - done)
            slab = self._noise_buf(B)
            if noise is not None:
                slab[:n].copy_(noise[done:done + n].to(self.device, non_blocking=True))
            else:
                self._draw_noise(slab, n, B, generator)
            _lib.check(self.lib.sfb200_ar_steps(self.handle, n, _lib.ptr(slab), int(bool(use_graph)), stream),
                       "sfb200_ar_steps")
            if on_chunk is not None:
                on_chunk(done, n)
            done += n
            if stop_early:
                _, ended = self._read_status()
                if ended >= 0:
                    steps = ended + 1
                    break
        if steps_cap < max_steps and ended < 0:
            if not stop_early:
                _, ended = self._read_status()
            if ended < 0:
                raise _lib.Sfb200Error(
                    f"context window exhausted: L_cond {L_c} + {steps_cap} generated tuples reached block_size "
                    f"{self.spec['block_size']} before every row ended; the reference's overflow crop "
                    f"(shapeformer.py:73-76) is not reproduced (SURVEY.md App. C-3)")
            steps = ended + 1
        x = self.tokens[:B, L_c:L_c + steps]
        # log-softmax(masked logits)[token] of every sampled tuple element, accumulated by the sampling kernel: the input of
        # the reference's ranking (compute_log_probs, shapeformer.py:407-418) without the logits history
        self.last_log_prob = self.logp[:B, :steps]
        hist = None
        if self.keep_history:
            n0 = self.max_rows * self.max_steps * V[0]
            h0 = self.hist[:n0].view(self.max_rows, self.max_steps, V[0])[:B, :steps]
            h1 = self.hist[n0:n0 + self.max_rows * self.max_steps * V[1]].view(self.max_rows, self.max_steps, V[1])[:B, :steps]
            hist = [h0, h1]
        return x, hist
