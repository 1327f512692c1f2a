"""Thin Python wrappers over the individual libsfb200 operators (the same kernels the AR engine enqueues); used by the
parity tests and for profiling single kernels.  Device tensors in, device tensors out, current CUDA stream."""
import ctypes

import torch

from . import _lib


def linear(x, W, bias=None, residual=None, act=None):
    """y = act(x @ W.T + bias) + residual  (nn.Linear; act in {None, "gelu"})."""
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.sfb200_linear(_lib.ptr(x), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), M, N, K,
                                 1 if act == "gelu" else 0, _lib.stream_ptr()), "sfb200_linear")
    return y


def linear_tc(x, W, bias=None, residual=None, act=None):
    """linear() on the tcgen05 tensor cores (3xTF32 with periodic fp32 promotion: fp32-level accuracy)."""
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.sfb200_linear_tc(_lib.ptr(x), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), M, N, K,
                                    1 if act == "gelu" else 0, _lib.stream_ptr()), "sfb200_linear_tc")
    return y


def split_lo(x):
    """Low part of the two-term TF32 split whose high part is the tensor core's truncation of x (what tc_big reads)."""
    lib = _lib.load()
    lo = torch.empty_like(x)
    _lib.check(lib.sfb200_split_lo(_lib.ptr(x), _lib.ptr(lo), x.numel(), _lib.stream_ptr()), "sfb200_split_lo")
    return lo


def linear_big(x, W, bias=None, residual=None, act=None, want_lo=False, split_k=True):
    """linear() through the TMA-fed large-M tensor-core kernel (csrc/tc_big.cu); lo parts of x and W made here."""
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    y_lo = torch.empty_like(y) if want_lo else None
    part = torch.empty(lib.sfb200_big_partial_floats(), dtype=torch.float32, device=x.device) if split_k else None
    cnt = torch.zeros(2048, dtype=torch.int32, device=x.device) if split_k else None
    x_lo, W_lo = split_lo(x), split_lo(W)      # keep the temporaries alive across the launch
    _lib.check(lib.sfb200_linear_big(_lib.ptr(x), _lib.ptr(x_lo), _lib.ptr(W), _lib.ptr(W_lo), _lib.ptr(bias),
                                     _lib.ptr(residual), _lib.ptr(y), _lib.ptr(y_lo), M, N, K, 1 if act == "gelu" else 0,
                                     _lib.ptr(part), _lib.ptr(cnt), _lib.stream_ptr()), "sfb200_linear_big")
    if split_k:
        assert int(cnt.abs().sum()) == 0      # the arrival counters reset themselves
    return (y, y_lo) if want_lo else y


def linear_tc_ps(x, W, bias=None, residual=None, act=None):
    """linear() for M <= 64 rows from pre-split TF32 weight tiles (pretiles W on every call: test helper)."""
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    wt = torch.empty(lib.sfb200_tc_pretiled_floats(N, K), dtype=torch.float32, device=x.device)
    _lib.check(lib.sfb200_tc_pretile(_lib.ptr(W), _lib.ptr(wt), N, K, _lib.stream_ptr()), "sfb200_tc_pretile")
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.sfb200_linear_tc_ps(_lib.ptr(x), _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), M, N, K,
                                       1 if act == "gelu" else 0, _lib.stream_ptr()), "sfb200_linear_tc_ps")
    return y


_chain_ws = {}


def linear_chain(x, W, bias=None, residual=None, act=None, ln=None):
    """y = act(LN?(x) @ W.T + bias) + residual for M <= 64 rows through the persistent GEMM-chain kernel the sampler's decode
    step uses (fp32 weights streamed by TMA, 3xTF32 on tcgen05, split-K reduced through L2).  ln = (weight, bias) applies a
    LayerNorm (eps 1e-5) to x on load.  `residual` may alias the returned tensor's role (in-place update is what the engine
    does); here a fresh y is returned."""
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    ws = _chain_ws.get(x.device)
    if ws is None:
        ws = torch.zeros(lib.sfb200_chain_workspace_bytes(), dtype=torch.uint8, device=x.device)
        _chain_ws[x.device] = ws
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.sfb200_chain_linear(_lib.ptr(x), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(residual), _lib.ptr(y), M, N, K,
                                       1 if act == "gelu" else 0, _lib.ptr(ln[0]) if ln else None,
                                       _lib.ptr(ln[1]) if ln else None, _lib.ptr(ws), _lib.stream_ptr()),
               "sfb200_chain_linear")
    return y


def layernorm(x, w, b):
    lib = _lib.load()
    rows, d = x.shape
    y = torch.empty_like(x)
    _lib.check(lib.sfb200_layernorm(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), rows, d, _lib.stream_ptr()),
               "sfb200_layernorm")
    return y


def attn_decode_grouped(qkv, kcache, vcache, pos, group, shared_len):
    """attn_decode for contiguous groups of `group` rows whose first `shared_len` cached positions are identical (read from
    the group's first row).  Returns (B, d); caches updated in place at `pos`."""
    lib = _lib.load()
    B = qkv.shape[0]
    _, H, max_len, hd = kcache.shape
    assert hd == 64
    out = torch.empty(B, H * 64, dtype=torch.float32, device=qkv.device)
    part = torch.empty(B * H * 3 * 66, dtype=torch.float32, device=qkv.device)
    cnt = torch.zeros(B * H, dtype=torch.int32, device=qkv.device)
    _lib.check(lib.sfb200_attn_decode_grouped(_lib.ptr(qkv), _lib.ptr(kcache), _lib.ptr(vcache), _lib.ptr(out), _lib.ptr(part),
                                              _lib.ptr(cnt), B, H, max_len, pos, group, shared_len, _lib.stream_ptr()),
               "sfb200_attn_decode_grouped")
    assert int(cnt.abs().sum()) == 0      # the arrival counters reset themselves
    return out


def attn_decode(qkv, kcache, vcache, pos, n_split=1):
    """qkv (B,3d); caches (B,H,max_len,64) updated in place at `pos`; returns (B,d)."""
    lib = _lib.load()
    B = qkv.shape[0]
    _, H, max_len, hd = kcache.shape
    assert hd == 64
    out = torch.empty(B, H * 64, dtype=torch.float32, device=qkv.device)
    part = torch.empty(B * H * n_split * 66, dtype=torch.float32, device=qkv.device) if n_split > 1 else None
    _lib.check(lib.sfb200_attn_decode(_lib.ptr(qkv), _lib.ptr(kcache), _lib.ptr(vcache), _lib.ptr(out), _lib.ptr(part), B, H,
                                      max_len, pos, None, n_split, _lib.stream_ptr()), "sfb200_attn_decode")
    return out


def attn_prefill(qkv, kcache, vcache):
    """qkv (B,T,3d); fills caches [0,T); returns (B,T,d)."""
    lib = _lib.load()
    B, T, _ = qkv.shape
    _, H, max_len, _ = kcache.shape
    out = torch.empty(B, T, H * 64, dtype=torch.float32, device=qkv.device)
    _lib.check(lib.sfb200_attn_prefill(_lib.ptr(qkv), _lib.ptr(kcache), _lib.ptr(vcache), _lib.ptr(out), B, H, T, max_len,
                                       _lib.stream_ptr()), "sfb200_attn_prefill")
    return out


def ar_sample(logits, tokens, L, L_cond, tuple_i, noise_sample, noise_best, end_tokens=(4096, 4096), top_k=100, top_p=0.8,
              temperature=1.0, best_in_first=False, mask_invalid=True, mask_invalid_completion=False, want_hist=True):
    """One masked + filtered categorical draw per row; writes tokens[:, L, tuple_i] in place.  Returns masked logits."""
    lib = _lib.load()
    B, V = logits.shape
    max_len = tokens.shape[1]
    hist = torch.empty(B, V, dtype=torch.float32, device=logits.device) if want_hist else None
    et = (ctypes.c_int64 * 2)(int(end_tokens[0]), int(end_tokens[1]))
    sp = _lib.ArSampling(int(top_k), float(top_p), float(temperature), int(bool(best_in_first)), int(bool(mask_invalid)),
                         int(bool(mask_invalid_completion)))
    _lib.check(lib.sfb200_ar_sample(_lib.ptr(logits), _lib.ptr(tokens), _lib.ptr(hist), _lib.ptr(noise_sample),
                                    _lib.ptr(noise_best), B, V, max_len, L, L_cond, tuple_i,
                                    ctypes.cast(et, ctypes.c_void_p), ctypes.byref(sp), _lib.stream_ptr()),
               "sfb200_ar_sample")
    return hist
