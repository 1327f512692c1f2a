"""GPU parity: every libsfb200 operator against the CPU oracle on the same seeded inputs (called through the C-ABI)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import sf_oracle as O
from shapeformer_b200 import ops, synth

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("M,N,K", [(1, 64, 128), (5, 1024, 1024), (16, 3072, 1024), (17, 4097, 1024), (33, 1024, 4096),
                                   (64, 4096, 1024), (64, 4097, 1024), (200, 384, 128), (1000, 1024, 1024), (64, 128, 512)])
@pytest.mark.parametrize("mode", ["plain", "bias_gelu", "bias_res"])
def test_linear(cuda, M, N, K, mode):
    x, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N, seed=4)
    if mode == "plain":
        ref = F.linear(x, W)
        out = ops.linear(x.to(cuda), W.to(cuda))
    elif mode == "bias_gelu":
        ref = F.gelu(F.linear(x, W, b))
        out = ops.linear(x.to(cuda), W.to(cuda), b.to(cuda), act="gelu")
    else:
        ref = r + F.linear(x, W, b)
        out = ops.linear(x.to(cuda), W.to(cuda), b.to(cuda), residual=r.to(cuda))
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err


def test_linear_in_place_residual_and_determinism(cuda):
    x, W, b = rnd(64, 1024, seed=1).to(cuda), rnd(1024, 1024, seed=2, scale=0.03).to(cuda), rnd(1024, seed=3).to(cuda)
    r = rnd(64, 1024, seed=4).to(cuda)
    a = ops.linear(x, W, b, residual=r)
    for _ in range(3):
        assert torch.equal(a, ops.linear(x, W, b, residual=r))   # cluster split-K sums in rank order: bitwise stable


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (64, 1024, 1024), (16, 3072, 1024), (17, 4097, 1024), (33, 1024, 4096),
                                   (64, 4096, 1024), (64, 4097, 1024), (9, 256, 128), (1, 1024, 1024), (3, 384, 128),
                                   (64, 128, 512), (12, 512, 128)])
@pytest.mark.parametrize("mode", ["plain", "bias_gelu", "bias_res", "ln_bias", "ln_plain"])
def test_linear_chain(cuda, M, N, K, mode):
    """The persistent GEMM-chain kernel of the decode step, one phase at a time: fp32-level accuracy against an fp64
    reference, with the LayerNorm fused into the activation load (row statistics merged from 512-column pieces)."""
    x, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N, seed=4)
    g, be = rnd(K, seed=5, scale=0.2) + 1, rnd(K, seed=6, scale=0.1)
    xd, Wd = x.double(), W.double()
    if mode == "plain":
        ref = xd @ Wd.t()
        out = ops.linear_chain(x.to(cuda), W.to(cuda))
    elif mode == "bias_gelu":
        ref = F.gelu(xd @ Wd.t() + b.double())
        out = ops.linear_chain(x.to(cuda), W.to(cuda), b.to(cuda), act="gelu")
    elif mode == "bias_res":
        ref = r.double() + xd @ Wd.t() + b.double()
        out = ops.linear_chain(x.to(cuda), W.to(cuda), b.to(cuda), residual=r.to(cuda))
    else:
        xs = x * 3.0 + 0.5          # non-trivial mean / variance
        h = F.layer_norm(xs.double(), (K,), g.double(), be.double(), 1e-5)
        ref = h @ Wd.t() + (b.double() if mode == "ln_bias" else 0.0)
        out = ops.linear_chain(xs.to(cuda), W.to(cuda), b.to(cuda) if mode == "ln_bias" else None, ln=(g.to(cuda), be.to(cuda)))
    err = (out.cpu().double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    assert err < 2.5e-6 * scale * max(1.0, (K / 1024) ** 0.5), (err, scale)   # single-pass TF32 would be ~1e-3


def test_linear_chain_deterministic_and_repeatable(cuda):
    """Split-K partials are reduced through L2 in split order: bitwise stable; the barrier counters reset themselves."""
    x, W, b = rnd(64, 1024, seed=1).to(cuda), rnd(1024, 1024, seed=2, scale=0.03).to(cuda), rnd(1024, seed=3).to(cuda)
    r = rnd(64, 1024, seed=4).to(cuda)
    a = ops.linear_chain(x, W, b, residual=r)
    for _ in range(20):
        assert torch.equal(a, ops.linear_chain(x, W, b, residual=r))


@pytest.mark.parametrize("M,N,K", [(65, 128, 32), (256, 1024, 1024), (512, 3072, 1024), (100, 4097, 1024), (256, 1024, 4096),
                                   (300, 4096, 1024), (2048, 1024, 1024), (128, 256, 96 * 32)])
@pytest.mark.parametrize("mode", ["plain", "bias_gelu", "bias_res"])
def test_linear_big(cuda, M, N, K, mode):
    """The TMA-fed large-M tensor-core GEMM (csrc/tc_big.cu: raw fp32 = truncated hi operand, lo parts read from memory,
    split-K through L2 with the last arriver reducing): fp32-level accuracy against an fp64 reference, and the lo part
    of the output it hands to a consuming GEMM."""
    x, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N, seed=4)
    xd, Wd = x.double(), W.double()
    if mode == "plain":
        ref = xd @ Wd.t()
        out, lo = ops.linear_big(x.to(cuda), W.to(cuda), want_lo=True)
    elif mode == "bias_gelu":
        ref = F.gelu(xd @ Wd.t() + b.double())
        out, lo = ops.linear_big(x.to(cuda), W.to(cuda), b.to(cuda), act="gelu", want_lo=True)
    else:
        ref = r.double() + xd @ Wd.t() + b.double()
        out, lo = ops.linear_big(x.to(cuda), W.to(cuda), b.to(cuda), residual=r.to(cuda), want_lo=True)
    err = (out.cpu().double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    assert err < 2.5e-6 * scale * max(1.0, (K / 1024) ** 0.5), (err, scale)   # single-pass TF32 would be ~1e-3
    # lo = rna_tf32(y - trunc_tf32(y)): hi + lo reproduces y to 2^-21 relative, and lo fits TF32 (low 13 bits clear)
    o = out.cpu()
    hi = (o.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert ((hi + lo.cpu()) - o).abs().max() <= (o.abs() * 2.0 ** -20).max()
    assert int((lo.cpu().view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_linear_big_deterministic_and_no_splitk_variant(cuda):
    x, W, b = rnd(256, 1024, seed=1).to(cuda), rnd(1024, 1024, seed=2, scale=0.03).to(cuda), rnd(1024, seed=3).to(cuda)
    a = ops.linear_big(x, W, b)
    for _ in range(5):
        assert torch.equal(a, ops.linear_big(x, W, b))          # partials are summed in split order
    c = ops.linear_big(x, W, b, split_k=False)
    assert (a - c).abs().max() < 1e-5


@pytest.mark.parametrize("rows,d", [(1, 128), (64, 1024), (777, 1024)])
def test_layernorm(cuda, rows, d):
    x, w, b = rnd(rows, d, seed=1, scale=3.0) + 0.5, rnd(d, seed=2) + 1, rnd(d, seed=3)
    out = ops.layernorm(x.to(cuda), w.to(cuda), b.to(cuda)).cpu()
    assert (out - F.layer_norm(x, (d,), w, b, 1e-5)).abs().max() < 1e-5


def _ref_attn(q, K, V):
    """q (B,H,hd), K/V (B,H,T,hd) -> (B,H*hd) exactly as CausalSelfAttention for the newest query."""
    att = (q[:, :, None] @ K.transpose(-2, -1)) * (1.0 / math.sqrt(q.shape[-1]))
    return (F.softmax(att, -1) @ V)[:, :, 0].reshape(q.shape[0], -1)


@pytest.mark.parametrize("B,H,G,shared,pos", [(4, 2, 4, 0, 0), (4, 2, 4, 5, 5), (8, 16, 4, 256, 511), (6, 3, 2, 33, 40),
                                               (8, 2, 8, 17, 100), (12, 4, 6, 31, 32), (64, 16, 4, 256, 300), (4, 2, 2, 100, 60),
                                               (256, 16, 4, 64, 100), (128, 16, 2, 33, 33)])
def test_attn_decode_grouped(cuda, B, H, G, shared, pos):
    """Grouped decode attention (shared prefix scored once per group from the leader's cache rows, TMA-staged tiles, partial
    states merged by the last arriver) against torch fp32 on the full per-row caches; appends the new K/V."""
    max_len, d = 520, H * 64
    qkv = rnd(B, 3 * d, seed=1)
    Kc, Vc = rnd(B, H, max_len, 64, seed=2), rnd(B, H, max_len, 64, seed=3)
    sh = min(shared, pos)
    lead = (torch.arange(B) // G) * G
    Kc[:, :, :sh] = Kc[lead][:, :, :sh]        # the group's rows hold identical prefix K/V
    Vc[:, :, :sh] = Vc[lead][:, :, :sh]
    kc, vc = Kc.to(cuda), Vc.to(cuda)
    # poison the siblings' prefix copies: the kernel must read the LEADER's rows only
    sib = torch.arange(B) % G != 0
    kc[sib.to(cuda), :, :sh] = float("nan")
    out = ops.attn_decode_grouped(qkv.to(cuda), kc, vc, pos, G, shared).cpu()
    q, k, v = (qkv[:, i * d:(i + 1) * d].view(B, H, 64) for i in range(3))
    Kr, Vr = Kc.clone(), Vc.clone()
    Kr[:, :, pos], Vr[:, :, pos] = k, v
    ref = _ref_attn(q, Kr[:, :, :pos + 1], Vr[:, :, :pos + 1])
    assert (out - ref).abs().max() < 2e-5
    assert torch.equal(kc[:, :, pos].cpu(), k) and torch.equal(vc[:, :, pos].cpu(), v)


@pytest.mark.parametrize("B,H,pos,n_split", [(1, 2, 0, 1), (3, 2, 1, 1), (2, 16, 7, 1), (2, 4, 100, 3), (1, 16, 511, 8),
                                              (64, 16, 300, 1), (2, 2, 37, 8)])
def test_attn_decode(cuda, B, H, pos, n_split):
    max_len, d = 520, H * 64
    qkv = rnd(B, 3 * d, seed=1)
    Kc, Vc = rnd(B, H, max_len, 64, seed=2), rnd(B, H, max_len, 64, seed=3)
    kc, vc = Kc.to(cuda), Vc.to(cuda)
    out = ops.attn_decode(qkv.to(cuda), kc, vc, pos, n_split).cpu()
    q, k, v = (qkv[:, i * d:(i + 1) * d].view(B, H, 64) for i in range(3))
    Kr, Vr = Kc.clone(), Vc.clone()
    Kr[:, :, pos], Vr[:, :, pos] = k, v
    ref = _ref_attn(q, Kr[:, :, :pos + 1], Vr[:, :, :pos + 1])
    assert (out - ref).abs().max() < 2e-5
    assert torch.equal(kc.cpu(), Kr) and torch.equal(vc.cpu(), Vr)   # append only touched `pos`


@pytest.mark.parametrize("B,H,T", [(1, 2, 1), (2, 2, 9), (2, 16, 64), (3, 4, 255), (2, 3, 32), (2, 16, 256), (1, 2, 300), (2, 2, 129)])
def test_attn_prefill(cuda, B, H, T):
    """Causal prefill attention + cache fill against torch fp32: the FFMA kernel below 32 positions, the tcgen05 kernel
    (QK^T and PV as 3xTF32 tensor-core GEMMs, P kept in tensor memory) from 32 positions on, incl. ragged last tiles."""
    max_len, d = 300, H * 64
    qkv = rnd(B, T, 3 * d, seed=5)
    kc, vc = torch.zeros(B, H, max_len, 64, device=cuda), torch.zeros(B, H, max_len, 64, device=cuda)
    out = ops.attn_prefill(qkv.to(cuda), kc, vc).cpu()
    q, k, v = (qkv[..., i * d:(i + 1) * d].view(B, T, H, 64).transpose(1, 2) for i in range(3))
    att = (q @ k.transpose(-2, -1)) / 8.0
    att = att.masked_fill(torch.triu(torch.ones(T, T, dtype=torch.bool), 1), float("-inf"))
    ref = (F.softmax(att, -1) @ v).transpose(1, 2).reshape(B, T, d)
    assert (out - ref).abs().max() < 2e-5
    assert torch.equal(kc[:, :, :T].cpu(), k.contiguous()) and torch.equal(vc[:, :, :T].cpu(), v.contiguous())
    assert float(kc[:, :, T:].abs().max()) == 0.0 if T < max_len else True


SAMPLE_CASES = [(100, 0.4, 1.0, True, True, True), (50, 0.0, 1.0, True, False, False), (1, 0.001, 1.0, False, False, False),
                (0, 0.9, 0.7, False, True, False), (5000, 1.0, 1.3, False, True, True), (0, 0.0, 1.0, False, False, False),
                (3, 0.5, 2.0, True, True, True)]


@pytest.mark.parametrize("top_k,top_p,T,bif,mi,mic", SAMPLE_CASES)
@pytest.mark.parametrize("tuple_i", [0, 1])
def test_ar_sample_bit_exact(cuda, top_k, top_p, T, bif, mi, mic, tuple_i):
    """mask -> filter -> draw: tokens identical to the oracle (= the reference's sampling_masker + sample_logits) for the
    same logits and noise; masked logits identical bit for bit."""
    B, V, Lc, L, max_len = 9, 4097, 6, 9, 16
    g = torch.Generator().manual_seed(top_k * 7 + tuple_i)
    logits = torch.randn(B, V, generator=g) * 2.5
    logits[4, :70] = logits[4, 70]                   # a tie group
    tokens = torch.zeros(B, max_len, 2, dtype=torch.int64)
    tokens[:, :Lc] = synth.cond_indices(B, Lc, seed=9)
    for b in range(B):                                # three generated tuples, increasing positions
        p = torch.sort(torch.randint(0, 4000, (3,), generator=g))[0] + torch.arange(3)
        tokens[b, Lc:L, 0], tokens[b, Lc:L, 1] = p, torch.randint(0, 4096, (3,), generator=g)
    tokens[2, L - 1, 0] = 4096                        # a row that already ended
    if tuple_i == 1:
        tokens[:, L, 0] = torch.randint(0, 4096, (B,), generator=g)
        tokens[5, L, 0] = 4096                        # forced end value
    qs, qb = torch.empty(B, V).exponential_(1.0, generator=g), torch.empty(B, V).exponential_(1.0, generator=g)
    # oracle
    masked = O.sampling_masker(logits, tokens[:, :L + 1], Lc, L - Lc, tuple_i, (4096, 4096), mi, mic)
    new = O.sample_rows(masked, qs, top_k, top_p, T)
    best = O.sample_rows(masked, qb, 1, 0.001, T)
    if bif:
        new[0] = best[0]
    tk = tokens.to(cuda)
    hist = ops.ar_sample(logits.to(cuda), tk, L, Lc, tuple_i, qs.to(cuda), qb.to(cuda), (4096, 4096), top_k, top_p, T, bif,
                         mi, mic)
    got = tk[:, L, tuple_i].cpu()
    assert torch.equal(hist.cpu(), masked)
    assert torch.equal(got, new), (got.tolist(), new.tolist())
    untouched = tk.cpu().clone()
    untouched[:, L, tuple_i] = tokens[:, L, tuple_i]
    assert torch.equal(untouched, tokens)


def test_ar_sample_small_vocab(cuda):
    B, V = 4, 37
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(B, V, generator=g)
    tokens = torch.zeros(B, 8, 2, dtype=torch.int64)
    tokens[:, 0] = torch.tensor([36, 36])
    qs, qb = torch.empty(B, V).exponential_(1.0, generator=g), torch.empty(B, V).exponential_(1.0, generator=g)
    masked = O.sampling_masker(logits, tokens[:, :2], 1, 0, 0, (36, 36), True, False)
    new = O.sample_rows(masked, qs, 5, 0.9, 1.0)
    tk = tokens.to(cuda)
    ops.ar_sample(logits.to(cuda), tk, 1, 1, 0, qs.to(cuda), qb.to(cuda), (36, 36), 5, 0.9, 1.0, False, True, False)
    assert torch.equal(tk[:, 1, 0].cpu(), new)


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (64, 1024, 1024), (16, 3072, 1024), (17, 4097, 1024), (33, 1024, 4096),
                                   (64, 4096, 1024), (64, 4097, 1024), (1, 256, 128), (200, 384, 128), (1000, 1024, 1024),
                                   (4096, 1024, 1024)])
@pytest.mark.parametrize("mode", ["plain", "bias_gelu", "bias_res"])
def test_linear_tc(cuda, M, N, K, mode):
    """tcgen05 3xTF32 GEMM: fp32-level accuracy (not TF32-level) against an fp64 reference."""
    x, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N, seed=4)
    xd, Wd = x.double(), W.double()
    if mode == "plain":
        ref = xd @ Wd.t()
        out = ops.linear_tc(x.to(cuda), W.to(cuda))
    elif mode == "bias_gelu":
        ref = F.gelu(xd @ Wd.t() + b.double())
        out = ops.linear_tc(x.to(cuda), W.to(cuda), b.to(cuda), act="gelu")
    else:
        ref = r.double() + xd @ Wd.t() + b.double()
        out = ops.linear_tc(x.to(cuda), W.to(cuda), b.to(cuda), residual=r.to(cuda))
    err = (out.cpu().double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    assert err < 2.5e-6 * scale * max(1.0, (K / 1024) ** 0.5), (err, scale)   # single-pass TF32 would be ~1e-3


def test_linear_tc_deterministic_and_in_place(cuda):
    x, W, b = rnd(64, 1024, seed=1).to(cuda), rnd(1024, 1024, seed=2, scale=0.03).to(cuda), rnd(1024, seed=3).to(cuda)
    r = rnd(64, 1024, seed=4).to(cuda)
    a = ops.linear_tc(x, W, b, residual=r)
    for _ in range(5):
        assert torch.equal(a, ops.linear_tc(x, W, b, residual=r))   # split-K partials are summed in split order


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (64, 1024, 1024), (16, 3072, 1024), (17, 4097, 1024), (33, 1024, 4096),
                                   (64, 4096, 1024), (64, 4097, 1024), (9, 256, 128)])
@pytest.mark.parametrize("mode", ["plain", "bias_gelu", "bias_res"])
def test_linear_tc_presplit(cuda, M, N, K, mode):
    """Decode-step GEMM from pre-split TF32 weight tiles (TMA bulk copies, both operands from shared memory)."""
    x, W, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N, seed=4)
    xd, Wd = x.double(), W.double()
    if mode == "plain":
        ref = xd @ Wd.t()
        out = ops.linear_tc_ps(x.to(cuda), W.to(cuda))
    elif mode == "bias_gelu":
        ref = F.gelu(xd @ Wd.t() + b.double())
        out = ops.linear_tc_ps(x.to(cuda), W.to(cuda), b.to(cuda), act="gelu")
    else:
        ref = r.double() + xd @ Wd.t() + b.double()
        out = ops.linear_tc_ps(x.to(cuda), W.to(cuda), b.to(cuda), residual=r.to(cuda))
    err = (out.cpu().double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    assert err < 2.5e-6 * scale * max(1.0, (K / 1024) ** 0.5), (err, scale)


# ---- conv prologue kernels (csrc/conv_tc.cu) ---------------------------------------------------------------------------
def _conv_tc(cuda, x_cl, w, bias=None, relu=False):
    """x_cl (B,R,R,R,Cin) CPU, w (Cout,Cin,k,k,k) CPU -> (out (B,R,R,R,Cout), stats (B,Cout,2)) through sfb200_conv3d_tc."""
    from shapeformer_b200 import _lib
    lib = _lib.load()
    B, R, Cin = x_cl.shape[0], x_cl.shape[1], x_cl.shape[-1]
    Cout, taps = w.shape[0], w.shape[2] ** 3
    hi = x_cl.to(cuda).contiguous()
    lo = ops.split_lo(hi)
    wp = w.reshape(Cout, Cin, taps).permute(2, 0, 1).contiguous().reshape(taps * Cout, Cin).to(cuda)
    wl = ops.split_lo(wp)
    out = torch.empty(B, R, R, R, Cout, device=cuda)
    st = torch.zeros(B, Cout, 2, dtype=torch.float64, device=cuda)
    bb = bias.to(cuda) if bias is not None else None
    _lib.check(lib.sfb200_conv3d_tc(_lib.ptr(hi), _lib.ptr(lo), _lib.ptr(wp), _lib.ptr(wl), _lib.ptr(bb), _lib.ptr(out), _lib.ptr(st),
                                    B, R, R, R, Cin, Cout, taps, int(relu), _lib.stream_ptr()), "conv3d_tc")
    return out.cpu(), st.cpu()


@pytest.mark.parametrize("B,R,Cin,Cout,k", [(2, 16, 128, 128, 3), (3, 8, 128, 256, 3), (3, 4, 256, 512, 3), (2, 8, 768, 256, 3),
                                            (1, 32, 64, 64, 3), (1, 64, 32, 32, 3), (2, 16, 128, 128, 1), (1, 32, 128, 64, 3)])
def test_conv3d_tc(cuda, B, R, Cin, Cout, k):
    """tcgen05 implicit-GEMM conv3d (TMA boxes with zero-filled halo = the padding; 3xTF32) against torch fp64 conv3d, incl.
    the per-channel sums of the epilogue (the next GroupNorm's statistics), boxes spanning two samples (4^3) and odd B."""
    x = rnd(B, R, R, R, Cin, seed=1)
    w = rnd(Cout, Cin, k, k, k, seed=2, scale=1.0 / math.sqrt(Cin * k ** 3))
    bias = rnd(Cout, seed=3) if k == 1 else None
    relu = k == 3
    out, st = _conv_tc(cuda, x, w, bias, relu)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w.double(), bias.double() if bias is not None else None, padding=k // 2)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 4, 1)
    err = (out.double() - ref).abs().max().item()
    assert err < 4e-6 * max(1.0, ref.abs().max().item()), err        # single-pass TF32 would be ~1e-3
    s_ref = torch.stack([ref.sum((1, 2, 3)), (ref * ref).sum((1, 2, 3))], -1)
    assert (st - s_ref).abs().max() < 1e-3 * max(1.0, s_ref.abs().max().item())


@pytest.mark.parametrize("B,R,Cin,Cout", [(2, 16, 128, 64), (1, 32, 64, 32), (3, 4, 64, 128)])
def test_conv3d_tc_subpixel_after_upsample(cuda, B, R, Cin, Cout):
    """[nearest x2 -> 3x3x3 conv pad 1 -> ReLU] in sub-pixel form (8 phases x 8 summed taps on the low-resolution input) against
    torch fp64 on the materialised upsampled tensor, incl. the epilogue sums."""
    from shapeformer_b200 import _lib
    from shapeformer_b200.decoder import pack_subpixel_weights
    lib = _lib.load()
    x = rnd(B, R, R, R, Cin, seed=1)
    w = rnd(Cout, Cin, 3, 3, 3, seed=2, scale=1.0 / math.sqrt(Cin * 27))
    hi = x.to(cuda).contiguous()
    lo = ops.split_lo(hi)
    wp = pack_subpixel_weights(w.to(cuda)).reshape(64 * Cout, Cin).contiguous()
    wl = ops.split_lo(wp)
    out = torch.empty(B, 2 * R, 2 * R, 2 * R, Cout, device=cuda)
    st = torch.zeros(B, Cout, 2, dtype=torch.float64, device=cuda)
    _lib.check(lib.sfb200_conv3d_tc(_lib.ptr(hi), _lib.ptr(lo), _lib.ptr(wp), _lib.ptr(wl), None, _lib.ptr(out), _lib.ptr(st),
                                    B, R, R, R, Cin, Cout, 8, 1, _lib.stream_ptr()), "conv3d_tc")
    up = F.interpolate(x.permute(0, 4, 1, 2, 3).double(), scale_factor=2, mode="nearest")
    ref = F.relu(F.conv3d(up, w.double(), padding=1)).permute(0, 2, 3, 4, 1)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 4e-6 * max(1.0, ref.abs().max().item()), err
    s_ref = torch.stack([ref.sum((1, 2, 3)), (ref * ref).sum((1, 2, 3))], -1)
    assert (st.cpu() - s_ref).abs().max() < 1e-3 * max(1.0, s_ref.abs().max().item())


def test_conv_prep_groupnorm_upsample_concat_and_pool(cuda):
    """conv_prep: GroupNorm(8) from per-channel sums over the CONCATENATION of a skip tensor and a x2-upsampled tensor, operand
    split on store; pool_stats: max-pool 2 + sums — against torch."""
    from shapeformer_b200 import _lib
    lib = _lib.load()
    B, R, C0, C1 = 2, 8, 64, 128
    a, b = rnd(B, R, R, R, C0, seed=1, scale=2.0) + 0.3, rnd(B, R // 2, R // 2, R // 2, C1, seed=2) - 0.5
    gamma, beta = rnd(C0 + C1, seed=3) + 1, rnd(C0 + C1, seed=4)
    def sums(t):
        st = torch.zeros(B, t.shape[-1], 2, dtype=torch.float64, device=cuda)
        _lib.check(lib.sfb200_pool_stats(_lib.ptr(t), None, _lib.ptr(st), B, t.shape[1], t.shape[1], t.shape[1], t.shape[-1], 1,
                                         _lib.stream_ptr()), "pool_stats")
        return st
    ad, bd = a.to(cuda), b.to(cuda)
    sa, sb = sums(ad), sums(bd)
    dst, lo = torch.empty(B, R, R, R, C0 + C1, device=cuda), torch.empty(B, R, R, R, C0 + C1, device=cuda)
    gd, btd = gamma.to(cuda), beta.to(cuda)       # keep the device copies alive across the launch
    _lib.check(lib.sfb200_conv_prep(_lib.ptr(ad), C0, 0, _lib.ptr(sa), float(R ** 3), _lib.ptr(bd), C1, 1, _lib.ptr(sb),
                                    float((R // 2) ** 3), _lib.ptr(gd), _lib.ptr(btd), 8, _lib.ptr(dst),
                                    _lib.ptr(lo), B, R, R, R, _lib.stream_ptr()), "conv_prep")
    cat = torch.cat([a.permute(0, 4, 1, 2, 3), F.interpolate(b.permute(0, 4, 1, 2, 3), scale_factor=2, mode="nearest")], 1)
    ref = F.group_norm(cat, 8, gamma, beta, 1e-5).permute(0, 2, 3, 4, 1)
    assert (dst.cpu() - ref).abs().max() < 2e-5
    d = dst.cpu()
    hi = (d.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert ((hi + lo.cpu()) - d).abs().max() <= (d.abs() * 2.0 ** -20).max()
    # max-pool 2 + sums
    pooled = torch.empty(B, R // 2, R // 2, R // 2, C0, device=cuda)
    st = torch.zeros(B, C0, 2, dtype=torch.float64, device=cuda)
    _lib.check(lib.sfb200_pool_stats(_lib.ptr(ad), _lib.ptr(pooled), _lib.ptr(st), B, R // 2, R // 2, R // 2, C0, 2, _lib.stream_ptr()),
               "pool_stats")
    pref = F.max_pool3d(a.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1)
    assert torch.equal(pooled.cpu(), pref.contiguous())
    assert (st.cpu()[..., 0] - pref.double().sum((1, 2, 3))).abs().max() < 1e-6
