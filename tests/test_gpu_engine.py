"""GPU parity of the whole hot path against the oracle and the committed reference fixtures."""
import pytest
import torch

from oracle import sf_oracle as O
from shapeformer_b200 import ar, decoder, synth
from tests import util

pytestmark = pytest.mark.gpu
END = (4096, 4096)


def make_sampler(cuda, cfg, sd, B, Lc, steps, **kw):
    blob = ar.pack_gpt_weights(sd, cfg, cuda)
    return ar.ARSampler(blob, cfg, END, max_rows=B, max_cond=Lc, max_steps=steps, keep_history=True, **kw)


@pytest.mark.parametrize("path", util.sampler_goldens())
@pytest.mark.parametrize("use_graph", [False, True])
def test_sampler_reproduces_reference_golden(cuda, path, use_graph):
    """Token sequences bit-exact vs the REFERENCE's sample_indices (fixture), logits history within 5e-5."""
    g = torch.load(path)
    cfg = synth.TINY_GPT
    wseed, cseed, rseed = g["seeds"]
    sd = synth.gpt_state_dict(cfg, seed=wseed, peaky=True)
    c = synth.cond_indices(g["B"], g["L_c"], seed=cseed, shared=True)
    noise = util.noise_from_seed(rseed, g["steps"], g["B"], 4097)
    s = make_sampler(cuda, cfg, sd, g["B"], g["L_c"], g["steps"], chunk_steps=5)
    x, hist = s.sample(c, g["steps"], top_k=g["top_k"], top_p=g["top_p"], temperature=g["temperature"],
                       best_in_first=g["best_in_first"], mask_invalid=g["masks"][0], mask_invalid_completion=g["masks"][1],
                       noise=noise, use_graph=use_graph)
    x = x.cpu()
    assert x.shape == g["tokens"].shape, (x.shape, g["tokens"].shape)   # incl. early termination at the same step
    assert torch.equal(x, g["tokens"])
    util.check_history_summary(x, hist, g["hist"])


@pytest.mark.parametrize("B,Lc,steps,masks,chunk", [(5, 1, 6, (True, False), 32), (2, 40, 23, (False, False), 7),
                                                    (7, 17, 9, (True, True), 4)])
def test_sampler_matches_oracle(cuda, B, Lc, steps, masks, chunk):
    """Distinct conditioning per row, L_cond = 1 (unconditional start), ragged chunking; full logits history compared."""
    cfg = dict(synth.TINY_GPT, n_layers=(3, 2))
    sd = synth.gpt_state_dict(cfg, seed=21, peaky=True)
    c = synth.cond_indices(B, Lc, seed=4) if Lc > 1 else torch.tensor([[list(END)]]).repeat(B, 1, 1)
    noise = util.noise_from_seed(5, steps, B, 4097)
    ox, oh = O.sample_indices(sd, O.GPTSpec(**cfg), c, c[:, :0], steps, END, True, 30, 0.7, 1.0, masks[0], masks[1],
                              noise=O.ListNoise(noise.reshape(-1, B, 4097)), cached=True)
    s = make_sampler(cuda, cfg, sd, B, Lc, steps, chunk_steps=chunk, prefill_tokens=3 * max(Lc, 1))
    for use_graph in (False, True):
        x, hist = s.sample(c, steps, top_k=30, top_p=0.7, best_in_first=True, mask_invalid=masks[0],
                           mask_invalid_completion=masks[1], noise=noise, use_graph=use_graph)
        assert torch.equal(x.cpu(), ox)
        for a, b in zip(hist, oh):
            a = a.cpu()
            assert torch.equal(torch.isfinite(a), torch.isfinite(b))
            fin = torch.isfinite(b)
            assert (a[fin] - b[fin]).abs().max() < 5e-5


def test_sampler_device_rng_is_the_multinomial_stream(cuda):
    """Without explicit noise the sampler draws Exp(1) with the same torch call torch.multinomial uses, so a re-seeded run
    repeats itself and equals a run fed with the tensor drawn by hand."""
    cfg = synth.TINY_GPT
    sd = synth.gpt_state_dict(cfg, seed=8, peaky=True)
    B, Lc, steps = 3, 10, 6
    c = synth.cond_indices(B, Lc, seed=1)
    s = make_sampler(cuda, cfg, sd, B, Lc, steps)
    torch.manual_seed(123); torch.cuda.manual_seed(123)
    x1, _ = s.sample(c, steps, top_k=50, top_p=0.0, mask_invalid=False, stop_early=False)
    x1 = x1.clone()
    torch.manual_seed(123); torch.cuda.manual_seed(123)
    noise = torch.stack([torch.stack([torch.empty(B, 4097, device=cuda).exponential_(1.0) for _ in range(4)])
                         for _ in range(steps)])
    x2, _ = s.sample(c, steps, top_k=50, top_p=0.0, mask_invalid=False, noise=noise, stop_early=False)
    assert torch.equal(x1, x2)


@pytest.mark.parametrize("impl", [0, 1])
def test_decoder_golden_and_oracle(cuda, impl):
    g = torch.load(util.GOLDEN + "/decoder.pt")
    sd = synth.vqdif_state_dict(seed=g["wseed"])
    code = synth.code_grids(1, seed=g["code_seed"])
    gen = torch.Generator().manual_seed(g["pts_seed"])
    Xtg = torch.rand(1, g["n"], 3, generator=gen) * 2 - 1
    dec = decoder.ImplicitDecoder(sd, cuda, impl=impl)
    out = dec.decode_index(code, Xtg)["logits"]
    assert out.shape == (1, g["n"], 1)
    err = (out[0, :, 0].cpu() - g["logits"]).abs().max().item()
    occ = torch.sigmoid(out[0, :, 0].cpu())
    assert (occ - torch.sigmoid(g["logits"])).abs().max() < 1e-4   # north-star tolerance: occupancy within 1e-4 fp32
    # logits: the cuDNN conv stack alone differs from the CPU convs by ~6e-5 (fp32 mode, summation order over K <= 20k);
    # the point kernel in isolation is held to 2e-5 in test_decoder_pieces_vs_oracle
    assert err < 2e-4, err


@pytest.mark.parametrize("impl", [0, 1])
def test_decoder_pieces_vs_oracle(cuda, impl):
    sd = synth.vqdif_state_dict(seed=6)
    dec = decoder.ImplicitDecoder(sd, cuda, impl=impl)
    code = synth.code_grids(2, seed=1)
    # get_code: exact
    assert torch.equal(dec.get_code(code).cpu(), O.get_code(sd, code))
    # per-point kernel on an oracle-made feature grid (isolates the fused trilinear + MLP kernel), incl. out-of-range points
    grid = O.feature_grid(sd, code)
    gen = torch.Generator().manual_seed(2)
    Xtg = torch.rand(2, 5000, 3, generator=gen) * 2.6 - 1.3
    Xtg[:, :4] = torch.tensor([[-1., -1, -1], [1, 1, 1], [1.101, 1.101, 1.101], [-1.101, 0, 1.2]])
    ref = O.decode_points(sd, grid, Xtg)[..., 0]
    gcl = grid.permute(0, 2, 3, 4, 1).contiguous().to(cuda)
    out = dec.decode_points(gcl, Xtg).cpu()
    assert (out - ref).abs().max() < 2e-5
    # shared query set (stride 0) == per-shape copies
    out_shared = dec.decode_points(gcl, Xtg[:1]).cpu()
    assert torch.equal(out_shared[1], dec.decode_points(gcl[1:], Xtg[:1]).cpu()[0])
    # layout kernel
    g2 = torch.randn(2, 32, 4, 5, 6)
    cl = torch.empty(2, 4, 5, 6, 32, device=cuda)
    from shapeformer_b200 import _lib
    _lib.check(_lib.load().sfb200_grid_to_channels_last(_lib.ptr(g2.to(cuda)), _lib.ptr(cl), 2, 32, 120, _lib.stream_ptr()))
    assert torch.equal(cl.cpu(), g2.permute(0, 2, 3, 4, 1).contiguous())


def test_tokens_to_dense(cuda):
    sd = synth.vqdif_state_dict(seed=6)
    dec = decoder.ImplicitDecoder(sd, cuda, impl=1)
    g = torch.Generator().manual_seed(0)
    toks = torch.stack([torch.randint(0, 4097, (3, 40), generator=g), torch.randint(0, 4097, (3, 40), generator=g)], -1)
    toks[1, 5:] = 4096
    toks[2, 3, 0] = toks[2, 1, 0]   # duplicate position: the later tuple wins
    empty = torch.tensor([7, 4000, 0])
    out = dec.tokens_to_dense(toks, empty).cpu()
    for b in range(3):
        assert torch.equal(out[b], O.tokens_to_dense(toks[b], empty[b]))


@pytest.mark.parametrize("impl", [0, 1])
def test_full_64cubed_decode_properties(cuda, impl):
    """BASELINE size (262,144 query points): size-independent checks — a shape decoded alone equals the same shape inside a
    batch, and a random 4,096-point subset matches the oracle."""
    sd = synth.vqdif_state_dict(seed=6)
    dec = decoder.ImplicitDecoder(sd, cuda, impl=impl)
    code = synth.code_grids(2, seed=5)
    Xtg = synth.make_grid(64)[None]
    full = dec.decode_index(code, Xtg)["logits"]
    assert full.shape == (2, 64 ** 3, 1)
    if impl == 0:   # the two point kernels agree with each other on the full grid (same feature grid, same points)
        other = dec.decode_points(dec.feature_grid(dec.get_code(code)), Xtg, impl=1)
        assert (full[..., 0] - other).abs().max() < 2e-5
        occ = dec.occupancy(code, Xtg)
        assert (occ - torch.sigmoid(full[..., 0])).abs().max() < 1e-6
    alone = dec.decode_index(code[1:], Xtg)["logits"]
    assert (full[1] - alone[0]).abs().max() < 5e-5     # cuDNN may pick another algorithm for B=1; the point kernel is bitwise
    sel = torch.randperm(64 ** 3, generator=torch.Generator().manual_seed(1))[:4096]
    ref = O.decode_index(sd, code[:1], Xtg[:, sel])["logits"]
    got = full[0, sel.to(cuda)].cpu()
    assert (torch.sigmoid(got) - torch.sigmoid(ref[0])).abs().max() < 1e-4
    assert (got - ref[0]).abs().max() < 2e-4


def test_sampler_shipped_model_matches_oracle(cuda):
    """The SHIPPED transformer size (20+4 layers, d=1024, 16 heads, 325M parameters): tokens identical to the oracle and
    logits history within 1.5e-4 absolute (3e-6 of their range) — exercises the tcgen05 3xTF32 GEMMs at K = 1024 / 4096 and
    cluster split-K."""
    cfg = synth.SHIPPED_GPT
    sd = synth.gpt_state_dict(cfg, seed=314, peaky=True)
    B, Lc, steps = 12, 24, 6
    c = synth.cond_indices(B, Lc, seed=7)
    noise = util.noise_from_seed(9, steps, B, 4097)
    ox, oh = O.sample_indices(sd, O.GPTSpec(**cfg), c, c[:, :0], steps, END, True, 50, 0.9, 1.0, True, True,
                              noise=O.ListNoise(noise.reshape(-1, B, 4097)), cached=True)
    s = make_sampler(cuda, cfg, sd, B, Lc, steps)
    x, hist = s.sample(c, steps, top_k=50, top_p=0.9, best_in_first=True, mask_invalid=True, mask_invalid_completion=True,
                       noise=noise, use_graph=True)
    worst = 0.0
    for a, b in zip(hist, oh):
        a = a.cpu()
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin)
        worst = max(worst, (a[fin] - b[fin]).abs().max().item())
    print("shipped-size max |dlogit| =", worst)
    assert torch.equal(x.cpu(), ox)
    assert worst < 1.5e-4, worst      # logits span +-20 here (peaky weights): measured 4-5e-5 absolute = 3e-6 relative


@pytest.mark.parametrize("pattern", [[0, 0, 1, 0, 1, 2, 2],          # scattered groups: shared prefill, per-row attention
                                     [0, 0, 0, 0, 1, 1, 1, 1],       # contiguous groups of 4: grouped attention kernel
                                     [0, 0, 1, 1, 2, 2]])            # contiguous groups of 2
def test_shared_conditioning_prefill_equals_per_row_prefill(cuda, pattern):
    """Rows with identical conditioning share one prefill (leader + K/V prefix copy) and, when the groups are contiguous,
    one pass over the conditioning-prefix K/V in the decode attention: same tokens, same logits as the per-row path."""
    cfg = dict(synth.TINY_GPT, n_layers=(2, 2))
    sd = synth.gpt_state_dict(cfg, seed=33, peaky=True)
    base = synth.cond_indices(3, 21, seed=6)
    c = base[pattern]
    B, steps = c.shape[0], 8
    noise = util.noise_from_seed(2, steps, B, 4097)
    s = make_sampler(cuda, cfg, sd, B, 21, steps, prefill_tokens=2 * 21)
    outs = []
    for share in (False, True):
        x, hist = s.sample(c, steps, top_k=40, top_p=0.8, best_in_first=True, mask_invalid=True,
                           mask_invalid_completion=True, noise=noise, share_prefix=share, stop_early=False)
        outs.append((x.cpu().clone(), [h.cpu().clone() for h in hist]))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        fin = torch.isfinite(a)
        assert torch.equal(fin, torch.isfinite(b)) and (a[fin] - b[fin]).abs().max() < 2e-5
    ox, _ = O.sample_indices(sd, O.GPTSpec(**cfg), c, c[:, :0], steps, END, True, 40, 0.8, 1.0, True, True,
                             noise=O.ListNoise(noise.reshape(-1, B, 4097)), cached=True)
    assert torch.equal(outs[1][0][:, :ox.shape[1]], ox)


def test_conv_prologue_tc_kernels(cuda):
    """The shipped conv prologue (csrc/conv_tc.cu: tcgen05 3xTF32 implicit-GEMM convs, fused GroupNorm / upsample / concat passes)
    through decode_index: occupancy within the north-star 1e-4 of the oracle, and close to the cuDNN fp32 path."""
    sd = synth.vqdif_state_dict(seed=4)
    code = synth.code_grids(3, seed=3)
    Xtg = torch.rand(1, 20000, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    ref = O.decode_index(sd, code, Xtg.expand(3, -1, -1))["logits"][..., 0]
    dec = decoder.ImplicitDecoder(sd, cuda, prologue="tc")
    assert dec.conv_tc is not None
    out = dec.decode_index(code, Xtg)["logits"][..., 0].cpu()
    err = (out - ref).abs().max().item()
    occ_err = (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item()
    grid = dec.feature_grid_from_codes(code).cpu()
    ref_grid = decoder.ImplicitDecoder(sd, cuda, prologue="cudnn", unet_mode="fp32", up_mode="fp32").feature_grid_from_codes(code).cpu()
    gerr = (grid - ref_grid).abs().max().item()
    print(f"tc conv prologue: max |dlogit| = {err:.2e}, max |docc| = {occ_err:.2e}, feature grid vs cuDNN fp32 {gerr:.2e} "
          f"(max |grid| {ref_grid.abs().max().item():.2f})")
    assert occ_err < 1e-4 and err < 2e-4, (occ_err, err)
    assert gerr < 2e-4 * max(1.0, ref_grid.abs().max().item()), gerr
    # decode (NCDHW features in) takes the same path
    out2 = dec.decode(dec.get_code(code), Xtg)["logits"][..., 0].cpu()
    assert (out2 - out).abs().max() < 1e-6


@pytest.mark.parametrize("unet_mode,up_mode", [("fp32", "fp32"), ("fp32", "3xtf32"), ("3xtf32", "3xtf32")])
def test_conv_prologue_precision_modes(cuda, unet_mode, up_mode):
    """The cuDNN cross-check path: 3xTF32 tensor-core convolutions keep the decoded logits within the 1e-4 budget (plain TF32
    does not)."""
    sd = synth.vqdif_state_dict(seed=4)
    code = synth.code_grids(2, seed=3)
    Xtg = torch.rand(1, 20000, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    ref = O.decode_index(sd, code, Xtg.expand(2, -1, -1))["logits"][..., 0]
    dec = decoder.ImplicitDecoder(sd, cuda, unet_mode=unet_mode, up_mode=up_mode, prologue="cudnn")
    out = dec.decode_index(code, Xtg)["logits"][..., 0].cpu()
    err = (out - ref).abs().max().item()
    occ_err = (torch.sigmoid(out) - torch.sigmoid(ref)).abs().max().item()
    print(f"conv modes unet={unet_mode} upsampler={up_mode}: max |dlogit| = {err:.2e}, max |docc| = {occ_err:.2e}")
    if unet_mode == "fp32":      # shipped modes: occupancy within the north-star 1e-4; logits measured 6e-5 / 7e-5
        assert occ_err < 1e-4 and err < 2e-4, (occ_err, err)
    else:                        # all-3xTF32 is NOT shipped (measured 3e-4 on the logits): documented bound only
        assert occ_err < 2e-4, occ_err
