#!/usr/bin/env python
"""bench.py — completed shapes/sec at 64^3 (512-tuple AR sampling + 262,144-point decode) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference algorithm's CPU arm (oracle port), same metric

One "step" = one pass of the hot path over one batch per GPU: prefill (L_cond = 256 conditioning tuples) + 512 AR steps
(KV-cached, top-k = 50, fixed-length mode) for 64 rows (16 shapes x sample_n 4) -> tokens -> dense 16^3 code grids ->
VQDIF decode of the full 64^3 query grid.  Rows shard across GPUs with no data-path collective inside the loop (weak
scaling: 64 rows per GPU); weights are broadcast once over NCCL and per-row outputs are all-gathered at the end of a step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Keep stdout to the single JSON line the driver parses: C libraries (NCCL prints "NCCL version ..." on fd 1) and stray
# prints are diverted to stderr by pointing fd 1 at fd 2; the JSON line is written to the saved original stdout at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())

import torch  # noqa: E402

END = (4096, 4096)
METRIC = "completed shapes/sec at 64^3 (512-tok AR + 262k-pt decode)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=64, help="rows (sampled completions) per GPU")
    ap.add_argument("--sample-n", type=int, default=4, help="rows sharing one conditioning (VisShapeFormer.sample_n)")
    ap.add_argument("--lcond", type=int, default=256)
    ap.add_argument("--ar-steps", type=int, default=512)
    ap.add_argument("--grid", type=int, default=64, help="query grid resolution per axis")
    ap.add_argument("--top-k", type=int, default=50)
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--tiny", action="store_true", help="tiny transformer (smoke runs; NOT a valid bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(alg_bytes_per_launch):
    """Per-launch DRAM bytes of the attention kernel: the committed `ncu --set full` capture measured
    dram__bytes_read+write at one context length; its ratio to the algorithmic bytes (1.03: no re-reads) is applied to this
    run's average algorithmic bytes per launch."""
    p = os.path.join(ROOT, "profiles", "attn_decode_ncu.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["traffic_over_algorithmic"] * alg_bytes_per_launch
        except Exception:
            pass
    return None


# ----------------------------------------------------------------------------------------------------------------------
def pick_cpu_threads(gpt_sd, cfg):
    """torch's CPU kernels do not scale to every core of a big host (128 threads were 20x slower than 8 on the bench box):
    time one transformer block of the reference algorithm at L = 256 for several thread counts and keep the fastest."""
    from oracle import sf_oracle as O
    n = os.cpu_count() or 1
    x = torch.randn(1, 256, cfg["n_embd"])
    best, best_t = 1, float("inf")
    for t in sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n}):
        torch.set_num_threads(t)
        with torch.no_grad():
            O.gpt_block(gpt_sd, "blocks.0.0.", x, cfg["n_head"])
            t0 = time.perf_counter()
            for _ in range(3):
                O.gpt_block(gpt_sd, "blocks.0.0.", x, cfg["n_head"])
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_sample(args, gpt_sd, vq_sd, cfg, threads):
    """The reference algorithm on the host cores (oracle port, faithful = UNCACHED like the reference): time one row's AR
    step at three context lengths and one 64^3 decode, then integrate over the 512-step schedule.  Returns a dict."""
    from oracle import sf_oracle as O
    torch.set_num_threads(threads)
    spec = O.GPTSpec(**cfg)
    from shapeformer_b200 import synth
    Lc, S = args.lcond, args.ar_steps
    lens = sorted({Lc, Lc + S // 2, Lc + S - 1})
    t_at = {}
    g = torch.Generator().manual_seed(0)
    for L in lens:
        c = synth.cond_indices(1, Lc, seed=1)
        z = torch.stack([torch.sort(torch.randperm(4096, generator=g)[:L - Lc])[0],
                         torch.randint(0, 4096, (L - Lc,), generator=g)], -1)[None]
        idx = torch.cat([c, z], 1)
        extra = O.extra_indices(c, z, END[0])
        t0 = time.perf_counter()
        with torch.no_grad():
            x = O.gpt_embed(gpt_sd, idx, extra, Lc)
            x = O.gpt_group(gpt_sd, spec, 0, x)
            l0 = O.gpt_head(gpt_sd, 0, x)[:, -1]
            l0 = O.sampling_masker(l0, torch.cat([idx, idx[:, -1:]], 1), Lc, L - Lc, 0, END, False, False)
            q = torch.empty(1, 4097).exponential_(1.0)
            O.sample_rows(l0, q, args.top_k, 0.0, 1.0); O.sample_rows(l0, q, 1, 0.001, 1.0)
            x = x + gpt_sd["tok_embs.0.weight"][idx[:, :, 0]]
            x = O.gpt_group(gpt_sd, spec, 1, x)
            l1 = O.gpt_head(gpt_sd, 1, x)[:, -1]
            O.sample_rows(l1, q, args.top_k, 0.0, 1.0); O.sample_rows(l1, q, 1, 0.001, 1.0)
        t_at[L] = time.perf_counter() - t0
    # piecewise-linear integral of t(L) over L = Lc .. Lc+S-1
    ar = 0.0
    for j in range(S):
        L = Lc + j
        lo = max(l for l in lens if l <= L)
        hi = min(l for l in lens if l >= L)
        ar += t_at[lo] if hi == lo else t_at[lo] + (t_at[hi] - t_at[lo]) * (L - lo) / (hi - lo)
    code = synth.code_grids(1, seed=2)
    Xtg = synth.make_grid(args.grid)[None]
    t0 = time.perf_counter()
    with torch.no_grad():
        O.decode_index(vq_sd, code, Xtg)
    dec = time.perf_counter() - t0
    per_row = ar + dec
    return {"value": 1.0 / per_row, "unit": "shapes/s", "cores": threads, "host_cores": os.cpu_count(), "kind": "port",
            "sample": (f"oracle port of the reference (uncached full forward per step, torch CPU fp32), 1 row: AR steps timed "
                       f"at L={lens} ({', '.join(f'{t_at[l]:.2f}s' for l in lens)}) integrated over {S} steps = {ar:.0f}s, "
                       f"+ one {args.grid}^3 decode = {dec:.2f}s; cost is linear in rows"),
            "sample_seconds": sum(t_at.values()) + dec}


# ----------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from shapeformer_b200 import synth
    cfg = synth.TINY_GPT if args.tiny else synth.SHIPPED_GPT
    if args.tiny:
        cfg = dict(cfg, block_size=812)
    workload = {"workload": (f"cfg4/5-style completion batch per GPU: {args.rows} rows ({args.rows // args.sample_n} shapes x "
                             f"sample_n {args.sample_n}), L_cond {args.lcond}, {args.ar_steps} AR steps fixed-length "
                             f"(masks off), top_k {args.top_k}, top_p 0, T 1, best_in_first, + {args.grid}^3 decode"),
                "rows_per_gpu": args.rows, "l_cond": args.lcond, "ar_steps": args.ar_steps, "grid": args.grid,
                "transformer": "tiny (INVALID as a bench number)" if args.tiny else "shipped 20+4 x 1024 (325M)",
                "weights": "synthetic seed 314 (reference init)", "parallelism": f"rows sharded x{world}",
                "l2_policy": "inputs larger than L2 (KV cache 9.7 GB, feature grids 2.1 GB per batch)"}

    if args.impl == "reference":
        if rank != 0:
            return
        gpt_sd = synth.gpt_state_dict(cfg, seed=314, peaky=False)
        vq_sd = synth.vqdif_state_dict(seed=314)
        threads = pick_cpu_threads(gpt_sd, cfg)
        vals, secs = [], []
        for i in range(args.warmup + args.steps):
            r = cpu_reference_sample(args, gpt_sd, vq_sd, cfg, threads)
            if i >= args.warmup:
                vals.append(r["value"]); secs.append(r["sample_seconds"])
        v = statistics.mean(vals)
        r["value"] = v
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": "shapes/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": workload, "cpu_baseline": r,
                          "e2e": {"value": v, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from shapeformer_b200 import _lib, ar, decoder
    from shapeformer_b200.models.shapeformer.shapeformer import ShapeFormer
    from shapeformer_b200.models.vqdif.vqdif import VQDIF
    lib = _lib.load()

    # ---- weights: generated on rank 0, ONE NCCL broadcast of the flattened state dicts
    gpt_sd = synth.gpt_state_dict(cfg, seed=314, peaky=False) if rank == 0 else None
    vq_sd = synth.vqdif_state_dict(seed=314)
    pre = "shapeformer_b200.models."
    model = ShapeFormer(tuple_n=2, block_size=cfg["block_size"], end_tokens=list(END), vocab_sizes=list(cfg["vocab_sizes"]),
                        extra_vocab_sizes=list(cfg["extra_vocab_sizes"]),
                        transformer_opt={"class": pre + "shapeformer.transformer.mingpt.CondTupleGPT",
                                         "kwargs": dict(tuple_n=2, vocab_sizes=cfg["vocab_sizes"],
                                                        extra_vocab_sizes=cfg["extra_vocab_sizes"],
                                                        n_layers=cfg["n_layers"], block_size=cfg["block_size"],
                                                        n_head=cfg["n_head"], n_embd=cfg["n_embd"])},
                        representer_opt={"class": pre + "shapeformer.representers.AR_N",
                                         "kwargs": dict(block_size=cfg["block_size"], end_tokens=list(END),
                                                        mask_invalid=False, mask_invalid_completion=False, vqvae_opt=None)})
    vq = VQDIF(decoder_opt={"class": pre + "vqdif.dec.LocalDecoder",
                            "kwargs": dict(sample_mode="bilinear", hidden_size=32, c_dim=32, unet3d=True,
                                           unet3d_kwargs=dict(num_levels=3, f_maps=128, in_channels=128, out_channels=128),
                                           upsampler=True, upsampler_kwargs=dict(in_channels=128, upsampler_steps=2))},
               quantizer_opt={"class": pre + "vqdif.quantizer.Quantizer", "kwargs": dict(vocab_size=4096, n_embd=128)})
    if rank == 0:
        model.transformer.load_state_dict(gpt_sd)
        vq.load_state_dict(vq_sd, strict=False)
    model.to(dev); vq.to(dev)
    from shapeformer_b200 import dist as sdist
    sdist.broadcast_parameters([p.data for p in list(model.transformer.parameters()) + list(vq.parameters())], src=0)
    model.representer.vqvae_model = vq
    model.history_device = None     # value arm: history stays off; the e2e arm turns the reference's CPU history on

    B, Lc, S, R = args.rows, args.lcond, args.ar_steps, args.grid
    # distinct conditioning per shape, repeated sample_n times (shapeformer.py:229), different per rank
    c_host = synth.cond_indices(B // args.sample_n, Lc, seed=1000 + rank).repeat_interleave(args.sample_n, 0).pin_memory()
    c_dev = c_host.to(dev)
    xtg_host = synth.make_grid(R)[None].pin_memory()
    xtg_dev = xtg_host.to(dev)
    empty = torch.full((B,), 17, dtype=torch.int64, device=dev)
    eng = vq.engine()
    use_graph = {"auto": True, "on": True, "off": False}[args.graph]
    sampler = model.transformer.sampler(B, Lc, S, END, keep_history=False)
    gather_tok = [torch.empty(B, S, 2, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None
    gather_occ = [torch.empty(B, R ** 3, dtype=torch.float32, device=dev) for _ in range(world)] if world > 1 else None

    def step_value():
        x, _ = sampler.sample(c_dev, S, top_k=args.top_k, top_p=0.0, temperature=1.0, best_in_first=True,
                              mask_invalid=False, mask_invalid_completion=False, use_graph=use_graph, stop_early=False)
        dense = eng.tokens_to_dense(x, empty)
        occ = eng.occupancy(dense, xtg_dev)
        if world > 1:
            dist.all_gather(gather_tok, x.contiguous())
            dist.all_gather(gather_occ, occ)
        return x, occ

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    for _ in range(args.warmup):
        step_value()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = lib.sfb200_launch_count()
    ms = timed(step_value, args.steps)
    launches = lib.sfb200_launch_count() - l0
    clk = clocks.stop()

    # ---- roofline of the dominant kernel (single-query attention + KV append): one more pass over the SAME batch with eager
    #      launches so that every attention launch can be bracketed by CUDA events on the launching stream
    import ctypes
    roof = None
    _lib.check(lib.sfb200_ar_profile(sampler.handle, 1))
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    sampler.sample(c_dev, S, top_k=args.top_k, top_p=0.0, temperature=1.0, best_in_first=True, mask_invalid=False,
                   mask_invalid_completion=False, use_graph=False, stop_early=False)
    pe1.record()
    torch.cuda.synchronize()
    a_ms, a_n, a_b, a_br = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.sfb200_ar_profile_read(sampler.handle, ctypes.byref(a_ms), ctypes.byref(a_n), ctypes.byref(a_b),
                                          ctypes.byref(a_br)))
    _lib.check(lib.sfb200_ar_profile(sampler.handle, 0))
    peak, how = measured_peaks()
    if a_n.value:
        ach = a_b.value / (a_ms.value * 1e-3) / 1e9
        roof = {"kernel": "attn_decode_kernel (single-query attention + KV append)", "bound": "hbm", "achieved": ach,
                "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic(a_b.value / a_n.value),
                "peak_source": how + "; sustained-style figure (kernel timed inside a long step)",
                "launches": a_n.value, "avg_launch_us": 1e3 * a_ms.value / a_n.value,
                "algorithmic_bytes_per_launch": a_b.value / a_n.value,
                "per_row_formula_bytes_per_launch": a_br.value / a_n.value,
                "per_row_formula_gbs": a_br.value / (a_ms.value * 1e-3) / 1e9,
                "note": "achieved = bytes that must move / time: the conditioning-prefix K/V of the sample_n rows of a shape is "
                        "read once per shape (SURVEY §7 item 5); per_row_formula_* applies SURVEY §8d's per-row byte count",
                "share_of_ar_pass": a_ms.value / pe0.elapsed_time(pe1),
                "how": "CUDA events around every attention launch of one eager AR pass over the same batch, right after the "
                       "timed region (the timed region replays the step as a CUDA graph, where events cannot be read)"}
    rows_total = B * world
    value = rows_total * args.steps / (ms * 1e-3)

    # ---- e2e: through the reference-facing model API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        model.history_device = "cpu"
        tok_host = torch.empty(B, S, 2, dtype=torch.int64).pin_memory()
        occ_host = torch.empty(B, R ** 3, dtype=torch.float32).pin_memory()

        def step_e2e():
            c = c_host.to(dev, non_blocking=True)
            out_x, x, hist = model.sample(c_indices=c, z_indices=c[:, :0], max_steps=S, temperature=1.0, sample=True,
                                          best_in_first=True, top_k=args.top_k, top_p=0.0)
            # NB: early exit is part of the API; with masks off and random weights no row ends, so all S steps run
            dense = eng.tokens_to_dense(out_x, empty)
            occ = vq.engine().occupancy(dense, xtg_host.to(dev, non_blocking=True))
            tok_host.copy_(out_x, non_blocking=True)
            occ_host.copy_(occ, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        step_e2e()
        n_e2e = max(1, min(args.steps, 3))
        ms_e = timed(step_e2e, n_e2e)
        hist_bytes = 2 * B * S * 4097 * 4
        e2e = {"value": rows_total * n_e2e / (ms_e * 1e-3), "unit": "shapes/s", "steps": n_e2e,
               "h2d_bytes_per_step": int(c_host.numel() * 8 + xtg_host.numel() * 4),
               "d2h_bytes_per_step": int(tok_host.numel() * 8 + occ_host.numel() * 4 + hist_bytes),
               "api": "ShapeFormer.sample(...) [tokens + CPU logits history like the reference] -> tokens_to_dense -> "
                      "VQDIF occupancy; pinned host buffers"}
        model.history_device = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd_cpu = {k: v.detach().cpu() for k, v in model.transformer.state_dict().items()}
        vq_cpu = {k: v.detach().cpu() for k, v in vq.state_dict().items()}
        cpu = cpu_reference_sample(args, sd_cpu, vq_cpu, cfg, pick_cpu_threads(sd_cpu, cfg))

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "shapes/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload, "clocks": clk,
               "gpu_launches": int(launches), "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
               "stepping": "cuda graph replay of one AR step" if use_graph else "eager launches"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
