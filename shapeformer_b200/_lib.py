"""ctypes binding of libsfb200.so (include/sfb200.h).  There is no CPU or PyTorch fallback: if the library is missing or
a call fails, this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFB200_LIB: development override (e.g. the timeline-probe build of scripts/build_probe.sh)
LIB_PATH = os.environ.get("SFB200_LIB") or os.path.join(_HERE, "lib", "libsfb200.so")

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
vp = ctypes.c_void_p


class ArConfig(ctypes.Structure):
    """struct sfb200_ar_config"""
    _fields_ = [
        ("n_embd", ctypes.c_int), ("n_head", ctypes.c_int), ("n_layers", ctypes.c_int * 2),
        ("block_size", ctypes.c_int), ("vocab", ctypes.c_int * 2), ("extra_vocab", ctypes.c_int),
        ("end_tokens", ctypes.c_int64 * 2),
        ("max_rows", ctypes.c_int), ("max_len", ctypes.c_int), ("max_steps", ctypes.c_int),
        ("prefill_rows", ctypes.c_int), ("max_cond", ctypes.c_int), ("keep_history", ctypes.c_int),
    ]


class ArSampling(ctypes.Structure):
    """struct sfb200_ar_sampling"""
    _fields_ = [
        ("top_k", ctypes.c_int), ("top_p", ctypes.c_float), ("temperature", ctypes.c_float),
        ("best_in_first", ctypes.c_int), ("mask_invalid", ctypes.c_int), ("mask_invalid_completion", ctypes.c_int),
    ]


class EncWeights(ctypes.Structure):
    """struct sfb200_enc_weights"""
    _fields_ = [
        ("fc_pos_w", c_f32p), ("fc_pos_b", c_f32p),
        ("fc0_w", c_f32p * 5), ("fc0_b", c_f32p * 5), ("fc1_w", c_f32p * 5), ("fc1_b", c_f32p * 5), ("sc_w", c_f32p * 5),
        ("fcc_w", c_f32p), ("fcc_b", c_f32p),
        ("ds_wT", c_f32p * 4), ("ds_gn_w", c_f32p * 4), ("ds_gn_b", c_f32p * 4),
        ("codebook", c_f32p), ("n_codes", ctypes.c_int),
    ]


# tensor ids of sfb200_ar_weight_offset
(W_POS_EMB, W_COND_POS_EMB, W_TOK_EMB0, W_TOK_EMB1, W_EXTRA_EMB, W_HEAD_LN_W, W_HEAD_LN_B, W_HEAD_W, W_LN1_W, W_LN1_B,
 W_QKV_W, W_QKV_B, W_PROJ_W, W_PROJ_B, W_LN2_W, W_LN2_B, W_FC1_W, W_FC1_B, W_FC2_W, W_FC2_B, W_COUNT) = range(21)

DEC_MLP_FLOATS = 32 * 3 + 32 + 5 * (3 * (32 * 32 + 32)) + 32 + 1

# name -> (restype, argtypes); mirrors include/sfb200.h one to one (tests/test_cabi.py checks the export list)
SIGNATURES = {
    "sfb200_version": (ctypes.c_int, []),
    "sfb200_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "sfb200_last_cuda_error": (ctypes.c_char_p, []),
    "sfb200_launch_count": (ctypes.c_int64, []),
    "sfb200_code_gather": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_grid_to_channels_last": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, vp]),
    "sfb200_decoder_set_weights": (ctypes.c_int, [vp, vp]),
    "sfb200_decoder_points": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_conv3d_tc": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_conv_prep": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_double, vp, ctypes.c_int, ctypes.c_int, vp,
                                        ctypes.c_double, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, vp]),
    "sfb200_pool_stats": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_gather_codes_cl": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_mesh_mark_edges": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_float, vp, vp]),
    "sfb200_mesh_emit_vertices": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_float, vp, vp, vp, vp]),
    "sfb200_mesh_count_faces": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_float, vp, vp]),
    "sfb200_mesh_emit_faces": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_float, vp, vp, vp, vp, vp]),
    "sfb200_tokens_to_dense": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                              ctypes.c_int64, vp]),
    "sfb200_encoder_workspace_bytes": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "sfb200_encode_cloud": (ctypes.c_int, [ctypes.POINTER(EncWeights), vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]),
    "sfb200_dense_to_tokens": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                              ctypes.c_int64, vp, vp, vp, vp, vp, vp]),
    "sfb200_ar_weight_floats": (ctypes.c_int64, [ctypes.POINTER(ArConfig)]),
    "sfb200_ar_weight_offset": (ctypes.c_int64, [ctypes.POINTER(ArConfig), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "sfb200_ar_kv_bytes": (ctypes.c_int64, [ctypes.POINTER(ArConfig)]),
    "sfb200_ar_workspace_bytes": (ctypes.c_int64, [ctypes.POINTER(ArConfig)]),
    "sfb200_ar_history_floats": (ctypes.c_int64, [ctypes.POINTER(ArConfig)]),
    "sfb200_ar_pretiled_floats": (ctypes.c_int64, [ctypes.POINTER(ArConfig)]),
    "sfb200_ar_set_pretiled": (ctypes.c_int, [vp, vp, vp]),
    "sfb200_ar_set_lo_weights": (ctypes.c_int, [vp, vp, vp]),
    "sfb200_big_partial_floats": (ctypes.c_int64, []),
    "sfb200_split_lo": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp]),
    "sfb200_linear_big": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         vp, vp, vp]),
    "sfb200_ar_create": (ctypes.c_int, [ctypes.POINTER(ArConfig), vp, vp, vp, vp, vp, ctypes.POINTER(vp)]),
    "sfb200_ar_destroy": (None, [vp]),
    "sfb200_ar_begin": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ArSampling), vp]),
    "sfb200_ar_begin_shared": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ArSampling), vp, vp]),
    "sfb200_ar_steps": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_int, vp]),
    "sfb200_ar_status_ptr": (vp, [vp]),
    "sfb200_ar_logprob_ptr": (vp, [vp]),
    "sfb200_ar_profile": (ctypes.c_int, [vp, ctypes.c_int]),
    "sfb200_ar_profile_read": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "sfb200_linear": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_linear_tc": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_tc_pretiled_floats": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int]),
    "sfb200_tc_pretile": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_linear_tc_ps": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_chain_workspace_bytes": (ctypes.c_int64, []),
    "sfb200_chain_linear": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                           vp, vp]),
    "sfb200_debug_chain_timeline": (ctypes.c_int, [vp]),
    "sfb200_debug_ps_timeline": (ctypes.c_int, [vp]),
    "sfb200_layernorm": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_attn_decode": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                          ctypes.c_int, vp]),
    "sfb200_attn_decode_grouped": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_attn_prefill": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]),
    "sfb200_ar_sample": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(ArSampling), vp]),
}


class Sfb200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libsfb200.so (once).  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sfb200Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(or ./build.sh). shapeformer_b200 has no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sfb200_version() != 100:
        raise Sfb200Error(f"libsfb200 version mismatch: {lib.sfb200_version()}")
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        lib = load()
        msg = lib.sfb200_error_string(code).decode()
        if code == -2:
            msg += ": " + lib.sfb200_last_cuda_error().decode()
        raise Sfb200Error(f"{what} failed: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be contiguous and on a CUDA device."""
    if t is None:
        return None
    if not t.is_cuda:
        raise Sfb200Error("libsfb200 takes device pointers only: got a CPU tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise Sfb200Error("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
