"""ctypes binding of libllc.so (include/llc.h). This is the ONLY compute backend: if the library is
missing or a call fails a RuntimeError is raised — there is no eager/CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LLC_LIB") or os.path.join(_HERE, "libllc.so")  # LLC_LIB: dev A/B builds

LORA_PAD = 16   # K columns appended for the rank-r factors
LORA_LD = 64    # row-pitch growth of an augmented buffer (keeps rows 128 B aligned)

c_void = C.c_void_p
c_int = C.c_int
c_float = C.c_float
c_fp = C.c_void_p  # float* passed as raw address


class GemmEpi(C.Structure):
    _fields_ = [
        ("bias", c_void), ("resid", c_void), ("ld_resid", c_int), ("act", c_int),
        ("aux", c_void), ("ld_aux", c_int), ("out", c_void), ("ld_out", c_int),
        ("out_fp32", c_int), ("out2", c_void), ("ld_out2", c_int),
        ("ws", c_void), ("ws_bytes", C.c_size_t),
    ]


class HeadArgs(C.Structure):
    _fields_ = [
        ("x", c_void), ("cls_stride", c_int), ("ld_x", c_int),
        ("ln_g", c_void), ("ln_b", c_void), ("proj", c_void), ("text", c_void),
        ("cls_idx", c_void), ("add_mask", c_void), ("logit_scale", c_float),
        ("N", c_int), ("D", c_int), ("E", c_int), ("C", c_int),
        ("labels", c_void), ("d_feat", c_void), ("skip_logit_grad", c_int),
        ("double_softmax", c_int), ("inv_batch", c_float),
        ("feat", c_void), ("fnorm", c_void), ("logits", c_void), ("probs", c_void),
        ("loss_rows", c_void), ("pred", c_void),
        ("row_idx", c_void), ("d_fnorm", c_void), ("dlogits", c_void), ("d_is_logits", c_int),
    ]


class VitCfg(C.Structure):
    _fields_ = [
        ("image_size", c_int), ("patch", c_int), ("width", c_int), ("layers", c_int),
        ("heads", c_int), ("mlp_dim", c_int), ("embed_dim", c_int), ("lora_r", c_int),
        ("lora_scale", c_float),
    ]


class VitLayer(C.Structure):
    _fields_ = [(n, c_void) for n in (
        "wqkv_aug", "wo_aug", "wfc", "wproj", "wqkvT_aug", "woT_aug", "wfcT", "wprojT",
        "f_out_A", "f_in_B", "f_out_B", "bqkv", "bo", "bfc", "bproj", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
        "in_A", "in_B", "out_A", "out_B", "g_in_A", "g_in_B", "g_out_A", "g_out_B")]


class VitWeights(C.Structure):
    _fields_ = [("wpatch", c_void), ("class_emb", c_void), ("pos_emb", c_void),
                ("ln_pre_g", c_void), ("ln_pre_b", c_void), ("layers", C.POINTER(VitLayer))]


class TextWeights(C.Structure):
    _fields_ = [("tok_emb", c_void), ("pos_emb", c_void), ("layers", C.POINTER(VitLayer)),
                ("vocab", c_int), ("context", c_int)]


class ImgTransform(C.Structure):
    _fields_ = [("src", c_void), ("src_u8", c_int), ("h", c_int), ("w", c_int),
                ("out_size", c_int), ("pad", c_int), ("crop_i", c_int), ("crop_j", c_int),
                ("flip", c_int), ("dyn_params", c_void), ("mean", c_float * 3),
                ("std", c_float * 3)]


class BlockBufs(C.Structure):
    _fields_ = [(n, c_void) for n in (
        "x_in", "h1", "qkv", "lse", "o", "x_mid", "h2", "z", "g", "x_out")]


class BlockBwdBufs(C.Structure):
    _fields_ = [(n, c_void) for n in (
        "dx", "dxb", "dz", "dh", "d_o", "dqkv", "partial", "delta", "dy")]


class Adapter(C.Structure):
    _fields_ = [("scale", c_float), ("dropout", c_float), ("seed", C.c_ulonglong)] + \
        [(n, c_void) for n in ("down_w", "down_b", "up_w", "up_b", "g_down_w", "g_down_b",
                               "g_up_w", "g_up_b", "wd", "wu", "wdT", "wuT", "bu_s", "woT_ad",
                               "wprojT_ad")] + [("mlp_dim", c_int)]


class AdapterBufs(C.Structure):
    _fields_ = [(n, c_void) for n in ("ya", "a1", "m", "a2", "mask1", "mask2", "da", "d_branch",
                                      "partial")]


ADAPTER_DIM = 64


class FinishJob(C.Structure):
    _fields_ = [("partial", c_void), ("n_partials", c_int), ("C", c_int), ("scale", c_float),
                ("out", c_void), ("o_sc", c_int), ("o_sj", c_int)]


class ProfRec(C.Structure):
    _fields_ = [("kind", c_int), ("m", c_int), ("n", c_int), ("k", c_int), ("ms", c_float),
                ("flops", C.c_double), ("bytes", C.c_double)]


PROF_KINDS = ("gemm", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "lora_side", "head", "embed",
              "other")

# name -> (restype, argtypes); every symbol llc.h declares
SIGNATURES = {
    "llc_prof_enable": (c_int, [c_int]),
    "llc_prof_read": (c_int, [C.POINTER(ProfRec), c_int]),
    "llc_version": (c_int, []),
    "llc_last_error": (C.c_char_p, []),
    "llc_check_device": (c_int, [c_int]),
    "llc_launch_count": (C.c_ulonglong, []),
    "llc_set_pdl_trigger": (c_int, [c_int]),
    "llc_set_traversal": (c_int, [c_int]),
    "llc_gemm_ws_bytes": (C.c_size_t, []),
    "llc_gemm_set_stream_k": (c_int, [c_int]),
    "llc_gemm_bf16_tn": (c_int, [c_void, c_int, c_void, c_int, c_int, c_int, c_int,
                                 C.POINTER(GemmEpi), c_void]),
    "llc_ln_fwd": (c_int, [c_void, c_int, c_void, c_void, c_int, c_int, c_void, c_int, c_void,
                           c_int, c_void]),
    "llc_ln_bwd": (c_int, [c_void, c_int, c_void, c_void, c_int, c_void, c_void, c_int, c_int,
                           c_void, c_int, c_void, c_int, c_float, c_void]),
    "llc_attn_fwd": (c_int, [c_void, c_int, c_void, c_int, c_void, c_int, c_int, c_int, c_int,
                             c_int, c_int, c_void]),
    "llc_attn_bwd": (c_int, [c_void, c_int, c_void, c_int, c_void, c_int, c_void, c_void, c_int,
                             c_int, c_int, c_int, c_int, c_int, c_int, c_void, c_void]),
    "llc_lora_side": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_int, c_int, c_float,
                              c_void, c_int, c_void, C.POINTER(c_int), c_void]),
    "llc_lora_colsum_finish": (c_int, [c_void, c_int, c_int, c_int, c_float, c_void, c_int, c_int,
                                       c_void]),
    "llc_lora_side_max_partials": (c_int, []),
    "llc_lora_side_fused": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_int, c_void,
                                    c_int, c_void, c_int, c_void, C.POINTER(c_int), c_void]),
    "llc_lora_colsum_finish_multi": (c_int, [C.POINTER(FinishJob), c_int, c_int, c_void]),
    "llc_pack_weight": (c_int, [c_void, c_int, c_int, c_int, c_void, c_int, c_void]),
    "llc_pack_lora_cols": (c_int, [c_void, c_int, c_int, c_int, c_int, c_float, c_void, c_int,
                                   c_int, c_void]),
    "llc_pack_factor_rows": (c_int, [c_void, c_int, c_int, c_int, c_int, c_float, c_void, c_int,
                                     c_void]),
    "llc_patchify": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_int, c_void]),
    "llc_embed_ln_pre": (c_int, [c_void, c_int, c_void, c_void, c_void, c_void, c_int, c_int,
                                 c_int, c_void, c_void]),
    "llc_head_fwd": (c_int, [C.POINTER(HeadArgs), c_void]),
    "llc_head_bwd": (c_int, [C.POINTER(HeadArgs), c_void, c_float, c_void, c_int, c_void]),
    "llc_head_dtext": (c_int, [c_void, c_void, c_int, c_int, c_int, c_float, c_void, c_void]),
    "llc_eval_accum": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void, c_void, c_void]),
    "llc_l2norm_rows": (c_int, [c_void, c_int, c_int, c_int, c_float, c_void, c_int, c_void, c_int,
                                c_void]),
    "llc_softmax_argmax": (c_int, [c_void, c_int, c_int, c_int, c_void, c_void, c_int, c_void,
                                   c_void]),
    "llc_transform_images": (c_int, [C.POINTER(ImgTransform), c_int, c_void, c_void]),
    "llc_transform_patchify": (c_int, [C.POINTER(ImgTransform), c_int, c_int, c_void, c_int,
                                       c_void]),
    "llc_vit_forward_tx": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights),
                                   C.POINTER(ImgTransform), c_int, c_void, c_int, c_int,
                                   C.POINTER(c_void), c_void]),
    "llc_mha_forward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer), C.POINTER(BlockBufs),
                                c_int, c_int, c_int, c_int, c_int, c_void]),
    "llc_mha_backward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer), C.POINTER(BlockBufs),
                                 C.POINTER(BlockBwdBufs), c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void]),
    "llc_adapter_refresh": (c_int, [C.POINTER(Adapter), c_int, c_void]),
    "llc_adapter_partial_floats": (C.c_size_t, [c_int]),
    "llc_adapter_forward": (c_int, [C.POINTER(Adapter), c_void, c_int, c_void, c_int, c_void,
                                    c_void, C.c_uint, c_int, c_void, c_int, c_int, c_void]),
    "llc_adapter_backward": (c_int, [C.POINTER(Adapter), c_void, c_int, c_void, c_void, c_void,
                                     c_int, c_void, c_int, c_void, c_void, c_int, c_int, c_int,
                                     c_int, c_void]),
    "llc_adapter_block_forward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer),
                                          C.POINTER(Adapter), C.POINTER(BlockBufs),
                                          C.POINTER(AdapterBufs), c_int, c_int, c_int, c_int,
                                          c_int, c_int, c_void]),
    "llc_adapter_block_backward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer),
                                           C.POINTER(Adapter), C.POINTER(BlockBufs),
                                           C.POINTER(AdapterBufs), C.POINTER(BlockBwdBufs), c_int,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_void]),
    "llc_text_arena_bytes": (C.c_size_t, [C.POINTER(VitCfg), c_int, c_int, c_int]),
    "llc_text_forward": (c_int, [C.POINTER(VitCfg), C.POINTER(TextWeights), c_void, c_int, c_void,
                                 c_int, C.POINTER(c_void), c_void]),
    "llc_text_backward": (c_int, [C.POINTER(VitCfg), C.POINTER(TextWeights), c_int, c_void,
                                  c_void, c_void]),
    "llc_label_remap": (c_int, [c_void, c_void, c_int, c_void, c_int, c_void]),
    "llc_loss_acc": (c_int, [c_void, c_void, c_void, c_int, c_void, c_void]),
    "llc_adamw": (c_int, [c_void, c_void, c_void, c_void, c_int, c_float, c_float, c_float,
                          c_float, c_float, c_int, c_float, c_void]),
    "llc_vit_arena_bytes": (C.c_size_t, [C.POINTER(VitCfg), c_int, c_int]),
    "llc_vit_forward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_void, c_int, c_void,
                                c_int, C.POINTER(c_void), c_void]),
    "llc_vit_backward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_int, c_void, c_void,
                                 c_void]),
    "llc_vit_forward_cls": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_void, c_int, c_void,
                                    c_int, C.POINTER(c_void), c_void]),
    "llc_vit_backward_cls": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_int, c_void,
                                     c_void, c_void]),
    "llc_cast_bf16": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void]),
    "llc_vit_refresh_lora": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_void]),
    "llc_block_forward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer), C.POINTER(BlockBufs),
                                  c_int, c_int, c_int, c_int, c_int, c_void]),
    "llc_block_backward": (c_int, [C.POINTER(VitCfg), C.POINTER(VitLayer), C.POINTER(BlockBufs),
                                   C.POINTER(BlockBwdBufs), c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void]),
}

_lib = None


def load() -> C.CDLL:
    """Load libllc.so and bind every declared symbol. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). lifelong_clip_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("LLC_STREAM_K") == "0":     # A/B measurements of the GEMM schedule
        lib.llc_gemm_set_stream_k(0)
    if os.environ.get("LLC_TRAVERSAL") is not None:
        lib.llc_set_traversal(int(os.environ["LLC_TRAVERSAL"]))
    if os.environ.get("LLC_PDL_TRIGGER") is not None:
        lib.llc_set_pdl_trigger(int(os.environ["LLC_PDL_TRIGGER"]))
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().llc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libllc {what} failed (rc={rc}): {msg}")


def ptr(t) -> int | None:
    """Raw device address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
