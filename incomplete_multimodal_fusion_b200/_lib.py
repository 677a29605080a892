"""ctypes binding of libmmf_b200.so (C ABI declared in include/mmf_b200.h).

The library is the ONLY compute path of this package: if it cannot be loaded the import of any op
fails loudly -- there is no PyTorch / CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MMF_LIB: another build of the same library (e.g. scratch/new_libmmf.so from a work-in-progress revision) for A/B runs of
# tools/ and tests/ on one box; unset, the in-tree library is the only one ever loaded
LIB_PATH = os.path.abspath(os.environ["MMF_LIB"]) if os.environ.get("MMF_LIB") else os.path.join(HERE, "libmmf_b200.so")

_lib = None

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", c_vp), ("b", c_vp), ("out", c_vp), ("bias", c_vp), ("residual", c_vp), ("residual2", c_vp),
        ("res_split", c_i64), ("res_row_map", c_vp),
        ("M", c_i64), ("N", c_i64), ("K", c_i64),
        ("lda", c_i64), ("ldb", c_i64), ("ldo", c_i64), ("ldr", c_i64),
        ("a_mn", c_i32), ("b_mn", c_i32), ("out_f32", c_i32), ("act", c_i32), ("split_k", c_i32),
        ("res_period", c_i32), ("out_period", c_i32), ("out_batch_rows", c_i32), ("block_n", c_i32),
        ("alpha", c_f32), ("out2", c_vp), ("ldo2", c_i64), ("accumulate", c_i32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", c_vp), ("k", c_vp), ("v", c_vp), ("o", c_vp), ("lse", c_vp),
        ("ldq", c_i64), ("ldk", c_i64), ("ldv", c_i64), ("ldo", c_i64),
        ("B", c_i32), ("H", c_i32), ("Nq", c_i32), ("Nk", c_i32), ("dh", c_i32),
        ("n_head_q", c_i32), ("n_tail_q", c_i32), ("n_head_k", c_i32), ("n_tail_k", c_i32),
        ("scale", c_f32), ("seg", c_vp), ("nseg", c_i32),
        ("d_o", c_vp), ("lddo", c_i64), ("delta", c_vp),
        ("dq", c_vp), ("dk", c_vp), ("dv", c_vp),
        ("lddq", c_i64), ("lddk", c_i64), ("lddv", c_i64),
    ]


class SlotAttnArgs(C.Structure):
    _fields_ = [
        ("q", c_vp), ("kv_tok", c_vp), ("kv_me", c_vp), ("slotmap", c_vp), ("seg", c_vp), ("out", c_vp), ("probs", c_vp),
        ("ldq", c_i64), ("ldkv", c_i64), ("ldme", c_i64), ("ldo", c_i64),
        ("B", c_i32), ("F", c_i32), ("H", c_i32), ("S", c_i32), ("dh", c_i32), ("n_head", c_i32),
        ("scale", c_f32),
        ("dout", c_vp), ("dq", c_vp), ("dkv_tok", c_vp), ("dkv_me", c_vp),
        ("lddout", c_i64), ("lddq", c_i64), ("lddkv", c_i64), ("lddme", c_i64),
        ("me_scratch", c_vp),
    ]


class PoolAttnArgs(C.Structure):
    _fields_ = [
        ("q", c_vp), ("kv", c_vp), ("mask", c_vp), ("mode", c_vp), ("out", c_vp), ("stat", c_vp),
        ("q_bstride", c_i64), ("ldkv", c_i64),
        ("B", c_i32), ("R", c_i32), ("H", c_i32), ("N", c_i32), ("dh", c_i32), ("n_head", c_i32), ("n_tail", c_i32),
        ("scale", c_f32),
        ("dout", c_vp), ("dq", c_vp), ("dq_bstride", c_i64), ("dkv", c_vp), ("lddkv", c_i64),
    ]


class AdamWTensor(C.Structure):
    _fields_ = [
        ("p", c_vp), ("g", c_vp), ("m", c_vp), ("v", c_vp), ("w16a", c_vp), ("w16b", c_vp),
        ("n", c_i64), ("pitch16a", c_i64), ("pitch16b", c_i64),
        ("cols", c_i32), ("bias_correction1", c_f32), ("bias_correction2", c_f32),
    ]


# name -> argtypes (restype is int unless listed in _RESTYPES).  Must list every symbol of mmf_b200.h.
SIGNATURES = {
    "mmf_abi_version": [],
    "mmf_launch_count": [],
    "mmf_reset_launch_count": [],
    "mmf_set_gemm_reserved_sms": [c_i32],
    "mmf_gemm_bf16": [C.POINTER(GemmArgs), c_vp],
    "mmf_layernorm_fwd": [c_vp, c_vp, c_i64, c_i64, c_i32, c_i64, c_vp, c_vp, c_f32, c_vp, c_f32, c_vp, c_i64, c_i32, c_vp,
                          c_vp, c_i64, c_i64, c_vp, c_i64, c_vp],
    "mmf_layernorm_bwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                          c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp],
    "mmf_attn_fwd": [C.POINTER(AttnArgs), c_vp],
    "mmf_attn_bwd": [C.POINTER(AttnArgs), c_vp],
    "mmf_slot_attn_fwd": [C.POINTER(SlotAttnArgs), c_vp],
    "mmf_slot_attn_bwd": [C.POINTER(SlotAttnArgs), c_vp],
    "mmf_pool_attn_fwd": [C.POINTER(PoolAttnArgs), c_vp],
    "mmf_pool_attn_bwd": [C.POINTER(PoolAttnArgs), c_vp],
    "mmf_masked_loss_fwd": [c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mmf_masked_loss_bwd": [c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mmf_masked_ce_fwd": [c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp],
    "mmf_masked_ce_bwd": [c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp],
    "mmf_cast_f32_bf16": [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_f32, c_vp],
    "mmf_geglu_bwd": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp],
    "mmf_gelu_bwd": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "mmf_colsum": [c_vp, c_i32, c_i64, c_i64, c_i64, c_vp, c_vp],
    "mmf_bcast_rows": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "mmf_reduce_batch": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "mmf_im2col_gather": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64, c_vp],
    "mmf_onehot_im2col": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64, c_vp],
    "mmf_raster_prep": [c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                        c_vp, c_vp, c_i32, c_i32, c_vp, c_vp],
    "mmf_trunc_standardize": [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp],
    "mmf_unpatchify_bf16": [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "mmf_gather_rows": [c_vp, c_i32, c_i64, c_i64, c_i64, c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_i32, c_vp],
    "mmf_add_inplace_f32": [c_vp, c_vp, c_i64, c_vp],
    "mmf_add_bf16_f32": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "mmf_dino_loss": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp, c_vp, c_vp],
    "mmf_hardneg_workspace_floats": [c_i32, c_i32],
    "mmf_hardneg_loss": [c_vp, c_i64, c_vp, c_i64, c_i32, c_i32, c_f32, c_f32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "mmf_adamw_step": [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_f32, c_f32, c_f32, c_f32, c_vp, c_vp],
    "mmf_grad_norm": [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_vp],
    "mmf_mask_build": [c_vp, c_vp, c_vp, c_i32, C.POINTER(c_i32), c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "mmf_mask_explicit": [c_vp, c_i32, C.POINTER(c_i32), c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "mmf_im2col_tokens": [C.POINTER(c_vp), C.POINTER(c_i32), C.POINTER(c_i32), C.POINTER(c_i32), c_i32, c_vp, c_i32, c_vp, c_i64, c_i32,
                          c_i64, c_i32, c_i32, c_i32, c_vp],
}
_RESTYPES = {"mmf_hardneg_workspace_floats": c_i64, "mmf_launch_count": c_i64, "mmf_reset_launch_count": None, "mmf_set_gemm_reserved_sms": None}


def lib_path() -> str:
    return LIB_PATH


def load(build_if_missing: bool = True):
    """dlopen libmmf_b200.so (building it in-tree first if it is absent) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError("libmmf_b200.so is missing (%s): run `python -m incomplete_multimodal_fusion_b200._build`" % LIB_PATH)
        from . import _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.mmf_abi_version() != 1:
        raise RuntimeError("libmmf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError("%s: invalid argument (code %d)" % (what, -rc))
    raise RuntimeError("%s: CUDA error %d" % (what, rc))
