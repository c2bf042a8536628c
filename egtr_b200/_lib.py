"""ctypes binding of `libegtr_b200.so` (declared in `include/egtr_b200.h`).

The library is built in-tree by `__graft_entry__.build()` / `make -C egtr_b200/csrc`.  There is
no CPU or PyTorch fallback: if the shared object is missing, importing the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EGTR_B200_LIB", os.path.join(_HERE, "csrc", "libegtr_b200.so"))  # override: dev A/B builds only


class EgtrError(RuntimeError):
    pass


class ASrc(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a2", C.c_void_p), ("mode", C.c_int), ("lda", C.c_int),
                ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
                ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("aux", C.c_void_p), ("fmt", C.c_int)]


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("res", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int),
                ("ldr", C.c_int), ("relu", C.c_int), ("rows_per_b", C.c_int), ("bstride", C.c_int),
                ("off", C.c_int), ("row_keep", C.c_void_p),
                ("pair_n", C.c_int), ("dot_w", C.c_void_p), ("dot_out", C.c_void_p), ("dot_b", C.c_float), ("dot_col0", C.c_int),
                ("fin", C.c_int), ("fin_n", C.c_int), ("cls", C.c_void_p), ("triplet", C.c_void_p), ("adj", C.c_void_p), ("k1", C.c_int),
                ("out_fmt", C.c_int), ("res_fmt", C.c_int),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_addend", C.c_void_p), ("ln_out2", C.c_void_p)]


class RelheadWeights(C.Structure):
    """egtr_relhead_weights_t (include/egtr_b200.h)."""
    _fields_ = [("layers", C.c_int), ("uv_planes", C.c_void_p), ("uv_bias", C.c_void_p), ("uv_npad", C.c_int),
                ("b1", C.c_void_p), ("w2g", C.c_void_p), ("b2", C.c_void_p), ("w3g", C.c_void_p), ("b3", C.c_void_p),
                ("w3c", C.c_void_p), ("b3c", C.c_float)]


class DecoderWeights(C.Structure):
    """egtr_decoder_weights_t (include/egtr_b200.h)."""
    _fields_ = [("layers", C.c_int), ("n_queries", C.c_int), ("w_qkv", C.c_void_p), ("w_o", C.c_void_p), ("w_offaw", C.c_void_p),
                ("w_out", C.c_void_p), ("w_fc1", C.c_void_p), ("w_fc2", C.c_void_p), ("vec", C.c_void_p), ("qkv_pos", C.c_void_p),
                ("off_pos", C.c_void_p), ("tgt", C.c_void_p), ("ref_points", C.c_void_p)]


FMT_F32, FMT_P32, FMT_H16PAIR = 0, 1, 2


_p, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes (every function returns int status unless listed in _RESTYPES)
SIGNATURES = {
    "egtr_last_error": [],
    "egtr_abi_version": [],
    "egtr_launch_count": [],
    "egtr_launch_count_reset": [],
    "egtr_set_scratch_slot": [_i],
    "egtr_set_splitk_max": [_i],
    "egtr_set_grid_div": [_i],
    "egtr_set_grid_balance": [_i],
    "egtr_set_pdl_mode": [_i],
    "egtr_set_debug_flags": [_i],
    "egtr_split_weight_bf16": [_p, _i, _i, _i, _p, _p],
    "egtr_gemm_sbf16": [C.POINTER(ASrc), _p, _i, _i, _i, _i, C.POINTER(Epilogue), _p],
    "egtr_gemm_sbf16_grouped": [_p, _p, _p, C.POINTER(_i), _i, C.POINTER(_i), _p, _i, _i, _i, _i, _i, C.POINTER(Epilogue), _p],
    "egtr_gemm_f32_grouped": [_p, _p, _p, C.POINTER(_i), _i, C.POINTER(_i), _p, _i, _i, _i, C.POINTER(Epilogue), _p],
    "egtr_gemm_f32": [C.POINTER(ASrc), _p, _i, _i, _i, C.POINTER(Epilogue), _p],
    "egtr_gemm_f32_splitk": [_p, _p, _i, _p, _i, _i, _i, _i, _p, _p],
    "egtr_sum_layernorm_f32": [_p, _i, _ll, _p, _p, _p, _p, _i, _i, _p, _p, _i, _ll, _p],
    "egtr_msda_fwd_f32": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "egtr_msda_fused_fwd_f32": [_p, _i, C.POINTER(_i), _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "egtr_add_layernorm_f32": [_p, _p, _p, _p, _i, _i, _p, _p],
    "egtr_add_layernorm_p32": [_p, _p, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "egtr_rows_to_p32": [_p, _p, _i, _i, _i, _p, _p],
    "egtr_p32_to_rows": [_p, _i, _i, _p, _i, _p],
    "egtr_msda_fused_fwd_ex": [_p, _i, C.POINTER(_i), _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "egtr_msda_fused_fwd_h16": [_p, _ll, _i, _i, C.POINTER(_i), _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "egtr_mask_rows_f32": [_p, _i, _i, _p, _i, _p],
    "egtr_pad_nchw3_to_nhwc4_f32": [_p, _i, _i, _i, _i, _p, _p],
    "egtr_maxpool3x3s2_nhwc_f32": [_p, _i, _i, _i, _i, _p, _p],
    "egtr_maxpool3x3s2_nhwc_ex": [_p, _i, _i, _i, _i, _p, _i, _p],
    "egtr_groupnorm_f32": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "egtr_groupnorm_ex": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "egtr_groupnorm_scratch_doubles": [_i, _i],
    "egtr_levels_geometry_f32": [_p, _i, _i, _i, C.POINTER(_i), _i, _p, _p, _i, _p, _p, _p, _p, _p],
    "egtr_mha_core_f32": [_p, _i, _i, _i, _i, _i, _p, _p],
    "egtr_small_linear_f32": [_p, _i, _p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p],
    "egtr_relation_pair_hidden_f32": [_p, _p, _i, _p, _i, _i, _i, _p, _p],
    "egtr_triplets_scratch_bytes": [_i, _i, _i, _i, _i],
    "egtr_triplets_f32": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "egtr_argmax_rows_f32": [_p, _i, _i, _p, _p],
    "egtr_resample_h_u8": [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    "egtr_resample_v_normalize_f32": [_p, _i, _i, _i, _i, _p, _p, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _ll, _i, _p, _i, _p],
    "egtr_pack_weight_p32g": [_p, _i, _i, _i, _p, _p, _p],
    "egtr_relation_pairs_fused_f32": [_p, _p, _i, _i, C.POINTER(RelheadWeights), _p, _p, _i, _p, _f, _i, _i, _i, _p, _p, _p],
    "egtr_relation_head_fwd_f32": [_p, _p, _i, _p, _i, _p, _i, C.POINTER(RelheadWeights), _p, _p, _f, _i, _i, _i, _i, _i,
                                   _p, _p, _p, _p, _p, _p],
    "egtr_stem_planes_bytes": [_i, _i, _i],
    "egtr_stem_krow": [],
    "egtr_stem_layout": [],
    "egtr_stem_pad_split_bf16": [_p, _i, _i, _i, _p, _p],
    "egtr_stem_conv7x7s2_bf16x3": [_p, _i, _i, _i, _p, _p, _p, _p],
    "egtr_decoder_scratch_bytes": [_i, _i],
    "egtr_decoder_fault": [],
    "egtr_decoder_debug_profile": [_p],
    "egtr_decoder_fused_f32": [C.POINTER(DecoderWeights), _p, _p, _ll, C.POINTER(_i), _i, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p],
    "egtr_relation_finish_f32": [_p, _i, _p, _i, _p, _i, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p],
}
_RESTYPES = {
    "egtr_last_error": C.c_char_p,
    "egtr_launch_count": _ll,
    "egtr_launch_count_reset": None,
    "egtr_groupnorm_scratch_doubles": _ll,
    "egtr_triplets_scratch_bytes": _ll,
    "egtr_decoder_scratch_bytes": _ll,
    "egtr_stem_planes_bytes": _ll,
}
_NO_STATUS = set(_RESTYPES) | {"egtr_abi_version", "egtr_decoder_fault", "egtr_stem_krow", "egtr_stem_layout"}

_lib = None


def load():
    """Load the shared object (once).  Raises EgtrError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EgtrError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(no CPU fallback exists for the EGTR hot path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        if "EGTR_B200_LIB" in os.environ and not hasattr(lib, name):
            continue  # dev A/B builds of older kernels may lack newer entry points
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an entry point and raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name not in _NO_STATUS and rc != 0:
        raise EgtrError(f"{name} failed (status {rc}): {lib.egtr_last_error().decode()}")
    return rc
