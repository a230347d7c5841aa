"""ctypes binding of libcoma_b200.so (the C ABI declared in include/coma_b200.h).

There is NO CPU fallback: if the shared library is missing or a kernel cannot run, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COMA_B200_LIB") or os.path.join(_HERE, "libcoma_b200.so")

_c = ctypes
_vp, _i64, _f32, _f64, _int = _c.c_void_p, _c.c_int64, _c.c_float, _c.c_double, _c.c_int
_f32p = _c.POINTER(_c.c_float)

# name -> argtypes, exactly the prototypes of include/coma_b200.h (device pointers travel as void*)
SIGNATURES = {
    "coma_layernorm_stats_f16": [_vp, _i64, _i64, _i64, _f32, _vp, _vp],
    "coma_host_stage_rows_f64_f32": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp],
    "coma_host_rows_equal_f64": [_vp, _i64, _vp, _i64, _vp],
    "coma_vertex_normals_f64": [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _f64, _vp, _vp],
    "coma_nearest_vertex_f64": [_vp, _i64, _vp, _i64, _vp, _vp],
    "coma_nearest_distance_f32": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    "coma_nearest_distance_backward_f32": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "coma_pair_accumulate_f32": [_vp, _vp, _i64, _i64, _i64, _f32, _f32, _vp, _vp, _vp],
    "coma_pair_accumulate_order_f32": [_vp, _vp, _i64, _i64, _i64, _f32, _f32, _int, _vp, _vp, _vp],
    "coma_orient_accumulate_f32": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _f64, _f64, _f32p, _f32p, _vp, _vp, _vp],
    "coma_orient_accumulate_cone_f32": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _f64, _f64, _f32p, _f32p, _vp, _int, _int, _vp, _vp, _vp],
    "coma_orient_accumulate_cone_ws_f32": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _f64, _f64, _f32p, _f32p, _vp, _int, _int, _vp, _vp, _vp, _vp],
    "coma_orient_bin_patches": [_c.POINTER(_c.c_double), _i64, _c.POINTER(_c.c_int32)],
    "coma_canonicalize_f32": [_vp, _i64, _vp, _i64, _f32p, _f32p, _f32, _vp, _vp],
    "coma_canonicalize_order_f32": [_vp, _i64, _vp, _i64, _f32p, _f32p, _f32, _int, _vp, _vp],
    "coma_occupancy_accumulate": [_vp, _i64, _i64, _vp, _i64, _f64, _vp, _vp],
    "coma_normalize_contact_readout_f32": [_vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _vp],
    "coma_significant_pairs": [_vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp],
    "coma_masked_max_f32": [_vp, _i64, _i64, _vp, _int, _vp, _vp],
    "coma_entropy_readout_f32": [_vp, _i64, _i64, _f32, _vp, _vp],
    "coma_entropy_readout_weighted_f32": [_vp, _i64, _i64, _f32, _vp, _f32, _vp, _vp],
    "coma_occupancy_readout_f32": [_vp, _i64, _i64, _vp, _i64, _vp, _vp],
    "coma_gemm_f16_tn": [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _vp, _i64, _vp],
    "coma_gemm_f16_ex": [_vp, _vp],
    "coma_gemm_plan": [_i64, _i64, _i64, _i64, _int, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int)],
    "coma_conv3x3_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _int, _vp, _vp, _i64, _vp],
    "coma_conv3x3_strided_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _int, _int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _int, _vp, _vp, _i64, _vp,
                                 _i64, _vp, _c.POINTER(_c.c_int), _vp],
    "coma_groupnorm_from_stats_f32": [_vp, _i64, _i64, _i64, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "coma_groupnorm_from_stats_rb_f32": [_vp, _i64, _i64, _i64, _i64, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "coma_conv3x3_f16_ws": [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _int, _vp, _vp, _i64, _vp, _i64, _vp],
    "coma_conv3x3_small_n_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp],
    "coma_conv3x3_halo_supported": [_i64, _i64, _i64, _i64, _i64],
    "coma_conv3x3_halo_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _int, _vp, _vp, _int, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _int, _vp, _i64, _vp, _vp],
    "coma_attention_fwd_f16": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp, _i64, _vp],
    "coma_attention_fwd_ex_f16": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp, _vp, _i64, _vp],
    "coma_attention_fwd_nt_f16": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp, _vp, _i64, _vp],
    "coma_groupnorm_workspace_doubles": [_i64, _int],
    "coma_groupnorm_affine_f16": [_vp, _i64, _i64, _i64, _i64, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "coma_affine_act_f16": [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _i64, _vp],
    "coma_upsample2x_affine_act_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _i64, _vp],
    "coma_im2col3x3_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _vp, _vp, _int, _vp, _i64, _vp],
    "coma_layernorm_f16": [_vp, _i64, _i64, _i64, _vp, _vp, _f32, _vp, _i64, _vp],
    "coma_softmax_rows_f16": [_vp, _i64, _i64, _i64, _vp],
    "coma_softmax_rows_causal_f16": [_vp, _i64, _i64, _i64, _i64, _vp],
    "coma_geglu_f16": [_vp, _i64, _i64, _i64, _vp, _i64, _vp],
    "coma_transpose_heads_f16": [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _vp],
    "coma_timestep_embedding_f16": [_vp, _i64, _i64, _vp, _vp],
    "coma_silu_f16": [_vp, _i64, _vp, _vp],
    "coma_cfg_ddim_step_f32": [_vp, _i64, _i64, _i64, _f32, _vp, _f64, _f64, _vp, _vp, _vp],
    "coma_assemble_unet_input_f16": [_vp, _vp, _vp, _i64, _vp, _i64, _vp],
    "coma_adaptive_mask_u8": [_vp, _vp, _i64, _i64, _i64, _int, _f32, _int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "coma_image_to_u8": [_vp, _i64, _i64, _vp, _vp],
    "coma_sample_latents_f32": [_vp, _vp, _vp, _i64, _f32, _vp, _vp],
}


class GemmArgs(ctypes.Structure):
    """struct coma_gemm_args (include/coma_b200.h)."""
    _fields_ = [("A", _vp), ("lda", _i64), ("a_s1", _i64), ("a_s2", _i64),
                ("W", _vp), ("ldw", _i64), ("w_s1", _i64), ("w_s2", _i64),
                ("out_f16", _vp), ("out_f32", _vp), ("ldo", _i64), ("o_s1", _i64), ("o_s2", _i64),
                ("residual", _vp), ("bias", _vp), ("bias_rows", _vp), ("rows_per_bias", _i64), ("bias_rows_ld", _i64),
                ("M", _i64), ("N", _i64), ("K", _i64), ("nb1", _i64), ("nb2", _i64),
                ("alpha", _f32), ("act", _int), ("workspace", _vp), ("workspace_elems", _i64), ("geglu", _int),
                ("ln_row_stats", _vp), ("ln_c1", _vp), ("ln_partials_in", _vp), ("ln_eps", _f32), ("ln_partials_out", _vp)]

_LIB = None


class ComaB200Error(RuntimeError):
    pass


def load():
    """Load the in-tree CUDA library; raise (never fall back) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ComaB200Error(
                f"{LIB_PATH} is missing: build it with `python -m coma_b200.build` (nvcc, sm_100a). "
                "coma_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.coma_b200_version.restype = _int
        lib.coma_b200_last_error.restype = _c.c_char_p
        lib.coma_b200_launch_count.restype = _i64
        lib.coma_b200_last_kernel.restype = _c.c_char_p
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _i64 if name == "coma_groupnorm_workspace_doubles" else _int
        _LIB = lib
    return _LIB


def launch_count():
    return int(load().coma_b200_launch_count())


def last_kernel():
    """Name of the kernel variant the last entry point called from this thread launched."""
    return load().coma_b200_last_kernel().decode()


def _ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ComaB200Error("coma_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise ComaB200Error("coma_b200 kernels need contiguous tensors")
    return t.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def _host3(v):
    return (ctypes.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise ComaB200Error(f"{name} failed (code {rc}): {lib.coma_b200_last_error().decode()}")
