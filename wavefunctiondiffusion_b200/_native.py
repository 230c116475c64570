"""ctypes binding of libwfd_b200.so (C ABI in include/wfd_b200.h).

There is no Python/CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwfd_b200.so")


class WfdError(RuntimeError):
    pass


class DecoderConfig(C.Structure):
    _fields_ = [
        ("latent_channels", C.c_int),
        ("out_channels", C.c_int),
        ("n_blocks", C.c_int),
        ("block_out_channels", C.c_int * 8),
        ("conv_groups", C.c_int * 8),
        ("block_down", C.c_int * 8),
        ("conv_init_groups", C.c_int),
        ("layers_per_block", C.c_int),
        ("norm_num_groups", C.c_int),
    ]


class TapGemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_C", C.c_int), ("a_W", C.c_int), ("a_H", C.c_int), ("a_B", C.c_int), ("a_ld", C.c_longlong),
        ("b", C.c_void_p), ("b_rows", C.c_int), ("ldb", C.c_longlong), ("b_zstride", C.c_longlong), ("b_Z", C.c_int),
        ("n_taps", C.c_int), ("cpt", C.c_int), ("tap_dx", C.c_int * 9), ("tap_dy", C.c_int * 9),
        ("a_c0", C.c_void_p),
        ("BN", C.c_int), ("N", C.c_int), ("W", C.c_int), ("H", C.c_int), ("B", C.c_int), ("w_box", C.c_int),
        ("h_box", C.c_int), ("in_stride", C.c_int),
        ("z_count", C.c_int), ("a_zmul", C.c_int), ("b_zmul", C.c_int), ("b_bbmul", C.c_int),
        ("out", C.c_void_p), ("out_fp32", C.c_int), ("ldc", C.c_longlong), ("c_zstride", C.c_longlong),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("alpha", C.c_float),
        ("rowdot", C.c_void_p), ("dotsrc", C.c_void_p), ("ld_dot", C.c_longlong),
        ("gn_sums", C.c_void_p), ("gn_cpg", C.c_int),
        ("dtype16", C.c_int),
        ("cg", C.c_int),
        ("n_group", C.c_int),
        ("reuse", C.c_int),
        ("dbg", C.c_void_p),
        ("a_mn_major", C.c_int),
        ("kc_list", C.c_void_p),
        ("kc_cnt", C.c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/wfd_b200.h and include/wfd_b200_test.h declare (tests check
# this against the headers)
SIGNATURES = {
    "wfd_version": (C.c_int, []),
    "wfd_last_error": (C.c_char_p, []),
    "wfd_launch_count": (C.c_longlong, []),
    "wfd_decoder_create": (C.c_int, [C.POINTER(DecoderConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_decoder_destroy": (None, [C.c_void_p]),
    "wfd_decoder_out_hw": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wfd_decoder_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "wfd_decoder_finalize_weights": (C.c_int, [C.c_void_p]),
    "wfd_decoder_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "wfd_decoder_bind_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "wfd_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "wfd_decoder_logits": (C.c_void_p, [C.c_void_p]),
    "wfd_decoder_export_logits": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wfd_decoder_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfd_decoder_dlogits": (C.c_void_p, [C.c_void_p]),
    "wfd_decoder_import_dlogits": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "wfd_encoder_create": (C.c_int, [C.POINTER(DecoderConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_encoder_destroy": (None, [C.c_void_p]),
    "wfd_encoder_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "wfd_encoder_finalize_weights": (C.c_int, [C.c_void_p]),
    "wfd_encoder_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "wfd_encoder_bind_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "wfd_encoder_forward_wave": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfd_encoder_forward_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "wfd_wfc_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_wfc_destroy": (None, [C.c_void_p]),
    "wfd_wfc_saved_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "wfd_wfc_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "wfd_wfc_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfd_wfc_saved_wave": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "wfd_guidance_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfd_guidance_loop": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                    C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "wfd_comm_unique_id": (C.c_int, [C.c_char_p, C.c_void_p]),
    "wfd_comm_create_nccl": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_loop_group_create": (C.c_void_p, [C.c_int]),
    "wfd_loop_group_destroy": (None, [C.c_void_p]),
    "wfd_comm_create_loopback": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_comm_destroy": (None, [C.c_void_p]),
    "wfd_comm_peer_handle": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "wfd_comm_peer_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wfd_decoder_create_band": (C.c_int, [C.POINTER(DecoderConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "wfd_decoder_attach_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wfd_wfc_attach_comm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "wfd_wfc_kc_kept": (C.c_double, [C.c_void_p]),
    "wfd_wfc_count_violations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfd_wfc_set_sparse": (C.c_int, [C.c_void_p, C.c_int]),
    "wfd_propagator_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "wfd_propagator_destroy": (None, [C.c_void_p]),
    "wfd_propagator_entries": (C.c_longlong, [C.c_void_p]),
    "wfd_propagator_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "wfd_wfc_allowed_count": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wfd_wave_to_allowed": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "wfd_tiles_finish": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wfd_argmax_tiles": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "wfd_nchw_to_nhwc16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wfd_nhwc16_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wfd_cvt16_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "wfd_test_tapgemm": (C.c_int, [C.POINTER(TapGemmDesc), C.c_int, C.c_void_p]),
    "wfd_profile_enable": (C.c_int, [C.c_int]),
    "wfd_profile_report": (C.c_int, [C.c_char_p, C.c_size_t]),
    "wfd_debug_set_ref_gemm": (C.c_int, [C.c_int]),
}

_lib = None


def lib():
    """Loads the shared library (once). Raises WfdError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WfdError(
                "libwfd_b200.so is not built (%s). Run `python -m wavefunctiondiffusion_b200.build`; "
                "there is no fallback implementation." % LIB_PATH
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = lib().wfd_last_error()
        raise WfdError("%s failed (%d): %s" % (what or "wfd call", rc, msg.decode() if msg else "?"))
    return rc


def stream_ptr():
    """Current torch CUDA stream as a void*."""
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
