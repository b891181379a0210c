"""ctypes loader for the C-ABI library.  There is NO CPU fallback: if the library is missing this raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvame_b200.so")

_lib = None

c_void_p, c_int, c_long, c_size_t, c_float, c_double = (ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_size_t,
                                                        ctypes.c_float, ctypes.c_double)

# name -> (restype, argtypes); mirrors include/vame_b200.h
SIGNATURES = {
    "vame_last_error": (ctypes.c_char_p, []),
    "vame_abi_version": (c_int, []),
    "vame_launch_count": (c_long, []),
    "vame_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "vame_get_option": (c_int, [ctypes.c_char_p]),
    "vame_set_debug_buffer": (c_int, [c_void_p]),
    "vame_debug_timeline": (c_int, [c_int]),
    "vame_debug_timeline_read": (c_int, [c_void_p, c_void_p, c_int]),
    "vame_debug_gru_sweep": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_p16_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vame_pack_p16": (c_int, [c_void_p, c_long, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vame_gemm_p16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_long, c_void_p, c_int, c_int, c_void_p]),
    "vame_param_tensors": (c_int, [c_void_p]),
    "vame_param_layout": (c_long, [c_void_p, c_void_p, c_void_p]),
    "vame_packed_weights_bytes": (c_size_t, [c_void_p]),
    "vame_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "vame_pack_weights_deferred": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "vame_arm_prior": (c_int, [c_void_p, c_void_p]),
    "vame_trainset_workspace_bytes": (c_size_t, [c_long, c_int]),
    "vame_trainset_zscore_clean": (c_int, [c_void_p, c_long, c_int, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_trainset_row_std": (c_int, [c_void_p, c_long, c_int, c_void_p, c_void_p]),
    "vame_trainset_savgol": (c_int, [c_void_p, c_long, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vame_pack_weights_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vame_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "vame_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_long, c_long, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_loss": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_void_p, c_int,
                          c_void_p, c_size_t, c_void_p]),
    "vame_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_adam_prepare": (c_int, [c_float, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p]),
    "vame_adam_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_float, c_float, c_float, c_float,
                                c_void_p]),
    "vame_pack_weights_train_part": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vame_grad_overlap": (c_int, [c_int]),
    "vame_grad_bucket_split": (c_long, [c_void_p]),
    "vame_wait_grads_ready": (c_int, [c_void_p]),
    "vame_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_float, c_void_p, c_void_p, c_void_p,
                               c_float, c_float, c_float, c_float, c_void_p]),
    "vame_embed_workspace_bytes": (c_size_t, [c_void_p, c_long, c_int]),
    "vame_embed_windows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_long, c_long, c_int, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "vame_encoder_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_long, c_long, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_lambda_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "vame_decoder_forward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_cluster_loss": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "vame_sample_windows": (c_int, [c_void_p, c_long, c_int, c_int, c_double, c_double, c_int, c_int, c_int, c_int, c_void_p,
                                    ctypes.c_ulonglong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vame_kmeans_workspace_bytes": (c_size_t, [c_long, c_int, c_int]),
    "vame_kmeans_lloyd": (c_int, [c_void_p, c_long, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "vame_kmeans_assign": (c_int, [c_void_p, c_long, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vame_kmeans_candidates": (c_int, [c_void_p, c_long, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vame_kmeans_sample": (c_int, [c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}


class VameB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VameB200Error(
                "vame_b200: %s not found. Build it with `python -m vame_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        # runtime switches (debugging / A-B measurements): VAME_B200_PDL, VAME_B200_STREAMS, VAME_B200_PERSISTENT = 0 | 1
        for opt in ("pdl", "streams", "persistent", "flags", "warps16", "slice16", "m64", "rw", "rw2", "rw_priv", "rw_sw", "rw_waves", "rw_exp", "rw_ng", "rows"):
            v = os.environ.get("VAME_B200_" + opt.upper())
            if v is not None:
                L.vame_set_option(opt.encode(), int(v))
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().vame_last_error()
        raise VameB200Error("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
