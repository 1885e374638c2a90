"""ctypes binding of libfsb200.so (C ABI declared in include/fsb200.h).

The shared library is built in-tree by `csrc/Makefile` (`fsb200.build()` / `__graft_entry__.build()`).
There is NO fallback: if the library is missing or the device is not sm_100, every entry point raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsb200.so")
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
c_ll, c_ull, c_char_p, c_double = ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_char_p, ctypes.c_double
MAX_BLOCKS = 8


class NetConfig(ctypes.Structure):
    """struct fsb_net_config (include/fsb200.h)."""
    _fields_ = [("two_d", c_int), ("feat_mode", c_int), ("n_fft", c_int), ("hop", c_int),
                ("n_features", c_int), ("num_blocks", c_int), ("depth", c_int * MAX_BLOCKS),
                ("start_deep_supervision_on", c_int), ("n_classes", c_int), ("dropout_p", c_float),
                ("precision", c_int), ("aggregation", c_int)]


_SIGNATURES = {
    "fsb_version": (c_int, []),
    "fsb_last_error": (c_char_p, []),
    "fsb_device_ok": (c_int, []),
    "fsb_feat_table_bytes": (c_size_t, [c_int]),
    "fsb_feat_init_tables": (c_int, [c_int, c_void_p, c_void_p]),
    "fsb_feat_forward": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_ll, c_ll, c_void_p]),
    "fsb_lsep_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fsb_lsep_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fsb_lsep_stable_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fsb_lsep_stable_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "fsb_lwlrap_scratch_bytes": (c_size_t, [c_int]),
    "fsb_lwlrap": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fsb_assemble_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_float, c_void_p, c_void_p,
                                   c_void_p]),
    "fsb_adam_chunk": (c_int, []),
    "fsb_adam_amsgrad_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                                      c_float, c_float, c_void_p]),
    "fsb_mixup_equal": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p]),
    "fsb_net_create": (c_int, [ctypes.POINTER(NetConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               ctypes.POINTER(c_void_p)]),
    "fsb_net_destroy": (None, [c_void_p]),
    "fsb_net_num_params": (c_int, [c_void_p]),
    "fsb_net_num_bn": (c_int, [c_void_p]),
    "fsb_net_param_numel": (c_ll, [c_void_p, c_int]),
    "fsb_net_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "fsb_net_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_ull, c_void_p, c_size_t, c_void_p, c_void_p]),
    "fsb_net_forward_features": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_ull, c_void_p, c_size_t, c_void_p, c_void_p]),
    "fsb_net_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fsb_net_read_activation": (c_int, [c_void_p, c_int, c_void_p, c_ll, ctypes.POINTER(c_ll), c_void_p, c_void_p]),
    "fsb_net_set_profiling": (c_int, [c_void_p, c_int]),
    "fsb_net_set_overlap": (c_int, [c_void_p, c_int]),
    "fsb_net_set_graphs": (c_int, [c_void_p, c_int]),
    "fsb_net_get_timings": (c_int, [c_void_p, c_int, ctypes.POINTER(c_char_p), ctypes.POINTER(c_float),
                                    ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
    "fsb_launch_count": (c_ll, [c_int]),
    "fsb_conv_workspace_bytes": (c_size_t, [c_int] * 7),
    "fsb_conv_forward": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "fsb_conv_backward": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 +
                          [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)
_lib = None


def build(verbose=False):
    """Compile libfsb200.so for sm_100a with nvcc (works without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libfsb200.so failed")
    return LIB_PATH


def lib():
    """Load the shared library (once) and attach the prototypes.  No CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libfsb200.so not found at %s -- build it with `make -C %s` (or __graft_entry__.build()); "
                "there is no CPU fallback" % (LIB_PATH, CSRC))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().fsb_last_error().decode("utf-8", "replace")
        raise RuntimeError("libfsb200 %s failed (code %d): %s" % (what, rc, msg))
