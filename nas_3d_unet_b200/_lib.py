"""ctypes binding of libnas3d_b200.so (the C-ABI declared in include/nas3d_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NAS3D_LIB: alternative build of the SAME library (e.g. the -DNAS3D_NO_FFMA2 variant made by
# `python -m nas_3d_unet_b200.build --variant noffma2`) for A/B measurements; never a fallback.
LIB_PATH = os.environ.get("NAS3D_LIB") or os.path.join(_HERE, "lib", "libnas3d_b200.so")

c_float_p = C.c_void_p   # device pointers travel as integers
c_ll = C.c_longlong
c_int = C.c_int
c_vp = C.c_void_p


class ConvDesc(C.Structure):
    """mirror of nas3d_conv_desc"""
    _fields_ = [(n, C.c_int) for n in (
        "N", "Db", "Hb", "Wb", "Cb", "ld_big", "Ds", "Hs", "Ws", "Cs", "ld_small",
        "k", "stride", "dil", "pad", "depthwise")]


class Nas3dError(RuntimeError):
    pass


# name -> argtypes (restype is int unless listed in _RESTYPES)
_PP = C.POINTER(C.c_void_p)   # host array of device pointers
_PI = C.POINTER(C.c_int)      # host array of ints
_SIGNATURES = {
    "nas3d_version": [],
    "nas3d_last_error": [],
    "nas3d_launch_count": [],
    "nas3d_launch_count_of": [C.c_char_p],
    "nas3d_launch_labels": [C.c_char_p, c_int],
    "nas3d_probe_fma": [c_vp, c_ll, c_int, c_vp],
    "nas3d_set_option": [C.c_char_p, c_int],
    "nas3d_get_option": [C.c_char_p],
    "nas3d_ncdhw_to_ndhwc": [c_vp, c_vp, c_int, c_int, c_ll, c_int, c_vp],
    "nas3d_conv_small_from_big": [C.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp,
                                  c_int, c_vp, c_vp],
    "nas3d_conv_big_from_small": [C.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp,
                                  c_int, c_vp, c_vp],
    "nas3d_conv_wgrad": [C.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp],
    "nas3d_conv_wgrad_ws": [C.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp],
    "nas3d_conv_wgrad_workspace_floats": [C.POINTER(ConvDesc), c_int],
    "nas3d_conv1x1_cat_fwd": [C.POINTER(ConvDesc), c_int, _PP, _PI, c_vp, c_vp, c_vp, c_int, c_int,
                              c_vp, c_vp, c_vp],
    "nas3d_conv1x1_cat_dgrad": [C.POINTER(ConvDesc), c_int, _PP, _PI, _PI, c_vp, c_vp, _PP, _PI, c_vp,
                                c_vp],
    "nas3d_conv1x1_cat_wgrad": [C.POINTER(ConvDesc), c_int, _PP, _PI, c_vp, c_vp, c_int, c_vp, c_vp,
                                c_vp],
    "nas3d_conv1x1_bwd_fused_supported": [C.POINTER(ConvDesc), c_int],
    "nas3d_conv1x1_bwd_fused": [C.POINTER(ConvDesc), c_int, _PP, _PI, _PP, _PI, _PI, c_vp, c_vp, c_vp,
                                c_vp, c_int, c_vp, c_vp, c_vp],
    "nas3d_umma_packed_floats": [C.POINTER(ConvDesc), c_int],
    "nas3d_umma_pack_weights": [C.POINTER(ConvDesc), c_vp, c_int, c_vp, c_vp],
    "nas3d_umma_pack_mode": [C.POINTER(ConvDesc), c_int],
    "nas3d_umma_pack_weights_batch": [c_int, _PI, _PI, _PP, _PP, c_vp],
    "nas3d_umma_conv": [C.POINTER(ConvDesc), c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp],
    "nas3d_moments_nc": [c_vp, c_int, c_ll, c_int, c_int, c_vp, c_vp],
    "nas3d_gn_coef": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_ll, C.c_float, c_vp, c_vp, c_vp, c_vp],
    "nas3d_gn_coef_batch": [c_int, _PP, _PP, _PP, c_int, c_int, c_int, c_ll, C.c_float, _PP, _PP, _PP,
                            c_vp],
    "nas3d_gn_bwd_coef_batch": [c_int, _PP, _PP, _PP, _PP, _PP, _PP, c_int, c_int, c_int, c_ll, _PP,
                                _PP, _PP, _PP, _PP, _PP, _PP, _PP, c_vp],
    "nas3d_stage_patches": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, _PI, c_int, c_vp, c_int, c_vp,
                            c_vp],
    "nas3d_stage_patches_seg8": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, _PI, c_int, c_vp, c_int, c_vp,
                            c_vp],
    "nas3d_adam_chunk_floats": [],
    "nas3d_adam_flat_step": [c_vp, c_vp, c_vp, _PP, c_int, _PI, c_vp, c_vp, c_vp],
    "nas3d_se_excite": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_vp],
    "nas3d_affine_sum_fwd": [c_int, _PP, _PI, _PP, _PP, _PP, _PI, c_vp, c_int, c_int, c_ll, c_int,
                             c_vp],
    "nas3d_affine_sum_fwd_gn": [c_int, _PP, _PI, _PP, _PP, _PP, _PI, _PP, _PP, _PP, _PP, c_int,
                                C.c_float, c_vp, c_int, c_int, c_ll, c_int, c_vp],
    "nas3d_affine_sum_bwd_apply_gn": [c_int, _PP, _PI, _PP, _PP, _PI, _PP, _PP, _PP, _PP, _PP, _PI,
                                      _PI, c_vp, c_int, _PP, _PP, _PP, _PP, _PP, _PP, _PP, _PP,
                                      c_int, c_int, c_ll, c_int, c_vp],
    "nas3d_affine_sum_bwd_reduce": [c_int, _PP, _PI, _PP, _PP, _PI, c_vp, c_int, _PP, c_int, c_ll,
                                    c_int, c_vp],
    "nas3d_gn_bwd_coef": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_ll, c_vp, c_vp,
                          c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "nas3d_se_bwd_coef": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp, c_vp,
                          c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "nas3d_plain_bwd_coef": [c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "nas3d_affine_sum_bwd_apply": [c_int, _PP, _PI, _PP, _PP, _PI, _PP, _PP, _PP, _PP, _PP, _PI,
                                   _PI, c_vp, c_int, c_int, c_ll, c_int, c_vp],
    "nas3d_pool2_fwd": [c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    "nas3d_pool2_bwd": [c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int,
                        c_int, c_int, c_vp],
    "nas3d_dice_fwd": [c_vp, c_ll, c_ll, c_ll, c_vp, c_int, c_ll, c_ll, c_ll, c_int, c_int, c_ll,
                       C.c_float, c_vp, c_vp, c_vp],
    "nas3d_dice_bwd": [c_vp, c_vp, c_vp, c_int, c_ll, c_ll, c_ll, c_vp, c_ll, c_ll, c_ll, c_int, c_int,
                       c_ll, C.c_float, c_vp],
    "nas3d_extract_patches": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp,
                              c_int, c_vp],
    "nas3d_patch_nonzero": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "nas3d_stitch_labels": [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_int, c_int, C.c_float, c_int, c_vp, c_vp, c_vp, c_vp],
    "nas3d_seg_to_masks": [c_vp, c_int, c_ll, c_int, c_vp, c_vp],
    "nas3d_sigmoid_bwd": [c_vp, c_vp, c_vp, c_ll, c_vp],
    "nas3d_add_inplace": [c_vp, c_vp, c_ll, c_vp],
}
_RESTYPES = {
    "nas3d_last_error": C.c_char_p,
    "nas3d_launch_count": C.c_ulonglong,
    "nas3d_launch_count_of": C.c_ulonglong,
    "nas3d_probe_fma": C.c_longlong,
    "nas3d_conv_wgrad_workspace_floats": C.c_longlong,
    "nas3d_umma_packed_floats": C.c_longlong,
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """dlopen the library once; raises Nas3dError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Nas3dError(
            "%s not found: build it with `python -m nas_3d_unet_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().nas3d_last_error()
        raise Nas3dError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


class option:
    """context manager: run a block with one kernel-selection option of the library changed
    (include/nas3d_b200.h: nas3d_set_option); parity tests A/B kernel variants with it"""

    def __init__(self, name, value):
        self.name, self.value = name.encode(), int(value)

    def __enter__(self):
        lib = load()
        self.prev = lib.nas3d_get_option(self.name)
        if self.prev < 0:
            check(self.prev, "get_option(%s)" % self.name.decode())
        check(lib.nas3d_set_option(self.name, self.value), "set_option(%s)" % self.name.decode())
        return self

    def __exit__(self, *exc):
        check(load().nas3d_set_option(self.name, self.prev), "set_option")
        return False


def launch_count():
    return int(load().nas3d_launch_count())


def launch_counts():
    """{kernel-variant label: launches so far} (nas3d_launch_labels / nas3d_launch_count_of)"""
    lib = load()
    n = lib.nas3d_launch_labels(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.nas3d_launch_labels(buf, n + 1)
    out = {}
    for name in buf.value.decode().split():
        out[name] = int(lib.nas3d_launch_count_of(name.encode()))
    return out


def ptr_array(ptrs):
    """host array of device pointers from a list of ints/None"""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])
