"""ctypes binding of libqtb200.so (the C ABI declared in include/qtb200.h).

There is no CPU fallback: if the library is missing, or a call is made with a
non-CUDA tensor, this module raises.  The oracle under ``oracle/`` is test
infrastructure and is never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libqtb200.so")

# ---- enums (mirror include/qtb200.h) -------------------------------------
Q_SIGN, Q_TERNARY, Q_DOREFA, Q_XNOR_ROW, Q_LOG, Q_LIN, Q_SPLIT = range(7)
W_SIGN, W_TERNARY, W_DOREFA, W_XNOR = range(4)
CODES_NONE, CODES_I8, CODES_U8, CODES_BF16, CODES_BF16X2, CODES_F16, CODES_BF16X3 = range(7)
CODES_F16_EXACT = 6      # qt_expand_weight out_kind 6 (fp16 exact integer codes)
CODES_F4 = 7             # fp4 (e2m1) codes, two per byte: the operand of qt_gemm_f4 (tcgen05 kind::mxf4)
FMT_BF16, FMT_FP16 = 0, 1
BACKEND_AUTO, BACKEND_TCGEN05, BACKEND_SIMT = 0, 1, 2

vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float


class QtActQuant(C.Structure):
    _fields_ = [("mode", i32), ("bit_width", i32), ("fsr", i32), ("with_sign", i32),
                ("x", vp), ("rows", i64), ("cols", i64), ("ld_x", i64),
                ("y", vp), ("ld_y", i64),
                ("codes", vp), ("codes_kind", i32), ("ld_codes", i64),
                ("bits", vp), ("ld_bits", i64),
                ("row_sum", vp), ("row_scale", vp), ("overflow", vp),
                ("pre_scale", vp), ("pre_shift", vp), ("pre_channels", i64), ("pre_hw", i64), ("pre_clamp", i32),
                ("pre_lo", f32), ("pre_hi", f32), ("nhwc_c", i64), ("row_parts", i32), ("max_ctas", i32), ("ready", vp), ("ready_rows", i32)]


class QtWeightPack(C.Structure):
    _fields_ = [("mode", i32), ("bit_width", i32), ("w", vp), ("n", i64), ("k", i64), ("ld_w", i64),
                ("packed", vp), ("ld_packed", i64), ("alpha", vp), ("alpha_is_input", i32),
                ("stats", vp), ("wq", vp)]


class QtWeightExpand(C.Structure):
    _fields_ = [("mode", i32), ("bit_width", i32), ("packed", vp), ("n", i64), ("k", i64), ("ld_packed", i64),
                ("alpha", vp), ("out", vp), ("out_kind", i32), ("ld_out", i64)]


class QtIm2col(C.Structure):
    _fields_ = [("x", vp), ("nhwc", i32), ("elem_bytes", i32), ("is_unsigned", i32),
                ("B", i64), ("C", i64), ("H", i64), ("W", i64),
                ("kh", i32), ("kw", i32), ("stride_h", i32), ("stride_w", i32), ("pad_h", i32), ("pad_w", i32),
                ("dil_h", i32), ("dil_w", i32), ("groups", i32), ("group", i32),
                ("OH", i64), ("OW", i64), ("out", vp), ("ld_out", i64), ("row_sum", vp), ("split3", i32)]


class QtConvGeom(C.Structure):
    _fields_ = [("B", i64), ("C", i64), ("H", i64), ("W", i64),
                ("kh", i32), ("kw", i32), ("stride_h", i32), ("stride_w", i32), ("pad_h", i32), ("pad_w", i32),
                ("dil_h", i32), ("dil_w", i32), ("groups", i32), ("group", i32), ("OH", i64), ("OW", i64),
                ("corner_mode", i32), ("lower_w", i32), ("upper_w", i32)]


class QtPoolGeom(C.Structure):
    _fields_ = [("B", i64), ("H", i64), ("W", i64), ("C", i64),
                ("kh", i32), ("kw", i32), ("stride_h", i32), ("stride_w", i32), ("pad_h", i32), ("pad_w", i32),
                ("OH", i64), ("OW", i64)]


class QtRequant(C.Structure):
    _fields_ = [("mode", i32), ("bit_width", i32), ("codes", vp), ("codes_kind", i32), ("ld_codes", i64),
                ("clamp", i32), ("lo", f32), ("hi", f32), ("row_part", vp), ("row_sum_part", vp),
                ("row_parts", i32), ("overflow", vp), ("cover", i64)]


class QtEpilogue(C.Structure):
    _fields_ = [("bias", vp), ("row_scale", vp), ("col_scale", vp), ("row_sum", vp),
                ("scale", f32), ("acc_mul", C.c_int32), ("rs_mul", C.c_int32),
                ("out", vp), ("ldo", i64), ("out_mode", i32), ("nchw_inner", i64), ("acc_out", vp),
                ("requant", C.POINTER(QtRequant)), ("row_scale_parts", i32), ("row_scale_mul", f32),
                ("row_sum_parts", i32), ("out_clamp", i32), ("out_lo", f32), ("out_hi", f32),
                ("residual", vp), ("ld_res", i64), ("a_ready", vp), ("a_ready_rows", i32), ("a_ready_target", i32)]


# every symbol include/qtb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "qt_version": (i32, []),
    "qt_sizeof": (i32, [C.c_char_p]),
    "qt_last_error": (C.c_char_p, []),
    "qt_requant_max_parts": (i32, [i64]),
    "qt_device_caps": (i32, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "qt_quant_act": (i32, [C.POINTER(QtActQuant), vp]),
    "qt_quant_xnor_parts": (i32, [i64, i32, i32]),
    "qt_pack_weight": (i32, [C.POINTER(QtWeightPack), vp]),
    "qt_col_absmean": (i32, [vp, i64, i64, i64, vp, vp]),
    "qt_expand_weight": (i32, [C.POINTER(QtWeightExpand), vp]),
    "qt_im2col": (i32, [C.POINTER(QtIm2col), vp]),
    "qt_gemm_b1b1": (i32, [vp, i64, vp, i64, i64, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_gemm_b1t2": (i32, [vp, i64, vp, vp, i64, i64, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_gemm_i8": (i32, [vp, i32, i64, vp, i32, i64, i64, i64, i64, C.POINTER(QtEpilogue), i32, vp]),
    "qt_gemm_f4": (i32, [vp, i64, vp, i64, i64, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_conv_i8": (i32, [vp, i32, C.POINTER(QtConvGeom), vp, i32, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_patch_rowsum": (i32, [vp, i32, C.POINTER(QtConvGeom), vp, vp, vp]),
    "qt_conv_bf16": (i32, [vp, C.POINTER(QtConvGeom), vp, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_image_planes": (i32, [vp, i64, i64, i64, i64, i32, i32, i32, i64, i64, i32, i32, vp, vp]),
    "qt_image_windows": (i32, [vp, i64, i64, i64, i64, i32, i32, i32, i32, i32, i64, i64, i32, vp, vp]),
    "qt_rowsum_codes": (i32, [vp, i32, i64, i64, vp, vp]),
    "qt_pool_codes": (i32, [vp, i32, C.POINTER(QtPoolGeom), vp, vp, vp]),
    "qt_pool_quant_f32": (i32, [vp, C.POINTER(QtPoolGeom), vp, i32, i32, vp, i32, vp, vp]),
    "qt_head_f32": (i32, [vp, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]),
    "qt_gemm_f16": (i32, [vp, i64, i64, vp, i64, i64, i32, i32, C.POINTER(i32), C.POINTER(i32),
                          i64, i64, i64, C.POINTER(QtEpilogue), i32, vp]),
    "qt_gemm_f32": (i32, [vp, i64, vp, i64, i64, i64, i64, C.POINTER(QtEpilogue), vp]),
    "qt_expand_loglin": (i32, [vp, i64, i64, i64, i32, i32, vp, i64, vp]),
    "qt_transpose_split": (i32, [vp, i64, i64, i64, vp, i64, i32, vp]),
    "qt_ste_clip": (i32, [vp, vp, f32, vp, i64, vp]),
    "qt_peer_push": (i32, [vp, C.POINTER(vp), i32, i64, i32, vp]),
    "qt_set_option": (i32, [C.c_char_p, i32]),
    "qt_launch_count": (i64, [i32]),
}

_lib = None


class QtError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raise loudly if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pytorch_quantize_impls_b200: %s is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for the quantized kernels." % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().qt_last_error().decode("utf-8", "replace")
        raise QtError("%s failed (code %d): %s" % (what, rc, msg))


def set_option(name, value):
    check(lib().qt_set_option(name.encode(), int(value)), "qt_set_option")


def launch_count(reset=False):
    return int(lib().qt_launch_count(1 if reset else 0))
