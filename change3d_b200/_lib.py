"""ctypes binding of libchange3d_b200.so (include/change3d_b200.h).

The product path has NO fallback: if the library is missing or a launch returns a non-zero
status, a RuntimeError is raised.  Structures mirror the C header field for field.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchange3d_b200.so")

STATUS = {0: "ok", 1: "bad argument", 2: "CUDA error", 3: "shared-memory budget exceeded"}

# prologue modes / row maps / epilogues (c3d_common.cuh)
PRO_NONE, PRO_BN_RELU, PRO_BN_GATE_SWISH, PRO_BNBWD, PRO_ABSDIFF, PRO_MASK_POS = range(6)
MAP_DENSE, MAP_SUB2 = range(2)
EPI_STORE, EPI_RELU_ADD, EPI_SWISH_BWD, EPI_ADD2, _EPI_RESERVED4, EPI_ABSDIFF_BWD = range(6)

GEMM_W_CONSTANT = 1     # C3D_GEMM_W_CONSTANT

_fp = C.c_void_p   # device pointers travel as integers


class Operand(C.Structure):
    _fields_ = [("A", _fp), ("A2", _fp), ("bnp", _fp), ("coef", _fp), ("gate", _fp),
                ("mode", C.c_int), ("map", C.c_int), ("ld", C.c_int),
                ("OH", C.c_int), ("OW", C.c_int), ("IH", C.c_int), ("IW", C.c_int),
                ("img_stride", C.c_longlong), ("img_stride2", C.c_longlong),
                ("frames_per_sample", C.c_int), ("seg0", C.c_int), ("nseg", C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [("a", Operand), ("W", _fp),
                ("w_sr", C.c_longlong), ("w_so", C.c_longlong), ("w_cls_stride", C.c_longlong),
                ("Kred", C.c_int), ("N", C.c_int), ("Ns", C.c_int), ("M", C.c_longlong),
                ("Y", _fp), ("out_img_stride", C.c_longlong), ("epi", C.c_int), ("stats", _fp),
                ("E1", _fp), ("e1_img_stride", C.c_longlong), ("E2", _fp), ("ebnp", _fp), ("egate", _fp),
                ("bias", _fp), ("Y2", _fp), ("rows_per_sample", C.c_longlong), ("flags", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("p", Operand), ("q", Operand), ("M", C.c_longlong), ("dW", _fp),
                ("dw_sn", C.c_longlong), ("dw_sk", C.c_longlong), ("N", C.c_int), ("K", C.c_int)]


class AttnDesc(C.Structure):
    _fields_ = [("q", _fp), ("k", _fp), ("v", _fp),
                ("q_ls", C.c_longlong), ("q_bs", C.c_longlong), ("k_ls", C.c_longlong), ("k_bs", C.c_longlong),
                ("v_ls", C.c_longlong), ("v_bs", C.c_longlong),
                ("o", _fp), ("o_ls", C.c_longlong), ("o_bs", C.c_longlong), ("P", _fp), ("keep", _fp),
                ("keep_scale", C.c_float), ("scale", C.c_float),
                ("B", C.c_int), ("nh", C.c_int), ("hd", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("causal", C.c_int)]


_i, _ll, _f = C.c_int, C.c_longlong, C.c_float

# name -> argtypes; every function returns int status.  Kept in one table so the CPU test-suite can
# check that the library exports every symbol the header declares.
SIGNATURES = {
    "c3d_pw_gemm": [C.POINTER(GemmDesc), _fp],
    "c3d_pw_wgrad": [C.POINTER(WgradDesc), _fp],
    "c3d_bn_finalize": [_fp, _i, _ll, _fp, _fp, _fp, _fp, _i, _i, _f, _f, _i, _fp, _fp],
    "c3d_bn_se_finalize": [_fp, _i, _ll, _fp, _fp, _fp, _fp, _i, _i, _f, _f, _i, _fp, _fp, _fp, _fp, _i,
                           _fp, _fp, _fp, _fp, _fp],
    "c3d_bn_add_relu": [_fp, _fp, _fp, _fp, _fp, _ll, _i, _fp],
    "c3d_dw_conv_fwd": [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
    "c3d_stem_fwd": [C.POINTER(_fp), C.POINTER(_ll), C.POINTER(_ll), _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp],
    "c3d_dec_head_fwd": [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _fp],
    "c3d_relu_bwd_stats": [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _ll, _i, _fp],
    "c3d_bn_bwd_finalize": [_fp, _i, _ll, _i, _i, _fp, _fp, _fp, _fp],
    "c3d_se_bn_bwd_finalize": [_fp, _i, _ll, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i,
                               _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp],
    "c3d_dw_conv_bwd": [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
    "c3d_colsum": [_fp, _ll, _i, _fp, _fp],
    "c3d_convt_col2im": [_fp, _fp, _ll, _fp, _fp, _i, _i, _i, _i, _fp],
    "c3d_convt_im2col": [_fp, _fp, _i, _i, _i, _i, _fp],
    "c3d_stem_bwd": [C.POINTER(_fp), C.POINTER(_ll), C.POINTER(_ll), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                     _i, _i, _i, _i, _i, _fp],
    "c3d_dec_head_bwd": [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _fp],
    "c3d_adam_step": [_fp, _fp, _fp, _fp, _ll, _f, _f, _f, _f, _f, _i, _f, _f, _fp],
    "c3d_bce_dice_fwd": [_fp, _fp, _ll, _fp, _fp, _fp, _fp],
    "c3d_bce_dice_bwd": [_fp, _fp, _fp, _fp, _f, _fp, _ll, _fp],
    "c3d_ce2d_fwd": [_fp, _fp, _i, _i, _ll, _ll, _ll, _fp, _fp, _fp, _fp, _fp],
    "c3d_ce2d_bwd": [_fp, _fp, _i, _i, _ll, _ll, _ll, _fp, _fp, _f, _fp, _fp],
    "c3d_change_similarity_fwd": [_fp, _fp, _fp, _i, _i, _ll, _ll, _ll, _fp, _fp, _fp],
    "c3d_change_similarity_bwd": [_fp, _fp, _fp, _i, _i, _ll, _ll, _ll, _fp, _f, _fp, _fp, _fp],
    "c3d_confusion_matrix": [_fp, _i, _fp, _ll, _i, _fp, _fp],
    "c3d_attention_fwd": [C.POINTER(AttnDesc), _fp],
    "c3d_attention_bwd": [C.POINTER(AttnDesc), _fp, _ll, _ll, _fp, _fp, _fp, _fp],
    "c3d_augment_pairs": [_fp, _i, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _fp, _fp, _fp, _fp],
}

LOSS_WS_BYTES = 128      # C3D_LOSS_WS_BYTES
LOSS_OUT_FLOATS = 8      # C3D_LOSS_OUT_FLOATS

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m change3d_b200.build` "
            "(or __graft_entry__.build()); change3d_b200 has no CPU/eager fallback")
    lib = C.CDLL(LIB_PATH)
    lib.c3d_version.restype = C.c_int
    lib.c3d_version.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


LAUNCHES = [0]      # kernels launched through the C ABI by this process (bench.py reports it)


def check(status: int, what: str) -> None:
    LAUNCHES[0] += 1
    if status != 0:
        raise RuntimeError(f"change3d_b200: {what} failed: {STATUS.get(status, status)}")
