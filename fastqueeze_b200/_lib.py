"""ctypes binding of the C ABI in include/seqarc_b200.h (fastqueeze_b200/libseqarc_b200.so).

The library is hand-written sm_100a CUDA; there is no CPU implementation behind this module.
Loading fails loudly when the shared object is missing (run tools/build.sh or
__graft_entry__.build()), and sq_create fails with SQ_ERR_NO_DEVICE without a GPU.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libseqarc_b200.so")

SQ_OK = 0
STATUS = {0: "SQ_OK", -1: "SQ_ERR_NO_DEVICE", -2: "SQ_ERR_CUDA", -3: "SQ_ERR_ARG", -4: "SQ_ERR_CAPACITY",
          -5: "SQ_ERR_INPUT", -6: "SQ_ERR_FORMAT", -7: "SQ_ERR_NOMEM", -8: "SQ_ERR_MD5_NAME",
          -9: "SQ_ERR_MD5_SEQ", -10: "SQ_ERR_MD5_QUAL"}

# every symbol include/seqarc_b200.h declares
EXPORTS = ["sq_create", "sq_destroy", "sq_last_error", "sq_load_model", "sq_model_size", "sq_train_model", "sq_encode_blocks",
           "sq_encode_blocks_dev", "sq_decode_blocks", "sq_decode_blocks_dev", "sq_rblock", "sq_last_timings",
           "sq_last_launches", "sq_last_wave_blocks", "sq_debug_fetch", "sq_device_count", "sq_version",
           "sq_side_encode", "sq_side_decode", "sq_set_reference", "sq_align_rebuild"]


class SqParams(C.Structure):
    _fields_ = [("device", C.c_int32), ("usemodel", C.c_int32), ("lossy", C.c_int32), ("lossy_r", C.c_double),
                ("md5", C.c_int32), ("slevel", C.c_int32), ("qlevel", C.c_int32), ("id_rule", C.c_int32),
                ("workspace_bytes", C.c_uint64), ("stream", C.c_void_p), ("ref_mode", C.c_int32)]


class SqBlockIn(C.Structure):
    _fields_ = [("nreads", C.c_uint32), ("longflag", C.c_int32), ("ids", C.c_void_p), ("seq", C.c_void_p),
                ("qual", C.c_void_p), ("idlen", C.c_void_p), ("seqlen", C.c_void_p),
                ("idbytes", C.c_uint64), ("seqbytes", C.c_uint64)]


class SqBlockOut(C.Structure):
    _fields_ = [("data", C.c_void_p), ("cap", C.c_uint64), ("len", C.c_uint64), ("status", C.c_int32)]


class SqBlockDec(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("idcap", C.c_uint64), ("seq", C.c_void_p), ("qual", C.c_void_p),
                ("seqcap", C.c_uint64), ("idlen", C.c_void_p), ("seqlen", C.c_void_p), ("readcap", C.c_uint32),
                ("nreads", C.c_uint32), ("idbytes", C.c_uint64), ("seqbytes", C.c_uint64), ("status", C.c_int32),
                ("longflag", C.c_int32)]


class SqSideStream(C.Structure):
    _fields_ = [("model", C.c_int32), ("nsym", C.c_uint32), ("sym", C.c_void_p), ("bytes", C.c_void_p),
                ("cap", C.c_uint64), ("len", C.c_uint64), ("status", C.c_int32)]


class SqAlignBlock(C.Structure):
    _fields_ = [("nreads", C.c_uint32), ("pe", C.c_int32), ("seqlen", C.c_void_p), ("order", C.c_void_p), ("pos", C.c_void_p),
                ("mis", C.c_void_p), ("rev", C.c_void_p), ("cigal", C.c_void_p), ("cigav", C.c_void_p), ("perel", C.c_void_p),
                ("npos", C.c_uint32), ("nalign", C.c_uint32), ("ncigal", C.c_uint32), ("ncigav", C.c_uint32),
                ("nperel", C.c_uint32), ("insbits", C.c_uint32), ("seq", C.c_void_p), ("seqbytes", C.c_uint64),
                ("status", C.c_int32)]


class SeqArcB200Error(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d)%s" % (STATUS.get(code, "?"), code, (": " + msg) if msg else ""))


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with tools/build.sh (nvcc, sm_100a). "
                          "fastqueeze_b200 has no CPU implementation." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.sq_create.restype = C.c_int
    L.sq_create.argtypes = [C.POINTER(SqParams), C.POINTER(C.c_void_p)]
    L.sq_destroy.restype = None
    L.sq_destroy.argtypes = [C.c_void_p]
    L.sq_last_error.restype = C.c_char_p
    L.sq_last_error.argtypes = [C.c_void_p]
    L.sq_load_model.restype = C.c_int
    L.sq_load_model.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.sq_model_size.restype = C.c_uint64
    L.sq_model_size.argtypes = [C.c_void_p]
    for n in ("sq_encode_blocks", "sq_encode_blocks_dev"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.POINTER(SqBlockIn), C.c_int, C.POINTER(SqBlockOut)]
    for n in ("sq_decode_blocks", "sq_decode_blocks_dev"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int, C.POINTER(SqBlockDec)]
    L.sq_train_model.restype = C.c_int
    L.sq_train_model.argtypes = [C.c_void_p, C.POINTER(SqBlockIn), C.c_int, C.c_void_p, C.c_uint64]
    L.sq_rblock.restype = C.c_int
    L.sq_rblock.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double]
    L.sq_last_timings.restype = C.c_int
    L.sq_last_timings.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.sq_last_launches.restype = C.c_uint64
    L.sq_last_launches.argtypes = [C.c_void_p]
    L.sq_last_wave_blocks.restype = C.c_int
    L.sq_last_wave_blocks.argtypes = [C.c_void_p]
    L.sq_debug_fetch.restype = C.c_int
    L.sq_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    for n in ("sq_side_encode", "sq_side_decode"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.POINTER(SqSideStream), C.c_int]
    L.sq_set_reference.restype = C.c_int
    L.sq_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.sq_align_rebuild.restype = C.c_int
    L.sq_align_rebuild.argtypes = [C.c_void_p, C.POINTER(SqAlignBlock), C.c_int]
    L.sq_device_count.restype = C.c_int
    L.sq_device_count.argtypes = []
    L.sq_version.restype = C.c_char_p
    L.sq_version.argtypes = []
    _lib = L
    return L
