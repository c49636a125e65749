"""ctypes binding of libdatok_b200.so (include/datok_b200.h).

The library holds the CUDA kernels and the C ABI.  There is no Python or CPU
implementation of the transduction: if the shared object is missing or no B200
is visible, loading / datok_load fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DATOK_B200_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("DATOK_B200_LIB") or os.path.join(HERE, "libdatok_b200.so")

TOKENS, SENTENCES, TOKEN_POS, SENTENCE_POS, NEWLINE_AFTER_EOT = 1, 2, 4, 8, 16
SIMPLE = TOKENS | SENTENCES
WRITER_USED = 256
NOT_FINAL = 512
COMPACT = 1024
COMPACT8 = 2048
FORMAT = 4096

OK = 0
ERR_BUFFER_OVERFLOW, ERR_SENT_NO_TOKEN, ERR_TEXT_NO_TOKEN, ERR_TEXT_NO_SENT, ERR_DEGENERATE = 1, 2, 3, 4, 5
ERR_IO, ERR_FORMAT, ERR_UNSUPPORTED_MODEL, ERR_NO_DEVICE, ERR_CUDA, ERR_TOO_LARGE, ERR_INVALID_ARG = 16, 17, 18, 19, 20, 21, 22
ERR_NOT_AT_BOUNDARY = 23
ERR_COMPACT_RANGE = 24


class Carry(C.Structure):
    _fields_ = [("state", C.c_uint32), ("sentence_end", C.c_uint32), ("text_end", C.c_uint32),
                ("reserved", C.c_uint32)]


class View(C.Structure):
    _fields_ = [
        ("n_tokens", C.c_uint64), ("n_sentences", C.c_uint64), ("n_texts", C.c_uint64),
        ("n_sent_pos", C.c_uint64), ("n_runes", C.c_uint64),
        ("tok_bytes", C.POINTER(C.c_uint32)), ("tok_pos", C.POINTER(C.c_int32)),
        ("sent_pos", C.POINTER(C.c_int32)), ("sent_tok", C.POINTER(C.c_uint32)),
        ("text_tok_end", C.POINTER(C.c_uint32)), ("text_sent_end", C.POINTER(C.c_uint32)),
        ("text_sentpos_end", C.POINTER(C.c_uint32)), ("text_byte_end", C.POINTER(C.c_uint32)),
        ("carry_out", Carry),
        ("has_invalid_utf8", C.c_uint32),
        ("ms_h2d", C.c_float), ("ms_kernels", C.c_float), ("ms_d2h", C.c_float),
        ("tok_delta", C.POINTER(C.c_uint16)),
        ("tok_delta8", C.POINTER(C.c_uint8)), ("tok_esc", C.POINTER(C.c_uint32)), ("n_esc", C.c_uint64),
        ("text", C.POINTER(C.c_uint8)), ("text_len", C.c_uint64),
    ]


TOKEN_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_size_t, C.c_int32)
EVENT_CB = C.CFUNCTYPE(None, C.c_void_p)


class Callbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("token", TOKEN_CB), ("sentence_end", EVENT_CB), ("text_end", EVENT_CB)]


# every symbol include/datok_b200.h declares
EXPORTS = ["datok_load", "datok_load_image", "datok_load_foma", "datok_compile_foma", "datok_save", "datok_write_image", "datok_free", "datok_type", "datok_model_type", "datok_model_info", "datok_transduce",
           "datok_transduce_device", "datok_result_view", "datok_result_free", "datok_expand", "datok_format", "datok_replay",
           "datok_last_kernel_times", "datok_last_launch_count", "datok_last_stats", "datok_measure_gather_bound", "datok_host_alloc", "datok_host_free",
           "datok_last_error", "datok_strerror", "datok_stream_open", "datok_stream_push", "datok_stream_finish",
           "datok_stream_bytes_done", "datok_stream_close", "datok_plan_shards", "datok_transduce_sharded", "datok_sharded_last_error",
           "datok_sharded_last_info"]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m datok_b200.build` "
                           "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    L.datok_load.restype = C.c_void_p
    L.datok_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    L.datok_load_image.restype = C.c_void_p
    L.datok_load_image.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    L.datok_free.argtypes = [C.c_void_p]
    L.datok_load_foma.restype = C.c_void_p
    L.datok_load_foma.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    L.datok_compile_foma.restype = C.c_int
    L.datok_compile_foma.argtypes = [C.c_char_p, C.c_char_p]
    L.datok_save.restype = C.c_int
    L.datok_save.argtypes = [C.c_void_p, C.c_char_p]
    L.datok_write_image.restype = C.c_size_t
    L.datok_write_image.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.datok_type.restype = C.c_char_p
    L.datok_model_type.restype = C.c_char_p
    L.datok_model_type.argtypes = [C.c_void_p]
    L.datok_model_info.restype = C.c_int
    L.datok_model_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 6
    for f in (L.datok_transduce, L.datok_transduce_device):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(Carry), C.POINTER(C.c_void_p)]
    L.datok_result_view.restype = C.POINTER(View)
    L.datok_result_view.argtypes = [C.c_void_p]
    L.datok_result_free.argtypes = [C.c_void_p]
    L.datok_expand.restype = C.c_int
    L.datok_expand.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.datok_format.restype = C.c_size_t
    L.datok_format.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t]
    L.datok_replay.restype = C.c_int
    L.datok_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Callbacks)]
    L.datok_last_kernel_times.restype = C.c_int
    L.datok_last_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.datok_last_launch_count.restype = C.c_int
    L.datok_last_launch_count.argtypes = [C.c_void_p]
    L.datok_last_stats.restype = C.c_int
    L.datok_last_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4
    L.datok_measure_gather_bound.restype = C.c_int
    L.datok_measure_gather_bound.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.datok_host_alloc.restype = C.c_void_p
    L.datok_host_alloc.argtypes = [C.c_size_t]
    L.datok_host_free.argtypes = [C.c_void_p]
    L.datok_stream_open.restype = C.c_void_p
    L.datok_stream_open.argtypes = [C.c_void_p, C.c_uint32]
    for f in (L.datok_stream_push,):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.datok_stream_finish.restype = C.c_int
    L.datok_stream_finish.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.datok_stream_bytes_done.restype = C.c_uint64
    L.datok_stream_bytes_done.argtypes = [C.c_void_p]
    L.datok_stream_close.argtypes = [C.c_void_p]
    L.datok_plan_shards.restype = C.c_int
    L.datok_plan_shards.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]
    L.datok_transduce_sharded.restype = C.c_int
    L.datok_transduce_sharded.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_size_t, C.c_uint32,
                                          C.POINTER(Carry), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.datok_sharded_last_error.restype = C.c_char_p
    L.datok_sharded_last_info.restype = C.c_int
    L.datok_sharded_last_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.datok_last_error.restype = C.c_char_p
    L.datok_strerror.restype = C.c_char_p
    L.datok_strerror.argtypes = [C.c_int]
    _lib = L
    return L
