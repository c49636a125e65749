"""ctypes binding of the CPU oracle (oracle/datok_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by datok_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdatok_oracle.so")

TOKENS, SENTENCES, TOKEN_POS, SENTENCE_POS, NEWLINE_AFTER_EOT = 1, 2, 4, 8, 16
SIMPLE = TOKENS | SENTENCES
WRITER_USED = 256

STATUS = {0: "ok", 1: "panic: buffer overflow (>1024 runes without a token boundary)",
          2: "panic: SentenceEnd before any token (pos flags)",
          3: "panic: TextEnd on a token-less text (TOKEN_POS)",
          4: "panic: TextEnd on a sentence-less text (SENTENCE_POS)",
          5: "panic: token slice out of range", 6: "panic: empty token buffer",
          7: "endless epsilon loop"}


def build(force=False):
    src = os.path.join(HERE, "datok_oracle.c")
    hdr = os.path.join(HERE, "datok_oracle.h")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "-B", "libdatok_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


class _Carry(C.Structure):
    _fields_ = [("state", C.c_uint32), ("ok", C.c_int32), ("sentence_end", C.c_int32),
                ("text_end", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("text", C.POINTER(C.c_uint8)), ("text_len", C.c_size_t),
        ("n_tokens", C.c_size_t),
        ("tok_byte_start", C.POINTER(C.c_uint32)), ("tok_byte_end", C.POINTER(C.c_uint32)),
        ("tok_buf_start", C.POINTER(C.c_uint32)), ("tok_offset", C.POINTER(C.c_int32)),
        ("n_tok_pos", C.c_size_t), ("tok_pos", C.POINTER(C.c_int32)),
        ("n_sent_events", C.c_size_t), ("sent_tok_idx", C.POINTER(C.c_uint64)),
        ("n_sent_pos", C.c_size_t), ("sent_pos", C.POINTER(C.c_int32)),
        ("n_texts", C.c_size_t),
        ("text_tok_end", C.POINTER(C.c_uint64)), ("text_sent_end", C.POINTER(C.c_uint64)),
        ("text_sentpos_end", C.POINTER(C.c_uint64)), ("text_byte_end", C.POINTER(C.c_uint32)),
        ("carry_out", _Carry),
        ("n_runes", C.c_uint64), ("n_iterations", C.c_uint64), ("n_backtracks", C.c_uint64),
        ("n_backtrack_runes", C.c_uint64), ("n_hardfail", C.c_uint64),
        ("max_window", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.ora_load.restype = C.c_void_p
        L.ora_load.argtypes = [C.c_char_p]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_load_foma.restype = C.c_void_p
        L.ora_load_foma.argtypes = [C.c_char_p]
        L.ora_write_matrix.restype = C.c_void_p
        L.ora_write_matrix.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        for f in ("ora_epsilon", "ora_unknown", "ora_identity", "ora_state_count", "ora_sigma_count"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.ora_array.restype = C.POINTER(C.c_uint32)
        L.ora_array.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.ora_sigma_ascii.restype = C.POINTER(C.c_int32)
        L.ora_sigma_ascii.argtypes = [C.c_void_p]
        L.ora_sigma_lookup.restype = C.c_int
        L.ora_sigma_lookup.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int)]
        L.ora_transduce.restype = C.POINTER(_Result)
        L.ora_transduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(_Carry)]
        L.ora_result_free.argtypes = [C.POINTER(_Result)]
        L.ora_transduce_docs_mt.restype = C.c_uint64
        L.ora_transduce_docs_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_uint64)]
        L.ora_token_writer_replay.restype = C.POINTER(C.c_uint8)
        L.ora_token_writer_replay.argtypes = [C.c_uint32, C.POINTER(C.c_int32), C.c_size_t,
                                              C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.ora_free_bytes.argtypes = [C.POINTER(C.c_uint8)]
        L.ora_state_histogram.restype = C.c_int
        L.ora_state_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ora_decode_rune.restype = C.c_int32
        L.ora_decode_rune.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class OracleResult:
    """Plain-Python copy of ora_result."""

    def __init__(self, r):
        self.status = r.status
        # (a numpy copy: ctypes.string_at takes a C int length, and the text of a full-size corpus exceeds 2 GiB)
        self.text_np = np.ctypeslib.as_array(r.text, shape=(r.text_len,)).copy() if r.text_len else np.zeros(0, np.uint8)
        self._text = None
        nt = r.n_tokens
        self.n_tokens = nt
        self.tok_byte_start = _arr(r.tok_byte_start, nt, np.uint32)
        self.tok_byte_end = _arr(r.tok_byte_end, nt, np.uint32)
        self.tok_buf_start = _arr(r.tok_buf_start, nt, np.uint32)
        self.tok_offset = _arr(r.tok_offset, nt, np.int32)
        # tok_pos only exists when a position flag is set
        self.tok_pos = _arr(r.tok_pos, r.n_tok_pos, np.int32)
        self.n_sent_events = r.n_sent_events
        self.sent_tok_idx = _arr(r.sent_tok_idx, r.n_sent_events, np.uint64)
        self.sent_pos = _arr(r.sent_pos, r.n_sent_pos, np.int32)
        self.n_texts = r.n_texts
        self.text_tok_end = _arr(r.text_tok_end, r.n_texts, np.uint64)
        self.text_sent_end = _arr(r.text_sent_end, r.n_texts, np.uint64)
        self.text_sentpos_end = _arr(r.text_sentpos_end, r.n_texts, np.uint64)
        self.text_byte_end = _arr(r.text_byte_end, r.n_texts, np.uint32)
        self.carry_out = dict(state=r.carry_out.state, ok=r.carry_out.ok,
                              sentence_end=r.carry_out.sentence_end, text_end=r.carry_out.text_end)
        self.stats = dict(runes=r.n_runes, iterations=r.n_iterations, backtracks=r.n_backtracks,
                          backtrack_runes=r.n_backtrack_runes, hardfail=r.n_hardfail,
                          max_window=r.max_window)


    @property
    def text(self):
        if self._text is None:
            self._text = self.text_np.tobytes()
        return self._text


class OracleModel:
    def __init__(self, path, foma=None):
        """foma=True: `path` is a foma file, compiled in memory like LoadFomaFile(path).ToMatrix();
        default: by extension (.fst)"""
        if foma is None:
            foma = str(path).endswith(".fst")
        self._h = (lib().ora_load_foma if foma else lib().ora_load)(os.fsencode(path))
        if not self._h:
            raise ValueError(f"oracle: cannot load {path}")
        L = lib()
        self.epsilon = L.ora_epsilon(self._h)
        self.unknown = L.ora_unknown(self._h)
        self.identity = L.ora_identity(self._h)
        self.state_count = L.ora_state_count(self._h)
        self.sigma_count = L.ora_sigma_count(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ora_free(self._h)
            self._h = None

    def write_matrix(self):
        """WriteTo (matrix.go:126-210): the uncompressed MATOK image"""
        n = C.c_size_t()
        p = lib().ora_write_matrix(self._h, C.byref(n))
        try:
            return C.string_at(p, n.value)
        finally:
            lib().ora_free_bytes(C.cast(p, C.POINTER(C.c_uint8)))

    def array(self):
        n = C.c_size_t()
        p = lib().ora_array(self._h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def sigma_ascii(self):
        return np.ctypeslib.as_array(lib().ora_sigma_ascii(self._h), shape=(256,)).copy()

    def sigma_lookup(self, rune):
        ok = C.c_int(-1)
        a = lib().ora_sigma_lookup(self._h, rune, C.byref(ok))
        return a, ok.value

    def transduce(self, data, flags=SIMPLE, carry_in=None):
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
        cin = None
        if carry_in is not None:
            cin = _Carry(**carry_in)
        rp = lib().ora_transduce(self._h, buf, len(data), flags, C.byref(cin) if cin else None)
        try:
            return OracleResult(rp.contents)
        finally:
            lib().ora_result_free(rp)

    def transduce_np(self, arr, flags=SIMPLE, carry_in=None):
        """same, zero-copy over a contiguous uint8 numpy array"""
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        cin = _Carry(**carry_in) if carry_in is not None else None
        rp = lib().ora_transduce(self._h, arr.ctypes.data, arr.size, flags, C.byref(cin) if cin else None)
        try:
            return OracleResult(rp.contents)
        finally:
            lib().ora_result_free(rp)

    def state_histogram(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        hist = np.zeros(self.state_count + 1, dtype=np.uint64)
        lib().ora_state_histogram(self._h, arr.ctypes.data, arr.size, hist.ctypes.data)
        return hist

    def transduce_docs_mt(self, arr, flags, nthreads):
        """CPU baseline: one worker per EOT-delimited document. Returns dict of counts."""
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        ob, ns, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        nt = lib().ora_transduce_docs_mt(self._h, arr.ctypes.data, arr.size, flags, nthreads,
                                         C.byref(ob), C.byref(ns), C.byref(nd))
        return dict(tokens=nt, sentences=ns.value, docs=nd.value, out_bytes=ob.value)


def token_writer_replay(flags, ops):
    a = (C.c_int32 * len(ops))(*ops)
    n, st = C.c_size_t(), C.c_int()
    p = lib().ora_token_writer_replay(flags, a, len(ops), C.byref(n), C.byref(st))
    out = bytes(C.string_at(p, n.value))
    lib().ora_free_bytes(p)
    return out, st.value


def decode_rune(b):
    w = C.c_int()
    buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")
    r = lib().ora_decode_rune(buf, len(b), C.byref(w))
    return r, w.value
