"""ctypes binding of tests/emul (CPU emulation of the kernel bodies; test infrastructure)
and the comparison of a structured result against the oracle's."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "emul", "libdatok_emul.so")
SRCS = [os.path.join(HERE, "emul", "emul.cpp"), os.path.join(ROOT, "datok_b200", "csrc", "model.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "datok_b200", "csrc", f) for f in
               ("walk_core.cuh", "chunk_core.cuh", "compact_core.cuh", "fast_core.cuh", "format_core.cuh", "model.hpp")]


def build():
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(p) for p in DEPS):
        return LIB
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-o", LIB] + SRCS + ["-lz"])
    return LIB


class _Res(C.Structure):
    _fields_ = [("status", C.c_int),
                ("n_tokens", C.c_uint64), ("n_sentences", C.c_uint64), ("n_texts", C.c_uint64),
                ("n_sent_pos", C.c_uint64), ("n_runes", C.c_uint64),
                ("tok_bytes", C.POINTER(C.c_uint32)), ("tok_pos", C.POINTER(C.c_int32)),
                ("sent_pos", C.POINTER(C.c_int32)), ("sent_tok", C.POINTER(C.c_uint32)),
                ("text_tok_end", C.POINTER(C.c_uint32)), ("text_sent_end", C.POINTER(C.c_uint32)),
                ("text_sentpos_end", C.POINTER(C.c_uint32)), ("text_byte_end", C.POINTER(C.c_uint32)),
                ("carry_state", C.c_uint32), ("has_invalid", C.c_uint32),
                ("rounds", C.c_uint32), ("n_rewalks", C.c_uint32), ("n_stitch_mismatch", C.c_uint32),
                ("tok_delta", C.POINTER(C.c_uint16)), ("tok_delta8", C.POINTER(C.c_uint8)),
                ("esc", C.POINTER(C.c_uint32)), ("n_esc", C.c_uint32),
                ("eot_rewind", C.c_uint32), ("text", C.POINTER(C.c_uint8)), ("text_len", C.c_uint64),
                ("delta_range", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emul_load.restype = C.c_void_p
        L.emul_load.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.emul_free.argtypes = [C.c_void_p]
        L.emul_n_classes.restype = C.c_uint32
        L.emul_n_classes.argtypes = [C.c_void_p]
        L.emul_transduce.restype = C.POINTER(_Res)
        L.emul_transduce.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emul_result_free.argtypes = [C.POINTER(_Res)]
        L.emul_calibrate.restype = C.c_int
        L.emul_calibrate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.emul_hot_cols.restype = C.c_uint32
        L.emul_hot_cols.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _arr(p, n, dt):
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)


class Structured:
    """flat result arrays in the layout of include/datok_b200.h's datok_view"""
    pass


class EmulModel:
    def __init__(self, path):
        err = C.c_int()
        self._h = lib().emul_load(os.fsencode(path), C.byref(err))
        if not self._h:
            raise ValueError(f"emul: cannot load {path}: {err.value}")
        self.n_classes = lib().emul_n_classes(self._h)

    def calibrate(self, data, force_cols=0):
        """class ids by frequency in `data`, compact rows with the frequent classes only (the product's calibration)"""
        a = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        rc = lib().emul_calibrate(self._h, a.ctypes.data if a.size else None, a.size, force_cols)
        assert rc == 0, rc
        return lib().emul_hot_cols(self._h)

    def transduce(self, data, flags, chunk=64, order=0, carry_state=0, sentence_end=0, text_end=0, mode=0):
        """mode 0: exact walker only; mode n > 0: fused fast path with n hot table rows"""
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        a = np.ascontiguousarray(a)
        ptr = a.ctypes.data if a.size else None
        rp = lib().emul_transduce(self._h, ptr, a.size, flags, chunk, carry_state, sentence_end, text_end, order, mode)
        r = rp.contents
        s = Structured()
        s.status = r.status
        s.stats = dict(rounds=r.rounds, rewalks=r.n_rewalks, mismatches=r.n_stitch_mismatch)
        if r.status == 0:
            s.n_tokens, s.n_sentences, s.n_texts = r.n_tokens, r.n_sentences, r.n_texts
            s.n_sent_pos, s.n_runes = r.n_sent_pos, r.n_runes
            s.tok_bytes = _arr(r.tok_bytes, 2 * r.n_tokens, np.uint32)
            s.tok_pos = _arr(r.tok_pos, 2 * r.n_tokens, np.int32)
            s.sent_pos = _arr(r.sent_pos, r.n_sent_pos, np.int32)
            s.sent_tok = _arr(r.sent_tok, r.n_sentences, np.uint32)
            s.text_tok_end = _arr(r.text_tok_end, r.n_texts, np.uint32)
            s.text_sent_end = _arr(r.text_sent_end, r.n_texts, np.uint32)
            s.text_sentpos_end = _arr(r.text_sentpos_end, r.n_texts, np.uint32)
            s.text_byte_end = _arr(r.text_byte_end, r.n_texts, np.uint32)
            s.carry_state = r.carry_state
            s.has_invalid = r.has_invalid
            # (a double-array model has no delta-coded forms: its cursors do not restart at a text)
            # (... and a delta that does not fit 16 bits -- DATOK_ERR_COMPACT_RANGE for a DATOK_COMPACT call -- leaves none)
            have_delta = r.eot_rewind and not r.delta_range
            s.delta_range = bool(r.delta_range)
            s.tok_delta = _arr(r.tok_delta, 4 * r.n_tokens, np.uint16) if have_delta else None
            s.tok_delta8 = _arr(r.tok_delta8, 4 * r.n_tokens, np.uint8) if have_delta else None
            s.tok_esc = _arr(r.esc, 2 * r.n_esc, np.uint32)
            s.text = bytes(C.string_at(r.text, r.text_len)) if r.text else None  # the device formatter's bodies
        lib().emul_result_free(rp)
        return s


def unescape_delta8(tok_delta8, tok_esc):
    """DATOK_COMPACT8 -> the u16 form: 255 means "see the escape list" ({token, field << 16 | value} pairs, any order)"""
    d = tok_delta8.astype(np.uint16).reshape(-1, 4).copy()
    e = tok_esc.reshape(-1, 2)
    assert int((d == 255).sum()) == e.shape[0], "every 255 has exactly one escape entry"
    for tok, fv in e:
        assert d[tok, fv >> 16] == 255
        d[tok, fv >> 16] = fv & 0xFFFF
    return d.reshape(-1)


def expand_delta(tok_delta, text_tok_end, text_byte_end):
    """numpy restatement of datok_expand() (format.cpp): DATOK_COMPACT deltas -> absolute (tok_bytes, tok_pos)"""
    d = tok_delta.astype(np.int64).reshape(-1, 4)
    n = d.shape[0]
    tb, tp = np.zeros(2 * n, np.int64), np.zeros(2 * n, np.int64)
    lo = 0
    for t in range(len(text_tok_end)):
        hi = int(text_tok_end[t])
        if hi > lo:
            seg = d[lo:hi]
            byte0 = int(text_byte_end[t - 1]) if t else 0
            ends = byte0 + np.cumsum(seg[:, 0] + seg[:, 1])
            tb[2 * lo:2 * hi:2] = ends - seg[:, 1]
            tb[2 * lo + 1:2 * hi:2] = ends
            rends = np.cumsum(seg[:, 2] + seg[:, 3])
            tp[2 * lo:2 * hi:2] = rends - seg[:, 3]
            tp[2 * lo + 1:2 * hi:2] = rends
        lo = hi
    return tb, tp


# oracle status -> DATOK_ERR_* code
ORACLE_TO_ERR = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 5, 7: 5}


def assert_matches_oracle(s, o, flags, ctx=""):
    """s: Structured (emulation or CUDA result), o: OracleResult for the same input and flags."""
    assert s.status == ORACLE_TO_ERR[o.status], f"{ctx}: status {s.status} vs oracle {o.status}"
    if o.status != 0:
        return
    assert s.n_tokens == o.n_tokens, f"{ctx}: tokens {s.n_tokens} vs {o.n_tokens}"
    assert s.n_sentences == o.n_sent_events, f"{ctx}: sentence events {s.n_sentences} vs {o.n_sent_events}"
    assert s.n_texts == o.n_texts, f"{ctx}: texts {s.n_texts} vs {o.n_texts}"
    assert s.n_runes == o.stats["runes"], f"{ctx}: runes"
    # the C ABI only produces the arrays whose flag was requested (the emulation always does)
    if (flags & 1) or s.tok_bytes.size:
        np.testing.assert_array_equal(s.tok_bytes[0::2], o.tok_byte_start, err_msg=f"{ctx}: token byte starts")
        np.testing.assert_array_equal(s.tok_bytes[1::2], o.tok_byte_end, err_msg=f"{ctx}: token byte ends")
    if (flags & 2) or s.sent_tok.size:
        np.testing.assert_array_equal(s.sent_tok, o.sent_tok_idx.astype(np.uint32), err_msg=f"{ctx}: sentence token index")
    np.testing.assert_array_equal(s.text_tok_end, o.text_tok_end.astype(np.uint32), err_msg=f"{ctx}: text token bounds")
    np.testing.assert_array_equal(s.text_sent_end, o.text_sent_end.astype(np.uint32), err_msg=f"{ctx}: text sentence bounds")
    np.testing.assert_array_equal(s.text_byte_end, o.text_byte_end, err_msg=f"{ctx}: text byte ends")
    if flags & 12:  # a position flag: the TokenWriter kept pos / sent
        if (flags & 4) or s.tok_pos.size:
            np.testing.assert_array_equal(s.tok_pos, o.tok_pos, err_msg=f"{ctx}: token rune offsets")
        if flags & 8:  # `sent` is only maintained consistently under SENTENCE_POS (token_writer.go:103-116,144-153)
            np.testing.assert_array_equal(s.sent_pos, o.sent_pos, err_msg=f"{ctx}: sentence rune offsets")
            np.testing.assert_array_equal(s.text_sentpos_end, o.text_sentpos_end.astype(np.uint32),
                                          err_msg=f"{ctx}: text sent-list bounds")
    assert s.carry_state == o.carry_out["state"], f"{ctx}: carry-out state"
    if getattr(s, "text", None) is not None and not (flags & ~31):  # the emulation runs the device formatter's bodies too
        assert s.text == o.text, f"{ctx}: device-formatted text"
    delta = getattr(s, "tok_delta", None)
    if delta is not None and delta.size and (flags & 12):  # the emulation also fills the DATOK_COMPACT form
        tb, tp = expand_delta(delta, s.text_tok_end, s.text_byte_end)
        np.testing.assert_array_equal(tb[0::2], o.tok_byte_start, err_msg=f"{ctx}: delta-coded byte starts")
        np.testing.assert_array_equal(tb[1::2], o.tok_byte_end, err_msg=f"{ctx}: delta-coded byte ends")
        np.testing.assert_array_equal(tp, o.tok_pos, err_msg=f"{ctx}: delta-coded rune offsets")
        d8 = getattr(s, "tok_delta8", None)
        if d8 is not None and d8.size and not hasattr(s, "expand"):  # the emulation fills the one-byte form as well
            np.testing.assert_array_equal(unescape_delta8(d8, s.tok_esc), delta, err_msg=f"{ctx}: one-byte deltas")
