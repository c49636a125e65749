"""Host side of the drop-in boundary: the reference's `Tokenizer` interface
(fomafile.go:29-33) and TokenWriter (token_writer.go:27-36) over the C ABI.

Go names are kept (LoadTokenizerFile, NewTokenWriter, Transduce,
TransduceTokenWriter, Type) so that the parity tests read like the reference's.
All transduction happens in libdatok_b200.so on the GPU; this module only moves
bytes in and formatted text out.
"""
import ctypes as C
import io
import sys

import numpy as np

from . import _lib
from ._lib import (COMPACT, COMPACT8, FORMAT, NEWLINE_AFTER_EOT, SENTENCE_POS, SENTENCES, SIMPLE, TOKEN_POS, TOKENS, WRITER_USED, Callbacks,
                   Carry, EVENT_CB, TOKEN_CB)


class ReferencePanic(RuntimeError):
    """The Go reference panics on this input (outside the parity domain); `code` is DATOK_ERR_*."""

    def __init__(self, code, msg):
        super().__init__(f"datok error {code}: {msg}")
        self.code = code


class DatokError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"datok error {code}: {msg}")
        self.code = code


def _raise(code):
    L = _lib.lib()
    msg = (L.datok_last_error() or b"").decode("utf-8", "replace") or L.datok_strerror(code).decode()
    if 1 <= code <= 5:
        raise ReferencePanic(code, msg)
    raise DatokError(code, msg)


class TokenWriter:
    """token_writer.go:27-33: a struct of four public callables.  Users may supply
    their own; NewTokenWriter builds the stock, flag-driven one."""

    def __init__(self, Token=None, SentenceEnd=None, TextEnd=None, Flush=None):
        self.Token = Token or (lambda offset, buf: None)
        self.SentenceEnd = SentenceEnd or (lambda n: None)
        self.TextEnd = TextEnd or (lambda n: None)
        self.Flush = Flush or (lambda: None)
        self._stock = None  # (writer, flags) for NewTokenWriter instances: formatted natively


def NewTokenWriter(w, flags):
    """token_writer.go:36.  `w` is a binary file-like object (has .write)."""
    tw = TokenWriter()
    tw._stock = {"w": w, "flags": int(flags), "init": True}
    tw.Flush = lambda: (w.flush() if hasattr(w, "flush") else None)
    return tw


def _go_runes(b: bytes):
    """[]rune(string(b)) with Go's one-U+FFFD-per-bad-byte rule"""
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        pass
    out = []
    i, n = 0, len(b)
    while i < n:
        c = b[i]
        need = 1 if c < 0x80 else 2 if 0xC2 <= c <= 0xDF else 3 if 0xE0 <= c <= 0xEF else 4 if 0xF0 <= c <= 0xF4 else 0
        if need:
            try:
                out.append(b[i:i + need].decode("utf-8"))
                if len(b[i:i + need]) == need:
                    i += need
                    continue
                out.pop()
            except UnicodeDecodeError:
                pass
        out.append("�")
        i += 1
    return "".join(out)


class Result:
    """Offset arrays of one transduction (include/datok_b200.h: datok_view) as numpy views.
    Keeps the native result alive; .close() or garbage collection frees it."""

    status = 0

    def __init__(self, handle, device=False):
        self._h = handle
        L = _lib.lib()
        v = L.datok_result_view(handle).contents
        self.device = device
        self.n_tokens, self.n_sentences, self.n_texts = v.n_tokens, v.n_sentences, v.n_texts
        self.n_sent_pos, self.n_runes = v.n_sent_pos, v.n_runes
        self.has_invalid_utf8 = bool(v.has_invalid_utf8)
        self.carry_state = v.carry_out.state
        self.carry = Carry(v.carry_out.state, v.carry_out.sentence_end, v.carry_out.text_end, 0)
        self.ms_h2d, self.ms_kernels, self.ms_d2h = v.ms_h2d, v.ms_kernels, v.ms_d2h
        self._ptrs = dict(tok_bytes=v.tok_bytes, tok_pos=v.tok_pos, sent_pos=v.sent_pos, sent_tok=v.sent_tok,
                          text_tok_end=v.text_tok_end, text_sent_end=v.text_sent_end,
                          text_sentpos_end=v.text_sentpos_end, text_byte_end=v.text_byte_end)
        if not device:
            def arr(p, n):
                if not p or n == 0:
                    return np.zeros(0, dtype=np.ctypeslib.as_array(p, shape=(1,)).dtype if p else np.uint32)
                return np.ctypeslib.as_array(p, shape=(int(n),))
            self.tok_bytes = arr(v.tok_bytes, 2 * v.n_tokens)
            self.tok_pos = arr(v.tok_pos, 2 * v.n_tokens)
            self.tok_delta = arr(v.tok_delta, 4 * v.n_tokens) if v.tok_delta else None
            self.tok_delta8 = arr(v.tok_delta8, 4 * v.n_tokens) if v.tok_delta8 else None
            self.tok_esc = arr(v.tok_esc, 2 * v.n_esc) if v.tok_delta8 else None
            if self.tok_delta is not None or self.tok_delta8 is not None:
                self.tok_bytes = self.tok_pos = None  # DATOK_COMPACT(8): the absolute arrays are rebuilt on demand
            self.sent_pos = arr(v.sent_pos, v.n_sent_pos)
            self.sent_tok = arr(v.sent_tok, v.n_sentences)
            self.text_tok_end = arr(v.text_tok_end, v.n_texts)
            self.text_sent_end = arr(v.text_sent_end, v.n_texts)
            self.text_sentpos_end = arr(v.text_sentpos_end, v.n_texts)
            self.text_byte_end = arr(v.text_byte_end, v.n_texts)
            # DATOK_FORMAT: the TokenWriter's text, formatted on the device (a numpy view of the result's pinned memory)
            self.text = arr(v.text, v.text_len) if v.text else None
        self.text_len = v.text_len

    def expand(self):
        """datok_expand(): absolute tok_bytes / tok_pos of a DATOK_COMPACT result (host-side decode)"""
        if (self.tok_delta is not None or self.tok_delta8 is not None) and self.tok_bytes is None:
            tb = np.empty(2 * self.n_tokens, dtype=np.uint32)
            tp = np.empty(2 * self.n_tokens, dtype=np.int32)
            rc = _lib.lib().datok_expand(self._h, tb.ctypes.data, tp.ctypes.data)
            if rc:
                _raise(rc)
            self.tok_bytes, self.tok_pos = tb, tp
        return self

    def device_ptr(self, name):
        p = self._ptrs[name]
        return C.cast(p, C.c_void_p).value if p else None

    def close(self):
        if self._h:
            _lib.lib().datok_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_buffer(data):
    """-> (address, nbytes, keepalive) without copying numpy arrays"""
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return (a.ctypes.data if a.size else None), a.size, a
    b = bytes(data)
    buf = C.create_string_buffer(b, len(b)) if b else None
    return (C.addressof(buf) if buf is not None else None), len(b), buf


class MatrixTokenizer:
    """MatrixTokenizer (matrix.go:16-26) resident on one B200."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        L = _lib.lib()
        vals = [C.c_uint32() for _ in range(6)]
        L.datok_model_info(handle, *[C.byref(x) for x in vals])
        (self.state_count, self.sigma_count, self.n_classes, self.epsilon, self.unknown,
         self.identity) = [x.value for x in vals]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().datok_free(self._h)
            self._h = None

    # -- Tokenizer interface (fomafile.go:29-33) -------------------------------
    def Type(self):
        return _lib.lib().datok_model_type(self._h).decode()

    def Save(self, file):
        """matrix.go:107-123: gzip(WriteTo).  Returns (bytes of the image, None) or (0, DatokError)."""
        L = _lib.lib()
        rc = L.datok_save(self._h, str(file).encode())
        if rc:
            return 0, DatokError(rc, (L.datok_last_error() or b"").decode("utf-8", "replace"))
        return L.datok_write_image(self._h, None, 0), None

    def WriteTo(self, w):
        """matrix.go:126-210: the uncompressed image into a binary writer; returns the byte count"""
        L = _lib.lib()
        n = L.datok_write_image(self._h, None, 0)
        if not n:
            raise DatokError(_lib.ERR_INVALID_ARG, (L.datok_last_error() or b"").decode("utf-8", "replace"))
        buf = C.create_string_buffer(n)
        L.datok_write_image(self._h, buf, n)
        w.write(buf.raw)
        return n

    def Transduce(self, r, w):
        """matrix.go:340-342"""
        return self.TransduceTokenWriter(r, NewTokenWriter(w, SIMPLE))

    def TransduceTokenWriter(self, r, w: TokenWriter, batch_bytes=None):
        """matrix.go:348-698.  `r`: bytes-like or binary file-like (.read()).

        batch_bytes: stream a file-like `r` in batches of about that many bytes, each cut after its last EOT
        (the stream front-end: bounded memory; state, sentenceEnd/textEnd and the writer's `init` flag are
        carried from batch to batch through the C ABI's carry / DATOK_NOT_FINAL / DATOK_WRITER_USED)."""
        if batch_bytes and hasattr(r, "read"):
            return self._transduce_stream(r, w, int(batch_bytes))
        data = r.read() if hasattr(r, "read") else r
        if isinstance(data, str):
            data = data.encode("utf-8")
        self._transduce_batch(data, w, None, True)
        w.Flush()  # matrix.go:374 defer w.Flush()
        return True

    def _transduce_batch(self, data, w, carry, final):
        """one datok_transduce call + the host half of the TokenWriter; returns the carry for the next batch"""
        L = _lib.lib()
        addr, n, keep = _as_buffer(data)
        extra = COMPACT8 | (0 if final else _lib.NOT_FINAL)  # the smallest transport form: the host decodes it anyway
        if w._stock is not None:
            st = w._stock
            # the stock writer's text is formatted on the device (DATOK_FORMAT) and comes back in one piece
            flags = st["flags"] | (0 if st["init"] else WRITER_USED) | FORMAT | (0 if final else _lib.NOT_FINAL)
            res = self.transduce_arrays_raw(addr, n, flags, carry)
            try:
                if res.n_tokens:
                    st["init"] = False
                st["w"].write(res.text.tobytes() if res.text is not None else b"")
                return res.carry
            finally:
                res.close()
        # custom TokenWriter: replay the events into its callables
        res = self.transduce_arrays_raw(addr, n, TOKENS | SENTENCES | extra, carry)
        try:
            def on_token(_u, buf, buf_bytes, _off_bytes, off_runes):
                w.Token(off_runes, list(_go_runes(C.string_at(buf, buf_bytes))))
            cb = Callbacks(None, TOKEN_CB(on_token), EVENT_CB(lambda _u: w.SentenceEnd(0)),
                           EVENT_CB(lambda _u: w.TextEnd(0)))
            rc = L.datok_replay(res._h, addr, n, C.byref(cb))
            if rc:
                _raise(rc)
            return res.carry
        finally:
            res.close()

    def _transduce_stream(self, r, w, batch_bytes):
        """the io.Reader front-end through the C ABI's datok_stream_* (what the Go shim calls): blocks are pushed as
        they are read, the library cuts them after EOT bytes and carries the state from batch to batch"""
        L = _lib.lib()
        stock = w._stock
        flags = (stock["flags"] | FORMAT | (0 if stock["init"] else WRITER_USED)) if stock is not None else (TOKENS | SENTENCES | COMPACT8)
        st = L.datok_stream_open(self._h, flags)
        if not st:
            raise DatokError(_lib.ERR_INVALID_ARG, "datok_stream_open failed")
        keep = []  # custom writers: the replay needs the batch's input bytes

        def deliver(handle, data):
            res = Result(handle)
            try:
                if stock is not None:
                    if res.n_tokens:
                        stock["init"] = False
                    stock["w"].write(res.text.tobytes() if res.text is not None else b"")
                else:
                    def on_token(_u, buf, buf_bytes, _off_bytes, off_runes):
                        w.Token(off_runes, list(_go_runes(C.string_at(buf, buf_bytes))))
                    cb = Callbacks(None, TOKEN_CB(on_token), EVENT_CB(lambda _u: w.SentenceEnd(0)), EVENT_CB(lambda _u: w.TextEnd(0)))
                    buf = C.create_string_buffer(data, len(data)) if data else None
                    rc = L.datok_replay(res._h, buf, len(data), C.byref(cb))
                    if rc:
                        _raise(rc)
            finally:
                res.close()

        try:
            pending = b""
            while True:
                block = r.read(batch_bytes)
                if not block:
                    break
                out = C.c_void_p()
                done0 = L.datok_stream_bytes_done(st)
                rc = L.datok_stream_push(st, block, len(block), C.byref(out))
                if rc:
                    _raise(rc)
                pending += block
                if out.value:
                    n = L.datok_stream_bytes_done(st) - done0
                    deliver(out.value, pending[:n])
                    pending = pending[n:]
            out = C.c_void_p()
            rc = L.datok_stream_finish(st, C.byref(out))
            if rc:
                _raise(rc)
            if out.value:
                deliver(out.value, pending)
        finally:
            L.datok_stream_close(st)
        w.Flush()  # matrix.go:374 defer w.Flush()
        return True

    # -- offset-array API ---------------------------------------------------------
    def transduce_arrays_raw(self, addr, n, flags, carry=None):
        out = C.c_void_p()
        cin = C.byref(carry) if carry is not None else None
        rc = _lib.lib().datok_transduce(self._h, addr, n, flags, cin, C.byref(out))
        if rc:
            _raise(rc)
        return Result(out.value)

    def transduce_arrays(self, data, flags=TOKENS | SENTENCES | TOKEN_POS | SENTENCE_POS, carry=None):
        """Offsets only: token byte spans, TokenWriter.pos / .sent entries, per-text bounds."""
        addr, n, keep = _as_buffer(data)
        return self.transduce_arrays_raw(addr, n, flags, carry)

    def transduce_device(self, d_ptr, n, flags=TOKENS | SENTENCES | TOKEN_POS | SENTENCE_POS, carry=None):
        """Input already in HBM (device pointer); offset arrays stay on the device."""
        out = C.c_void_p()
        cin = C.byref(carry) if carry is not None else None
        rc = _lib.lib().datok_transduce_device(self._h, d_ptr, n, flags, cin, C.byref(out))
        if rc:
            _raise(rc)
        return Result(out.value, device=True)

    def format(self, res: Result, data, flags):
        """datok_format: exact TokenWriter text for a host-resident result"""
        L = _lib.lib()
        addr, n, keep = _as_buffer(data)
        need = L.datok_format(res._h, addr, n, flags, None, 0)
        if need == C.c_size_t(-1).value:
            raise DatokError(_lib.ERR_INVALID_ARG, "result does not hold the arrays these flags need")
        out = C.create_string_buffer(max(1, need))
        L.datok_format(res._h, addr, n, flags, out, need)
        return out.raw[:need]

    def gather_bound(self):
        """byte steps per second of the bare shared-memory gather chain of the walk (measurement only)"""
        v = C.c_double()
        rc = _lib.lib().datok_measure_gather_bound(self._h, C.byref(v))
        if rc:
            _raise(rc)
        return v.value

    def kernel_times(self):
        L = _lib.lib()
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        k = L.datok_last_kernel_times(self._h, names, ms, 16)
        return {names[i].decode(): float(ms[i]) for i in range(k)}

    def launch_count(self):
        return _lib.lib().datok_last_launch_count(self._h)

    def stats(self):
        """fix-up rounds of the last call; resident table rows / class columns and chunk size of the current layout"""
        v = [C.c_uint32() for _ in range(4)]
        _lib.lib().datok_last_stats(self._h, *[C.byref(x) for x in v])
        return dict(zip(("fixup_rounds", "hot_rows", "hot_cols", "chunk_bytes"), (x.value for x in v)))


def LoadTokenizerFile(file, device=0):
    """fomafile.go:452-484 for the MATOK magic.  Returns None on any error, like the
    reference (which logs and returns nil); the reason goes to stderr."""
    L = _lib.lib()
    err = C.c_int()
    h = L.datok_load(str(file).encode(), device, C.byref(err))
    if not h:
        msg = (L.datok_last_error() or b"").decode("utf-8", "replace")
        print(f"datok: {msg or L.datok_strerror(err.value).decode()}", file=sys.stderr)
        return None
    return MatrixTokenizer(h, device)


LoadMatrixFile = LoadTokenizerFile  # matrix.go:214


class Automaton:
    """What LoadFomaFile returns (fomafile.go:36-52): the parsed foma file, waiting for ToMatrix().  The
    intermediate representation lives inside the library; this object only remembers the file."""

    def __init__(self, file):
        self.file = str(file)

    def ToMatrix(self, device=0):
        """matrix.go:30-99 -- the matrix model, in memory and resident on `device`.  None on error."""
        L = _lib.lib()
        err = C.c_int()
        h = L.datok_load_foma(self.file.encode(), device, C.byref(err))
        if not h:
            msg = (L.datok_last_error() or b"").decode("utf-8", "replace")
            print(f"datok: {msg or L.datok_strerror(err.value).decode()}", file=sys.stderr)
            return None
        return MatrixTokenizer(h, device)


def LoadFomaFile(file):
    """fomafile.go:56-72.  None (like the reference's nil) when the file cannot be read as gzip."""
    try:
        with open(file, "rb") as f:
            if f.read(2) != b"\x1f\x8b":
                print("datok: gzip: invalid header", file=sys.stderr)
                return None
    except OSError as e:
        print(f"datok: {e}", file=sys.stderr)
        return None
    return Automaton(file)


def convert(foma_file, matok_file):
    """`datok convert -i foma_file -o matok_file` (cmd/datok.go:63): compile without touching a device.
    Raises DatokError with the reference's message."""
    L = _lib.lib()
    rc = L.datok_compile_foma(str(foma_file).encode(), str(matok_file).encode())
    if rc:
        raise DatokError(rc, (L.datok_last_error() or b"").decode("utf-8", "replace"))


def load_error_code(file, device=0):
    """DATOK_ERR_* that loading `file` produces (0 = loads fine); for tests."""
    L = _lib.lib()
    err = C.c_int()
    h = L.datok_load(str(file).encode(), device, C.byref(err))
    if h:
        L.datok_free(h)
        return 0
    return err.value


def transduce_sharded(toks, data, flags=TOKENS | SENTENCES | TOKEN_POS | SENTENCE_POS, carry=None):
    """datok_transduce_sharded: one corpus over the GPUs the tokenizers `toks` live on (one model instance per device).
    Returns (results per shard, bases [n_shards x 5: bytes, tokens, sentences, texts, sent entries], bounds, info)."""
    L = _lib.lib()
    addr, n, keep = _as_buffer(data)
    nd = len(toks)
    models = (C.c_void_p * nd)(*[t._h for t in toks])
    devices = (C.c_int * nd)(*[t.device for t in toks])
    outs = (C.c_void_p * nd)()
    bases = (C.c_uint64 * (5 * nd))()
    bounds = (C.c_uint64 * (nd + 1))()
    cin = C.byref(carry) if carry is not None else None
    rc = L.datok_transduce_sharded(models, devices, nd, addr, n, flags, cin, outs, bases, bounds)
    if rc:
        msg = (L.datok_sharded_last_error() or b"").decode("utf-8", "replace")
        if 1 <= rc <= 5:
            raise ReferencePanic(rc, msg)
        raise DatokError(rc, msg)
    used, rew = C.c_int(), C.c_int()
    L.datok_sharded_last_info(C.byref(used), C.byref(rew))
    return ([Result(outs[i]) for i in range(nd)], np.array(list(bases), dtype=np.int64).reshape(nd, 5),
            [int(x) for x in bounds], {"used_nccl": bool(used.value), "shards_rewalked": rew.value})
