"""Parity tests proper: the CUDA path, called through the C ABI (datok_b200 ->
libdatok_b200.so), against the CPU oracle on identical inputs.  Bit-exact:
token byte spans, TokenWriter.pos / .sent rune offsets, sentence and text
structure, formatted output for every flag combination, and the error code on
inputs where the reference panics."""
import io
import os
import random

import numpy as np
import pytest

import parity_util as P
from golden_util import case_id, check_output, load_cases

pytestmark = pytest.mark.gpu

MODELS = ("tokenizer_de.matok", "tokenizer_en.matok", "simpletok.matok", "clitic_test.matok")


@pytest.fixture(scope="module")
def gpu_models(testdata):
    import datok_b200 as d
    ms = {n: d.LoadTokenizerFile(os.path.join(testdata, n)) for n in MODELS}
    assert all(m is not None for m in ms.values()), "CUDA extension failed to load a model (no fallback exists)"
    yield ms
    for m in ms.values():
        m.close()


class _Status:
    def __init__(self, code):
        self.status = code


def gpu_arrays(tok, data, flags, **kw):
    import datok_b200 as d
    try:
        return tok.transduce_arrays(data, flags, **kw)
    except d.ReferencePanic as e:
        return _Status(e.code)


def test_native_library_is_loaded(gpu_models):
    from datok_b200 import _lib
    assert os.path.basename(_lib.LIB_PATH) == "libdatok_b200.so" and os.path.exists(_lib.LIB_PATH)
    assert gpu_models["tokenizer_de.matok"].Type() == "MATOK"
    assert gpu_models["tokenizer_de.matok"].state_count == 18400
    loaded = open("/proc/self/maps").read()
    assert "libdatok_b200.so" in loaded


@pytest.mark.parametrize("case", load_cases(), ids=case_id)
def test_reference_vectors(case, gpu_models, oracle_models):
    """the reference's own golden vectors, through Transduce/TransduceTokenWriter"""
    import datok_b200 as d
    tok = gpu_models[case["model"]]
    data = bytes.fromhex(case["input_hex"])
    w = io.BytesIO()
    tw = d.NewTokenWriter(w, case["flags"] & 0xFF)
    if case["flags"] & d.WRITER_USED:
        tw._stock["init"] = False
    assert tok.TransduceTokenWriter(io.BytesIO(data), tw)
    check_output(case, w.getvalue())
    o = oracle_models[case["model"]].transduce(data, case["flags"])
    assert w.getvalue() == o.text
    for flags in (15, 31):
        o = oracle_models[case["model"]].transduce(data, flags)
        P.assert_matches_oracle(gpu_arrays(tok, data, flags), o, flags, case["src"])


from test_emul_parity import ODD, _fuzz_text  # noqa: E402  (same inputs as the CPU emulation tests)


@pytest.mark.parametrize("model", MODELS)
def test_edge_and_panic_inputs(model, gpu_models, oracle_models):
    for data in ODD:
        for flags in (0, 3, 15, 31, 4, 8, 8 | 16, 2 | 4):
            o = oracle_models[model].transduce(data, flags)
            P.assert_matches_oracle(gpu_arrays(gpu_models[model], data, flags), o, flags,
                                    f"{model} {data[:24]!r} flags={flags}")


@pytest.mark.parametrize("kind,model", [(2, "tokenizer_de.matok"), (3, "tokenizer_en.matok"),
                                        (1, "simpletok.matok"), (4, "tokenizer_de.matok")])
def test_synthetic_corpora(kind, model, gpu_models, oracle_models):
    from datok_b200 import corpus
    a = corpus.generate(kind, 4 << 20, seed=corpus.SEED + kind)
    for flags in (15, 31):
        o = oracle_models[model].transduce_np(a, flags)
        assert o.status == 0
        s = gpu_arrays(gpu_models[model], a, flags)
        P.assert_matches_oracle(s, o, flags, f"kind={kind} flags={flags}")
        assert gpu_models[model].format(s, a, flags & 31) == o.text


@pytest.mark.parametrize("model", MODELS)
def test_fuzz(model, gpu_models, oracle_models):
    rng = random.Random(7)
    for it in range(150):
        data = _fuzz_text(rng, rng.choice((7, 33, 64, 200, 1500, 9000)))
        flags = rng.choice((3, 15, 31, 4, 12, 28))
        o = oracle_models[model].transduce(data, flags)
        P.assert_matches_oracle(gpu_arrays(gpu_models[model], data, flags), o, flags, f"{model} it={it} {data[:40]!r}")


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 5, 6, 7, 8, 12, 15, 16 | 15, 16 | 12, 16 | 8, 0])
def test_format_every_flag_combination(flags, gpu_models, oracle_models):
    """host half of the TokenWriter (datok_format) == NewTokenWriter(w, flags) output"""
    import datok_b200 as d
    data = ("\nErste Zeile. Und <b>noch</b> eine!\n\x04\nZweiter Text – mit „Zitat“ usw. Ende?\x04"
            "Dritter.\n\x04\n").encode() + b"Kaputt \xff\xc3 hier.\x04"
    tok = gpu_models["tokenizer_de.matok"]
    o = oracle_models["tokenizer_de.matok"].transduce(data, flags)
    w = io.BytesIO()
    try:
        tok.TransduceTokenWriter(data, d.NewTokenWriter(w, flags))
        assert o.status == 0 and w.getvalue() == o.text
    except d.ReferencePanic as e:
        assert e.code == P.ORACLE_TO_ERR[o.status] != 0


def test_custom_token_writer_replay(gpu_models, oracle_models):
    """user-supplied TokenWriter callables (token_writer.go:27-33) see the reference's event stream"""
    import datok_b200 as d
    data = "  Der alte Mann.\nEr ging. \x04\nNeu hier?\x04".encode()
    ev = []
    tw = d.TokenWriter(Token=lambda off, buf: ev.append(("T", off, "".join(buf))),
                       SentenceEnd=lambda n: ev.append(("S",)), TextEnd=lambda n: ev.append(("X",)))
    assert gpu_models["tokenizer_de.matok"].TransduceTokenWriter(data, tw)
    o = oracle_models["tokenizer_de.matok"].transduce(data, 3)
    toks = [e for e in ev if e[0] == "T"]
    assert len(toks) == o.n_tokens
    for k, (_, off, buf) in enumerate(toks):
        assert off == o.tok_offset[k]
        assert buf.encode() == data[o.tok_buf_start[k]:o.tok_byte_end[k]]
    # same interleaving as the stock SIMPLE writer: Token -> "x\n", SentenceEnd/TextEnd -> "\n"
    text = "".join(b[off:] + "\n" if k == "T" else "\n" for (k, *rest) in ev for off, b in [rest or (0, "")])
    assert text.encode() == o.text


def test_carry_between_calls(gpu_models, oracle_models):
    import datok_b200 as d
    om, tok = oracle_models["tokenizer_de.matok"], gpu_models["tokenizer_de.matok"]
    a, b = "Erster Text. Zwei Sätze.\n\x04".encode(), "\nZweiter Text <a href=\"x y\">hier</a>.\x04".encode()
    ra = tok.transduce_arrays(a, 15)
    oa = om.transduce(a, 15)
    assert ra.carry_state == oa.carry_out["state"]
    ob = om.transduce(b, 15 | 256, carry_in=dict(state=oa.carry_out["state"], ok=0, sentence_end=1, text_end=1))
    rb = tok.transduce_arrays(b, 15 | 256, carry=d.Carry(ra.carry_state, 1, 1, 0))
    P.assert_matches_oracle(rb, ob, 15, "second shard")


def test_chunk_size_independence(testdata, oracle_models, monkeypatch):
    """the result must not depend on the speculation granularity"""
    import datok_b200 as d
    from datok_b200 import corpus
    a = corpus.generate(4, 1 << 20, seed=5)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    for chunk in ("32", "96", "1024", "8192"):
        monkeypatch.setenv("DATOK_CHUNK", chunk)
        tok = d.LoadTokenizerFile(os.path.join(testdata, "tokenizer_de.matok"))
        P.assert_matches_oracle(tok.transduce_arrays(a, 15), o, 15, f"chunk={chunk}")
        tok.close()


@pytest.mark.parametrize("name", ["longdoc_handoff_a.bin", "longdoc_handoff_b.bin"])
def test_hand_off_with_stale_bufft_at_the_chunk_end(name, gpu_models, oracle_models):
    """regression: see tests/test_emul_parity.py (a lane stored into its successor's words)"""
    import os
    a = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", name), dtype=np.uint8)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    for _ in range(20):  # (a race: not every run showed it)
        s = gpu_arrays(gpu_models["tokenizer_de.matok"], a, 15)
        P.assert_matches_oracle(s, o, 15, name)


def test_long_document_corpus_in_windows(gpu_models, oracle_models):
    """C4 (one long document, markup- and abbreviation-heavy: far backtracks, hand-off probes over long
    tokens, fix-up rounds): 256 MiB against the oracle, in 64 MiB windows"""
    from datok_b200 import corpus
    n, w = 256 << 20, 64 << 20
    a = np.empty(n, dtype=np.uint8)
    corpus.generate_into(corpus.GERMAN_LONGDOC, corpus.SEED + 3, a)
    tok = gpu_models["tokenizer_de.matok"]
    for k in range(n // w):
        part = np.ascontiguousarray(a[k * w:(k + 1) * w])
        o = oracle_models["tokenizer_de.matok"].transduce_np(part, 3)
        r = tok.transduce_arrays(part, 3)
        assert r.n_tokens == o.n_tokens, f"window {k}"
        np.testing.assert_array_equal(r.tok_bytes[0::2], o.tok_byte_start, err_msg=f"window {k}")
        np.testing.assert_array_equal(r.tok_bytes[1::2], o.tok_byte_end, err_msg=f"window {k}")
        np.testing.assert_array_equal(r.sent_tok, o.sent_tok_idx.astype(np.uint32), err_msg=f"window {k}")
        r.close()


@pytest.mark.parametrize("name", ["tokenizer_de.datok", "simpletok.datok"])
def test_double_array_models(name, testdata):
    """LoadTokenizerFile on a .datok file (fomafile.go:476-480): bit-exact against the double-array oracle, with and
    without EOT bytes -- the double-array loop does not rewind the buffer at an EOT (datok.go:1019-1030), so a text's
    first Token call reaches back to the last token of the text before (tests/test_emul_parity.py has the details)"""
    import json
    import datok_b200 as d
    from datok_b200 import _lib, corpus
    from oracle import pyoracle
    from test_emul_parity import ODD
    tok = d.LoadTokenizerFile(os.path.join(testdata, name))
    assert tok is not None and tok.Type() == "DATOK"
    om = pyoracle.OracleModel(os.path.join(testdata, name))
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_datok.json")))["cases"]
    n = 0
    eot_inputs = ["Erste.\n\n\n\n\x04\nNächst.\x04".encode(), b"\nThis.\n\x04\nAnd.\n\x04\n", b"This.\n\x04And.\n\x04\n", b"Tree\n\x04\n",
                  b"Ein Text. \x04 \n Noch einer, mit Rand . \x04\x04 Ende", b"<a href=\"x \x04 y\">z</a> . \x04"]
    for data in [bytes.fromhex(c["input_hex"]) for c in cases if c["model"] == name] + ODD + eot_inputs:
        for flags in (15, 31, 3):
            o = om.transduce(data, flags)
            s = gpu_arrays(tok, data, flags)
            P.assert_matches_oracle(s, o, flags, f"{name} {data[:30]!r} flags={flags}")  # (also when both say "the reference panics here")
            if o.status == 0:
                assert tok.format(s, data, flags) == o.text
                rf = tok.transduce_arrays(data, flags | d.FORMAT)
                assert rf.text.tobytes() == o.text
                # the delta-coded forms do not apply (their cursors restart at a text): absolute arrays come back
                rc = tok.transduce_arrays(data, flags | d.COMPACT8)
                assert rc.tok_delta8 is None and rc.n_tokens == o.n_tokens
        n += 1
    assert n >= (100 if name == "tokenizer_de.datok" else 3)
    if name == "tokenizer_de.datok":
        a = corpus.generate(corpus.GERMAN_LONGDOC, 4 << 20, seed=5)
        P.assert_matches_oracle(gpu_arrays(tok, a, 15), om.transduce_np(a, 15), 15, "long document")
        a = corpus.generate(corpus.GERMAN, 4 << 20, seed=6)   # ~10 KB documents, each ended by an EOT
        for flags in (15, 31):
            o = om.transduce_np(a, flags)
            P.assert_matches_oracle(gpu_arrays(tok, a, flags), o, flags, f"documents flags={flags}")
            rf = tok.transduce_arrays(a, flags | d.FORMAT)
            assert rf.text.tobytes() == o.text
        w = io.BytesIO()
        assert tok.TransduceTokenWriter(io.BytesIO(a.tobytes()), d.NewTokenWriter(w, 15))
        assert w.getvalue() == om.transduce_np(a, 15).text
    # its stream cannot be cut behind an EOT
    with pytest.raises(d.DatokError) as e:
        tok.transduce_arrays(b"Ein Text.\x04", 15 | d.NOT_FINAL)
    assert e.value.code == _lib.ERR_UNSUPPORTED_MODEL
    tok.close()


def test_gather_bound_measurement(gpu_models):
    """bench.py's second bound (the bare gather chain of the walk): runs and gives a plausible rate"""
    g = gpu_models["tokenizer_de.matok"].gather_bound()
    assert 1e11 < g < 1e14


def test_large_input_properties(gpu_models, oracle_models):
    """64 MiB German corpus: size-independent properties plus oracle parity on a prefix."""
    from datok_b200 import corpus
    n = 64 << 20
    a = np.empty(n, dtype=np.uint8)
    docs = corpus.generate_blocks_into(corpus.GERMAN, corpus.SEED, a, block=16 << 20)
    tok = gpu_models["tokenizer_de.matok"]
    r = tok.transduce_arrays(a, 15)
    tb = r.tok_bytes.astype(np.int64)
    starts, ends = tb[0::2], tb[1::2]
    assert r.n_texts == docs == int((a == 4).sum())
    assert (ends > starts).all() and (starts[1:] >= ends[:-1]).all() and ends[-1] <= n
    assert (np.diff(r.text_tok_end.astype(np.int64)) > 0).all() and r.text_tok_end[-1] == r.n_tokens
    assert r.text_sent_end[-1] == r.n_sentences and r.n_sent_pos == 2 * r.n_sentences
    tp = r.tok_pos.astype(np.int64)
    assert (tp[1::2] > tp[0::2]).all()
    # rune offsets restart at 0..few in every text and never exceed the text's rune count
    first = np.concatenate(([0], r.text_tok_end[:-1].astype(np.int64)))
    assert (tp[2 * first] >= 0).all() and (tp[2 * first] < 64).all()
    # every token surface is free of whitespace bytes the root state skips
    sample = np.random.default_rng(1).integers(0, r.n_tokens, 20000)
    for k in sample:
        s = a[starts[k]:ends[k]]
        assert s.size and s[0] not in (32, 10, 9, 4)
    # oracle parity on the first block (documents are independent given the carried state)
    cut = int(r.text_byte_end[np.searchsorted(r.text_byte_end, 8 << 20)])
    o = oracle_models["tokenizer_de.matok"].transduce_np(a[:cut], 15)
    k = o.n_tokens
    np.testing.assert_array_equal(r.tok_bytes[:2 * k:2], o.tok_byte_start)
    np.testing.assert_array_equal(r.tok_pos[:2 * k], o.tok_pos)
    np.testing.assert_array_equal(r.sent_pos[:o.sent_pos.size], o.sent_pos)


def test_device_resident_path(gpu_models, oracle_models):
    """datok_transduce_device: input and offset arrays stay in HBM"""
    import torch
    from datok_b200 import corpus
    a = corpus.generate(2, 2 << 20, seed=3)
    tok = gpu_models["tokenizer_de.matok"]
    d_in = torch.from_numpy(a).cuda()
    r = tok.transduce_device(d_in.data_ptr(), a.size, 15)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    assert (r.n_tokens, r.n_sentences, r.n_texts) == (o.n_tokens, o.n_sent_events, o.n_texts)
    import ctypes as C
    host = np.empty(2 * r.n_tokens, dtype=np.int32)
    torch.cuda.synchronize()
    cudart = C.CDLL("libcudart.so.12")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert cudart.cudaMemcpy(host.ctypes.data, r.device_ptr("tok_pos"), host.nbytes, 2) == 0  # cudaMemcpyDeviceToHost
    np.testing.assert_array_equal(host, o.tok_pos)


def test_shards_equal_one_stream(gpu_models, oracle_models):
    """EOT-aligned shards walked independently (DATOK_NOT_FINAL + carry), concatenated with
    their index bases, equal the single-stream result (the multi-GPU path, on one GPU)."""
    import datok_b200 as d
    from datok_b200 import corpus, shard
    a = corpus.generate(corpus.GERMAN, 2 << 20, seed=9)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    tok = gpu_models["tokenizer_de.matok"]
    plan = shard.plan_shards(a, 4)
    parts, carry, seen = [], None, False
    for r, (lo, hi) in enumerate(plan):
        f = 15 | (0 if r == len(plan) - 1 else d.NOT_FINAL) | (d.WRITER_USED if seen else 0)
        res = tok.transduce_arrays(a[lo:hi], f, carry=carry)
        parts.append(res)
        carry = d.Carry(res.carry_state, 1, 1, 0)
        seen = seen or res.n_tokens > 0
    tb = np.concatenate([p.tok_bytes.astype(np.int64) + lo for p, (lo, _) in zip(parts, plan)])
    np.testing.assert_array_equal(tb[0::2], o.tok_byte_start)
    np.testing.assert_array_equal(tb[1::2], o.tok_byte_end)
    np.testing.assert_array_equal(np.concatenate([p.tok_pos for p in parts]), o.tok_pos)
    np.testing.assert_array_equal(np.concatenate([p.sent_pos for p in parts]), o.sent_pos)
    assert sum(p.n_texts for p in parts) == o.n_texts
    # a non-final input that stops inside a token is refused
    with pytest.raises(d.DatokError) as ei:
        tok.transduce_arrays(b"mitten im Wor", 15 | d.NOT_FINAL)
    assert ei.value.code == 23


@pytest.mark.parametrize("model,kind", [("tokenizer_de.matok", 2), ("tokenizer_en.matok", 3)])
def test_pipelined_host_path(model, kind, testdata, oracle_models, monkeypatch):
    """large host inputs are cut after EOT bytes into pieces whose copies overlap the kernels
    (api.cu run_pipelined): same arrays as the oracle's single stream, for every flag set"""
    import datok_b200 as d
    from datok_b200 import corpus
    a = corpus.generate(kind, 6 << 20, seed=21)
    monkeypatch.setenv("DATOK_PIECE_MB", "1")
    tok = d.LoadTokenizerFile(os.path.join(testdata, model))
    for flags in (15, 31, 3, 12, 5):
        o = oracle_models[model].transduce_np(a, flags)
        r = tok.transduce_arrays(a, flags)
        P.assert_matches_oracle(r, o, flags, f"pipelined flags={flags}")
        assert tok.format(r, a, flags) == o.text
        rc = tok.transduce_arrays(a, flags | d.COMPACT)
        assert rc.tok_bytes is None and rc.tok_delta is not None
        assert tok.format(rc, a, flags) == o.text
        P.assert_matches_oracle(rc.expand(), o, flags, f"pipelined compact flags={flags}")
    # a reused TokenWriter and a carried state behave like in the single-pass path
    head = "Vorspann ohne Ende.\n\x04".encode()
    oa = oracle_models[model].transduce(head, 31)
    ob = oracle_models[model].transduce(a.tobytes(), 31 | 256, carry_in=dict(state=oa.carry_out["state"], ok=0, sentence_end=1, text_end=1))
    rb = tok.transduce_arrays(a, 31 | 256, carry=d.Carry(oa.carry_out["state"], 1, 1, 0))
    P.assert_matches_oracle(rb, ob, 31, "pipelined, writer used")
    # no EOT to cut at: falls back to one pass
    one = corpus.generate(4, 3 << 20, seed=2)
    P.assert_matches_oracle(tok.transduce_arrays(one, 15), oracle_models[model].transduce_np(one, 15), 15, "uncuttable")
    tok.close()
    monkeypatch.setenv("DATOK_NO_PIPELINE", "1")
    tok1 = d.LoadTokenizerFile(os.path.join(testdata, model))
    r1 = tok1.transduce_arrays(a, 31)
    monkeypatch.delenv("DATOK_NO_PIPELINE")
    tok2 = d.LoadTokenizerFile(os.path.join(testdata, model))
    r2 = tok2.transduce_arrays(a, 31)
    for f in ("tok_bytes", "tok_pos", "sent_pos", "sent_tok", "text_tok_end", "text_sent_end", "text_sentpos_end", "text_byte_end"):
        np.testing.assert_array_equal(getattr(r1, f), getattr(r2, f), err_msg=f)
    tok1.close(); tok2.close()


@pytest.mark.parametrize("model,kind", [("tokenizer_de.matok", 2), ("tokenizer_en.matok", 3), ("tokenizer_de.matok", 4),
                                        ("simpletok.matok", 1)])
def test_compact_token_deltas(model, kind, gpu_models, oracle_models):
    """DATOK_COMPACT: 8-byte delta-coded token spans; datok_expand / datok_format / datok_replay decode them"""
    import datok_b200 as d
    from datok_b200 import corpus
    tok = gpu_models[model]
    a = corpus.generate(kind, 3 << 20, seed=33)
    for flags in (15, 31, 1, 4, 5 | 16):
        o = oracle_models[model].transduce_np(a, flags)
        r = tok.transduce_arrays(a, flags | d.COMPACT)
        assert r.tok_delta is not None and r.tok_delta.size == 4 * r.n_tokens
        assert tok.format(r, a, flags) == o.text
        P.assert_matches_oracle(r.expand(), o, flags, f"compact {model} flags={flags}")
    # numpy restatement of the decode agrees with the library's
    r = tok.transduce_arrays(a, 15 | d.COMPACT)
    tb, tp = P.expand_delta(r.tok_delta, r.text_tok_end, r.text_byte_end)
    r.expand()
    np.testing.assert_array_equal(tb, r.tok_bytes)
    np.testing.assert_array_equal(tp, r.tok_pos)
    for data in (b"", b"abc", b"a.\x04b.\x04c", "ä ö ü ß „Zitat“ – …".encode(), b"\nThis.\n\x04\nAnd.\n\x04\n"):
        for flags in (3, 31):
            o = oracle_models[model].transduce(data, flags)
            if o.status:
                continue
            r = tok.transduce_arrays(data, flags | d.COMPACT)
            assert tok.format(r, data, flags) == o.text


def test_full_size_corpus_properties(gpu_models, oracle_models, testdata, monkeypatch):
    """BASELINE.json's C2 size (1 GiB, ~10 KB EOT-separated documents): the pipelined host path against
    size-independent properties, against the single-pass path, and against the oracle on sampled documents."""
    import datok_b200 as d
    from datok_b200 import corpus
    n = 1 << 30
    a = np.empty(n, dtype=np.uint8)
    docs = corpus.generate_blocks_into(corpus.GERMAN, corpus.SEED, a, block=16 << 20)
    tok = gpu_models["tokenizer_de.matok"]
    r = tok.transduce_arrays(a, 15 | d.COMPACT)          # >= 128 MiB: pieces, copies overlapped with the kernels
    assert r.n_texts == docs == int((a == 4).sum())
    assert r.tok_delta.size == 4 * r.n_tokens and r.n_sent_pos == 2 * r.n_sentences
    tte = r.text_tok_end.astype(np.int64)
    assert (np.diff(tte) > 0).all() and tte[-1] == r.n_tokens and r.text_sent_end[-1] == r.n_sentences
    assert int(r.text_byte_end[-1]) == n and (np.diff(r.text_byte_end.astype(np.int64)) > 0).all()
    dl = r.tok_delta.reshape(-1, 4).astype(np.int64)
    # every byte of a text is either skipped before a token or part of one, up to the trailing skip
    used = np.add.reduceat(dl[:, 0] + dl[:, 1], np.concatenate(([0], tte[:-1])))
    text_len = np.diff(np.concatenate(([0], r.text_byte_end.astype(np.int64))))
    assert (used <= text_len).all() and (text_len - used <= 4).all()
    assert (dl[:, 1] > 0).all() and (dl[:, 3] > 0).all() and (dl[:, 3] <= dl[:, 1]).all()
    r.expand()
    tb, tp = r.tok_bytes.copy(), r.tok_pos.copy()
    sp, st = r.sent_pos.copy(), r.sent_tok.copy()
    tbe = r.text_byte_end.copy()
    r.close()
    # the single-pass path (absolute arrays) gives the same stream
    monkeypatch.setenv("DATOK_NO_PIPELINE", "1")
    tok1 = d.LoadTokenizerFile(os.path.join(testdata, "tokenizer_de.matok"))
    r1 = tok1.transduce_arrays(a, 15)
    np.testing.assert_array_equal(r1.tok_bytes, tb)
    np.testing.assert_array_equal(r1.tok_pos, tp)
    np.testing.assert_array_equal(r1.sent_pos, sp)
    np.testing.assert_array_equal(r1.sent_tok, st)
    np.testing.assert_array_equal(r1.text_byte_end, tbe)
    r1.close(); tok1.close()
    # oracle on sampled documents (each starts in the root state: checked through the carry of its predecessor)
    om = oracle_models["tokenizer_de.matok"]
    rng = np.random.default_rng(7)
    starts = np.concatenate(([0], tbe[:-1].astype(np.int64)))
    for k in rng.integers(1, docs, 24):
        lo, hi = int(starts[k]), int(tbe[k])
        o = om.transduce_np(a[lo:hi], 15 | 256)
        t0, t1 = int(tte[k - 1]), int(tte[k])
        np.testing.assert_array_equal(tb[2 * t0:2 * t1:2].astype(np.int64) - lo, o.tok_byte_start)
        np.testing.assert_array_equal(tb[2 * t0 + 1:2 * t1:2].astype(np.int64) - lo, o.tok_byte_end)
        np.testing.assert_array_equal(tp[2 * t0:2 * t1], o.tok_pos)


def test_stream_front_end(gpu_models, oracle_models):
    """TransduceTokenWriter over a file-like reader in EOT-aligned batches (bounded memory): the same text
    as one call, for the stock writer and for a custom TokenWriter"""
    import datok_b200 as d
    from datok_b200 import corpus
    tok = gpu_models["tokenizer_de.matok"]
    a = corpus.generate(2, 2 << 20, seed=77).tobytes() + "Schluss ohne EOT. Noch ein Satz".encode()
    for flags in (3, 15, 31):
        o = oracle_models["tokenizer_de.matok"].transduce(a, flags)
        for batch in (64 << 10, 300_000, 8 << 20):
            w = io.BytesIO()
            assert tok.TransduceTokenWriter(io.BytesIO(a), d.NewTokenWriter(w, flags), batch_bytes=batch)
            assert w.getvalue() == o.text, (flags, batch)
    events, ref = [], []
    mk = lambda ev: d.TokenWriter(Token=lambda off, buf: ev.append(("T", off, "".join(buf))),
                                  SentenceEnd=lambda n: ev.append("S"), TextEnd=lambda n: ev.append("X"))
    tok.TransduceTokenWriter(io.BytesIO(a), mk(events), batch_bytes=128 << 10)
    tok.TransduceTokenWriter(a, mk(ref))
    assert events == ref and len(events) > 1000


@pytest.mark.parametrize("hot_rows,chunk,threads,hot_cols", [("8", "64", "256", ""), ("40", "96", "768", "12"),
                                                             ("300", "2048", "512", ""), ("2000", "640", "1024", "30")])
def test_stress_configurations(hot_rows, chunk, threads, hot_cols, testdata, oracle_models, monkeypatch):
    """few resident rows (most steps take the cold path through the full table), few columns in the compact rows
    (most classes take it), small and odd chunk sizes, other CTA sizes: the rare paths of the real kernels under load"""
    import datok_b200 as d
    from datok_b200 import corpus
    monkeypatch.setenv("DATOK_HOT_ROWS", hot_rows)
    if hot_cols:
        monkeypatch.setenv("DATOK_HOT_COLS", hot_cols)
    monkeypatch.setenv("DATOK_CHUNK", chunk)
    monkeypatch.setenv("DATOK_FUSED_THREADS", threads)
    rng = random.Random(int(hot_rows) * 7919 + int(chunk))
    for model, kinds in (("tokenizer_de.matok", (2, 4)), ("tokenizer_en.matok", (3,))):
        tok = d.LoadTokenizerFile(os.path.join(testdata, model))
        for kind in kinds:
            a = corpus.generate(kind, 1 << 20, seed=int(chunk) + kind)
            for flags in (15, 31 | d.COMPACT):
                o = oracle_models[model].transduce_np(a, flags & 31)
                r = tok.transduce_arrays(a, flags)
                if flags & d.COMPACT:
                    assert tok.format(r, a, flags & 31) == o.text
                    r.expand()
                P.assert_matches_oracle(r, o, flags & 31, f"{model} kind={kind} hot={hot_rows} chunk={chunk}")
        for it in range(60):
            data = _fuzz_text(rng, rng.choice((33, 200, 1500, 5000)))
            flags = rng.choice((3, 15, 31))
            o = oracle_models[model].transduce(data, flags)
            P.assert_matches_oracle(gpu_arrays(tok, data, flags), o, flags, f"{model} fuzz it={it} {data[:40]!r}")
        tok.close()


@pytest.mark.parametrize("model,kind", [("tokenizer_de.matok", 2), ("tokenizer_en.matok", 3), ("tokenizer_de.matok", 4)])
def test_compact8_token_deltas(model, kind, gpu_models, oracle_models, testdata, monkeypatch):
    """DATOK_COMPACT8: one byte per delta, values >= 255 through the escape list (long tokens, long gaps)"""
    import datok_b200 as d
    from datok_b200 import corpus
    tok = gpu_models[model]
    long_bits = ("\n" + "x" * 300 + " " * 400 + "é" * 280 + "\t" * 256 + "kurz " + "-" * 255 + " Ende.\x04").encode()
    a = np.concatenate([corpus.generate(kind, 2 << 20, seed=41), np.frombuffer(long_bits, dtype=np.uint8),
                        corpus.generate(kind, 1 << 20, seed=42)])
    for flags in (15, 31, 1, 4):
        o = oracle_models[model].transduce_np(a, flags)
        r = tok.transduce_arrays(a, flags | d.COMPACT8)
        assert r.tok_delta8 is not None and r.tok_delta8.size == 4 * r.n_tokens and r.tok_bytes is None
        assert r.tok_esc.size >= 2 * 5 and (r.tok_delta8 == 255).sum() == r.tok_esc.size // 2
        assert (np.diff(r.tok_esc[0::2].astype(np.int64)) >= 0).all()
        assert tok.format(r, a, flags) == o.text
        d16 = P.unescape_delta8(r.tok_delta8, r.tok_esc)
        r16 = tok.transduce_arrays(a, flags | d.COMPACT)
        np.testing.assert_array_equal(d16, r16.tok_delta)
        P.assert_matches_oracle(r.expand(), o, flags, f"compact8 {model} flags={flags}")
    # the pipelined path: escapes of several pieces, merged and sorted
    monkeypatch.setenv("DATOK_PIECE_MB", "1")
    tokp = d.LoadTokenizerFile(os.path.join(testdata, model))
    o = oracle_models[model].transduce_np(a, 15)
    rp = tokp.transduce_arrays(a, 15 | d.COMPACT8)
    assert tokp.format(rp, a, 15) == o.text
    P.assert_matches_oracle(rp.expand(), o, 15, "compact8 pipelined")
    tokp.close()


@pytest.mark.parametrize("flags", [1, 2, 3, 4, 5, 6, 7, 8, 12, 15, 16 | 15, 16 | 12, 16 | 8, 0])
def test_device_formatter_every_flag_combination(flags, gpu_models, oracle_models):
    """DATOK_FORMAT: the text half of the TokenWriter on the device (format_core.cuh) == NewTokenWriter(w, flags) output;
    malformed UTF-8 (surfaces re-encoded) takes the host formatter behind the same flag"""
    import datok_b200 as d
    tok = gpu_models["tokenizer_de.matok"]
    clean = ("\nErste Zeile. Und <b>noch</b> eine!\n\x04\nZweiter Text – mit „Zitat“ usw. Ende?\x04"
             "Dritter.\n\x04\n").encode()
    for data in (clean, clean + b"Kaputt \xff\xc3 hier.\x04", b"abc", b"a.\x04b.\x04c", "ä ö ü ß. Noch einer".encode()):
        o = oracle_models["tokenizer_de.matok"].transduce(data, flags)
        try:
            r = tok.transduce_arrays(data, flags | d.FORMAT)
        except d.ReferencePanic as e:
            assert e.code == P.ORACLE_TO_ERR[o.status] != 0
            continue
        assert o.status == 0
        assert r.text is not None and r.text.tobytes() == o.text, (flags, data[:20])
        assert (r.has_invalid_utf8 or r.tok_bytes.size == 0) and r.n_tokens == o.n_tokens and r.n_texts == o.n_texts
        np.testing.assert_array_equal(r.text_byte_end, o.text_byte_end)
        r.close()


@pytest.mark.parametrize("kind,model", [(2, "tokenizer_de.matok"), (3, "tokenizer_en.matok"), (1, "simpletok.matok"),
                                        (4, "tokenizer_de.matok")])
def test_device_formatter_on_corpora(kind, model, testdata, oracle_models, monkeypatch):
    """single pass and EOT-aligned pieces (texts of the pieces concatenated, per-text bounds rebased), a reused writer"""
    import datok_b200 as d
    from datok_b200 import corpus
    a = corpus.generate(kind, 5 << 20, seed=61 + kind)
    monkeypatch.setenv("DATOK_PIECE_MB", "1")
    tok = d.LoadTokenizerFile(os.path.join(testdata, model))
    for flags in (15, 31, 3, 12, 5, 2):
        o = oracle_models[model].transduce_np(a, flags)
        r = tok.transduce_arrays(a, flags | d.FORMAT)              # pieces (if the corpus has EOTs)
        assert r.text.tobytes() == o.text, f"kind={kind} flags={flags}"
        np.testing.assert_array_equal(r.text_tok_end, o.text_tok_end.astype(np.uint32))
        np.testing.assert_array_equal(r.text_sent_end, o.text_sent_end.astype(np.uint32))
        np.testing.assert_array_equal(r.text_byte_end, o.text_byte_end)
        r.close()
        half = a[:1 << 20]
        o = oracle_models[model].transduce_np(half, flags | 256)
        r = tok.transduce_arrays(half, flags | 256 | d.FORMAT)     # one pass, TokenWriter already used
        assert r.text.tobytes() == o.text
        r.close()
    # malformed UTF-8 in a piece: the call falls back to one pass and the host formatter
    bad = a.copy()
    bad[3 << 20] = 0xFF
    o = oracle_models[model].transduce_np(bad, 15)
    r = tok.transduce_arrays(bad, 15 | d.FORMAT)
    assert r.has_invalid_utf8 and r.text.tobytes() == o.text
    tok.close()


def test_stream_through_the_c_abi(gpu_models, oracle_models):
    """datok_stream_open / push / finish: arbitrary block sizes (blocks without any EOT, blocks ending inside a token),
    the carry inside the stream, DATOK_FORMAT text per batch == the reference's single stream"""
    import ctypes as C
    import datok_b200 as d
    from datok_b200 import _lib, corpus
    from datok_b200.tokenizer import Result
    L = _lib.lib()
    tok = gpu_models["tokenizer_de.matok"]
    a = corpus.generate(2, 3 << 20, seed=5).tobytes() + "Kein EOT am Ende. Letzter Satz".encode()
    for flags in (15, 31, 3):
        o = oracle_models["tokenizer_de.matok"].transduce(a, flags)
        for block in (1000, 70_000, 1 << 20, 8 << 20):
            st = L.datok_stream_open(tok._h, flags | d.FORMAT)
            text, n_tok, results = [], 0, 0
            for lo in range(0, len(a), block):
                out = C.c_void_p()
                assert L.datok_stream_push(st, a[lo:lo + block], len(a[lo:lo + block]), C.byref(out)) == 0
                if out.value:
                    r = Result(out.value); text.append(r.text.tobytes()); n_tok += r.n_tokens; results += 1; r.close()
            out = C.c_void_p()
            assert L.datok_stream_finish(st, C.byref(out)) == 0
            r = Result(out.value); text.append(r.text.tobytes()); n_tok += r.n_tokens; r.close()
            assert L.datok_stream_bytes_done(st) == len(a)
            L.datok_stream_close(st)
            assert b"".join(text) == o.text and n_tok == o.n_tokens, (flags, block)
            assert results >= (1 if block >= (1 << 20) else 10)


def test_sharded_call_on_one_device(testdata, oracle_models):
    """datok_transduce_sharded with several model instances on one GPU (the exchange then stays on the host): shards +
    bases == the single stream, including a cut inside a quoted XML attribute whose shard has to be walked again"""
    import datok_b200 as d
    from datok_b200 import corpus
    path = os.path.join(testdata, "tokenizer_de.matok")
    toks = [d.LoadTokenizerFile(path) for _ in range(3)]
    a = corpus.generate(2, 3 << 20, seed=12)
    at = a.size // 3 + 1
    evil = b' <a href="x \x04 y">z</a> . '
    a[at:at + len(evil)] = np.frombuffer(evil, dtype=np.uint8)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    res, bases, bounds, info = d.transduce_sharded(toks, a, 15)
    assert info["shards_rewalked"] >= 1 and bounds[0] == 0 and bounds[-1] == a.size
    tb = np.concatenate([r.tok_bytes.astype(np.int64) + bounds[i] for i, r in enumerate(res)])
    np.testing.assert_array_equal(tb[0::2], o.tok_byte_start)
    np.testing.assert_array_equal(tb[1::2], o.tok_byte_end)
    np.testing.assert_array_equal(np.concatenate([r.tok_pos for r in res]), o.tok_pos)
    np.testing.assert_array_equal(np.concatenate([r.sent_pos for r in res]), o.sent_pos)
    np.testing.assert_array_equal(np.concatenate([r.sent_tok.astype(np.int64) + bases[i][1] for i, r in enumerate(res)]), o.sent_tok_idx.astype(np.int64))
    np.testing.assert_array_equal(np.concatenate([r.text_tok_end.astype(np.int64) + bases[i][1] for i, r in enumerate(res)]), o.text_tok_end.astype(np.int64))
    assert [int(b[0]) for b in bases] == bounds[:-1] and int(bases[2][1]) == res[0].n_tokens + res[1].n_tokens
    # the formatted text of the shards, in order, is the stream's text
    res2, _, _, _ = d.transduce_sharded(toks, a, 15 | d.FORMAT)
    assert b"".join(r.text.tobytes() for r in res2) == o.text
    for r in res + res2:
        r.close()
    for t in toks:
        t.close()


def test_text_without_sync_points(gpu_models, oracle_models):
    """no whitespace the root state skips (minified markup, CSV, URL lists): every chunk depends on its predecessor's
    exit state; the chain is followed on the device (chain_kernel) instead of one host round per chunk"""
    import time
    tok, om = gpu_models["tokenizer_de.matok"], oracle_models["tokenizer_de.matok"]
    data = (b"http://example.com/a/b/c?d=e&f=g," * 60000)[: 2 << 20]
    o = om.transduce(data, 15)
    t0 = time.perf_counter()
    r = tok.transduce_arrays(data, 15)
    dt = time.perf_counter() - t0
    P.assert_matches_oracle(r, o, 15, "no sync points")
    rounds = tok.stats()["fixup_rounds"]
    assert rounds < 64, rounds          # (one host round per chunk would be ~3300)
    assert dt < 5.0, dt
    data = (b'{"k":[1,2,3],"s":"x"},' * 40000)[: 1 << 20]
    P.assert_matches_oracle(gpu_arrays(tok, data, 3), om.transduce(data, 3), 3, "minified json")


def _markup_text(nbytes, seed):
    """text whose whitespace mostly sits inside quoted attributes: the guess at a chunk's sync point (root state, nothing
    pending) is wrong for most chunks, so the re-walk list is long"""
    rnd = random.Random(seed)
    words = "ein langer Titel mit vielen Worten hier und dort aber nicht immer so wie gedacht".split()
    out, n = [], 0
    while n < nbytes:
        s = '<a href="x y" title="%s">%s</a> %s. ' % (" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 14))),
                                                     rnd.choice(words).capitalize(),
                                                     " ".join(rnd.choice(words) for _ in range(rnd.randint(0, 6))))
        out.append(s)
        n += len(s)
    return "".join(out)[:nbytes].encode()


@pytest.mark.parametrize("nbytes,chunk", [(3000, "64"), (96 << 10, "64"), (2 << 20, "64"), (12 << 20, "64"), (2 << 20, "640")])
def test_rewalk_list_shapes(nbytes, chunk, testdata, oracle_models, monkeypatch):
    """the re-walk kernel spreads its list over the warps of the grid (one chunk per warp for a short list, k lanes per
    warp for a longer one, one chunk per lane beyond that): lists of a few, thousands and >100 000 wrong guesses"""
    import datok_b200 as d
    data = _markup_text(nbytes, 11)
    o = oracle_models["tokenizer_de.matok"].transduce(data, 15)
    monkeypatch.setenv("DATOK_CHUNK", chunk)
    tok = d.LoadTokenizerFile(os.path.join(testdata, "tokenizer_de.matok"))
    P.assert_matches_oracle(tok.transduce_arrays(data, 15), o, 15, f"markup {nbytes} B, chunk {chunk}")
    tok.close()


def test_concurrent_calls_on_one_model(gpu_models, oracle_models):
    """the reference's model is immutable and shareable across goroutines (matrix.go:16-26): calls from several threads
    on one model run side by side, each on an execution context of its own, and give the single-call results"""
    import threading
    import datok_b200 as d
    from datok_b200 import corpus
    tok, om = gpu_models["tokenizer_de.matok"], oracle_models["tokenizer_de.matok"]
    big = corpus.generate(2, 8 << 20, seed=91)
    tok.transduce_arrays(big, 15).close()      # (the one-time calibration happens here)
    inputs = [corpus.generate(2, (1 << 20) + 4096 * i, seed=100 + i) for i in range(6)]
    want = [om.transduce_np(a, 15) for a in inputs]
    errors = []

    def work(i):
        try:
            for rep in range(4):
                r = tok.transduce_arrays(inputs[i], 15)
                P.assert_matches_oracle(r, want[i], 15, f"thread {i} rep {rep}")
                rf = tok.transduce_arrays(inputs[i], 15 | d.FORMAT)
                assert rf.text.tobytes() == want[i].text
                r.close(); rf.close()
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)[:300]))

    th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("name", ["simpletok.fst", "bauamt.fst", "wahlamt.fst", "ignorable_mcs.fst", "clitic_test.fst",
                                  "tokenizer_de.fst"])
def test_foma_compiled_models(name, testdata, tmp_path):
    """The compile path on the device: LoadFomaFile(x).ToMatrix() (fomafile.go:56-450, matrix.go:30-99) gives a model
    that transduces bit-exactly like the oracle's compile of the same file -- the cases matrix_test.go / datok_test.go
    run on foma-built models (bauamt / wahlamt have no identity symbol: every rune outside sigma is a hard failure),
    the odd inputs, random strings -- and Save / WriteTo (matrix.go:107-210) reproduce the shipped .matok files."""
    import gzip
    import json
    import datok_b200 as d
    from datok_b200 import corpus
    from oracle import pyoracle
    path = os.path.join(testdata, name)
    auto = d.LoadFomaFile(path)
    assert auto is not None
    tok = auto.ToMatrix()
    assert tok is not None and tok.Type() == "MATOK"
    om = pyoracle.OracleModel(path)
    assert (tok.state_count, tok.sigma_count, tok.epsilon) == (om.state_count, om.sigma_count, om.epsilon)
    # WriteTo == the oracle's image; for the models the reference ships, == the shipped file
    w = io.BytesIO()
    n = tok.WriteTo(w)
    assert n == len(w.getvalue()) and w.getvalue() == om.write_matrix()
    shipped = os.path.join(testdata, name[:-4] + ".matok")
    if os.path.exists(shipped):
        saved = tmp_path / "saved.matok"
        nbytes, err = tok.Save(saved)
        assert err is None and nbytes == n
        assert gzip.open(saved).read() == gzip.open(shipped).read()
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_foma.json")))["cases"]
    for c in cases:
        if c["model"] != name:
            continue
        data = bytes.fromhex(c["input_hex"])
        out = io.BytesIO()
        assert tok.TransduceTokenWriter(io.BytesIO(data), d.NewTokenWriter(out, c["flags"] & 0xFF))
        check_output(c, out.getvalue())
        assert out.getvalue() == om.transduce(data, c["flags"]).text
    rng = random.Random(len(name))
    alphabet = ["a", "b", "m", "t", "u", "i", "d", "w", "h", "l", " ", "\n", "\t", ".", "!", "?", "<", ">", "'", "ä", "中",
                "bau", "bauamt", "wahl", "wahlamt", "<ab>", "\x04", "He's", "don't "]
    inputs = list(ODD) + ["".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 4000))).encode() for _ in range(40)]
    if name == "tokenizer_de.fst":
        inputs = inputs[:60] + [corpus.generate(corpus.GERMAN, 2 << 20, seed=11).tobytes()]
    for data in inputs:
        for flags in (15, 31, 3):
            o = om.transduce(data, flags)
            if o.status == 8:   # the reference indexes outside its matrix (identity without unknown symbol): no parity domain
                continue
            s = gpu_arrays(tok, data, flags)
            P.assert_matches_oracle(s, o, flags, f"{name} {data[:30]!r} flags={flags}")
            if o.status == 0 and not getattr(s, "has_invalid_utf8", 0):
                rf = tok.transduce_arrays(data, flags | d.FORMAT)
                assert rf.text.tobytes() == o.text
    tok.close()
    # `datok convert`: host only, the file it writes loads like any shipped model
    if os.path.exists(shipped):
        conv = tmp_path / "conv.matok"
        d.convert(path, conv)
        t2 = d.LoadTokenizerFile(conv)
        assert t2 is not None
        for data in (b"bau", b"wald gehen", "Der alte Mann. Er geht.".encode()):
            o = om.transduce(data, 15)
            P.assert_matches_oracle(gpu_arrays(t2, data, 15), o, 15, f"{name} converted")
        t2.close()
