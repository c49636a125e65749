"""CPU tests of the kernel bodies (datok_b200/csrc/*.cuh) run through tests/emul, a
sequential emulation of the kernels' grid decomposition, against the oracle.
They exercise the speculative-chunk algorithm (sync points, probe hand-off, stitch,
re-walk rounds) and the bit compaction without a GPU.  The CUDA kernels themselves
are checked by the -m gpu tests."""
import os
import random

import numpy as np
import pytest

import parity_util as P
from golden_util import case_id, load_cases

MODELS = ("tokenizer_de.matok", "tokenizer_en.matok", "simpletok.matok", "clitic_test.matok")


@pytest.fixture(scope="module")
def emul_models(testdata):
    return {n: P.EmulModel(os.path.join(testdata, n)) for n in MODELS}


@pytest.fixture(scope="module")
def corpus_lib():
    from datok_b200 import corpus
    return corpus


@pytest.mark.parametrize("case", load_cases(), ids=case_id)
def test_reference_vectors(case, oracle_models, emul_models):
    data = bytes.fromhex(case["input_hex"])
    for flags in {case["flags"], 15, 31}:
        o = oracle_models[case["model"]].transduce(data, flags)
        # mode 0: exact walker only; mode n: fused fast path with n shared-memory table rows
        for chunk, order, mode in ((32, 0, 0), (96, 1, 0), (32, 0, 4), (64, 1, 100000)):
            s = emul_models[case["model"]].transduce(data, flags, chunk, order, mode=mode)
            P.assert_matches_oracle(s, o, flags, f'{case["src"]} chunk={chunk} mode={mode}')


ODD = [b"", b" ", b"\x04a", b"a\x04\x04b", b"\x04", b"\x04\x04", b" " * 1025, b"a" * 1100, b"a" * 1024 + b" b",
       b" " * 1023 + b"a", b" " * 1024 + b"a", b"x\x04 y", b"\n\x04\n\x04\n", b"abc", b"abc\x04",
       b"\xff\xfe abc \xc3", b"\xc3\xa4\xc3", b"\xe2\x82", b"a\xe2\x82\xacb \xed\xa0\x80 \xf4\x90\x80\x80 \xc0\xaf",
       "ä ö ü ß „Zitat“ – …".encode(), b"\x80\x80\x80", b"a\x80b", b"\xf0\x9f\x98\x80 \xf0\x9f\x98",
       "日本語のテキスト。次の文。".encode(), b"\x00\x01\x02 \x7f", b"a.\x04b.\x04c", b"Hallo.\x04\nWelt.\x04\n\nEnde",
       b"...", b". . .", b"!!! ??? ...", b"<x y=\"a b", b"\"\"\"", b"\n\n\n", b"a\n\nb", ("é" * 1030).encode(),
       ("é" * 600).encode(), ("日" * 600).encode(), ("日" * 1100).encode()]


@pytest.mark.parametrize("model", MODELS)
def test_edge_and_panic_inputs(model, oracle_models, emul_models):
    """empty / EOT-only / token-less texts, invalid UTF-8, and the inputs on which the
    reference panics: the error code must match the panic."""
    seen = set()
    for data in ODD:
        for flags in (0, 3, 15, 31, 4, 8, 8 | 16, 2 | 4):
            o = oracle_models[model].transduce(data, flags)
            seen.add(o.status)
            for chunk, mode in ((32, 0), (64, 300), (1024, 2)):
                s = emul_models[model].transduce(data, flags, chunk, 0, mode=mode)
                P.assert_matches_oracle(s, o, flags, f"{model} {data[:24]!r} flags={flags} chunk={chunk} mode={mode}")
    assert {0, 1, 2, 3}.issubset(seen)


@pytest.mark.parametrize("kind,model", [(2, "tokenizer_de.matok"), (3, "tokenizer_en.matok"),
                                        (1, "simpletok.matok"), (4, "tokenizer_de.matok")])
def test_synthetic_corpora(kind, model, oracle_models, emul_models, corpus_lib):
    """the bench corpora, scaled down, for several chunk sizes"""
    a = corpus_lib.generate(kind, 1 << 19, seed=11 + kind)
    for flags in (15, 31):
        o = oracle_models[model].transduce_np(a, flags)
        assert o.status == 0
        for chunk, order, mode in ((64, 1, 0), (256, 0, 256), (2048, 0, 16), (32, 1, 100000)):
            s = emul_models[model].transduce(a, flags, chunk, order, mode=mode)
            P.assert_matches_oracle(s, o, flags, f"kind={kind} chunk={chunk} mode={mode}")


def test_tiny_chunks_on_markup_heavy_text(oracle_models, emul_models, corpus_lib):
    """32-byte chunks on the long-token corpus: tokens span up to 14 chunks, hard-fail tokens end
    exactly on sync points (regression: a pending END bit must survive the re-walk)"""
    a = corpus_lib.generate(4, 1 << 20, seed=5)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    for mode in (0, 256):
        s = emul_models["tokenizer_de.matok"].transduce(a, 15, 32, 0, mode=mode)
        P.assert_matches_oracle(s, o, 15, f"kind=4 chunk=32 mode={mode}")
        assert s.stats["rounds"] > 5


@pytest.mark.parametrize("name", ["longdoc_handoff_a.bin", "longdoc_handoff_b.bin"])
def test_hand_off_with_stale_bufft_at_the_chunk_end(name, oracle_models, emul_models):
    """regression (16 KiB windows of the C4 corpus): the hand-off probe of the chunk ending at byte 8192
    backtracks to a SentenceEnd point one byte below the chunk end with a stale bufft (matrix.go:573-576).
    The lane must not store anything beyond its chunk on the way back (a race on the GPU; the emulation
    aborts on such a store, chunk_core.cuh store_seg_bits)."""
    import os
    a = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", name), dtype=np.uint8)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, 15)
    assert o.status == 0
    for chunk, mode in ((512, 886), (512, 64), (256, 886), (1024, 300)):
        for order in (0, 1):
            s = emul_models["tokenizer_de.matok"].transduce(a, 15, chunk, order, mode=mode)
            P.assert_matches_oracle(s, o, 15, f"{name} chunk={chunk} mode={mode} order={order}")


@pytest.mark.parametrize("name", ["tokenizer_de.datok", "simpletok.datok"])
def test_double_array_models(name, testdata, corpus_lib):
    """.datok models (datok.go): the product's loader converts the double array to the matrix layout and the
    kernel bodies walk it.  The double-array loop differs from the matrix loop at an EOT (no buffer rewind,
    datok.go:1019-1030): a text's first Token call reaches back to the last token of the text before.  Every
    golden case of datok_test.go, EOT or not, and multi-document corpora against the double-array oracle
    (tests/test_oracle_golden_datok.py pins that oracle)."""
    import json
    from oracle import pyoracle
    em = P.EmulModel(os.path.join(testdata, name))
    om = pyoracle.OracleModel(os.path.join(testdata, name))
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_datok.json")))["cases"]
    n = n_eot = 0
    for c in cases:
        data = bytes.fromhex(c["input_hex"])
        if c["model"] != name:
            continue
        for flags in {15, 31, c["flags"] & 31}:
            o = om.transduce(data, flags)
            for chunk, mode in ((32, 0), (64, 256), (640, 886)):
                P.assert_matches_oracle(em.transduce(data, flags, chunk, 0, mode=mode), o, flags, f"{c['src']} chunk={chunk} mode={mode}")
        n += 1
        n_eot += b"\x04" in data
    assert n >= (100 if name == "tokenizer_de.datok" else 3)
    eot_inputs = ["Erste.\n\n\n\n\x04\nNächst.\x04".encode(), b"\nThis.\n\x04\nAnd.\n\x04\n", b"This.\n\x04And.\n\x04\n", b"Tree\n\x04\n",
                  b"Ein Text. \x04 \n Noch einer, mit Rand . \x04\x04 Ende", b"<a href=\"x \x04 y\">z</a> . \x04"]
    for data in ODD + eot_inputs:
        for flags in (3, 15, 31):
            o = om.transduce(data, flags)
            P.assert_matches_oracle(em.transduce(data, flags, 64, 0, mode=300), o, flags, f"{name} {data[:24]!r} flags={flags}")
    if name == "tokenizer_de.datok":
        a = corpus_lib.generate(4, 1 << 20, seed=5)  # the long-document corpus has no EOT
        P.assert_matches_oracle(em.transduce(a, 15, 640, 0, mode=886), om.transduce_np(a, 15), 15, "long document")
        a = corpus_lib.generate(2, 1 << 19, seed=8)  # ~10 KB documents, each ended by an EOT (+ "\n")
        for flags in (15, 31, 3):
            o = om.transduce_np(a, flags)
            for chunk, mode in ((640, 886), (64, 0), (256, 40)):
                P.assert_matches_oracle(em.transduce(a, flags, chunk, 1, mode=mode), o, flags, f"documents flags={flags} chunk={chunk} mode={mode}")
        rng = random.Random(3)
        for it in range(150):
            data = _fuzz_text(rng, rng.choice((33, 200, 1500)))
            flags = rng.choice((3, 15, 31))
            o = om.transduce(data, flags)
            s = em.transduce(data, flags, rng.choice((32, 96)), 0, mode=rng.choice((0, 64)))
            if s.status == 5 and o.status == 0:
                continue  # a backtrack across a consumed EOT: the reference fires that TextEnd twice (not representable: reported)
            P.assert_matches_oracle(s, o, flags, f"{name} fuzz it={it} {data[:40]!r}")


def _fuzz_text(rng, n):
    alphabet = [b" ", b" ", b" ", b"\n", b"\t", b".", b",", b"!", b"?", b"\x04", b"<", b">", b"\"", b"'", b"&", b";",
                b"-", b"/", b":", b"@", b"a", b"e", b"n", b"r", b"S", b"T", b"1", b"9", "ä".encode(), "ß".encode(),
                "„".encode(), "“".encode(), "…".encode(), "日".encode(), "\U0001F600".encode(), b"\xc3", b"\xa4",
                b"\xe2", b"\x80", b"\xff", b"z.B.", b"Dr.", b"usw.", b"http://", b"www.", b".de", b"&amp;", b"<b>",
                b"</b>", b"...", b"'s", b"n't", b"I."]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(alphabet)
    return bytes(out[:n])


@pytest.mark.parametrize("model", MODELS)
def test_fuzz(model, oracle_models, emul_models):
    """differential fuzzing over a markup/abbreviation/invalid-UTF-8 heavy alphabet"""
    rng = random.Random(20261018)
    for it in range(400):
        data = _fuzz_text(rng, rng.choice((7, 33, 64, 200, 1500, 5000)))
        flags = rng.choice((3, 15, 31, 4, 12, 28))
        o = oracle_models[model].transduce(data, flags)
        chunk = rng.choice((32, 64, 128, 512))
        mode = rng.choice((0, 1, 16, 256, 100000))
        s = emul_models[model].transduce(data, flags, chunk, rng.randint(0, 1), mode=mode)
        P.assert_matches_oracle(s, o, flags, f"{model} it={it} {data[:40]!r} chunk={chunk} mode={mode}")


def test_carry_between_calls(oracle_models, emul_models):
    """a stream cut after an EOT continues from the carried state (shard hand-over)"""
    om, em = oracle_models["tokenizer_de.matok"], emul_models["tokenizer_de.matok"]
    a, b = "Erster Text. Zwei Sätze.\n\x04".encode(), "\nZweiter Text <a href=\"x y\">hier</a>.\x04".encode()
    oa = om.transduce(a, 15)
    ob = om.transduce(b, 15 | 256, carry_in=dict(state=oa.carry_out["state"], ok=0, sentence_end=1, text_end=1))
    sb = em.transduce(b, 15 | 256, 32, 0, carry_state=oa.carry_out["state"], sentence_end=1, text_end=1)
    P.assert_matches_oracle(sb, ob, 15, "second shard")
    whole = om.transduce(a + b, 15)
    assert whole.text == oa.text + ob.text


@pytest.mark.parametrize("model,kind", [("tokenizer_de.matok", 2), ("tokenizer_en.matok", 3), ("tokenizer_de.matok", 4),
                                        ("simpletok.matok", 1)])
def test_frequency_ordered_classes_and_narrow_rows(model, kind, testdata, oracle_models, corpus_lib):
    """the calibrated layout: class ids ordered by frequency, compact rows holding the frequent classes only --
    rarer classes go through the full table (fast_run's rare path), whatever the text they were measured on"""
    em = P.EmulModel(os.path.join(testdata, model))
    a = corpus_lib.generate(kind, 1 << 19, seed=3 + kind)
    o = oracle_models[model].transduce_np(a, 15)
    rng = random.Random(kind)
    fuzz = [_fuzz_text(rng, rng.choice((33, 200, 1500, 5000))) for _ in range(60)]
    n_cls = em.n_classes
    for sample, force in ((a[:1 << 16], 0), (b"aaaa bbbb. ", 0), (a[:1 << 16], 5), (a[:1 << 16], 24), (b"", 0)):
        cols = em.calibrate(sample, force)
        assert 3 <= cols <= n_cls
        for chunk, mode in ((640, 300), (64, 8), (256, 100000)):
            P.assert_matches_oracle(em.transduce(a, 15, chunk, 0, mode=mode), o, 15, f"{model} cols={cols} chunk={chunk} mode={mode}")
        for data in fuzz + ODD:
            oo = oracle_models[model].transduce(data, 31)
            P.assert_matches_oracle(em.transduce(data, 31, 96, 1, mode=64), oo, 31, f"{model} cols={cols} {data[:30]!r}")
    if model == "tokenizer_de.matok" and kind == 2:
        assert em.calibrate(a[:1 << 16], 0) < n_cls  # a German sample does not need every class in the compact rows
