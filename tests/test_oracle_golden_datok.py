"""Pins the oracle's restatement of the DOUBLE-ARRAY path (datok.go:621-729 loader, datok.go:781-1135
transduction) against the reference's own golden vectors: every case of
datok_test.go that runs on a shipped .datok model, and testdata/de/{dontsplit,split}.txt, which the
reference runs against tokenizer_de.datok (datok_test.go:1201-1236).  Extracted by
tests/golden/make_golden.py into tests/golden/reference_vectors_datok.json.

The oracle converts the double array into the dense layout of the matrix at load time; the two
transduction loops differ in one statement (no buffer rewind at an EOT in datok.go).  There is no
CUDA path for these models yet.

The shipped testdata/tokenizer_de.datok -- like tokenizer_de.matok -- predates part of the grammar that
datok_test.go of 0.3.1 tests (gender forms, "ver.di", ...: `Changes`, 0.3.1): 57 of the 155 cases pin a
newer model, not the algorithm.  For exactly those the oracle must give the SAME output from the
double array as the independently pinned matrix oracle gives from tokenizer_de.matok (two encodings
of one automaton, two loaders), and the set of them is fixed (STALE_MODEL_CASES)."""
import json
import os

import pytest

from golden_util import case_id, check_output

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_vectors_datok.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.fixture(scope="module")
def datok_models(testdata):
    from oracle import pyoracle
    return {n: pyoracle.OracleModel(os.path.join(testdata, n)) for n in ("tokenizer_de.datok", "simpletok.datok")}


def test_vector_inventory():
    assert len(CASES) >= 150
    assert sum(len(c["checks"]) for c in CASES) >= 590
    assert {c["model"] for c in CASES} == {"tokenizer_de.datok", "simpletok.datok"}


STALE_MODEL_CASES = 57


def _stale(case, datok_models, oracle_models):
    """True if the case's checks fail only because the shipped model is older than the test"""
    data = bytes.fromhex(case["input_hex"])
    r = datok_models[case["model"]].transduce(data, case["flags"])
    assert r.status == 0
    try:
        check_output(case, r.text)
        return False
    except AssertionError:
        assert case["model"] == "tokenizer_de.datok", case["src"]
        m = oracle_models["tokenizer_de.matok"].transduce(data, case["flags"])
        assert m.status == 0 and m.text == r.text, case["src"] + ": differs from the matrix oracle as well"
        return True


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_reference_vector(case, datok_models, oracle_models):
    _stale(case, datok_models, oracle_models)


def test_stale_model_cases_are_the_known_ones(datok_models, oracle_models):
    assert sum(_stale(c, datok_models, oracle_models) for c in CASES) == STALE_MODEL_CASES


def test_same_automaton_as_the_matrix_model(datok_models, oracle_models):
    """tokenizer_de.datok and tokenizer_de.matok encode one automaton: without position flags the two
    paths write the same text for the synthetic German corpus (EOTs included) and for every input of
    the matrix path's own golden vectors"""
    from datok_b200 import corpus
    from golden_util import load_cases
    a = corpus.generate(corpus.GERMAN, 1 << 20, seed=99)
    assert datok_models["tokenizer_de.datok"].transduce_np(a, 3).text == oracle_models["tokenizer_de.matok"].transduce_np(a, 3).text
    b = corpus.generate(corpus.GERMAN_LONGDOC, 1 << 20, seed=98)
    assert datok_models["tokenizer_de.datok"].transduce_np(b, 3).text == oracle_models["tokenizer_de.matok"].transduce_np(b, 3).text
    for c in load_cases():
        if c["model"] == "tokenizer_de.matok" and not (c["flags"] & 12):
            data = bytes.fromhex(c["input_hex"])
            assert datok_models["tokenizer_de.datok"].transduce(data, c["flags"]).text == \
                oracle_models["tokenizer_de.matok"].transduce(data, c["flags"]).text, c["src"]


def test_eot_does_not_rewind_the_buffer(datok_models, oracle_models):
    """the one difference between the two loops: after an EOT the double-array walk keeps its buffer
    (datok.go:1019-1030), so the first token of the next text starts its rune count at the end of the
    previous text's last token; the matrix walk rewinds (matrix.go:603)"""
    data = "Der alte Mann.\x04Der junge.".encode()
    da = datok_models["tokenizer_de.datok"].transduce(data, 15)
    ma = oracle_models["tokenizer_de.matok"].transduce(data, 15)
    assert da.status == 0 and ma.status == 0
    assert ma.text.split(b"\n")[-3].split()[0] == b"0"   # pos line of the second text starts at 0
    assert da.text.split(b"\n")[-3].split()[0] == b"1"   # ... at 1: the EOT rune is still in the buffer
    assert [t for t in da.text.split(b"\n") if t and not t[:1].isdigit()] == \
           [t for t in ma.text.split(b"\n") if t and not t[:1].isdigit()]


def test_fuzz_against_the_matrix_oracle(datok_models, oracle_models):
    """differential fuzzing of the two oracles on the two encodings of the German automaton (no EOT in the
    alphabet: an EOT inside a pending token is where the two loops legitimately part, see above)"""
    import random
    from test_emul_parity import _fuzz_text
    rng = random.Random(4711)
    da, ma = datok_models["tokenizer_de.datok"], oracle_models["tokenizer_de.matok"]
    n = 0
    for _ in range(400):
        data = _fuzz_text(rng, rng.choice((7, 33, 200, 1500))).replace(b"\x04", b" ")
        a, b = da.transduce(data, 3), ma.transduce(data, 3)
        assert a.status == b.status, data[:60]
        if a.status == 0:
            assert a.text == b.text, data[:60]
            n += 1
    assert n > 300
