"""The compile path: LoadFomaFile / ParseFoma (fomafile.go:56-450), ToMatrix (matrix.go:30-99), WriteTo / Save
(matrix.go:107-210).

Pins, without a GPU:
  * the oracle's restatement and the library's compiler (datok_compile_foma, host only) against the reference's
    own compiled artefacts: the image compiled from testdata/X.fst is byte-identical to the gunzipped shipped
    testdata/X.matok (simpletok, clitic_test, tokenizer_de, tokenizer_en), and has the 230 bytes
    matrix_test.go:167 asserts for simpletok;
  * the oracle on every case of matrix_test.go / datok_test.go that builds its model from a foma file
    (tests/golden/reference_vectors_foma.json): bauamt.fst and wahlamt.fst have no identity symbol,
    ignorable_mcs.fst has multi-character symbols in sigma;
  * the kernel bodies (tests/emul) on those in-memory models against the oracle.
The CUDA path on the same models: tests/test_gpu_parity.py::test_foma_compiled_models."""
import ctypes as C
import gzip
import json
import os
import random

import pytest

import parity_util as P
from golden_util import HERE, case_id, check_output

FOMA = ("simpletok.fst", "bauamt.fst", "wahlamt.fst", "ignorable_mcs.fst", "clitic_test.fst")
SHIPPED = ("simpletok", "clitic_test", "tokenizer_de", "tokenizer_en")

with open(os.path.join(HERE, "golden", "reference_vectors_foma.json")) as _f:
    CASES = json.load(_f)["cases"]


@pytest.fixture(scope="module")
def foma_oracles(testdata):
    from oracle import pyoracle
    return {n: pyoracle.OracleModel(os.path.join(testdata, n)) for n in FOMA}


def test_vector_inventory():
    assert len(CASES) >= 33 and sum(len(c["checks"]) for c in CASES) >= 94
    assert {c["model"] for c in CASES} == {"simpletok.fst", "bauamt.fst", "wahlamt.fst", "ignorable_mcs.fst"}


@pytest.mark.parametrize("name", SHIPPED)
def test_oracle_compiles_to_the_shipped_model(name, testdata):
    from oracle import pyoracle
    img = pyoracle.OracleModel(os.path.join(testdata, name + ".fst")).write_matrix()
    assert img == gzip.open(os.path.join(testdata, name + ".matok")).read()
    if name == "simpletok":
        assert len(img) == 230  # matrix_test.go:167


def test_oracle_header_of_models_without_identity(foma_oracles):
    b = foma_oracles["bauamt.fst"]
    # sigma of bauamt.fst: epsilon 0, a b m t u = 3..7, all shifted by one; 7 states
    assert (b.epsilon, b.unknown, b.identity, b.state_count, b.sigma_count) == (1, -1, -1, 7, 9)
    assert b.sigma_lookup(ord("a")) == (4, -1) and b.sigma_lookup(ord("i"))[0] == 0
    assert b.sigma_lookup(0x4E2D) == (0, 0)  # not in sigma, no identity: symbol 0 (matrix.go:430,459)


@pytest.mark.parametrize("case", CASES, ids=lambda c: case_id(c) + ":" + c["model"])
def test_oracle_on_reference_vectors(case, foma_oracles):
    r = foma_oracles[case["model"]].transduce(bytes.fromhex(case["input_hex"]), case["flags"])
    assert r.status == 0
    check_output(case, r.text)


# ---------------------------------------------------------------- the library's compiler (host side, no device)

def _lib():
    from datok_b200 import _lib
    return _lib.lib()


def _convert(src, dst):
    L = _lib()
    rc = L.datok_compile_foma(str(src).encode(), str(dst).encode())
    return rc, (L.datok_last_error() or b"").decode("utf-8", "replace")


@pytest.mark.parametrize("name", SHIPPED)
def test_convert_is_byte_identical_to_the_shipped_model(name, testdata, tmp_path):
    out = tmp_path / (name + ".matok")
    rc, why = _convert(os.path.join(testdata, name + ".fst"), out)
    assert rc == 0, why
    assert gzip.open(out).read() == gzip.open(os.path.join(testdata, name + ".matok")).read()


@pytest.mark.parametrize("name", ("bauamt.fst", "wahlamt.fst", "ignorable_mcs.fst"))
def test_convert_agrees_with_the_oracle(name, testdata, tmp_path, foma_oracles):
    out = tmp_path / "m.matok"
    rc, why = _convert(os.path.join(testdata, name), out)
    assert rc == 0, why
    assert gzip.open(out).read() == foma_oracles[name].write_matrix()


def _foma(tmp_path, text, name="x.fst"):
    p = tmp_path / name
    with gzip.open(p, "wb") as f:
        f.write(text.encode("utf-8"))
    return p


GOOD = ("##foma-net 1.0##\n##props##\n1 6 7 8 2 2 1 1 1 1 1 2 5B57D486\n##sigma##\n0 @_EPSILON_SYMBOL_@\n3 a\n4 b\n"
        "##states##\n0 3 1 0\n1 4 0 1\n-1 -1 -1 -1 -1\n##end##\n")


def test_convert_errors(tmp_path, testdata):
    from datok_b200 import _lib as B
    # not gzip (fomafile.go:64-68), missing file (fomafile.go:57-60)
    raw = tmp_path / "raw.fst"
    raw.write_text(GOOD)
    assert _convert(raw, tmp_path / "o")[0] == B.ERR_IO
    assert _convert(tmp_path / "missing.fst", tmp_path / "o")[0] == B.ERR_IO
    # the well-formed file compiles
    assert _convert(_foma(tmp_path, GOOD), tmp_path / "o")[0] == 0
    # fomafile.go:159-167: deterministic / epsilon free
    rc, why = _convert(_foma(tmp_path, GOOD.replace("2 2 1 1 1 1 1 2", "2 2 0 1 1 1 1 2")), tmp_path / "o")
    assert rc == B.ERR_FORMAT and why == "The FST needs to be deterministic"
    rc, why = _convert(_foma(tmp_path, GOOD.replace("2 2 1 1 1 1 1 2", "2 2 1 1 1 0 1 2")), tmp_path / "o")
    assert rc == B.ERR_FORMAT and why == "The FST needs to be epsilon free"
    # fomafile.go:291-313: a:b arcs are refused; fomafile.go:317-319: epsilon:epsilon arcs are refused
    rc, why = _convert(_foma(tmp_path, GOOD.replace("0 3 1 0\n", "0 3 4 1 0\n")), tmp_path / "o")
    assert rc == B.ERR_FORMAT and why.startswith("Unsupported transition")
    rc, why = _convert(_foma(tmp_path, GOOD.replace("0 3 1 0\n", "0 0 1 0\n")), tmp_path / "o")
    assert rc == B.ERR_FORMAT and why == "General epsilon transitions are not supported"
    # an arc into a state beyond the state count (matrix.go:78 panics "stateCount is smaller")
    rc, why = _convert(_foma(tmp_path, GOOD.replace("1 4 0 1\n", "1 4 9 1\n")), tmp_path / "o")
    assert rc == B.ERR_FORMAT
    # the oracle refuses the same files
    from oracle import pyoracle
    for bad in (GOOD.replace("2 2 1 1 1 1 1 2", "2 2 0 1 1 1 1 2"), GOOD.replace("0 3 1 0\n", "0 3 4 1 0\n"),
                GOOD.replace("0 3 1 0\n", "0 0 1 0\n"), GOOD.replace("1 4 0 1\n", "1 4 9 1\n")):
        with pytest.raises(ValueError):
            pyoracle.OracleModel(str(_foma(tmp_path, bad, "bad.fst")))


def test_parser_quirks_agree_with_the_oracle(tmp_path):
    """line feed as a symbol (fomafile.go:425-439), multi-character symbols (dropped arcs), a repeated arc (the map
    keeps the later one), an unknown ## line (ends the parse), an unterminated last line (dropped)"""
    from oracle import pyoracle
    text = ("##foma-net 1.0##\n##props##\n2 9 4 9 1 -1 1 1 1 1 0 2 X\n##sigma##\n0 @_EPSILON_SYMBOL_@\n2 @_IDENTITY_SYMBOL_@\n"
            "3 \n\n4  \n5 +MCS\n6 @_TOKEN_BOUND_@\n7 a\n8 ä\n9 中\n"
            "##states##\n0 7 1 0\n3 0 0\n4 0 0\n5 2\n2 1\n8 1\n9 1\n1 0 6 0 0\n7 1\n7 2\n2 7 1 1\n"
            "-1 -1 -1 -1 -1\n##end##\n##bogus##\n0 7 3 0")
    p = _foma(tmp_path, text, "quirks.fst")
    out = tmp_path / "quirks.matok"
    rc, why = _convert(p, out)
    assert rc == 0, why
    o = pyoracle.OracleModel(str(p))
    assert gzip.open(out).read() == o.write_matrix()
    # ... and the saved model (it has an identity symbol, so it survives the u16 header) walks like the in-memory one
    o2 = pyoracle.OracleModel(str(out))
    for data in ("aä 中a\na", "aaa a\taa", "a\n\na 中"):
        assert o2.transduce(data.encode()).text == o.transduce(data.encode()).text
    # (a rune outside sigma would make the reference index the matrix with the missing unknown symbol and panic)
    assert o.transduce(b"ax").status == 8 and o2.transduce(b"ax").status == 8
    assert o.transduce(b"aaa a").text == b"aaa\na\n\n\n"


# ---------------------------------------------------------------- kernel bodies on the compiled models (tests/emul)

@pytest.fixture(scope="module")
def foma_emul(testdata):
    return {n: P.EmulModel(os.path.join(testdata, n)) for n in FOMA}


@pytest.mark.parametrize("case", CASES, ids=lambda c: case_id(c) + ":" + c["model"])
def test_emulation_on_reference_vectors(case, foma_oracles, foma_emul):
    data = bytes.fromhex(case["input_hex"])
    for flags in {case["flags"], 15, 31}:
        o = foma_oracles[case["model"]].transduce(data, flags)
        for chunk, order, mode in ((32, 0, 0), (96, 1, 0), (32, 0, 4), (64, 1, 100000)):
            s = foma_emul[case["model"]].transduce(data, flags, chunk, order, mode=mode)
            P.assert_matches_oracle(s, o, flags, f'{case["src"]} {case["model"]} chunk={chunk} mode={mode}')


@pytest.mark.parametrize("name", FOMA)
def test_emulation_fuzz(name, foma_oracles, foma_emul):
    """random strings over the model's own alphabet plus strangers (no identity: every stranger is a hard failure)"""
    rng = random.Random(hash(name) & 0xFFFF)
    alphabet = ["a", "b", "m", "t", "u", "i", "d", "w", "h", "l", " ", "\n", "\t", ".", "!", "?", "<", ">", "'", "ä", "中",
                "bau", "bauamt", "wahl", "wahlamt", "<ab>", "\x04"]
    for it in range(60):
        data = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 120))).encode()
        for flags in (3, 15, 31):
            o = foma_oracles[name].transduce(data, flags)
            for chunk, mode in ((32, 0), (32, 3), (64, 100000)):
                s = foma_emul[name].transduce(data, flags, chunk, it & 1, mode=mode)
                P.assert_matches_oracle(s, o, flags, f"{name} {data!r} flags={flags} chunk={chunk} mode={mode}")
