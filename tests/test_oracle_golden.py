"""Pins the CPU oracle (oracle/datok_oracle.c) against the reference's own golden
vectors: every case of matrix_test.go / token_writer_test.go that runs on a shipped
.matok model (extracted by tests/golden/make_golden.py)."""
import pytest

from golden_util import case_id, check_output, load_cases

CASES = load_cases()


def test_vector_inventory():
    # the extraction covers the whole path: >= 150 cases / >= 650 assertions
    assert len(CASES) >= 110
    assert sum(len(c["checks"]) for c in CASES) >= 570
    assert {c["model"] for c in CASES} == {"tokenizer_de.matok", "tokenizer_en.matok",
                                           "simpletok.matok", "clitic_test.matok"}


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_reference_vector(case, oracle_models):
    r = oracle_models[case["model"]].transduce(bytes.fromhex(case["input_hex"]), case["flags"])
    assert r.status == 0
    check_output(case, r.text)


def test_token_writer_simple():
    # token_writer_test.go:11-32
    from oracle import pyoracle
    out, st = pyoracle.token_writer_replay(
        pyoracle.SIMPLE, [0, 0, 3, ord("a"), ord("b"), ord("c"), 0, 1, 3, ord("d"), ord("e"), ord("f"), 1, 2])
    assert st == 0 and out == b"abc\nef\n\n\n"


def test_model_header(oracle_models):
    # SURVEY.md section 8 table (parsed from the shipped files)
    de, en = oracle_models["tokenizer_de.matok"], oracle_models["tokenizer_en.matok"]
    assert (de.state_count, de.sigma_count, de.epsilon, de.unknown, de.identity) == (18400, 171, 1, 2, 3)
    assert (en.state_count, en.sigma_count, en.epsilon, en.unknown, en.identity) == (14768, 172, 1, 2, 3)
    st = oracle_models["simpletok.matok"]
    assert (st.state_count, st.sigma_count) == (4, 10)
    assert de.array().size == (18400 + 1) * 171
