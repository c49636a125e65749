"""Shared helpers: load the reference's golden vectors and evaluate their checks."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)["cases"]


def case_id(c):
    return c["src"].replace("testdata/de/", "")


def check_output(case, out: bytes):
    """Apply every check of a golden case to the formatted output `out`.
    Mirrors the Go helpers: ttokenize (datok_test.go:23-33) and strings.Split."""
    views = {
        "tokens": lambda: re.split(rb"\n+", out)[:-1],
        "split1": lambda: out.split(b"\n"),
        "split2": lambda: out.split(b"\n\n"),
    }
    for ck in case["checks"]:
        kind = ck["kind"]
        exp = bytes.fromhex(ck["eq_hex"]) if "eq_hex" in ck else ck.get("eq")
        where = f'{case["src"]} {kind} idx={ck.get("idx")}'
        if kind == "full":
            assert out == exp, where
        elif kind == "contains":
            assert exp in out, where
        elif kind == "tokens_joined":
            assert b"\n".join(views["tokens"]()) == exp, where
        elif kind in views:
            v = views[kind]()
            assert ck["idx"] < len(v), where
            assert v[ck["idx"]] == exp, where
        elif kind.endswith("_len_gt"):
            assert len(views[kind[:-7]]()) > exp, where
        elif kind.endswith("_len"):
            assert len(views[kind[:-4]]()) == exp, where
        else:
            raise AssertionError("unknown check kind " + kind)
