"""Deterministic synthetic corpora of the shapes BASELINE.json names.

Thin wrapper over csrc/corpus_gen.c.  Bench / test tooling, not on the transduction path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20261018

SIMPLE, GERMAN, ENGLISH, GERMAN_LONGDOC = 1, 2, 3, 4
MODEL_FOR_KIND = {SIMPLE: "simpletok.matok", GERMAN: "tokenizer_de.matok", ENGLISH: "tokenizer_en.matok",
                  GERMAN_LONGDOC: "tokenizer_de.matok"}

_lib = None
_ABBR_KEEP = []


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libdatok_corpus.so")
        if not os.path.exists(path):
            raise RuntimeError(path + " is missing: run `python -m datok_b200.build`")
        _lib = C.CDLL(path)
        _lib.datok_corpus_generate.restype = C.c_size_t
        _lib.datok_corpus_generate.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t]
        _lib.datok_corpus_set_abbreviations.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        # abbreviations are sampled from fixture copies of the reference's lists (src/de/abbrv.txt, 5743 forms;
        # src/en/abbrv.txt, 346): SURVEY.md 8d
        for english, name in ((0, "de"), (1, "en")):
            f = os.path.join(os.path.dirname(HERE), "testdata", name, "abbrv.txt")
            if os.path.exists(f):
                blob = open(f, "rb").read()
                _ABBR_KEEP.append(blob)  # the library keeps the pointer
                _lib.datok_corpus_set_abbreviations(english, blob, len(blob))
    return _lib


def generate_into(kind, seed, out: np.ndarray):
    """fill the uint8 array `out` completely; returns the number of documents"""
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    return _load().datok_corpus_generate(kind, seed, out.ctypes.data, out.size)


def generate(kind, nbytes, seed=SEED):
    out = np.empty(nbytes, dtype=np.uint8)
    generate_into(kind, seed, out)
    return out


def generate_blocks_into(kind, seed, out: np.ndarray, block=64 << 20):
    """large corpora: independent blocks with distinct seeds (every block ends a document)"""
    docs = 0
    for i, lo in enumerate(range(0, out.size, block)):
        docs += generate_into(kind, seed + 7919 * i, out[lo:lo + block])
    return docs
