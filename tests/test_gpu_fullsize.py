"""BASELINE.json's full-size configurations against the oracle, every array and the formatted text:
C2 (1 GiB German, ~10 KB EOT-separated documents), C3 (1 GiB English), C4 (one 1 GiB document, no EOT).

The multi-document corpora are checked in EOT-aligned slabs that the oracle walks in parallel from the
carry a finished text leaves (state 1, sentenceEnd, textEnd: matrix.go:593-605); every slab's real carry-out
is compared with what its successor assumed and the successor is redone if they differ, so the slabs
together are the oracle's single stream.  The long document has no such cut: the oracle walks it whole."""
import concurrent.futures as cf
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLAGS = 15
ROOT_CARRY = dict(state=1, ok=0, sentence_end=1, text_end=1)


def _cuts(a, slab):
    """slab boundaries right after EOT bytes"""
    cuts = [0]
    n = a.size
    while n - cuts[-1] > slab + slab // 2:
        want = cuts[-1] + slab
        k = np.flatnonzero(a[want:want + (1 << 20)] == 4)
        assert k.size, "no EOT within 1 MiB of a slab boundary"
        cuts.append(want + int(k[0]) + 1)
    cuts.append(n)
    return cuts


def _oracle_slabs(om, a, flags, slab=16 << 20, threads=None):
    cuts = _cuts(a, slab)
    threads = threads or os.cpu_count() or 4

    def run(k, carry):
        f = flags | (256 if k else 0)  # the TokenWriter has seen a token before every slab but the first
        return om.transduce_np(a[cuts[k]:cuts[k + 1]], f, carry_in=carry if k else None)

    with cf.ThreadPoolExecutor(threads) as ex:
        futs = [ex.submit(run, k, ROOT_CARRY) for k in range(len(cuts) - 1)]
        prev = None
        for k, fu in enumerate(futs):
            o = fu.result()
            if prev is not None and any(prev.carry_out[f] != ROOT_CARRY[f] for f in ("state", "sentence_end", "text_end")):
                o = run(k, prev.carry_out)  # the text before ended in another state (EOT inside markup): redo from there
            assert o.status == 0
            yield cuts[k], cuts[k + 1], o
            prev = o


def _format_np(tok, r, a, flags):
    from datok_b200 import _lib
    L = _lib.lib()
    need = L.datok_format(r._h, a.ctypes.data, a.size, flags, None, 0)
    out = np.empty(need, dtype=np.uint8)
    L.datok_format(r._h, a.ctypes.data, a.size, flags, out.ctypes.data, need)
    return out


def _compare_stream(r, text, slabs):
    T = S = SP = X = B = 0  # tokens, sentence events, sent entries, texts, text bytes so far
    tb, tp = r.tok_bytes.astype(np.int64), r.tok_pos
    for lo, hi, o in slabs:
        ctx = f"slab [{lo}, {hi})"
        t1 = T + o.n_tokens
        np.testing.assert_array_equal(tb[2 * T:2 * t1:2] - lo, o.tok_byte_start, err_msg=ctx + ": token starts")
        np.testing.assert_array_equal(tb[2 * T + 1:2 * t1:2] - lo, o.tok_byte_end, err_msg=ctx + ": token ends")
        np.testing.assert_array_equal(tp[2 * T:2 * t1], o.tok_pos, err_msg=ctx + ": token rune offsets")
        s1, sp1, x1 = S + o.n_sent_events, SP + o.sent_pos.size, X + o.n_texts
        np.testing.assert_array_equal(r.sent_tok[S:s1].astype(np.int64) - T, o.sent_tok_idx.astype(np.int64), err_msg=ctx + ": sentence token index")
        np.testing.assert_array_equal(r.sent_pos[SP:sp1], o.sent_pos, err_msg=ctx + ": sentence rune offsets")
        np.testing.assert_array_equal(r.text_tok_end[X:x1].astype(np.int64) - T, o.text_tok_end.astype(np.int64), err_msg=ctx + ": text token bounds")
        np.testing.assert_array_equal(r.text_sent_end[X:x1].astype(np.int64) - S, o.text_sent_end.astype(np.int64), err_msg=ctx + ": text sentence bounds")
        np.testing.assert_array_equal(r.text_sentpos_end[X:x1].astype(np.int64) - SP, o.text_sentpos_end.astype(np.int64), err_msg=ctx + ": text sent-list bounds")
        np.testing.assert_array_equal(r.text_byte_end[X:x1].astype(np.int64) - lo, o.text_byte_end, err_msg=ctx + ": text byte ends")
        ot = o.text_np
        assert np.array_equal(text[B:B + ot.size], ot), ctx + ": formatted text"
        T, S, SP, X, B = t1, s1, sp1, x1, B + ot.size
    assert (T, S, SP, X) == (r.n_tokens, r.n_sentences, r.n_sent_pos, r.n_texts) and B == text.size


@pytest.mark.parametrize("kind,model", [(2, "tokenizer_de.matok"), (3, "tokenizer_en.matok")])
def test_full_size_multi_document_corpus(kind, model, testdata, oracle_models):
    """C2 / C3 at 1 GiB: all arrays and the whole formatted text equal the oracle's"""
    import datok_b200 as d
    from datok_b200 import corpus
    n = 1 << 30
    a = np.empty(n, dtype=np.uint8)
    docs = corpus.generate_blocks_into(kind, corpus.SEED, a, block=64 << 20)
    tok = d.LoadTokenizerFile(os.path.join(testdata, model))
    r = tok.transduce_arrays(a, FLAGS)           # (>= 128 MiB: the pipelined host path)
    assert r.n_texts == docs
    text = _format_np(tok, r, a, FLAGS)
    _compare_stream(r, text, _oracle_slabs(oracle_models[model], a, FLAGS))
    r.close()
    tok.close()


def test_full_size_single_document(testdata, oracle_models):
    """C4 at 1 GiB (one document, no EOT: intra-document speculation, far backtracks, fix-up rounds)"""
    import datok_b200 as d
    from datok_b200 import corpus
    n = 1 << 30
    a = np.empty(n, dtype=np.uint8)
    corpus.generate_into(corpus.GERMAN_LONGDOC, corpus.SEED, a)
    assert not (a == 4).any()
    tok = d.LoadTokenizerFile(os.path.join(testdata, "tokenizer_de.matok"))
    r = tok.transduce_arrays(a, FLAGS)
    text = _format_np(tok, r, a, FLAGS)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, FLAGS)   # the whole stream, one thread
    assert o.status == 0
    _compare_stream(r, text, [(0, n, o)])
    r.close()
    tok.close()
