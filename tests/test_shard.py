"""Multi-GPU path on the CPU: shard planning, the count exchange (torch.distributed, gloo,
world_size 2) and the equivalence 'shards walked independently + bases == one stream'.
The per-shard transduction runs through tests/emul (kernel bodies on the CPU) here; the
-m gpu suite repeats the equivalence on the device."""
import os
import socket

import numpy as np
import pytest

import parity_util as P

FLAGS = 15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plan_shards_eot_aligned_and_balanced():
    from datok_b200 import corpus, shard
    a = corpus.generate(corpus.GERMAN, 1 << 20, seed=3)
    for n in (1, 2, 4, 8):
        plan = shard.plan_shards(a, n)
        assert len(plan) == n and plan[0][0] == 0 and plan[-1][1] == a.size
        for (lo, hi), (lo2, _) in zip(plan, plan[1:]):
            assert hi == lo2 and (hi == a.size or a[hi - 1] == 4)
        sizes = [hi - lo for lo, hi in plan]
        assert max(sizes) - min(sizes) < 40000  # documents are 6-14 KB
    # no EOT at all: everything stays in the first shard
    b = np.frombuffer(b"kein eot hier " * 100, dtype=np.uint8)
    plan = shard.plan_shards(b, 4)
    assert plan[0] == (0, b.size) and all(lo == hi for lo, hi in plan[1:])


def _run_shards(model, a, plan, flags):
    """walk every shard on its own (as one rank each would), return per-shard results"""
    from datok_b200 import NOT_FINAL, WRITER_USED
    out = []
    carry, seen_token = 0, False
    for r, (lo, hi) in enumerate(plan):
        last = r == len(plan) - 1
        f = flags | (0 if last else NOT_FINAL) | (WRITER_USED if seen_token else 0)
        s = model.transduce(a[lo:hi], f, 256, 0, carry_state=carry, sentence_end=1 if r else 0,
                            text_end=1 if r else 0, mode=256)
        assert s.status == 0
        out.append(s)
        carry = s.carry_state
        seen_token = seen_token or s.n_tokens > 0
    return out


@pytest.mark.parametrize("n", [2, 3, 8])
def test_shards_equal_one_stream(n, oracle_models, testdata):
    from datok_b200 import corpus, shard
    a = corpus.generate(corpus.GERMAN, 1 << 19, seed=9)
    o = oracle_models["tokenizer_de.matok"].transduce_np(a, FLAGS)
    em = P.EmulModel(os.path.join(testdata, "tokenizer_de.matok"))
    plan = shard.plan_shards(a, n)
    parts = _run_shards(em, a, plan, FLAGS)
    tok_base = np.cumsum([0] + [p.n_tokens for p in parts])
    sent_base = np.cumsum([0] + [p.n_sentences for p in parts])
    tb = np.concatenate([p.tok_bytes.astype(np.int64) + lo for p, (lo, _) in zip(parts, plan)])
    np.testing.assert_array_equal(tb[0::2], o.tok_byte_start)
    np.testing.assert_array_equal(tb[1::2], o.tok_byte_end)
    np.testing.assert_array_equal(np.concatenate([p.tok_pos for p in parts]), o.tok_pos)
    np.testing.assert_array_equal(np.concatenate([p.sent_pos for p in parts]), o.sent_pos)
    np.testing.assert_array_equal(
        np.concatenate([p.sent_tok.astype(np.int64) + tok_base[i] for i, p in enumerate(parts)]), o.sent_tok_idx)
    np.testing.assert_array_equal(
        np.concatenate([p.text_tok_end.astype(np.int64) + tok_base[i] for i, p in enumerate(parts)]), o.text_tok_end)
    np.testing.assert_array_equal(
        np.concatenate([p.text_sent_end.astype(np.int64) + sent_base[i] for i, p in enumerate(parts)]),
        o.text_sent_end)
    assert all(p.carry_state == 1 for p in parts[:-1])  # every shard ended in the guessed state


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from datok_b200 import corpus, shard
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        a = corpus.generate(corpus.GERMAN, 1 << 18, seed=21)
        lo, hi = shard.plan_shards(a, world)[rank]
        o = pyoracle.OracleModel(os.path.join(root, "testdata", "tokenizer_de.matok")).transduce_np(a[lo:hi], FLAGS)
        counts = [hi - lo, o.n_tokens, o.n_sent_events, o.n_texts, o.sent_pos.size, o.carry_out["state"]]
        allc, bases = shard.exchange_counts(counts)
        q.put((rank, counts, allc.tolist(), bases.tolist(), shard.carry_mismatch(allc, rank)))
    finally:
        dist.destroy_process_group()


def test_count_exchange_gloo_world2():
    """the path's only collective: all-gather of per-shard counts -> global index bases"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, all0, b0, mm0), (r1, c1, all1, b1, mm1) = res
    assert all0 == all1 == [c0, c1]
    assert b0 == [0] * 6 and b1 == c0
    assert not mm0 and not mm1  # shard 0 ended in the root state
    assert c0[0] + c1[0] == 1 << 18
