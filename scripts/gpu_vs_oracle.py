"""debug tool: CUDA path vs the oracle on windows of a synthetic corpus; saves the first differing window"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import datok_b200 as d
from datok_b200 import corpus
import pyoracle
kind = {"de": corpus.GERMAN, "en": corpus.ENGLISH, "longdoc": corpus.GERMAN_LONGDOC}[sys.argv[1]]
size = int(sys.argv[2]); W = int(sys.argv[3])
A = np.empty(size, dtype=np.uint8)
if kind == corpus.GERMAN_LONGDOC: corpus.generate_into(kind, corpus.SEED, A)
else: corpus.generate_blocks_into(kind, corpus.SEED, A)
model = "testdata/" + corpus.MODEL_FOR_KIND[kind]
tok = d.LoadTokenizerFile(model)
om = pyoracle.OracleModel(model)
bad = 0
for k in range(size // W):
    a = np.ascontiguousarray(A[k * W:(k + 1) * W])
    o = om.transduce_np(a, 3)
    r = tok.transduce_arrays(a, 3)
    g = np.array(r.tok_bytes, copy=True).reshape(-1, 2); nt = r.n_tokens
    r.close()
    ob = np.stack([o.tok_byte_start, o.tok_byte_end], axis=1)
    n = min(len(ob), len(g))
    diff = np.nonzero((ob[:n] != g[:n]).any(axis=1))[0]
    print(k, o.n_tokens, nt, "first diff", diff[:1], flush=True)
    if len(diff):
        j = int(diff[0]); p = int(ob[j][0])
        print(" oracle", ob[j - 1:j + 3].tolist(), "\n gpu   ", g[j - 1:j + 3].tolist())
        print(" chunk offset of the token", p % 512, "text:", bytes(a[max(0, p - 700):p + 120]))
        os.makedirs("gpurun_out", exist_ok=True)
        lo = max(0, (p - 8192) & ~511)
        np.save(f"gpurun_out/repro_{k}.npy", a[lo:lo + 16384])
        bad += 1
        if bad >= 2: break
