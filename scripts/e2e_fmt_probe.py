"""the headline leg alone: datok_transduce(DATOK_FORMAT) with pinned host input, wall clock per call.

usage: python scripts/e2e_fmt_probe.py [bytes] [reps]     (env: DATOK_PIECE_MB, DATOK_PIPE_TRACE, DATOK_B200_LIB)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import datok_b200 as d
from datok_b200 import corpus, _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib()
ptr = L.datok_host_alloc(size)
arr = np.frombuffer((C.c_uint8 * size).from_address(ptr), dtype=np.uint8)
corpus.generate_blocks_into(corpus.GERMAN, corpus.SEED, arr, block=64 << 20)
tok = d.LoadTokenizerFile("testdata/tokenizer_de.matok")
FLAGS = 15 | d.FORMAT
trace = os.environ.pop("DATOK_PIPE_TRACE", None)
for _ in range(2):
    tok.transduce_arrays(arr, FLAGS).close()
torch.cuda.synchronize()
ts = []
for i in range(reps):
    if trace and i == reps - 1:
        os.environ["DATOK_PIPE_TRACE"] = "1"
    t0 = time.perf_counter()
    r = tok.transduce_arrays(arr, FLAGS)
    ts.append(time.perf_counter() - t0)
    n_text = int(r.text_len)
    r.close()
print(f"e2e formatted piece_mb={os.environ.get('DATOK_PIECE_MB','auto')}: best {min(ts)*1e3:.2f} ms, all {[round(t*1e3,2) for t in ts]}, {size/min(ts)/1e9:.2f} GB/s, text {n_text} B", flush=True)
