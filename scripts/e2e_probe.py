"""e2e leg alone (datok_transduce with pinned host input, DATOK_COMPACT8) + the box's raw PCIe rates.

usage: python scripts/e2e_probe.py [bytes] [reps]     (env: DATOK_PIECE_MB, DATOK_PIPE_TRACE, DATOK_B200_LIB)
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
FLAGS = 15 | d.COMPACT8

if os.environ.get("PROBE_PCIE", "1") == "1":
    # raw copy rates of this box: H2D alone, D2H alone, both at once (pinned memory, one 1 GiB copy each)
    h = torch.from_numpy(arr)
    dev = torch.empty(size, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(size, dtype=torch.uint8, device="cuda")
    back = torch.empty(size, dtype=torch.uint8).pin_memory()
    torch.cuda.cudart().cudaHostRegister(ptr, size, 0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t0
    def h2d():
        with torch.cuda.stream(s1): dev.copy_(h, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): back.copy_(dev2, non_blocking=True)
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", lambda: (h2d(), d2h()))):
        fn(); torch.cuda.synchronize()
        t = min(timed(fn) for _ in range(3))
        print(f"pcie {name}: {t*1e3:.2f} ms per GiB-copy = {size/t/1e9:.1f} GB/s per direction", flush=True)
    del dev, dev2, back

for _ in range(2):
    tok.transduce_arrays_raw(ptr, size, FLAGS).close()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    r = tok.transduce_arrays_raw(ptr, size, FLAGS)
    ts.append(time.perf_counter() - t0)
    n_tok = r.n_tokens
    kt = tok.kernel_times()
    r.close()
print("phase sums of the last call (ms):", {k: round(v, 3) for k, v in kt.items()}, "sum", round(sum(kt.values()), 3), flush=True)
print(f"e2e piece_mb={os.environ.get('DATOK_PIECE_MB','64')}: best {min(ts)*1e3:.2f} ms, mean {sum(ts)/len(ts)*1e3:.2f} ms, {size/min(ts)/1e9:.1f} GB/s, tokens {n_tok}", flush=True)
