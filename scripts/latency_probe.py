"""times the host-side phases of small and medium calls (diagnostics)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
t0 = time.perf_counter()
import datok_b200 as d
from datok_b200 import corpus
print("import", time.perf_counter() - t0, flush=True)
t0 = time.perf_counter()
tok = d.LoadTokenizerFile("testdata/tokenizer_de.matok")
print("load", time.perf_counter() - t0, flush=True)
for size in (16, 16, 4096, 1 << 20, 1 << 20, 16 << 20, 16 << 20, 64 << 20, 64 << 20):
    a = corpus.generate(2, size) if size > 64 else np.frombuffer(b"Der alte Mann. ", dtype=np.uint8)
    t0 = time.perf_counter()
    r = tok.transduce_arrays(a, 15)
    dt = time.perf_counter() - t0
    print(size, "call %.4f s" % dt, "h2d %.3f kern %.3f d2h %.3f ms" % (r.ms_h2d, r.ms_kernels, r.ms_d2h),
          "launches", tok.launch_count(), {k: round(v, 3) for k, v in tok.kernel_times().items()}, flush=True)
    t0 = time.perf_counter(); r.close(); print("  close %.4f" % (time.perf_counter() - t0), flush=True)
