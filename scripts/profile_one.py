"""one warm-up + one device-resident transduction of a synthetic German corpus (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import datok_b200 as d
from datok_b200 import corpus
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
kind = {"de": corpus.GERMAN, "en": corpus.ENGLISH, "longdoc": corpus.GERMAN_LONGDOC}[sys.argv[2] if len(sys.argv) > 2 else "de"]
a = np.empty(size, dtype=np.uint8)
if kind == corpus.GERMAN_LONGDOC:
    corpus.generate_into(kind, corpus.SEED, a)   # one document, no EOT (C4)
else:
    corpus.generate_blocks_into(kind, corpus.SEED, a)
tok = d.LoadTokenizerFile("testdata/" + corpus.MODEL_FOR_KIND[kind])
d_in = torch.from_numpy(a).cuda()
torch.cuda.synchronize()
for i in range(2):
    r = tok.transduce_device(d_in.data_ptr(), size, 15 | d.COMPACT)
    print(i, r.n_tokens, r.ms_kernels, {k: round(v, 3) for k, v in tok.kernel_times().items()}, flush=True)
    r.close()
print("layout", tok.stats())
