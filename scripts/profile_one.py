"""one warm-up + one device-resident transduction of a synthetic German corpus (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import datok_b200 as d
from datok_b200 import corpus
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
a = np.empty(size, dtype=np.uint8)
corpus.generate_blocks_into(corpus.GERMAN, corpus.SEED, a)
tok = d.LoadTokenizerFile("testdata/tokenizer_de.matok")
d_in = torch.from_numpy(a).cuda()
torch.cuda.synchronize()
for i in range(2):
    r = tok.transduce_device(d_in.data_ptr(), size, 15 | d.COMPACT)
    print(i, r.n_tokens, r.ms_kernels, tok.kernel_times(), flush=True)
    r.close()
