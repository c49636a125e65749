import io, os, sys, time, faulthandler
faulthandler.enable()
faulthandler.dump_traceback_later(20, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datok_b200 as d
tok = d.LoadTokenizerFile(os.path.join(ROOT, "testdata", "tokenizer_de.matok"))
for data in (b"a", b"", b" "):
    print("arrays", repr(data), flush=True)
    r = tok.transduce_arrays(data, 15) if data else None
    try:
        r = tok.transduce_arrays(data, 3)
        print(" ok", r.n_tokens, r.n_sentences, r.n_texts, flush=True)
    except Exception as e:
        print(" exc", e, flush=True)
    w = io.BytesIO()
    print("ttw", flush=True)
    tok.TransduceTokenWriter(io.BytesIO(data), d.NewTokenWriter(w, 3))
    print(" ->", w.getvalue(), flush=True)
