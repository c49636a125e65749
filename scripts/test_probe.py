import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.perf_counter()
import datok_b200 as d
from oracle import pyoracle
import parity_util as P
from golden_util import load_cases, check_output
print("imports", time.perf_counter() - t0, flush=True)
t0 = time.perf_counter()
ms = {n: d.LoadTokenizerFile(os.path.join(ROOT, "testdata", n)) for n in ("tokenizer_de.matok", "tokenizer_en.matok", "simpletok.matok", "clitic_test.matok")}
print("gpu models", time.perf_counter() - t0, flush=True)
t0 = time.perf_counter()
om = {n: pyoracle.OracleModel(os.path.join(ROOT, "testdata", n)) for n in ms}
print("oracle models", time.perf_counter() - t0, flush=True)
for case in load_cases()[:12]:
    tok = ms[case["model"]]; data = bytes.fromhex(case["input_hex"])
    t0 = time.perf_counter()
    w = io.BytesIO(); tw = d.NewTokenWriter(w, case["flags"] & 0xFF)
    tok.TransduceTokenWriter(io.BytesIO(data), tw)
    t1 = time.perf_counter()
    o = om[case["model"]].transduce(data, case["flags"])
    t2 = time.perf_counter()
    r = tok.transduce_arrays(data, 15)
    t3 = time.perf_counter()
    print(case["src"], "ttw %.4f oracle %.4f arrays %.4f" % (t1 - t0, t2 - t1, t3 - t2), flush=True)
