#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the matrix transduction path.

Reads (never copies) the Go test sources of KorAP/Datok under /root/reference:

    matrix_test.go          every ttokenize / Transduce known-answer case that
                            runs on a shipped .matok model
    token_writer_test.go    the TokenWriter flag / offset / EOT vectors
    (testdata/de/dontsplit.txt / split.txt are deliberately left out, see main())

and writes tests/golden/reference_vectors.json: a list of cases

    {"src": "matrix_test.go:478", "model": "tokenizer_de.matok", "flags": 3,
     "input_hex": "...", "checks": [ {"kind": "tokens", "idx": 0, "eq": "Der"},
                                      {"kind": "tokens_len", "eq": 3}, ... ]}

`tokens`   = ttokenize():  output split on /\\n+/ minus the last field (datok_test.go:23-33)
`split1`   = strings.Split(out, "\\n"),  `split2` = strings.Split(out, "\\n\\n")
`full`     = the whole output string, `contains` = substring.

Only the *vectors* (inputs and expected strings) are extracted; they are data.
Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

REF = os.environ.get("DATOK_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")

FLAGS = {"TOKENS": 1, "SENTENCES": 2, "TOKEN_POS": 4, "SENTENCE_POS": 8,
         "NEWLINE_AFTER_EOT": 16, "SIMPLE": 3}
WRITER_USED = 256

# models a test may reference -> shipped .matok (None = needs the foma parser, out of scope)
FOMA_EQUIV = {"testdata/simpletok.fst": "simpletok.matok"}
# ... and -> shipped .datok (double-array path, datok_test.go)
FOMA_EQUIV_DA = {"testdata/simpletok.fst": "simpletok.datok"}
OUT_DA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors_datok.json")
OUT_FOMA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors_foma.json")
# foma files small enough to commit next to the shipped models (testdata/); the others (abbr_bench, simple_bench)
# are only used by benchmarks
FOMA_FIXTURES = {"simpletok.fst", "bauamt.fst", "wahlamt.fst", "ignorable_mcs.fst", "clitic_test.fst",
                 "tokenizer_de.fst", "tokenizer_en.fst"}


class Tok:
    def __init__(self, kind, val, line):
        self.kind, self.val, self.line = kind, val, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


def go_unquote(body):
    out = bytearray()
    i = 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        e = body[i + 1]
        simple = {"n": 10, "t": 9, "r": 13, "\\": 92, '"': 34, "'": 39, "a": 7, "b": 8, "f": 12, "v": 11}
        if e in simple:
            out.append(simple[e]); i += 2
        elif e == "x":
            out.append(int(body[i + 2:i + 4], 16)); i += 4
        elif e == "u":
            out += chr(int(body[i + 2:i + 6], 16)).encode("utf-8"); i += 6
        elif e == "U":
            out += chr(int(body[i + 2:i + 10], 16)).encode("utf-8"); i += 10
        elif e in "01234567":
            out.append(int(body[i + 1:i + 4], 8)); i += 4
        else:
            raise ValueError("escape \\" + e)
    return bytes(out)


def lex(src):
    toks = []
    i, line, n = 0, 1, len(src)
    while i < n:
        c = src[i]
        if c == "\n":
            toks.append(Tok("nl", "\n", line)); line += 1; i += 1
        elif c in " \t\r":
            i += 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.index("*/", i)
            line += src.count("\n", i, j)
            i = j + 2
        elif c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("str", go_unquote(src[i + 1:j]), line)); i = j + 1
        elif c == "`":
            j = src.index("`", i + 1)
            raw = src[i + 1:j].replace("\r", "")
            toks.append(Tok("str", raw.encode("utf-8"), line))
            line += raw.count("\n"); i = j + 1
        elif c == "'":
            j = src.index("'", i + 1)
            while src[j - 1] == "\\" and src[j - 2] != "\\":
                j = src.index("'", j + 1)
            toks.append(Tok("rune", go_unquote(src[i + 1:j]).decode("utf-8"), line)); i = j + 1
        elif c.isalpha() or c == "_":
            j = i
            while j < n and (src[j].isalnum() or src[j] == "_"):
                j += 1
            toks.append(Tok("id", src[i:j], line)); i = j
        elif c.isdigit():
            j = i
            while j < n and src[j].isalnum():
                j += 1
            toks.append(Tok("num", int(src[i:j], 0), line)); i = j
        elif src.startswith(":=", i):
            toks.append(Tok("op", ":=", line)); i += 2
        else:
            toks.append(Tok("op", c, line)); i += 1
    return toks


def statements(toks):
    """split into statements at newlines outside (), [], {} -- except that a
    trailing '{' keeps the block header as its own statement."""
    cur, depth = [], 0
    for t in toks:
        if t.kind == "nl":
            if depth == 0 and cur:
                yield cur
                cur = []
            continue
        if t.kind == "op" and t.val in "([":
            depth += 1
        elif t.kind == "op" and t.val in ")]":
            depth -= 1
        elif t.kind == "op" and t.val in "{}":
            if cur:
                yield cur
            yield [t]
            cur = []
            continue
        cur.append(t)
    if cur:
        yield cur


def sig(st):
    return " ".join(t.val if t.kind in ("id", "op") else t.kind.upper() for t in st)


class Extractor:
    def __init__(self, fname, double_array=False, foma=False):
        self.fname = fname
        self.double_array = double_array  # resolve .datok models (datok_test.go) instead of skipping them
        # foma: a model built by LoadFomaFile(x).ToMatrix() / .ToDoubleArray() resolves to the foma file itself
        # (the compile path: the harness compiles testdata/<x>.fst instead of loading a shipped model)
        self.foma = foma
        self.cases = []
        self.skipped = []

    def run(self):
        src = open(os.path.join(REF, self.fname), encoding="utf-8").read()
        toks = lex(src)
        # global string vars (var s string = `...`)
        self.gvars = {}
        sts = list(statements(toks))
        # split per top-level func
        i = 0
        depth = 0
        func = None
        body = []
        for st in sts:
            s = sig(st)
            if depth == 0 and s.startswith("var ") and any(t.kind == "str" for t in st):
                self.gvars[st[1].val] = [t for t in st if t.kind == "str"][0].val
            if depth == 0 and s.startswith("func "):
                func = st[1].val
                body = []
                continue
            if st[0].kind == "op" and st[0].val == "{":
                depth += 1
                continue
            if st[0].kind == "op" and st[0].val == "}":
                depth -= 1
                if depth == 0 and func:
                    if func.startswith("Test"):
                        self.do_func(func, body)
                    func = None
                continue
            if func:
                body.append((depth, st))
        return self.cases

    # -- expression helpers -------------------------------------------------
    def strval(self, toks, env):
        """evaluate a string-valued token list, or None"""
        if len(toks) == 1 and toks[0].kind == "str":
            return toks[0].val
        if len(toks) == 1 and toks[0].kind == "id" and toks[0].val in env:
            return env[toks[0].val]
        return None

    def flagval(self, toks):
        v = 0
        for t in toks:
            if t.kind == "id":
                v |= FLAGS[t.val]
        return v

    def do_func(self, func, body):
        env = dict(self.gvars)   # string variables
        models = {}              # var -> model file or None
        writers = {}             # tws var -> {"flags":..., "used": bool}
        readers = {}             # r var -> bytes
        cur = None               # current case
        views = {}               # var -> ("tokens"|"split1"|"split2"|"full")
        wdirty = False           # w holds output of an earlier call

        def split_args(toks):
            args, curr, d = [], [], 0
            for t in toks:
                if t.kind == "op" and t.val in "([":
                    d += 1
                if t.kind == "op" and t.val in ")]":
                    d -= 1
                if t.kind == "op" and t.val == "," and d == 0:
                    args.append(curr); curr = []
                else:
                    curr.append(t)
            if curr:
                args.append(curr)
            return args

        def new_case(line, model, flags, data, kind):
            nonlocal cur, wdirty
            if model is None:
                cur = None
                self.skipped.append(f"{self.fname}:{line} (model needs the foma parser)")
                return
            if data is None:
                cur = None
                self.skipped.append(f"{self.fname}:{line} (input is not a literal)")
                return
            if wdirty and kind != "ttokenize":
                raise RuntimeError(f"{self.fname}:{line}: output buffer not reset")
            cur = {"src": f"{self.fname}:{line}", "func": func, "model": model, "flags": flags,
                   "input_hex": data.hex(), "checks": []}
            self.cases.append(cur)
            wdirty = True

        def find_call(st, name):
            """index of identifier `name` followed by '(' ; returns arg token list"""
            for k, t in enumerate(st):
                if t.kind == "id" and t.val == name and k + 1 < len(st) and st[k + 1].val == "(":
                    d, j = 0, k + 1
                    while True:
                        if st[j].kind == "op" and st[j].val == "(":
                            d += 1
                        if st[j].kind == "op" and st[j].val == ")":
                            d -= 1
                            if d == 0:
                                break
                        j += 1
                    return k, st[k + 2:j]
            return -1, None

        for depth, st in body:
            s = sig(st)
            line = st[0].line
            # model loading
            k, args = find_call(st, "LoadMatrixFile")
            if k >= 0 and st[0].kind == "id":
                p = args[0].val.decode()
                models[st[0].val] = os.path.basename(p)
                continue
            k, args = find_call(st, "LoadFomaFile")
            if k >= 0 and st[0].kind == "id":
                p = args[0].val.decode()
                models[st[0].val] = ("foma", p)
                continue
            if "ToMatrix" in s and st[0].kind == "id" and st[1].val in (":=", "="):
                srcv = st[2].val
                m = models.get(srcv)
                if self.foma and isinstance(m, tuple):
                    models[st[0].val] = os.path.basename(m[1])
                    continue
                models[st[0].val] = FOMA_EQUIV.get(m[1]) if isinstance(m, tuple) else None
                continue
            if "ToDoubleArray" in s and st[0].kind == "id" and st[1].val in (":=", "="):
                m = models.get(st[2].val)
                if self.foma and isinstance(m, tuple):
                    models[st[0].val] = os.path.basename(m[1])
                    continue
                models[st[0].val] = FOMA_EQUIV_DA.get(m[1]) if isinstance(m, tuple) and self.double_array else None
                continue
            k, args = find_call(st, "LoadDatokFile")
            if k < 0:
                k, args = find_call(st, "LoadTokenizerFile")
            if k >= 0 and st[0].kind == "id":
                p = args[0].val.decode()
                models[st[0].val] = os.path.basename(p) if (self.double_array or p.endswith(".matok")) else None
                continue
            # w.Reset()
            if s == "w . Reset ( )":
                wdirty = False
                continue
            # string variable
            if len(st) == 3 and st[0].kind == "id" and st[1].val in (":=", "=") and st[2].kind == "str":
                env[st[0].val] = st[2].val
                continue
            # writers
            k, args = find_call(st, "NewTokenWriter")
            if k >= 0 and st[0].kind == "id" and st[1].val in (":=", "="):
                a = split_args(args)
                writers[st[0].val] = {"flags": self.flagval(a[1]), "used": False}
                continue
            # readers
            if st[0].kind == "id" and len(st) > 2 and st[1].val in (":=", "=") and "strings . NewReader" in s:
                k, args = find_call(st, "NewReader")
                readers[st[0].val] = self.strval(args, env)
                continue
            if s.startswith("r . Reset ("):
                k, args = find_call(st, "Reset")
                readers["r"] = self.strval(args, env)
                continue
            # views
            if st[0].kind == "id" and st[1].val in (":=", "=") and "strings . Split ( w . String ( )" in s:
                sep = [t for t in st if t.kind == "str"][-1].val
                views[st[0].val] = {b"\n": "split1", b"\n\n": "split2"}[sep]
                continue
            if st[0].kind == "id" and st[1].val in (":=", "=") and s.endswith("w . String ( )") and len(st) == 7:
                views[st[0].val] = "full"
                continue
            # ttokenize
            k, args = find_call(st, "ttokenize")
            if k >= 0 and st[0].kind == "id" and st[1].val in (":=", "="):
                a = split_args(args)
                data = self.strval(a[2], env)
                mv = a[0][0].val
                if models.get(mv) is None and mv in ("dat",):
                    cur = None
                    continue
                new_case(line, models.get(mv), FLAGS["SIMPLE"], data, "ttokenize")
                views[st[0].val] = "tokens"
                continue
            # TransduceTokenWriter
            k, args = find_call(st, "TransduceTokenWriter")
            if k >= 0:
                mv = st[k - 2].val
                a = split_args(args)
                kk, ra = find_call(a[0], "NewReader")
                data = self.strval(ra, env) if kk >= 0 else readers[a[0][0].val]
                wv = writers[a[1][0].val]
                fl = wv["flags"] | (WRITER_USED if wv["used"] else 0)
                new_case(line, models.get(mv), fl, data, "transduce")
                wv["used"] = True  # every vector here emits at least one token
                continue
            # Transduce
            k, args = find_call(st, "Transduce")
            if k >= 0:
                mv = st[k - 2].val
                if models.get(mv) is None:
                    cur = None
                    if mv not in ("dat",):
                        self.skipped.append(f"{self.fname}:{line} (model needs the foma parser)")
                    continue
                a = split_args(args)
                kk, ra = find_call(a[0], "NewReader")
                data = self.strval(ra, env) if kk >= 0 else readers[a[0][0].val]
                new_case(line, models.get(mv), FLAGS["SIMPLE"], data, "transduce")
                continue
            # ttokenizeStr inside assert.Equal
            k, args = find_call(st, "ttokenizeStr")
            if k >= 0:
                a = split_args(args)
                mv = a[0][0].val
                if models.get(mv) is None:
                    self.skipped.append(f"{self.fname}:{line} (model needs the foma parser)")
                    continue
                data = self.strval(a[1], env)
                new_case(line, models.get(mv), FLAGS["SIMPLE"], data, "ttokenize")
                exp = [t for t in st if t.kind == "str"][-1].val
                cur["checks"].append({"kind": "tokens_joined", "eq_hex": exp.hex()})
                wdirty = False
                continue
            # assertions
            k, args = find_call(st, "Equal")
            if k >= 0 and st[0].val == "assert":
                if cur is None:
                    continue
                a = split_args(args)
                if len(a) != 2:
                    continue
                self.do_equal(cur, a, env, views, line)
                continue
            k, args = find_call(st, "Contains")
            if k >= 0 and cur is not None:
                a = split_args(args)
                cur["checks"].append({"kind": "contains", "eq_hex": self.strval(a[1], env).hex()})
                continue

    def do_equal(self, cur, a, env, views, line):
        def classify(toks):
            s = sig(toks)
            sv = self.strval(toks, env)
            if sv is not None and not (len(toks) == 1 and toks[0].kind == "id" and toks[0].val in views):
                return ("str", sv)
            if len(toks) == 1 and toks[0].kind == "num":
                return ("num", toks[0].val)
            if s == "w . String ( )":
                return ("view", "full", None)
            if len(toks) == 1 and toks[0].kind == "id" and views.get(toks[0].val) == "full":
                return ("view", "full", None)
            m = re.fullmatch(r"(\w+) \[ NUM \]", s)
            if m and m.group(1) in views:
                return ("view", views[m.group(1)], toks[2].val)
            m = re.fullmatch(r"len \( (\w+) \)", s)
            if m and m.group(1) in views:
                return ("len", views[m.group(1)])
            return ("?", s)

        x, y = classify(a[0]), classify(a[1])
        if x[0] in ("str", "num"):
            x, y = y, x
        if x[0] == "view" and y[0] == "str":
            c = {"kind": x[1], "eq_hex": y[1].hex()}
            if x[2] is not None:
                c["idx"] = x[2]
            cur["checks"].append(c)
        elif x[0] == "len" and y[0] == "num":
            cur["checks"].append({"kind": x[1] + "_len", "eq": y[1]})
        elif x[0] == "view" and y[0] == "view":
            pass  # datStr == matStr (needs the double-array path)
        else:
            self.skipped.append(f"{self.fname}:{line} assert.Equal on non-output values ({x[1]} / {y[1]})")


def file_list_cases():
    """datok_test.go:1201-1236 over testdata/de/*.txt (trimmed, '#' and blank skipped)"""
    cases = []
    for name, kind in (("dontsplit.txt", "dont"), ("split.txt", "split")):
        path = os.path.join(REF, "testdata", "de", name)
        for ln, raw in enumerate(open(path, encoding="utf-8"), 1):
            tok = raw.strip()
            if not tok or tok.startswith("#"):
                continue
            data = tok.encode("utf-8")
            c = {"src": f"testdata/de/{name}:{ln}", "func": "GenderFromFile", "model": "tokenizer_de.matok",
                 "flags": 3, "input_hex": data.hex(), "checks": []}
            if kind == "dont":
                c["checks"] += [{"kind": "tokens_len", "eq": 1}, {"kind": "tokens", "idx": 0, "eq_hex": data.hex()}]
            else:
                c["checks"].append({"kind": "tokens_len_gt", "eq": 1})
            cases.append(c)
    return cases


def main():
    cases = []
    skipped = []
    for f in ("matrix_test.go", "token_writer_test.go"):
        ex = Extractor(f)
        cases += ex.run()
        skipped += ex.skipped
    # testdata/de/{dontsplit,split}.txt are NOT included: the reference runs them
    # against tokenizer_de.datok only (datok_test.go:1201-1236).  The shipped
    # tokenizer_de.matok predates the gender-form grammar of 0.3.1 (Changes:4) and
    # splits 40 of the 46 "dontsplit" forms, so those files pin a different model.
    cases = [c for c in cases if c["checks"]]
    nchecks = sum(len(c["checks"]) for c in cases)
    json.dump({"reference": "KorAP/Datok 0.3.1", "generator": "tests/golden/make_golden.py",
               "cases": cases, "skipped": skipped}, open(OUT, "w"), indent=0, ensure_ascii=True)
    print(f"{len(cases)} cases, {nchecks} checks -> {OUT}")
    for s in skipped:
        print("skipped:", s)
    # ---- the double-array path (datok.go:781-1135): datok_test.go on the shipped .datok models, plus
    # testdata/de/{dontsplit,split}.txt, which the reference runs against tokenizer_de.datok
    ex = Extractor("datok_test.go", double_array=True)
    da = [c for c in ex.run() if c["checks"] and str(c.get("model", "")).endswith(".datok")]
    for c in file_list_cases():
        c["model"] = "tokenizer_de.datok"
        da.append(c)
    ncheck_da = sum(len(c["checks"]) for c in da)
    json.dump({"reference": "KorAP/Datok 0.3.1", "generator": "tests/golden/make_golden.py",
               "cases": da, "skipped": ex.skipped}, open(OUT_DA, "w"), indent=0, ensure_ascii=True)
    print(f"{len(da)} double-array cases, {ncheck_da} checks -> {OUT_DA}")
    for s in ex.skipped:
        print("skipped (double array):", s)
    # ---- the compile path (fomafile.go:56-450 + matrix.go:30-99): the cases of matrix_test.go that build their
    # model with LoadFomaFile(x).ToMatrix(), and those of datok_test.go that build it with .ToDoubleArray() from
    # the same kind of file and hold no EOT (the two walks only differ at an EOT, datok.go:1019-1030): both pin
    # ParseFoma, the first also ToMatrix.
    fo = []
    fskip = []
    for f, da in (("matrix_test.go", False), ("datok_test.go", True)):
        ex = Extractor(f, double_array=da, foma=True)
        for c in ex.run():
            if not c["checks"] or not str(c.get("model", "")).endswith(".fst") or c["model"] not in FOMA_FIXTURES:
                continue
            if da and b"\x04" in bytes.fromhex(c["input_hex"]):
                fskip.append(c["src"] + " (double-array case with an EOT)")
                continue
            c["via"] = "ToDoubleArray" if da else "ToMatrix"
            fo.append(c)
    json.dump({"reference": "KorAP/Datok 0.3.1", "generator": "tests/golden/make_golden.py",
               "cases": fo, "skipped": fskip}, open(OUT_FOMA, "w"), indent=0, ensure_ascii=True)
    print(f"{len(fo)} compile-path cases, {sum(len(c['checks']) for c in fo)} checks -> {OUT_FOMA}")


if __name__ == "__main__":
    sys.exit(main())
