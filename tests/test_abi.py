"""The C-ABI library loads on a CPU-only box and exports every symbol include/datok_b200.h
declares.  No compute calls here (those are the -m gpu tests)."""
import ctypes as C
import gzip
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "datok_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(datok_[a-z_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("datok_load", "datok_transduce", "datok_transduce_device", "datok_result_view", "datok_result_free",
              "datok_free", "datok_type", "datok_format", "datok_replay"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from datok_b200 import _lib
    L = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert missing == []
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_type_is_matok():
    from datok_b200 import _lib
    assert _lib.lib().datok_type() == b"MATOK"  # matrix.go:102-104


def test_flag_values_match_reference():
    import datok_b200 as d
    # token_writer.go:17-25
    assert (d.TOKENS, d.SENTENCES, d.TOKEN_POS, d.SENTENCE_POS, d.NEWLINE_AFTER_EOT, d.SIMPLE) == (1, 2, 4, 8, 16, 3)


def test_loader_errors_mirror_reference(tmp_path, testdata):
    """LoadTokenizerFile returns nil on any error (fomafile.go:452-484, matrix.go:214-337)."""
    from datok_b200 import _lib
    from datok_b200.tokenizer import LoadTokenizerFile, load_error_code
    assert LoadTokenizerFile(str(tmp_path / "nope.matok")) is None
    assert load_error_code(str(tmp_path / "nope.matok")) == _lib.ERR_IO
    p = tmp_path / "plain.matok"
    p.write_bytes(b"MATOK not gzipped")
    assert load_error_code(str(p)) == _lib.ERR_IO           # gzip.NewReader fails, matrix.go:222
    p = tmp_path / "magic.matok"
    p.write_bytes(gzip.compress(b"DATOK" + b"\0" * 64))
    assert load_error_code(str(p)) == _lib.ERR_FORMAT       # the double-array magic with an empty array (datok.go:667-725)
    p.write_bytes(gzip.compress(b"MATOX" + b"\0" * 64))
    assert load_error_code(str(p)) == _lib.ERR_FORMAT       # matrix.go:258
    raw = gzip.decompress(open(os.path.join(testdata, "simpletok.matok"), "rb").read())
    p = tmp_path / "version.matok"
    p.write_bytes(gzip.compress(raw[:5] + b"\x02\x00" + raw[7:]))
    assert load_error_code(str(p)) == _lib.ERR_FORMAT       # matrix.go:276
    p = tmp_path / "short.matok"
    p.write_bytes(gzip.compress(raw[:-8]))
    assert load_error_code(str(p)) == _lib.ERR_FORMAT       # matrix.go:327
    assert len(raw) == 230                                   # matrix_test.go:167 pins the image size


def test_no_cpu_fallback(testdata):
    """Without a CUDA device the product refuses to load a model instead of running on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from datok_b200 import _lib
    from datok_b200.tokenizer import load_error_code
    assert load_error_code(os.path.join(testdata, "tokenizer_de.matok")) == _lib.ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """Nothing under datok_b200/ may import, link or execute oracle/ or tests/emul."""
    pkg = os.path.join(ROOT, "datok_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".c", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "datok_oracle" not in txt and "libdatok_emul" not in txt, f


def test_double_array_models_load_by_magic(testdata, monkeypatch, tmp_path):
    """LoadTokenizerFile (fomafile.go:452-484) dispatches on the magic: .datok files load like .matok files (the loader
    converts the double array, datok.go:621-729, to the matrix layout; the kernels know that its walk does not rewind
    the buffer at an EOT)."""
    import torch
    from datok_b200 import _lib
    from datok_b200.tokenizer import load_error_code
    p = tmp_path / "empty.datok"
    p.write_bytes(gzip.compress(b"DATOK" + b"\0" * 64))
    assert load_error_code(str(p)) == _lib.ERR_FORMAT       # an empty double array (datok.go:703-725)
    if not torch.cuda.is_available():  # parsed; the load then stops at the device, like a .matok load
        assert load_error_code(os.path.join(testdata, "tokenizer_de.datok")) == _lib.ERR_NO_DEVICE
        assert load_error_code(os.path.join(testdata, "simpletok.datok")) == _lib.ERR_NO_DEVICE
