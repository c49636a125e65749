"""datok_b200 -- B200-native (sm_100a) implementation of KorAP/Datok's matrix-FSA
transduction path (MatrixTokenizer.TransduceTokenWriter over .matok models).

Public surface mirrors the reference's Go API for that path:

    tok = LoadTokenizerFile("tokenizer_de.matok")          # fomafile.go:452
    tok.Transduce(reader, writer)                           # matrix.go:340
    tok.TransduceTokenWriter(reader, NewTokenWriter(w, TOKENS | SENTENCES | TOKEN_POS))
    tok.Type() == "MATOK"
    mat = LoadFomaFile("tokenizer.fst").ToMatrix(); mat.Save("tokenizer.matok")   # fomafile.go:56, matrix.go:30,107

plus the offset-array API (`transduce_arrays`) the C ABI is built around.
"""
from ._lib import (COMPACT, COMPACT8, FORMAT, NEWLINE_AFTER_EOT, NOT_FINAL, SENTENCE_POS, SENTENCES, SIMPLE, TOKEN_POS, TOKENS, WRITER_USED, Carry)
from .tokenizer import (Automaton, DatokError, LoadFomaFile, LoadMatrixFile, convert, LoadTokenizerFile, MatrixTokenizer, NewTokenWriter,
                        ReferencePanic, Result, TokenWriter, transduce_sharded)

__all__ = ["LoadTokenizerFile", "LoadMatrixFile", "LoadFomaFile", "Automaton", "convert", "MatrixTokenizer", "NewTokenWriter", "TokenWriter", "Result",
           "TOKENS", "SENTENCES", "TOKEN_POS", "SENTENCE_POS", "NEWLINE_AFTER_EOT", "SIMPLE", "WRITER_USED", "NOT_FINAL", "COMPACT", "COMPACT8", "FORMAT",
           "Carry", "DatokError", "ReferencePanic", "transduce_sharded"]
