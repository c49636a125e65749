/*
 * datok_oracle.h -- CPU ORACLE for the Datok matrix-FSA transduction path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithm (KorAP/Datok 0.3.1):
 *     matrix.go:235-337   ParseMatrix            (.matok loader)
 *     matrix.go:348-698   TransduceTokenWriter   (greedy walk, single backtrack)
 *     token_writer.go:36-175 NewTokenWriter      (flag-driven formatter)
 *     fomafile.go:56-450  LoadFomaFile/ParseFoma, matrix.go:30-99 ToMatrix, matrix.go:126-210 WriteTo
 *                         (the compile path foma -> .matok; pinned by tests/test_foma_compile.py on the
 *                         reference's own compiled artefacts and matrix_test.go's vectors)
 * and, for the rows SURVEY.md section 8f marks "next", of the double-array path:
 *     datok.go:621-729    ParseDatok             (.datok loader; ora_load dispatches on the magic like
 *                                                 LoadTokenizerFile, fomafile.go:452-484, and converts
 *                                                 the double array to the matrix's dense layout)
 *     datok.go:781-1135   TransduceTokenWriter   (the same loop; no buffer rewind at an EOT)
 * pinned by tests/test_oracle_golden_datok.py on datok_test.go's vectors.
 * It exists to CHECK the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * datok_b200/ links, imports or executes it.
 *
 * Parity pin: tests/test_oracle_golden.py runs this oracle against every
 * golden vector of the reference's own tests for this path
 * (matrix_test.go, token_writer_test.go, testdata/de/{dontsplit,split}.txt),
 * extracted by tests/golden/make_golden.py.  The Go reference itself cannot be
 * built here (no Go toolchain), so there is no oracle/_ref.
 */
#ifndef DATOK_ORACLE_H
#define DATOK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* token_writer.go:17-25 */
enum {
  ORA_TOKENS = 1,
  ORA_SENTENCES = 2,
  ORA_TOKEN_POS = 4,
  ORA_SENTENCE_POS = 8,
  ORA_NEWLINE_AFTER_EOT = 16,
  ORA_SIMPLE = 3,
  /* not in the reference: the TokenWriter passed in has already seen a Token
   * call (token_writer.go:42,70 `init` is false).  Lets a test replay a
   * writer that is reused across Transduce calls (token_writer_test.go:52). */
  ORA_WRITER_USED = 256
};

/* status codes: 0 = ok, >0 = the Go reference would have panicked */
enum {
  ORA_OK = 0,
  ORA_PANIC_BUFFER_OVERFLOW = 1, /* matrix.go:365,406 buffer[1024] */
  ORA_PANIC_SENT_NO_TOKEN = 2,   /* token_writer.go:108 pos[len(pos)-1] */
  ORA_PANIC_TEXT_NO_TOKEN = 3,   /* token_writer.go:135 pos[0] */
  ORA_PANIC_TEXT_NO_SENT = 4,    /* token_writer.go:145 sent[0] */
  ORA_PANIC_TOKEN_SLICE = 5,     /* token_writer.go:85 buf[offset:], offset>len */
  ORA_PANIC_EMPTY_BUF = 6,       /* token_writer.go:66 buf[0] on empty buf */
  ORA_ERR_LOOP = 7,              /* endless epsilon loop (matrix.go:633 TODO) */
  ORA_PANIC_INDEX = 8            /* matrix.go:463 array index out of range (symbol outside the matrix) */
};

typedef struct ora_model ora_model;

/* LoadMatrixFile matrix.go:214-231.  NULL on any error (reference: nil). */
ora_model *ora_load(const char *path);
void ora_free(ora_model *m);
/* LoadFomaFile(path).ToMatrix()  (fomafile.go:56-450, matrix.go:30-99): the compile path.  NULL where the
 * reference returns nil or panics. */
ora_model *ora_load_foma(const char *path);
/* WriteTo (matrix.go:126-210): the uncompressed MATOK image; malloc'd, release with ora_free_bytes */
uint8_t *ora_write_matrix(const ora_model *m, size_t *out_len);

/* model introspection (for tests and for checking the GPU re-layout) */
int ora_epsilon(const ora_model *m);
int ora_unknown(const ora_model *m);
int ora_identity(const ora_model *m);
int ora_state_count(const ora_model *m);
int ora_sigma_count(const ora_model *m);
const uint32_t *ora_array(const ora_model *m, size_t *n);
const int32_t *ora_sigma_ascii(const ora_model *m);
/* rune -> symbol as matrix.go:421-435 does it; *ok mirrors the map's `ok`
 * (untouched for rune < 256). */
int ora_sigma_lookup(const ora_model *m, int32_t rune, int *ok);

/* carry state (an extension for shard tests; zero-initialised = reference) */
typedef struct {
  uint32_t state;       /* t; 0 means "use the initial state 1" */
  int32_t ok;           /* sticky map-lookup result, matrix.go:352 */
  int32_t sentence_end; /* matrix.go:360 */
  int32_t text_end;     /* matrix.go:363 */
} ora_carry;

typedef struct {
  int status;
  /* formatted output exactly as the TokenWriter would have written it */
  uint8_t *text;
  size_t text_len;
  /* structured events, stream order */
  size_t n_tokens;
  uint32_t *tok_byte_start; /* surface start, absolute byte offset */
  uint32_t *tok_byte_end;   /* surface end (exclusive) */
  uint32_t *tok_buf_start;  /* byte offset of buf[0] of the Token call */
  int32_t *tok_offset;      /* `offset` argument of the Token call (runes) */
  size_t n_tok_pos;         /* entries in tok_pos (2 per token when a pos flag is set) */
  int32_t *tok_pos;         /* TokenWriter.pos entries over all texts */
  size_t n_sent_events;     /* SentenceEnd calls */
  uint64_t *sent_tok_idx;   /* tokens emitted before each SentenceEnd call */
  size_t n_sent_pos;        /* TokenWriter.sent entries over all texts */
  int32_t *sent_pos;
  size_t n_texts;           /* TextEnd calls */
  uint64_t *text_tok_end;   /* tokens emitted before each TextEnd */
  uint64_t *text_sent_end;  /* SentenceEnd calls before each TextEnd */
  uint64_t *text_sentpos_end; /* sent entries flushed up to each TextEnd */
  uint32_t *text_byte_end;  /* byte position of the walk at each TextEnd */
  ora_carry carry_out;
  /* statistics (design input, not part of parity) */
  uint64_t n_runes, n_iterations, n_backtracks, n_backtrack_runes, n_hardfail;
  uint32_t max_window;      /* longest buffer fill (runes) seen */
} ora_result;

/* TransduceTokenWriter(bytes.NewReader(in), NewTokenWriter(sink, flags))
 * matrix.go:348-698 + token_writer.go:36-175.  carry_in may be NULL. */
ora_result *ora_transduce(const ora_model *m, const uint8_t *in, size_t n,
                          uint32_t flags, const ora_carry *carry_in);
void ora_result_free(ora_result *r);

/* CPU baseline: tokenises every EOT-delimited document of `in` independently
 * ("one goroutine per document"), `nthreads` workers, each formatting into a
 * private in-memory sink.  Returns total tokens; *out_bytes = bytes formatted.
 * want_structs=0 skips the structured arrays (pure reference work). */
uint64_t ora_transduce_docs_mt(const ora_model *m, const uint8_t *in, size_t n,
                               uint32_t flags, int nthreads,
                               uint64_t *out_bytes, uint64_t *out_sentences,
                               uint64_t *out_docs);

/* TokenWriter restatement alone; ops: 0,offset,len,rune.. = Token; 1 = SentenceEnd;
 * 2 = TextEnd.  Returns malloc'd output (ora_free_bytes). */
uint8_t *ora_token_writer_replay(uint32_t flags, const int32_t *ops, size_t nops, size_t *out_len,
                                 int *status);
void ora_free_bytes(uint8_t *p);

/* design input: visits per source state of the transition lookup (hist has stateCount+1 slots) */
int ora_state_histogram(const ora_model *m, const uint8_t *in, size_t n, uint64_t *hist);

/* Go's unicode/utf8.DecodeRune: returns rune, *width in bytes (0 only if n==0) */
int32_t ora_decode_rune(const uint8_t *p, size_t n, int *width);

#ifdef __cplusplus
}
#endif
#endif
