/*
 * datok_b200.h -- C ABI of the B200-native Datok matrix-FSA transduction path.
 *
 * Drop-in boundary for KorAP/Datok's `Tokenizer` interface (fomafile.go:29-33)
 * restricted to MatrixTokenizer (.matok models).  The reference has no FFI layer;
 * a Go shim type implementing `Tokenizer` binds these symbols through cgo
 * (see INTEGRATION.md and go/datokb200/).  Plain pointers and sizes only.
 *
 * Every transduction runs on the GPU (hand-written sm_100a kernels).  There is no
 * CPU fallback: without a usable CUDA device datok_load() fails with
 * DATOK_ERR_NO_DEVICE.
 */
#ifndef DATOK_B200_H
#define DATOK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TokenWriter flag bits -- same values as token_writer.go:17-25 (type Bits). */
enum {
  DATOK_TOKENS = 1,
  DATOK_SENTENCES = 2,
  DATOK_TOKEN_POS = 4,
  DATOK_SENTENCE_POS = 8,
  DATOK_NEWLINE_AFTER_EOT = 16,
  DATOK_SIMPLE = 3,
  /* Not a reference flag: the caller's TokenWriter has already received a Token
   * call (token_writer.go:42,70: `init` is false), i.e. the writer is being
   * reused across Transduce calls as in token_writer_test.go:52-56. */
  DATOK_WRITER_USED = 256,
  /* Not a reference flag: the input is a shard of a longer stream that continues in a
   * later call (another GPU, the next batch).  End-of-input processing (matrix.go:650-695:
   * final token flush, final SentenceEnd/TextEnd) is skipped; the input must end at a
   * text boundary (right after an EOT), else DATOK_ERR_NOT_AT_BOUNDARY. */
  DATOK_NOT_FINAL = 512,
  /* Not a reference flag: token spans come back delta-encoded (view.tok_delta, 8 bytes per token
   * instead of 16 over PCIe) in place of tok_bytes / tok_pos; see datok_view.tok_delta.  The
   * TokenWriter itself only ever sees such deltas: Token(offset, buf) gets the runes since the
   * previous token end (token_writer.go:59-95). */
  DATOK_COMPACT = 1024,
  /* Not a reference flag: like DATOK_COMPACT with one byte per value (view.tok_delta8, 4 bytes per
   * token over PCIe).  A value that does not fit is stored as 255 and listed in view.tok_esc. */
  DATOK_COMPACT8 = 2048,
  /* Not a reference flag: the result carries the text NewTokenWriter(w, flags) writes (view.text, view.text_len),
   * formatted on the device (token_writer.go:59-167 as parallel writes, format_core.cuh) and copied back in one
   * piece, instead of the token / sentence arrays (NULL then; the per-text bounds are still there).  The output of
   * Transduce / TransduceTokenWriter with a stock writer is exactly these bytes. */
  DATOK_FORMAT = 4096
};

/* Error codes.  1..6 mirror inputs on which the Go reference panics (they are
 * outside the parity domain and are reported, never undefined behaviour). */
enum {
  DATOK_OK = 0,
  DATOK_ERR_BUFFER_OVERFLOW = 1, /* >1024 runes without a token boundary (matrix.go:365,406) */
  DATOK_ERR_SENT_NO_TOKEN = 2,   /* SentenceEnd before any token of the text, SENTENCE_POS (token_writer.go:108) */
  DATOK_ERR_TEXT_NO_TOKEN = 3,   /* TextEnd on a token-less text, TOKEN_POS (token_writer.go:135) */
  DATOK_ERR_TEXT_NO_SENT = 4,    /* TextEnd on a sentence-less text, SENTENCE_POS (token_writer.go:145) */
  DATOK_ERR_DEGENERATE = 5,      /* empty/negative token slice or repeated SentenceEnd at one position:
                                    cannot happen with the shipped models */
  DATOK_ERR_IO = 16,             /* file cannot be read / not gzip */
  DATOK_ERR_FORMAT = 17,         /* not a MATOK v1 image (matrix.go:258,276,312,327) */
  DATOK_ERR_UNSUPPORTED_MODEL = 18, /* >= 32768 states, > 253 symbol classes, non-empty unknown column, ... */
  DATOK_ERR_NO_DEVICE = 19,      /* no CUDA device / wrong architecture */
  DATOK_ERR_CUDA = 20,           /* a CUDA call failed; see datok_last_error() */
  DATOK_ERR_TOO_LARGE = 21,      /* input >= 2^32 - 2^20 bytes in one call (split at EOT) */
  DATOK_ERR_INVALID_ARG = 22,
  DATOK_ERR_NOT_AT_BOUNDARY = 23, /* DATOK_NOT_FINAL input ends inside a token / pending epsilon point */
  DATOK_ERR_COMPACT_RANGE = 24    /* DATOK_COMPACT: a delta does not fit 16 bits (ask for the absolute arrays) */
};

typedef struct datok_model datok_model;
typedef struct datok_result datok_result;

/* Walk state carried between byte-adjacent calls / shards (zero = stream start).
 * `state` uses the reference's state numbering (matrix.go:351: initial state 1). */
typedef struct {
  uint32_t state;        /* 0 = initial state 1 */
  uint32_t sentence_end; /* matrix.go:360 */
  uint32_t text_end;     /* matrix.go:363 */
  uint32_t reserved;
} datok_carry;

/* Flat, host-resident view of one transduction.  All arrays live in pinned host
 * memory owned by the result; they stay valid until datok_result_free().
 * Arrays whose flag was not requested are NULL (see datok_transduce). */
typedef struct {
  uint64_t n_tokens;       /* Token events                      (matrix.go:528,569,675) */
  uint64_t n_sentences;    /* SentenceEnd events                (matrix.go:575,597,684) */
  uint64_t n_texts;        /* TextEnd events                    (matrix.go:600,691)     */
  uint64_t n_sent_pos;     /* entries of the TokenWriter's `sent` list over all texts   */
  uint64_t n_runes;        /* runes the reference's ReadRune would have produced        */
  /* per token k: surface = in[tok_bytes[2k] .. tok_bytes[2k+1])          (DATOK_TOKENS) */
  const uint32_t *tok_bytes;
  /* per token k: TokenWriter.pos entries (text-relative rune offsets, after the
   * NEWLINE_AFTER_EOT shift): start = tok_pos[2k], end = tok_pos[2k+1] (DATOK_TOKEN_POS) */
  const int32_t *tok_pos;
  /* TokenWriter.sent entries, flat over all texts                  (DATOK_SENTENCE_POS) */
  const int32_t *sent_pos;
  /* per SentenceEnd event: number of tokens emitted before it        (DATOK_SENTENCES) */
  const uint32_t *sent_tok;
  /* per TextEnd event d (always present): exclusive prefix bounds of text d */
  const uint32_t *text_tok_end;     /* tokens emitted up to TextEnd d          */
  const uint32_t *text_sent_end;    /* SentenceEnd events up to TextEnd d      */
  const uint32_t *text_sentpos_end; /* `sent` entries flushed up to TextEnd d  */
  const uint32_t *text_byte_end;    /* byte position of the walk at TextEnd d  */
  datok_carry carry_out;
  uint32_t has_invalid_utf8; /* some input byte decodes to U+FFFD (surface re-encoding needed) */
  /* timing of the last call, milliseconds (CUDA events on the call's stream) */
  float ms_h2d, ms_kernels, ms_d2h;
  /* DATOK_COMPACT: per token k four 16-bit values, relative to the end of the previous token of the
   * same text (or to the start of the text: byte text_byte_end[d-1], rune 0):
   *   tok_delta[4k+0] bytes skipped before the token     tok_delta[4k+1] bytes of the token
   *   tok_delta[4k+2] runes skipped before the token     tok_delta[4k+3] runes of the token
   * (the rune skip of a text's first token already includes the NEWLINE_AFTER_EOT shift).
   * tok_bytes / tok_pos are NULL then; datok_expand() rebuilds them. */
  const uint16_t *tok_delta;
  /* DATOK_COMPACT8: the same four values per token as one byte each.  255 means "look it up":
   * tok_esc holds n_esc pairs {token index, field << 16 | value}, sorted by token index and field. */
  const uint8_t *tok_delta8;
  const uint32_t *tok_esc;
  uint64_t n_esc;
  /* DATOK_FORMAT: the formatted text (pinned host memory; a device pointer from datok_transduce_device) */
  const uint8_t *text;
  uint64_t text_len;
} datok_view;

/* LoadTokenizerFile (fomafile.go:452-484) for the MATOK magic / LoadMatrixFile
 * (matrix.go:214-231): gunzip, ParseMatrix (matrix.go:235-337), build the GPU
 * layout on `device`.  NULL on error (the reference returns nil), *err says why. */
datok_model *datok_load(const char *path, int device, int *err);
/* same, from an in-memory gunzipped MATOK image (ParseMatrix, matrix.go:235) */
datok_model *datok_load_image(const uint8_t *image, size_t n, int device, int *err);
/* The compile path.  LoadFomaFile(path).ToMatrix() (fomafile.go:56-450, matrix.go:30-99): parses a gzipped foma
 * file that follows the tokenizer's conventions and builds the matrix model in memory, ready on `device`.
 * (A model without @_IDENTITY_SYMBOL_@ only exists in this in-memory form, as in the reference.) */
datok_model *datok_load_foma(const char *path, int device, int *err);
/* `datok convert -i foma_path -o matok_path` (cmd/datok.go:63: LoadFomaFile, ToMatrix, Save).  Host only: no
 * device is touched.  0 or a DATOK_ERR_* code (datok_last_error() has the reference's message). */
int datok_compile_foma(const char *foma_path, const char *matok_path);
/* MatrixTokenizer.Save (matrix.go:107-123): gzip(WriteTo).  DATOK_ERR_INVALID_ARG for a double-array model. */
int datok_save(const datok_model *m, const char *path);
/* MatrixTokenizer.WriteTo (matrix.go:126-210): the uncompressed image.  Returns its size; written to dst when
 * cap is large enough.  0 on error. */
size_t datok_write_image(const datok_model *m, uint8_t *dst, size_t cap);
void datok_free(datok_model *m);

/* Tokenizer.Type() (matrix.go:102-104) -> "MATOK" */
const char *datok_type(void);
/* Tokenizer.Type() of a loaded model: "MATOK" (matrix.go:102-104) or, for a double-array file,
 * "DATOK" (datok.go:252-254) -- LoadTokenizerFile dispatches on the magic (fomafile.go:476-480). */
const char *datok_model_type(const datok_model *m);

/* model introspection (reference numbering) */
int datok_model_info(const datok_model *m, uint32_t *state_count, uint32_t *sigma_count,
                     uint32_t *n_classes, uint32_t *epsilon, uint32_t *unknown, uint32_t *identity);

/* TransduceTokenWriter (matrix.go:348-698) over `n` bytes of host memory.
 * `flags` are TokenWriter Bits (plus DATOK_WRITER_USED, DATOK_NOT_FINAL, DATOK_COMPACT,
 * DATOK_COMPACT8); they select which arrays are produced and copied back.  carry_in may be
 * NULL (stream start).  Inputs of 128 MiB and more are cut after EOT bytes into pieces whose
 * host<->device copies overlap the kernels; the result is the same single set of arrays.
 * Synchronous; calls on one model are serialised. Returns DATOK_OK or an error;
 * on reference-panic inputs (codes 1..5) *out is still a valid, partial result
 * holding everything up to the failing event is NOT guaranteed -- treat as failed. */
int datok_transduce(datok_model *m, const uint8_t *in, size_t n, uint32_t flags,
                    const datok_carry *carry_in, datok_result **out);

/* Same, but `d_in` is a DEVICE pointer (input already resident in HBM) and the
 * offset arrays stay on the device; *view then holds device pointers (under
 * DATOK_COMPACT8 the escape pairs are left in the order the kernel appended them).
 * Used by the benchmark's kernel-only leg and by callers that post-process on the GPU. */
int datok_transduce_device(datok_model *m, const uint8_t *d_in, size_t n, uint32_t flags,
                           const datok_carry *carry_in, datok_result **out);

const datok_view *datok_result_view(const datok_result *r);
void datok_result_free(datok_result *r);

/* ---- streaming front-end: the reference reads any io.Reader (matrix.go:373,388-408; cmd/datok.go:108-132 incl.
 * STDIN).  The caller pushes blocks as they arrive; a push transduces everything up to the last EOT seen so far (a
 * text boundary, matrix.go:593-605) and hands back that batch's result (*out is NULL while no text has ended yet);
 * finish() transduces the rest with the end-of-input processing (matrix.go:650-695).  The walk state, sentenceEnd /
 * textEnd and the writer's `init` flag travel from batch to batch inside the stream; `flags` are those of
 * datok_transduce (DATOK_FORMAT: every result carries its part of the TokenWriter's text, to be written in order). */
typedef struct datok_stream datok_stream;
datok_stream *datok_stream_open(datok_model *m, uint32_t flags);
int datok_stream_push(datok_stream *s, const uint8_t *data, size_t n, datok_result **out);
int datok_stream_finish(datok_stream *s, datok_result **out);
uint64_t datok_stream_bytes_done(const datok_stream *s);
void datok_stream_close(datok_stream *s);

/* ---- one corpus over the GPUs of a box.  models[i] is the same model file loaded on devices[i].
 * The input is cut into ndev byte-balanced shards that end right after an EOT (datok_plan_shards; bounds has
 * n_shards + 1 entries), shard i is transduced on device i -- all concurrently, each from the guess "root state,
 * behind a finished text, a token has been seen" --, the per-shard counts {bytes, tokens, sentences, texts, sent
 * entries} and carry-out states are all-gathered over NCCL (NVLink / NVSwitch; bound at run time), and a shard whose
 * guess turns out wrong (matrix.go:593-605: the state behind an EOT is whatever the matrix says) is transduced again
 * from the true carry.  outs[i] is shard i's result with shard-relative indices and byte offsets; bases (5 per
 * shard, may be NULL) holds what to add: {bytes, tokens, sentences, texts, sent entries} before the shard. */
int datok_plan_shards(const uint8_t *in, size_t n, int n_shards, uint64_t *bounds);
int datok_transduce_sharded(datok_model *const *models, const int *devices, int ndev, const uint8_t *in, size_t n,
                            uint32_t flags, const datok_carry *carry_in, datok_result **outs, uint64_t *bases,
                            uint64_t *bounds_out);
const char *datok_sharded_last_error(void);
/* of the last sharded call on this thread: whether the exchange went through NCCL, how many shards were redone */
int datok_sharded_last_info(int *used_nccl, int *shards_rewalked);

/* Rebuilds the absolute token arrays of a DATOK_COMPACT / DATOK_COMPACT8 result on the host: tok_bytes (2 per token)
 * and / or tok_pos (2 per token); either may be NULL. */
int datok_expand(const datok_result *r, uint32_t *tok_bytes, int32_t *tok_pos);

/* Host half of the TokenWriter (token_writer.go:36-175): formats a host-resident
 * result exactly as NewTokenWriter(w, flags) would have written it.  Returns the
 * number of bytes needed; writes at most `cap` bytes to `dst`. */
size_t datok_format(const datok_result *r, const uint8_t *in, size_t n, uint32_t flags,
                    uint8_t *dst, size_t cap);

/* Replays one result into caller-supplied TokenWriter callbacks in stream order
 * (custom TokenWriters, token_writer.go:27-33).  token(): `buf` points at the
 * input bytes of the Token call's rune buffer, `offset_bytes` at the surface. */
typedef struct {
  void *user;
  void (*token)(void *user, const uint8_t *buf, size_t buf_bytes, size_t offset_bytes, int32_t offset_runes);
  void (*sentence_end)(void *user);
  void (*text_end)(void *user);
} datok_callbacks;
int datok_replay(const datok_result *r, const uint8_t *in, size_t n, const datok_callbacks *cb);

/* per-kernel device times (ms) of the last call on this model, for bench.py.
 * names: "classify","walk","stitch","rewalk","compact_reduce","compact_scan","compact_emit" */
int datok_last_kernel_times(const datok_model *m, const char **names, float *ms, int cap);
/* number of kernel launches issued by the last call */
int datok_last_launch_count(const datok_model *m);
/* of the last call: fix-up rounds of the speculative walk; of the model's current layout: table rows and class
 * columns resident in shared memory, bytes per speculative chunk (any pointer may be NULL) */
int datok_last_stats(const datok_model *m, uint32_t *fixup_rounds, uint32_t *hot_rows, uint32_t *hot_cols,
                     uint32_t *chunk_bytes);

/* Measurement only (bench.py, SURVEY.md 8d "R_gather"): byte steps per second of the bare dependent
 * shared-memory gather chain of the walk (class byte -> row entry -> next state) in the walk's own
 * configuration on this device: the bound of any lane-per-chunk table walk of this shape. */
int datok_measure_gather_bound(datok_model *m, double *byte_steps_per_s);

/* Page-locked host memory for inputs: host->device copies from it run at full
 * PCIe speed and asynchronously.  Any other host pointer works too, slower. */
void *datok_host_alloc(size_t bytes);
void datok_host_free(void *p);

const char *datok_last_error(void);
const char *datok_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif
