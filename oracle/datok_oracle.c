/*
 * datok_oracle.c -- CPU ORACLE (test infrastructure, see datok_oracle.h).
 *
 * Restates, in plain C, the Go reference KorAP/Datok 0.3.1:
 *   ParseMatrix            matrix.go:235-337
 *   TransduceTokenWriter   matrix.go:348-698
 *   NewTokenWriter         token_writer.go:36-175
 *   utf8.DecodeRune        Go stdlib (borrowed semantics of bufio.ReadRune)
 * Variable names follow the reference so the two can be read side by side.
 * Parity pin: tests/test_oracle_golden.py (reference golden vectors).
 */
#define _GNU_SOURCE
#include "datok_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define FIRSTBIT 0x80000000u /* datok.go:43 */
#define EOT 4                /* matrix.go:13 */
#define VERSION 1            /* datok.go:39 */
#define BUFSZ 1024           /* matrix.go:365 */
#define RUNE_ERROR 0xFFFD

/* ------------------------------------------------------------------ utf8 */

/* Go unicode/utf8: first[] classes and accept ranges. */
static void utf8_class(uint8_t b, int *size, uint8_t *lo, uint8_t *hi) {
  *lo = 0x80;
  *hi = 0xBF;
  if (b < 0x80) { *size = 1; return; }
  if (b < 0xC2) { *size = 0; return; }           /* xx: invalid lead */
  if (b < 0xE0) { *size = 2; return; }
  if (b == 0xE0) { *size = 3; *lo = 0xA0; return; }
  if (b == 0xED) { *size = 3; *hi = 0x9F; return; }
  if (b < 0xF0) { *size = 3; return; }
  if (b == 0xF0) { *size = 4; *lo = 0x90; return; }
  if (b < 0xF4) { *size = 4; return; }
  if (b == 0xF4) { *size = 4; *hi = 0x8F; return; }
  *size = 0;
}

int32_t ora_decode_rune(const uint8_t *p, size_t n, int *width) {
  if (n < 1) { *width = 0; return RUNE_ERROR; }
  uint8_t p0 = p[0];
  int sz; uint8_t lo, hi;
  utf8_class(p0, &sz, &lo, &hi);
  if (sz == 1) { *width = 1; return p0; }
  *width = 1;
  if (sz == 0) return RUNE_ERROR;
  if (n < (size_t)sz) return RUNE_ERROR;
  uint8_t b1 = p[1];
  if (b1 < lo || hi < b1) return RUNE_ERROR;
  if (sz == 2) { *width = 2; return ((int32_t)(p0 & 0x1F) << 6) | (b1 & 0x3F); }
  uint8_t b2 = p[2];
  if (b2 < 0x80 || 0xBF < b2) return RUNE_ERROR;
  if (sz == 3) {
    *width = 3;
    return ((int32_t)(p0 & 0x0F) << 12) | ((int32_t)(b1 & 0x3F) << 6) | (b2 & 0x3F);
  }
  uint8_t b3 = p[3];
  if (b3 < 0x80 || 0xBF < b3) return RUNE_ERROR;
  *width = 4;
  return ((int32_t)(p0 & 0x07) << 18) | ((int32_t)(b1 & 0x3F) << 12) |
         ((int32_t)(b2 & 0x3F) << 6) | (b3 & 0x3F);
}

/* Go string(rune): invalid runes become U+FFFD */
static int encode_rune(int32_t r, uint8_t *out) {
  if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = RUNE_ERROR;
  if (r < 0x80) { out[0] = (uint8_t)r; return 1; }
  if (r < 0x800) { out[0] = 0xC0 | (r >> 6); out[1] = 0x80 | (r & 0x3F); return 2; }
  if (r < 0x10000) {
    out[0] = 0xE0 | (r >> 12); out[1] = 0x80 | ((r >> 6) & 0x3F); out[2] = 0x80 | (r & 0x3F);
    return 3;
  }
  out[0] = 0xF0 | (r >> 18); out[1] = 0x80 | ((r >> 12) & 0x3F);
  out[2] = 0x80 | ((r >> 6) & 0x3F); out[3] = 0x80 | (r & 0x3F);
  return 4;
}

/* -------------------------------------------------------- growable arrays */

#define VEC(T) struct { T *p; size_t n, cap; }
#define VPUSH(v, x) do { if ((v).n == (v).cap) { (v).cap = (v).cap ? (v).cap * 2 : 1024; \
      (v).p = realloc((v).p, (v).cap * sizeof(*(v).p)); } (v).p[(v).n++] = (x); } while (0)

/* ----------------------------------------------------------------- model */

struct ora_model {
  int epsilon, unknown, identity, stateCount, sigmaCount;
  int32_t sigmaASCII[256];
  /* sigma map[rune]int as an open-addressing table */
  uint32_t hmask;
  int32_t *hkey; /* -1 = empty */
  int32_t *hval;
  uint32_t *array;
  size_t arraySize;
  /* 1: MatrixTokenizer (matrix.go), 0: DaTokenizer (datok.go), converted to the same dense layout at
   * load time.  The two transduction loops differ in one statement: the double-array one does not
   * rewind the buffer at an EOT (datok.go:1019-1030 has no `rewindBuffer = true`, matrix.go:603 has). */
  int eot_rewind;
};

static inline uint32_t hash_rune(int32_t r) { return (uint32_t)r * 2654435761u; }

static void sigma_put(ora_model *m, int32_t r, int32_t v) {
  uint32_t i = (hash_rune(r) >> 7) & m->hmask;
  while (m->hkey[i] != -1 && m->hkey[i] != r) i = (i + 1) & m->hmask;
  m->hkey[i] = r;
  m->hval[i] = v;
}

static inline int sigma_get(const ora_model *m, int32_t r, int *ok) {
  uint32_t i = (hash_rune(r) >> 7) & m->hmask;
  while (m->hkey[i] != -1) {
    if (m->hkey[i] == r) { *ok = 1; return m->hval[i]; }
    i = (i + 1) & m->hmask;
  }
  *ok = 0;
  return 0; /* Go map miss yields the zero value */
}

static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t le32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* ParseMatrix matrix.go:235-337 over an in-memory (already gunzipped) image */
static ora_model *parse_matrix(const uint8_t *d, size_t n) {
  size_t p = 0;
  if (n < 5 || memcmp(d, "MATOK", 5) != 0) return NULL; /* :258 */
  p = 5;
  if (n - p < 14) return NULL; /* :263-272 */
  if (le16(d + p) != VERSION) return NULL; /* :274-279 */
  ora_model *m = (ora_model *)calloc(1, sizeof(*m));
  m->epsilon = le16(d + p + 2);
  m->unknown = le16(d + p + 4);
  m->identity = le16(d + p + 6);
  m->stateCount = (int)le32(d + p + 8);
  m->sigmaCount = le16(d + p + 12);
  p += 14;
  m->arraySize = ((size_t)m->stateCount + 1) * (size_t)m->sigmaCount; /* :286 */
  for (int i = 0; i < 256; i++) m->sigmaASCII[i] = m->identity; /* :289-293 (identity is never -1 after a u16 read) */
  uint32_t cap = 64;
  while (cap < (uint32_t)m->sigmaCount * 4u) cap <<= 1;
  m->hmask = cap - 1;
  m->hkey = (int32_t *)malloc(cap * sizeof(int32_t));
  m->hval = (int32_t *)malloc(cap * sizeof(int32_t));
  for (uint32_t i = 0; i < cap; i++) m->hkey[i] = -1;
  for (int x = 0; x < m->sigmaCount; x++) { /* :295-303 */
    int w;
    int32_t sym = ora_decode_rune(d + p, n - p, &w);
    if (w == 0) continue; /* err != nil (EOF) */
    p += (size_t)w;
    if (sym != 0) {
      if (sym < 256) m->sigmaASCII[sym] = x;
      sigma_put(m, sym, x);
    }
  }
  if (p >= n || d[p] != 'M') { ora_free(m); return NULL; } /* :305-315 */
  p++;
  if (n - p < m->arraySize * 4) { ora_free(m); return NULL; } /* :327-330 */
  m->array = (uint32_t *)malloc(m->arraySize * 4 + 4);
  for (size_t x = 0; x < m->arraySize; x++) m->array[x] = le32(d + p + x * 4); /* :332-334 */
  m->eot_rewind = 1;
  return m;
}

/* ParseDatok datok.go:621-729 over an in-memory image, followed by a conversion of the double array
 * into the dense layout of the matrix (same cell meaning: target state, FIRSTBIT = non-token target).
 * A transition of the double array (datok.go:888-902,1058-1066):
 *     t = base(t0) + a;  valid iff t <= check(array[1]) and check(array[t]) == t0;
 *     non-token iff check-word FIRSTBIT of array[t];  the state that follows is base(array[t]) if
 *     array[t] is "separate" (base-word FIRSTBIT: it points to its representative), else t.
 * States are renumbered densely in breadth-first order from state 1 (which stays 1); state identity
 * is only observable through transitions, so the walk is the same. */
#define DA_RESTBIT 0x3FFFFFFFu
static ora_model *parse_datok(const uint8_t *d, size_t n) {
  size_t p = 0;
  if (n < 5 || memcmp(d, "DATOK", 5) != 0) return NULL; /* :651 */
  p = 5;
  if (n - p < 16) return NULL; /* :656-665 */
  if (le16(d + p) != VERSION) return NULL; /* :667-672 */
  ora_model *m = (ora_model *)calloc(1, sizeof(*m));
  m->epsilon = le16(d + p + 2);
  m->unknown = le16(d + p + 4);
  m->identity = le16(d + p + 6);
  /* final = le16(d + p + 8): not used by the transduction */
  m->sigmaCount = le16(d + p + 10);
  const size_t daSize = (size_t)le32(d + p + 12) / 2; /* :681 "legacy support" */
  p += 16;
  for (int i = 0; i < 256; i++) m->sigmaASCII[i] = m->identity; /* :686-691 */
  uint32_t cap = 64;
  while (cap < (uint32_t)m->sigmaCount * 4u) cap <<= 1;
  m->hmask = cap - 1;
  m->hkey = (int32_t *)malloc(cap * sizeof(int32_t));
  m->hval = (int32_t *)malloc(cap * sizeof(int32_t));
  for (uint32_t i = 0; i < cap; i++) m->hkey[i] = -1;
  for (int x = 0; x < m->sigmaCount; x++) { /* :693-701 */
    int w;
    int32_t sym = ora_decode_rune(d + p, n - p, &w);
    if (w == 0) continue;
    p += (size_t)w;
    if (sym != 0) {
      if (sym < 256) m->sigmaASCII[sym] = x;
      sigma_put(m, sym, x);
    }
  }
  if (p >= n || d[p] != 'T') { ora_free(m); return NULL; } /* :703-713 */
  p++;
  if (n - p < daSize * 8 || daSize < 2) { ora_free(m); return NULL; } /* :722-725 */
  const uint8_t *da = d + p; /* entry x: base = le32(da + 8x), check = le32(da + 8x + 4) */
#define DA_BASE(x) (le32(da + 8 * (size_t)(x)))
#define DA_CHECK(x) (le32(da + 8 * (size_t)(x) + 4))
  const uint32_t maxIndex = DA_CHECK(1) & DA_RESTBIT; /* :891 dat.array[1].getCheck() */
  /* breadth-first renumbering */
  uint32_t *id = (uint32_t *)calloc(daSize, sizeof(uint32_t));
  uint32_t *queue = (uint32_t *)malloc(daSize * sizeof(uint32_t));
  size_t qh = 0, qt = 0;
  id[1] = 1;
  queue[qt++] = 1;
  int ok = 1;
  while (qh < qt && ok) {
    const uint32_t t0 = queue[qh++];
    const uint32_t b = DA_BASE(t0) & DA_RESTBIT;
    for (int a = 1; a < m->sigmaCount; a++) {
      const uint64_t t = (uint64_t)b + (uint64_t)a;
      if (t > maxIndex || t >= daSize) continue;
      if ((DA_CHECK(t) & DA_RESTBIT) != t0) continue;
      uint32_t nx = (uint32_t)t;
      if (DA_BASE(t) & FIRSTBIT) nx = DA_BASE(t) & DA_RESTBIT; /* representative */
      if (nx >= daSize || nx == 0) { ok = 0; break; }
      if (!id[nx]) { id[nx] = (uint32_t)qt + 1; queue[qt++] = nx; }
    }
  }
  if (!ok) { free(id); free(queue); ora_free(m); return NULL; }
  m->stateCount = (int)qt;
  m->arraySize = ((size_t)m->stateCount + 1) * (size_t)m->sigmaCount;
  m->array = (uint32_t *)calloc(m->arraySize + 1, 4);
  const size_t S = (size_t)m->stateCount;
  for (size_t k = 0; k < qt; k++) {
    const uint32_t t0 = queue[k];
    const uint32_t b = DA_BASE(t0) & DA_RESTBIT;
    for (int a = 1; a < m->sigmaCount; a++) {
      const uint64_t t = (uint64_t)b + (uint64_t)a;
      if (t > maxIndex || t >= daSize) continue;
      if ((DA_CHECK(t) & DA_RESTBIT) != t0) continue;
      uint32_t nx = (uint32_t)t;
      if (DA_BASE(t) & FIRSTBIT) nx = DA_BASE(t) & DA_RESTBIT;
      m->array[(size_t)(a - 1) * S + id[t0]] = id[nx] | ((DA_CHECK(t) & FIRSTBIT) ? FIRSTBIT : 0u);
    }
  }
#undef DA_BASE
#undef DA_CHECK
  free(id);
  free(queue);
  m->eot_rewind = 0;
  return m;
}

/* gzip.NewReader + read to the end (matrix.go:214-231, fomafile.go:56-72) */
static uint8_t *read_gz(const char *path, size_t *len_out) {
  gzFile f = gzopen(path, "rb");
  if (!f) return NULL;
  /* gzopen transparently reads non-gzip files; the reference's gzip.NewReader
   * rejects them (matrix.go:222-226).  Check the magic ourselves. */
  FILE *raw = fopen(path, "rb");
  unsigned char mg[2] = {0, 0};
  if (raw) { if (fread(mg, 1, 2, raw) != 2) mg[0] = 0; fclose(raw); }
  if (mg[0] != 0x1f || mg[1] != 0x8b) { gzclose(f); return NULL; }
  size_t cap = 1 << 20, len = 0;
  uint8_t *buf = (uint8_t *)malloc(cap);
  for (;;) {
    if (len == cap) { cap *= 2; buf = (uint8_t *)realloc(buf, cap); }
    int got = gzread(f, buf + len, (unsigned)(cap - len > (1u << 30) ? (1u << 30) : cap - len));
    if (got < 0) { free(buf); gzclose(f); return NULL; }
    if (got == 0) break;
    len += (size_t)got;
  }
  gzclose(f);
  *len_out = len;
  return buf;
}

/* LoadMatrixFile matrix.go:214-231 (gzip -> ParseMatrix) */
ora_model *ora_load(const char *path) {
  size_t len = 0;
  uint8_t *buf = read_gz(path, &len);
  if (!buf) return NULL;
  /* LoadTokenizerFile fomafile.go:452-484 dispatches on the magic */
  ora_model *m = (len >= 5 && memcmp(buf, "DATOK", 5) == 0) ? parse_datok(buf, len) : parse_matrix(buf, len);
  free(buf);
  return m;
}

/* ------------------------------------------------- foma -> matrix (compile path)
 * LoadFomaFile fomafile.go:56-72, ParseFoma fomafile.go:77-450, ToMatrix matrix.go:30-99.
 * The Automaton's transitions ([]map[int]*edge) are kept as one flat list of edges; a later edge of the
 * same (state, symbol) replaces the earlier one like the map assignment does. */
typedef struct { int state, alpha, end; uint8_t nontoken; } foma_edge;

static int split_fields(char *line, char **f, int maxf) { /* strings.Split(line, " ") */
  int n = 0;
  f[n++] = line;
  for (char *c = line; *c; c++)
    if (*c == ' ') { *c = 0; if (n < maxf) f[n++] = c + 1; else return maxf + 1; }
  return n;
}
static int go_atoi(const char *s, int *out) { /* strconv.Atoi: optional sign, digits only */
  const char *c = s;
  if (*c == '+' || *c == '-') c++;
  if (!*c) return 0;
  long long v = 0;
  for (; *c; c++) { if (*c < '0' || *c > '9') return 0; v = v * 10 + (*c - '0'); if (v > 0x7fffffffLL) return 0; }
  *out = (int)(s[0] == '-' ? -v : v);
  return 1;
}

static ora_model *parse_foma(uint8_t *d, size_t n) {
  enum { M_NONE0 = 0, M_PROPS = 1, M_SIGMA = 2, M_STATES = 3, M_NONE = 4 }; /* fomafile.go:14-19 */
  int epsilon = -1, unknown = -1, identity = -1, final_sym = -1, tokenend = -1; /* :83-88 */
  int sigmaCount = 0, stateCount = -1;
  int state = 0, inSym = 0, outSym = 0, end = 0, final = 0; /* :91, live across lines */
  int mode = M_NONE0;
  /* sigmaRev / sigmaMCS, indexed by symbol number */
  size_t symcap = 1024;
  int32_t *sigmaRev = (int32_t *)malloc(symcap * sizeof(int32_t)); /* -1: no entry */
  uint8_t *isMCS = (uint8_t *)calloc(symcap, 1);
  for (size_t i = 0; i < symcap; i++) sigmaRev[i] = -1;
  VEC(foma_edge) edges = {0};
  ora_model *m = NULL;
  size_t p = 0;
  while (p < n) {
    uint8_t *nl = (uint8_t *)memchr(d + p, '\n', n - p);
    if (!nl) break; /* ReadString hits io.EOF: the partial last line is dropped (:101-107) */
    char *line = (char *)(d + p);
    size_t len = (size_t)(nl - (d + p)); /* without the '\n' */
    p += len + 1;
    *nl = 0;
    if (len >= 2 && line[0] == '#' && line[1] == '#') { /* :111-136 */
      if (!strncmp(line, "##props##", 9)) mode = M_PROPS;
      else if (!strncmp(line, "##states##", 10)) { mode = M_STATES; sigmaCount++; final_sym = sigmaCount; }
      else if (!strncmp(line, "##sigma##", 9)) mode = M_SIGMA;
      else if (!strncmp(line, "##end##", 7)) mode = M_NONE;
      else if (strncmp(line, "##foma-net", 10) != 0) break; /* "Unknown input line": leaves the loop */
      continue;
    }
    if (mode == M_PROPS) { /* :141-186 */
      char *f[16];
      int nf = split_fields(line, f, 16);
      if (nf < 13) goto fail; /* elem[12] would panic */
      if (strcmp(f[6], "1") != 0) goto fail; /* deterministic */
      if (strcmp(f[9], "1") != 0) goto fail; /* epsilon free */
      int v;
      if (!go_atoi(f[1], &v)) goto fail; /* arccount */
      if (!go_atoi(f[2], &v)) goto fail;
      stateCount = v;
      continue;
    }
    if (mode == M_STATES) { /* :187-372 */
      char *f[8];
      int nf = split_fields(line, f, 6);
      int e[5] = {0, 0, 0, 0, 0};
      if (!strcmp(f[0], "-1")) continue; /* :190 */
      int bad = 0;
      for (int k = 0; k < nf && k < 5; k++)
        if (!go_atoi(f[k], &e[k])) { bad = 1; break; }
      if (bad) continue; /* `break` leaves the switch: the line is skipped */
      switch (nf) { /* :232-279 */
        case 5: state = e[0]; inSym = e[1]; outSym = e[2]; end = e[3]; final = e[4]; break;
        case 4:
          if (e[1] == -1) {
            state = e[0]; final = e[3];
            if (stateCount < 0 || state + 1 < 0 || state + 1 > stateCount) goto fail;
            /* final state without outgoing edges: transitions[state+1][final] = &edge{} -- an edge to
             * state 0, which ToMatrix stores as 0: kept only for its index check */
            if (final == 1) { foma_edge fe = {state + 1, final_sym, 0, 0}; VPUSH(edges, fe); }
            continue;
          }
          state = e[0]; inSym = e[1]; end = e[2]; final = e[3]; outSym = inSym;
          break;
        case 3: inSym = e[0]; outSym = e[1]; end = e[2]; break;
        case 2: inSym = e[0]; end = e[1]; outSym = inSym; break;
        default: break; /* no case: the values of the previous line stand */
      }
      int nontoken = 0, tok_end = 0;
      inSym++; outSym++; /* :287-288 */
      if (inSym != outSym) { /* :291-313 */
        if (outSym == tokenend && inSym == epsilon) tok_end = 1;
        else if (outSym == epsilon) nontoken = 1;
        else goto fail; /* "Unsupported transition" */
      } else if (inSym == tokenend) {
        continue; /* :314-316 */
      } else if (inSym == epsilon) {
        goto fail; /* :317-319 */
      } else if (inSym >= 0 && (size_t)inSym < symcap && isMCS[inSym]) {
        continue; /* :320-324 */
      }
      (void)tok_end;
      if (stateCount < 0 || state + 1 < 0 || state + 1 > stateCount) goto fail; /* transitions[state+1]: index out of range */
      if (inSym >= 0) { foma_edge fe = {state + 1, inSym, end + 1, (uint8_t)nontoken}; VPUSH(edges, fe); } /* :336-344 */
      if (final == 1) { foma_edge fe = {state + 1, final_sym, 0, 0}; VPUSH(edges, fe); } /* :347-351 */
      continue;
    }
    if (mode == M_SIGMA) { /* :374-443 */
      char *sp = strchr(line, ' '); /* strings.SplitN(line, " ", 2) */
      if (!sp) goto fail; /* elem[1] would panic */
      *sp = 0;
      const char *sym = sp + 1;
      size_t symlen = len - (size_t)(sym - line);
      int number;
      if (!go_atoi(line, &number)) goto fail;
      number++;
      if (number < 0) goto fail;
      sigmaCount = number;
      if ((size_t)number >= symcap) {
        size_t nc = symcap;
        while ((size_t)number >= nc) nc *= 2;
        sigmaRev = (int32_t *)realloc(sigmaRev, nc * sizeof(int32_t));
        isMCS = (uint8_t *)realloc(isMCS, nc);
        for (size_t i = symcap; i < nc; i++) { sigmaRev[i] = -1; isMCS[i] = 0; }
        symcap = nc;
      }
      /* utf8.RuneCountInString */
      size_t nr = 0, q = 0;
      int32_t first = 0;
      while (q < symlen) {
        int w;
        int32_t r = ora_decode_rune((const uint8_t *)sym + q, symlen - q, &w);
        if (nr == 0) first = r;
        nr++; q += (size_t)w;
      }
      int32_t symbol;
      if (nr == 1) symbol = first;
      else if (nr > 1) {
        if (!strcmp(sym, "@_EPSILON_SYMBOL_@")) epsilon = number;
        else if (!strcmp(sym, "@_UNKNOWN_SYMBOL_@")) unknown = number;
        else if (!strcmp(sym, "@_IDENTITY_SYMBOL_@")) identity = number;
        else if (!strcmp(sym, "@_TOKEN_SYMBOL_@") || !strcmp(sym, "@_TOKEN_BOUND_@")) tokenend = number;
        else isMCS[number] = 1;
        continue;
      } else { /* the symbol is the line feed: its line ends right after the blank (:425-439) */
        if (p >= n) goto fail;
        uint8_t *nl2 = (uint8_t *)memchr(d + p, '\n', n - p);
        if (!nl2) goto fail;
        size_t l2 = (size_t)(nl2 - (d + p)) + 1;
        p += l2;
        if (l2 != 1) { isMCS[number] = 1; continue; }
        symbol = '\n';
      }
      sigmaRev[number] = symbol;
      continue;
    }
  }
  if (stateCount < 0) goto fail; /* (no ##props##: ToMatrix would run on an automaton without states) */
  {
    /* ToMatrix matrix.go:30-99 */
    m = (ora_model *)calloc(1, sizeof(*m));
    m->epsilon = epsilon; m->unknown = unknown; m->identity = identity; m->stateCount = stateCount;
    m->eot_rewind = 1;
    int max = 0;
    if (identity != -1) { for (int i = 0; i < 256; i++) m->sigmaASCII[i] = identity; max = identity; } /* :43-48 */
    uint32_t cap = 64;
    while (cap < (uint32_t)(sigmaCount + 2) * 4u) cap <<= 1;
    m->hmask = cap - 1;
    m->hkey = (int32_t *)malloc(cap * sizeof(int32_t));
    m->hval = (int32_t *)malloc(cap * sizeof(int32_t));
    for (uint32_t i = 0; i < cap; i++) m->hkey[i] = -1;
    for (int num = 0; num <= sigmaCount && (size_t)num < symcap; num++) { /* :50-65 */
      if (sigmaRev[num] < 0) continue;
      if (sigmaRev[num] < 256) m->sigmaASCII[sigmaRev[num]] = num;
      sigma_put(m, sigmaRev[num], num);
      if (num > max) max = num;
    }
    m->sigmaCount = max + 1;
    m->arraySize = ((size_t)stateCount + 1) * (size_t)(max + 1); /* :71 */
    m->array = (uint32_t *)calloc(m->arraySize + 1, 4);
    /* edges of a state, in file order; only the states reachable from state 1 are stored (:76-96) */
    const size_t S = (size_t)stateCount;
    size_t *first = (size_t *)calloc(S + 2, sizeof(size_t));
    for (size_t k = 0; k < edges.n; k++) first[edges.p[k].state + 1]++;
    for (size_t t = 1; t <= S + 1; t++) first[t] += first[t - 1];
    size_t *fill = (size_t *)malloc((S + 2) * sizeof(size_t));
    memcpy(fill, first, (S + 2) * sizeof(size_t));
    foma_edge *by_state = (foma_edge *)malloc((edges.n + 1) * sizeof(foma_edge));
    for (size_t k = 0; k < edges.n; k++) by_state[fill[edges.p[k].state]++] = edges.p[k];
    uint8_t *remember = (uint8_t *)calloc(S + 2, 1);
    int *stack = (int *)malloc((S + 2) * sizeof(int));
    size_t sp = 0;
    int okm = 1;
    if (S >= 1) { stack[sp++] = 1; remember[1] = 1; }
    while (sp && okm) {
      const int t = stack[--sp];
      for (size_t k = first[t]; k < first[t + 1]; k++) {
        const foma_edge *fe = &by_state[k];
        const long long idx = ((long long)fe->alpha - 1) * (long long)S + t;
        if (idx < 0 || (size_t)idx >= m->arraySize) { okm = 0; break; } /* Go: index out of range */
        m->array[idx] = (uint32_t)fe->end | (fe->nontoken ? FIRSTBIT : 0u); /* later edges replace earlier ones */
        if (fe->end > stateCount) { okm = 0; break; } /* :78 panic("stateCount is smaller") */
        if (fe->end >= 1 && !remember[fe->end]) { remember[fe->end] = 1; stack[sp++] = fe->end; }
      }
    }
    free(first); free(fill); free(by_state); free(remember); free(stack);
    if (!okm) { ora_free(m); m = NULL; }
  }
fail:
  free(sigmaRev); free(isMCS); free(edges.p);
  return m;
}

ora_model *ora_load_foma(const char *path) {
  size_t len = 0;
  uint8_t *buf = read_gz(path, &len);
  if (!buf) return NULL;
  ora_model *m = parse_foma(buf, len);
  free(buf);
  return m;
}

/* WriteTo matrix.go:126-210: the uncompressed image Save() gzips.  malloc'd, ora_free_bytes. */
uint8_t *ora_write_matrix(const ora_model *m, size_t *out_len) {
  int max = 0;
  for (uint32_t i = 0; i <= m->hmask; i++)
    if (m->hkey[i] != -1 && m->hval[i] > max) max = m->hval[i];
  int32_t *sigmalist = (int32_t *)calloc((size_t)max + 1, sizeof(int32_t));
  for (uint32_t i = 0; i <= m->hmask; i++)
    if (m->hkey[i] != -1) sigmalist[m->hval[i]] = m->hkey[i];
  uint8_t *out = (uint8_t *)malloc(5 + 14 + 4 * ((size_t)max + 1) + 1 + 4 * m->arraySize);
  size_t p = 0;
  memcpy(out, "MATOK", 5); p = 5;
#define PUT16(v) do { out[p++] = (uint8_t)((v) & 0xFF); out[p++] = (uint8_t)(((v) >> 8) & 0xFF); } while (0)
  PUT16(VERSION); PUT16((uint32_t)m->epsilon); PUT16((uint32_t)m->unknown); PUT16((uint32_t)m->identity);
  { uint32_t sc = (uint32_t)m->stateCount; PUT16(sc); PUT16(sc >> 16); }
  PUT16((uint32_t)(max + 1));
#undef PUT16
  for (int k = 0; k <= max; k++) p += (size_t)encode_rune(sigmalist[k], out + p);
  out[p++] = 'M';
  for (size_t x = 0; x < m->arraySize; x++) {
    const uint32_t v = m->array[x];
    out[p++] = (uint8_t)v; out[p++] = (uint8_t)(v >> 8); out[p++] = (uint8_t)(v >> 16); out[p++] = (uint8_t)(v >> 24);
  }
  free(sigmalist);
  *out_len = p;
  return out;
}

void ora_free(ora_model *m) {
  if (!m) return;
  free(m->hkey); free(m->hval); free(m->array); free(m);
}

int ora_epsilon(const ora_model *m) { return m->epsilon; }
int ora_unknown(const ora_model *m) { return m->unknown; }
int ora_identity(const ora_model *m) { return m->identity; }
int ora_state_count(const ora_model *m) { return m->stateCount; }
int ora_sigma_count(const ora_model *m) { return m->sigmaCount; }
const uint32_t *ora_array(const ora_model *m, size_t *n) { if (n) *n = m->arraySize; return m->array; }
const int32_t *ora_sigma_ascii(const ora_model *m) { return m->sigmaASCII; }

int ora_sigma_lookup(const ora_model *m, int32_t r, int *ok) {
  if (r < 256) return m->sigmaASCII[r];       /* matrix.go:421-425 */
  int a = sigma_get(m, r, ok);                /* :427 */
  if (!*ok && m->identity != -1) a = m->identity; /* :430-434 */
  return a;
}


typedef VEC(uint8_t) vec_u8;
typedef VEC(int32_t) vec_i32;
typedef VEC(uint32_t) vec_u32;
typedef VEC(uint64_t) vec_u64;

static void sink_write(vec_u8 *s, const uint8_t *d, size_t n) {
  if (s->n + n > s->cap) {
    size_t c = s->cap ? s->cap : 4096;
    while (c < s->n + n) c *= 2;
    s->p = (uint8_t *)realloc(s->p, c);
    s->cap = c;
  }
  memcpy(s->p + s->n, d, n);
  s->n += n;
}
static inline void sink_byte(vec_u8 *s, uint8_t b) { sink_write(s, &b, 1); }

static void sink_itoa(vec_u8 *s, long v) { /* strconv.Itoa */
  uint8_t tmp[24];
  int i = 24;
  unsigned long u = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
  do { tmp[--i] = (uint8_t)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) tmp[--i] = '-';
  sink_write(s, tmp + i, (size_t)(24 - i));
}

/* ------------------------------------------------------------ TokenWriter */

/* captured variables of NewTokenWriter, token_writer.go:37-42 */
typedef struct {
  uint32_t flags;
  vec_u8 *writer;
  long posC;
  vec_i32 pos;
  int sentB;
  vec_i32 sent;
  int init;
  int status;
  /* recorder (test infrastructure, not in the reference) */
  int record;
  vec_i32 rec_tok_pos;
  vec_i32 rec_sent_pos;
  size_t sent_flushed; /* sent entries already emitted by earlier TextEnds */
} token_writer;

static void tw_init(token_writer *tw, vec_u8 *sink, uint32_t flags, int record) {
  memset(tw, 0, sizeof(*tw));
  tw->flags = flags;
  tw->writer = sink;
  tw->posC = 0;
  tw->sentB = 1;
  tw->init = (flags & ORA_WRITER_USED) ? 0 : 1;
  tw->record = record;
}

static void tw_free(token_writer *tw) {
  free(tw->pos.p); free(tw->sent.p); free(tw->rec_tok_pos.p); free(tw->rec_sent_pos.p);
}

/* tw.Token, token_writer.go:59-100 */
static void tw_token(token_writer *tw, int offset, const int32_t *buf, int len) {
  uint32_t flags = tw->flags;
  if (flags & (ORA_TOKEN_POS | ORA_SENTENCE_POS)) {
    if (tw->posC == 0 && (flags & ORA_NEWLINE_AFTER_EOT)) { /* :66 */
      if (len < 1) { tw->status = ORA_PANIC_EMPTY_BUF; return; }
      if (buf[0] == '\n' && !tw->init) tw->posC--;
    }
    tw->init = 0;
    tw->posC += offset;
    VPUSH(tw->pos, (int32_t)tw->posC);
    if (tw->record) VPUSH(tw->rec_tok_pos, (int32_t)tw->posC);
    if (tw->sentB) {
      tw->sentB = 0;
      VPUSH(tw->sent, (int32_t)tw->posC);
      if (tw->record) VPUSH(tw->rec_sent_pos, (int32_t)tw->posC);
    }
    tw->posC += len - offset;
    VPUSH(tw->pos, (int32_t)tw->posC);
    if (tw->record) VPUSH(tw->rec_tok_pos, (int32_t)tw->posC);
    if (!(flags & ORA_TOKENS)) return;
  } else if (!(flags & ORA_TOKENS)) {
    return; /* :99 */
  }
  if (offset > len || offset < 0) { tw->status = ORA_PANIC_TOKEN_SLICE; return; }
  uint8_t enc[4];
  for (int i = offset; i < len; i++) { /* string(buf[offset:]) */
    int w = encode_rune(buf[i], enc);
    sink_write(tw->writer, enc, (size_t)w);
  }
  sink_byte(tw->writer, '\n');
}

/* tw.SentenceEnd, token_writer.go:103-127 */
static void tw_sentence_end(token_writer *tw) {
  uint32_t flags = tw->flags;
  if (flags & ORA_SENTENCE_POS) {
    if (tw->pos.n == 0) { tw->status = ORA_PANIC_SENT_NO_TOKEN; return; } /* :108 */
    int32_t v = tw->pos.p[tw->pos.n - 1];
    VPUSH(tw->sent, v);
    if (tw->record) VPUSH(tw->rec_sent_pos, v);
    tw->sentB = 1;
    if (flags & ORA_SENTENCES) sink_byte(tw->writer, '\n');
  } else if (flags & ORA_SENTENCES) {
    sink_byte(tw->writer, '\n');
  }
}

/* tw.TextEnd, token_writer.go:130-167 */
static void tw_text_end(token_writer *tw) {
  uint32_t flags = tw->flags;
  if (flags & (ORA_TOKEN_POS | ORA_SENTENCE_POS)) {
    if (flags & ORA_TOKEN_POS) {
      if (tw->pos.n == 0) { tw->status = ORA_PANIC_TEXT_NO_TOKEN; return; } /* :135 */
      sink_itoa(tw->writer, tw->pos.p[0]);
      for (size_t i = 1; i < tw->pos.n; i++) { sink_byte(tw->writer, ' '); sink_itoa(tw->writer, tw->pos.p[i]); }
      sink_byte(tw->writer, '\n');
    }
    if (flags & ORA_SENTENCE_POS) {
      if (tw->sent.n == 0) { tw->status = ORA_PANIC_TEXT_NO_SENT; return; } /* :145 */
      sink_itoa(tw->writer, tw->sent.p[0]);
      for (size_t i = 1; i < tw->sent.n; i++) { sink_byte(tw->writer, ' '); sink_itoa(tw->writer, tw->sent.p[i]); }
      sink_byte(tw->writer, '\n');
      tw->sent_flushed += tw->sent.n;
      tw->sent.n = 0;
      tw->sentB = 1;
    }
    tw->posC = 0;
    tw->pos.n = 0;
  } else {
    sink_byte(tw->writer, '\n');
  }
}

/* ------------------------------------------------------------------ walk */

typedef struct {
  vec_u32 tok_byte_start, tok_byte_end, tok_buf_start;
  vec_i32 tok_offset;
  vec_u64 sent_tok_idx;
  vec_u64 text_tok_end, text_sent_end, text_sentpos_end;
  vec_u32 text_byte_end;
} recorder;

typedef struct {
  uint64_t n_runes, n_iter, n_back, n_back_runes, n_hard;
  uint64_t n_tok, n_sent, n_text;
  uint32_t max_window;
  uint64_t *hist; /* optional: visits per (state) of the transition lookup, design input */
} walk_stats;

/* TransduceTokenWriter, matrix.go:348-698.  Returns status. */
static int transduce(const ora_model *mat, const uint8_t *in, size_t n, token_writer *w,
                     recorder *rec, const ora_carry *cin, ora_carry *cout, walk_stats *st) {
  int a = 0;
  uint32_t t0 = 0;
  uint32_t t = 1; /* :351 */
  int ok = 0;
  int rewindBuffer;
  uint32_t epsilonState = 0; /* :356 */
  int epsilonOffset = 0;
  int sentenceEnd = 0; /* :360 */
  int textEnd = 0;     /* :363 */
  if (cin) {
    if (cin->state) t = cin->state;
    ok = cin->ok; sentenceEnd = cin->sentence_end; textEnd = cin->text_end;
  }
  int32_t buffer[BUFSZ];      /* :365 */
  uint32_t boff[BUFSZ + 1];   /* byte offset of each buffered rune (recorder) */
  int bufft = 0, buffc = 0, buffi = 0;
  size_t rp = 0; /* reader cursor (bufio.Reader over `in`) */
  int32_t chr = 0;
  int eof = 0, eot = 0, newchar = 1;
  const int S = mat->stateCount;
  const uint32_t *array = mat->array;
  const int epsilon = mat->epsilon, identity = mat->identity, unknown = mat->unknown;
  uint64_t iter = 0, iter_cap = 64 * ((uint64_t)n + 64);
  uint64_t ntok = 0, nsent = 0;

#define BYTE_AT(i) ((i) < buffi ? boff[(i)] : (uint32_t)rp)
#define EMIT_TOKEN() do { \
    if (rec) { \
      if (bufft <= buffc) { VPUSH(rec->tok_byte_start, BYTE_AT(bufft)); } else { VPUSH(rec->tok_byte_start, BYTE_AT(buffc)); } \
      VPUSH(rec->tok_byte_end, BYTE_AT(buffc)); \
      VPUSH(rec->tok_buf_start, buffi > 0 ? boff[0] : (uint32_t)rp); \
      VPUSH(rec->tok_offset, bufft); } \
    ntok++; st->n_tok++; \
    tw_token(w, bufft, buffer, buffc); \
    if (w->status) return w->status; } while (0)
#define EMIT_SENT() do { \
    if (rec) VPUSH(rec->sent_tok_idx, ntok); \
    nsent++; st->n_sent++; \
    tw_sentence_end(w); \
    if (w->status) return w->status; } while (0)
#define EMIT_TEXT() do { \
    tw_text_end(w); st->n_text++; \
    if (w->status) return w->status; \
    if (rec) { VPUSH(rec->text_tok_end, ntok); VPUSH(rec->text_sent_end, nsent); \
      VPUSH(rec->text_sentpos_end, (uint64_t)w->sent_flushed); \
      VPUSH(rec->text_byte_end, BYTE_AT(buffc)); } } while (0)

  for (;;) { /* outer: re-entered by the `goto PARSECHARM` of :658,:667 */
    for (;;) { /* PARSECHARM :384 */
      if (++iter > iter_cap) return ORA_ERR_LOOP;
      if (newchar) {
        if (buffc >= buffi) { /* :388 */
          if (eof) break;
          if (rp >= n) { eof = 1; break; } /* io.EOF :396-399 */
          int width;
          chr = ora_decode_rune(in + rp, n - rp, &width); /* :392 */
          if (buffi >= BUFSZ) return ORA_PANIC_BUFFER_OVERFLOW; /* :406 */
          buffer[buffi] = chr;
          boff[buffi] = (uint32_t)rp;
          rp += (size_t)width;
          buffi++;
          st->n_runes++;
          if ((uint32_t)buffi > st->max_window) st->max_window = (uint32_t)buffi;
        }
        chr = buffer[buffc]; /* :410 */
        eot = 0;
        if (chr < 256) { /* :421 */
          eot = (chr == EOT);
          a = mat->sigmaASCII[chr];
        } else {
          a = sigma_get(mat, chr, &ok); /* :427 */
          if (!ok && identity != -1) a = identity; /* :430-434 */
        }
        t0 = t; /* :437 */
        if (array[(size_t)(epsilon - 1) * S + t0] != 0) { /* :442 */
          epsilonState = t0;
          epsilonOffset = buffc;
        }
      }
      if (a == 0) { /* :459 */
        t = 0;
      } else {
        /* Go checks the index: a model with an identity but without an unknown symbol (unknown == -1 in memory,
         * 65535 once saved and loaded) panics here on the retry of :478-485 */
        const long long ix = ((long long)a - 1) * (long long)S + (long long)t0;
        if (ix < 0 || (size_t)ix >= mat->arraySize) return ORA_PANIC_INDEX;
        t = array[(size_t)ix]; /* :463 */
      }
      st->n_iter++;
      if (st->hist) st->hist[t0]++;
      if (t == 0) { /* :472 */
        if (!ok && a == identity) { /* :478 */
          a = unknown;
        } else if (a != epsilon && epsilonState != 0) { /* :487 */
          t0 = epsilonState;
          epsilonState = 0;
          st->n_back++;
          st->n_back_runes += (uint64_t)(buffc - epsilonOffset);
          buffc = epsilonOffset;
          a = epsilon;
        } else { /* :499 */
          st->n_hard++;
          if (buffc - bufft <= 0) { /* :515 */
            buffc++;
            if (buffc == 0) { eof = 1; break; }
          }
          EMIT_TOKEN(); /* :528 */
          sentenceEnd = 0;
          textEnd = 0;
          memmove(buffer, buffer + buffc, (size_t)(buffi - buffc) * sizeof(int32_t)); /* :537 */
          memmove(boff, boff + buffc, (size_t)(buffi - buffc) * sizeof(uint32_t));
          buffi -= buffc;
          epsilonState = 0;
          buffc = 0;
          bufft = 0;
          a = epsilon; /* :545 */
          t = 1;
          newchar = 1;
          continue;
        }
        newchar = 0; /* :554 */
        eot = 0;
        continue;
      }
      rewindBuffer = 0; /* :560 */
      if (a == epsilon) { /* :563 */
        if (buffc - bufft > 0) {
          EMIT_TOKEN(); /* :569 */
          rewindBuffer = 1;
          sentenceEnd = 0;
          textEnd = 0;
        } else {
          sentenceEnd = 1;
          EMIT_SENT(); /* :575 */
        }
      } else {
        buffc++; /* :580 */
        if (buffc - bufft == 1 && (t & FIRSTBIT) != 0) bufft++; /* :584-588 */
      }
      if (eot) { /* :593 */
        eot = 0;
        if (!sentenceEnd) {
          sentenceEnd = 1;
          EMIT_SENT(); /* :597 */
        }
        textEnd = 1;
        EMIT_TEXT(); /* :600 */
        if (mat->eot_rewind) rewindBuffer = 1; /* matrix.go:603; absent in datok.go:1019-1030 */
      }
      if (rewindBuffer) { /* :608 */
        memmove(buffer, buffer + buffc, (size_t)(buffi - buffc) * sizeof(int32_t));
        memmove(boff, boff + buffc, (size_t)(buffi - buffc) * sizeof(uint32_t));
        buffi -= buffc;
        epsilonOffset = 0;
        epsilonState = 0;
        buffc = 0;
        bufft = 0;
      }
      t &= ~FIRSTBIT; /* :629 */
      newchar = 1;
    }
    if (!eof) return ORA_ERR_LOOP; /* :638-644 "should never happen" */
    /* final check :650-668 */
    t0 = t;
    t = array[(size_t)(epsilon - 1) * S + t0];
    a = epsilon;
    newchar = 0;
    if (t != 0) continue; /* goto PARSECHARM :658 */
    if (epsilonState != 0) { /* :660 */
      t0 = epsilonState;
      epsilonState = 0;
      st->n_back++;
      st->n_back_runes += (uint64_t)(buffc - epsilonOffset);
      buffc = epsilonOffset;
      continue; /* goto PARSECHARM :667 */
    }
    break;
  }
  if (buffc - bufft > 0) { /* :671 */
    EMIT_TOKEN();
    sentenceEnd = 0;
    textEnd = 0;
  }
  if (!sentenceEnd) EMIT_SENT(); /* :683 */
  if (!textEnd) EMIT_TEXT();     /* :690 */
  if (cout) {
    /* t is 0 here (the failed final epsilon probe, :652); the state a
     * continuation would start from is t0 */
    cout->state = t0; cout->ok = ok; cout->sentence_end = 1; cout->text_end = 1;
  }
  return ORA_OK;
#undef BYTE_AT
#undef EMIT_TOKEN
#undef EMIT_SENT
#undef EMIT_TEXT
}

/* ------------------------------------------------------------- public API */

ora_result *ora_transduce(const ora_model *m, const uint8_t *in, size_t n, uint32_t flags,
                          const ora_carry *carry_in) {
  ora_result *r = (ora_result *)calloc(1, sizeof(*r));
  vec_u8 sink = {0};
  token_writer tw;
  tw_init(&tw, &sink, flags, 1);
  recorder rec;
  memset(&rec, 0, sizeof rec);
  walk_stats st;
  memset(&st, 0, sizeof st);
  r->status = transduce(m, in, n, &tw, &rec, carry_in, &r->carry_out, &st);
  r->text = sink.p; r->text_len = sink.n;
  r->n_tokens = rec.tok_byte_end.n;
  r->tok_byte_start = rec.tok_byte_start.p; r->tok_byte_end = rec.tok_byte_end.p;
  r->tok_buf_start = rec.tok_buf_start.p; r->tok_offset = rec.tok_offset.p;
  r->n_tok_pos = tw.rec_tok_pos.n; r->tok_pos = tw.rec_tok_pos.p; tw.rec_tok_pos.p = NULL;
  r->n_sent_events = rec.sent_tok_idx.n; r->sent_tok_idx = rec.sent_tok_idx.p;
  r->n_sent_pos = tw.rec_sent_pos.n; r->sent_pos = tw.rec_sent_pos.p; tw.rec_sent_pos.p = NULL;
  r->n_texts = rec.text_tok_end.n;
  r->text_tok_end = rec.text_tok_end.p; r->text_sent_end = rec.text_sent_end.p;
  r->text_sentpos_end = rec.text_sentpos_end.p; r->text_byte_end = rec.text_byte_end.p;
  r->n_runes = st.n_runes; r->n_iterations = st.n_iter; r->n_backtracks = st.n_back;
  r->n_backtrack_runes = st.n_back_runes; r->n_hardfail = st.n_hard; r->max_window = st.max_window;
  tw_free(&tw);
  return r;
}

void ora_result_free(ora_result *r) {
  if (!r) return;
  free(r->text); free(r->tok_byte_start); free(r->tok_byte_end); free(r->tok_buf_start);
  free(r->tok_offset); free(r->tok_pos); free(r->sent_tok_idx); free(r->sent_pos);
  free(r->text_tok_end); free(r->text_sent_end); free(r->text_sentpos_end); free(r->text_byte_end);
  free(r);
}

/* ---- CPU baseline: one worker per EOT-delimited document ---------------- */

typedef struct {
  const ora_model *m;
  const uint8_t *in;
  const size_t *doc_start; /* ndocs+1 entries */
  size_t ndocs;
  uint32_t flags;
  atomic_size_t next;
  atomic_ullong tokens, sentences, out_bytes;
} mt_job;

static void *mt_worker(void *arg) {
  mt_job *job = (mt_job *)arg;
  vec_u8 sink = {0};
  unsigned long long tok = 0, sen = 0, ob = 0;
  for (;;) {
    size_t d0 = atomic_fetch_add(&job->next, 16);
    if (d0 >= job->ndocs) break;
    size_t d1 = d0 + 16 < job->ndocs ? d0 + 16 : job->ndocs;
    for (size_t d = d0; d < d1; d++) {
      sink.n = 0;
      token_writer tw;
      tw_init(&tw, &sink, job->flags, 0);
      walk_stats st;
      memset(&st, 0, sizeof st);
      transduce(job->m, job->in + job->doc_start[d], job->doc_start[d + 1] - job->doc_start[d], &tw,
                NULL, NULL, NULL, &st);
      tok += st.n_tok;
      sen += st.n_sent;
      ob += sink.n;
      tw_free(&tw);
    }
  }
  free(sink.p);
  atomic_fetch_add(&job->tokens, tok);
  atomic_fetch_add(&job->sentences, sen);
  atomic_fetch_add(&job->out_bytes, ob);
  return NULL;
}

uint64_t ora_transduce_docs_mt(const ora_model *m, const uint8_t *in, size_t n, uint32_t flags,
                               int nthreads, uint64_t *out_bytes, uint64_t *out_sentences,
                               uint64_t *out_docs) {
  /* document = bytes up to and including an EOT; a trailing EOT-less rest is a document too */
  size_t cap = 1024, nd = 0;
  size_t *ds = (size_t *)malloc(cap * sizeof(size_t));
  ds[0] = 0;
  for (size_t i = 0; i < n; i++) {
    if (in[i] == EOT) {
      if (nd + 2 >= cap) { cap *= 2; ds = (size_t *)realloc(ds, cap * sizeof(size_t)); }
      ds[++nd] = i + 1;
    }
  }
  if (ds[nd] < n) {
    if (nd + 2 >= cap) { cap *= 2; ds = (size_t *)realloc(ds, cap * sizeof(size_t)); }
    ds[++nd] = n;
  }
  mt_job job;
  memset(&job, 0, sizeof job);
  job.m = m; job.in = in; job.doc_start = ds; job.ndocs = nd; job.flags = flags;
  atomic_init(&job.next, 0);
  atomic_init(&job.tokens, 0); atomic_init(&job.sentences, 0); atomic_init(&job.out_bytes, 0);
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, mt_worker, &job);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
  free(ds);
  if (out_bytes) *out_bytes = atomic_load(&job.out_bytes);
  if (out_sentences) *out_sentences = atomic_load(&job.sentences);
  if (out_docs) *out_docs = nd;
  return atomic_load(&job.tokens);
}

/* Drives the TokenWriter restatement alone (token_writer_test.go:11-32).
 * ops: 0,offset,len,rune...  = Token ; 1 = SentenceEnd ; 2 = TextEnd */
uint8_t *ora_token_writer_replay(uint32_t flags, const int32_t *ops, size_t nops, size_t *out_len,
                                 int *status) {
  vec_u8 sink = {0};
  token_writer tw;
  tw_init(&tw, &sink, flags, 0);
  size_t i = 0;
  while (i < nops && !tw.status) {
    if (ops[i] == 0) {
      int offset = ops[i + 1], len = ops[i + 2];
      tw_token(&tw, offset, ops + i + 3, len);
      i += 3 + (size_t)len;
    } else if (ops[i] == 1) {
      tw_sentence_end(&tw); i++;
    } else {
      tw_text_end(&tw); i++;
    }
  }
  *status = tw.status;
  *out_len = sink.n;
  tw_free(&tw);
  if (!sink.p) sink.p = (uint8_t *)malloc(1);
  return sink.p;
}

void ora_free_bytes(uint8_t *p) { free(p); }

/* design input: how often each state is the source of a transition lookup */
int ora_state_histogram(const ora_model *m, const uint8_t *in, size_t n, uint64_t *hist) {
  vec_u8 sink = {0};
  token_writer tw;
  tw_init(&tw, &sink, 0, 0);
  walk_stats st;
  memset(&st, 0, sizeof st);
  st.hist = hist;
  int rc = transduce(m, in, n, &tw, NULL, NULL, NULL, &st);
  tw_free(&tw);
  free(sink.p);
  return rc;
}
