/*
 * corpus_gen.c -- deterministic synthetic corpora of the shapes BASELINE.json names
 * (SURVEY.md section 8d).  Bench/test tooling, not part of the transduction path:
 * the same bytes are fed to the CUDA path and to the CPU oracle.
 *
 *   kind 1  C1  simpletok: one document, no EOT, words [a-zà-ÿ]{1,12}
 *   kind 2  C2  German-like, ~10 KB documents (6-14 KB) each ended by EOT (+ "\n" p=.5)
 *   kind 3  C3  English-like, same shape, ASCII dominant, clitics
 *   kind 4  C4  German-like single document, no EOT, abbreviation/markup heavy
 *
 * All output is valid UTF-8, every document has at least one token, and no run of
 * more than ~500 runes is free of a token boundary (the reference panics beyond
 * 1024, matrix.go:365).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  uint64_t s;
  uint8_t *out;
  size_t n, cap;
  int kind;
  size_t open_quote; /* byte position by which an attribute value left open must meet a '<' (0 = none) */
} gen_t;

static inline uint64_t rnd(gen_t *g) { /* xorshift64* */
  uint64_t x = g->s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  g->s = x;
  return x * 2685821657736338717ull;
}
static inline uint32_t rndn(gen_t *g, uint32_t n) { return (uint32_t)((rnd(g) >> 33) % n); }
static inline int chance(gen_t *g, uint32_t per_mille10) { /* probability in 1/10000 */
  return rndn(g, 10000) < per_mille10;
}
static inline void put(gen_t *g, const char *s, size_t len) {
  if (g->n + len > g->cap) len = g->cap - g->n;
  memcpy(g->out + g->n, s, len);
  g->n += len;
}
static inline void puts_(gen_t *g, const char *s) { put(g, s, strlen(s)); }
static inline void putc_(gen_t *g, char c) { if (g->n < g->cap) g->out[g->n++] = (uint8_t)c; }

#define PICK(g, arr) (arr[rndn(g, (uint32_t)(sizeof(arr) / sizeof(arr[0])))])

static const char *DE_COMMON[] = {
  "der", "die", "und", "in", "den", "von", "zu", "das", "mit", "sich", "des", "auf", "f\xc3\xbcr", "ist", "im",
  "dem", "nicht", "ein", "eine", "als", "auch", "es", "an", "werden", "aus", "er", "hat", "dass", "sie", "nach",
  "wird", "bei", "einer", "um", "am", "sind", "noch", "wie", "einem", "\xc3\xbc" "ber", "einen", "so", "zum", "war",
  "haben", "nur", "oder", "aber", "vor", "zur", "bis", "mehr", "durch", "man", "sein", "wurde", "sei", "hatte",
  "kann", "gegen", "vom", "k\xc3\xb6nnen", "schon", "wenn", "habe", "seine", "ihre", "dann", "unter", "wir", "soll",
  "ich", "eines", "Jahr", "zwei", "Jahren", "diese", "dieser", "wieder", "keine", "Uhr", "seiner", "worden",
  "will", "zwischen", "immer", "Millionen", "was", "sagte", "gibt", "alle", "seit", "muss", "wurden", "beim",
  "doch", "jetzt", "waren", "drei", "Jahre", "neue", "neuen", "damit", "bereits", "da", "ab", "ohne", "sondern",
  "selbst", "ersten", "nun", "etwa", "heute", "weil", "ihm", "Menschen", "Deutschland", "anderen", "werde",
  "ihr", "ihrer", "viele", "dort", "Stra\xc3\x9f" "e", "gro\xc3\x9f" "e", "M\xc3\xa4nner", "K\xc3\xb6nig", "sch\xc3\xb6n",
  "Zeit", "Stadt", "Haus", "Frau", "Mann", "Kinder", "Land", "Welt", "Leben", "Arbeit", "Schule", "Wasser",
  "Regierung", "Unternehmen", "Geschichte", "Entwicklung", "M\xc3\xb6glichkeit", "Gesellschaft", "B\xc3\xbcrger",
};
static const char *DE_ONSET[] = {"b", "d", "f", "g", "h", "k", "l", "m", "n", "p", "r", "s", "t", "w", "z", "sch",
                                 "st", "sp", "tr", "br", "gr", "kl", "fr", "bl", "kr", "pf", "v", "j", ""};
static const char *DE_NUCLEUS[] = {"a", "e", "i", "o", "u", "e", "e", "a", "i", "ei", "au", "ie", "eu",
                                   "\xc3\xa4", "\xc3\xb6", "\xc3\xbc", "e", "a", "o", "u", "e", "i", "a", "e"};
static const char *DE_CODA[] = {"", "", "", "n", "r", "t", "s", "l", "m", "ch", "ng", "nd", "rt", "st", "\xc3\x9f",
                                "en", "er", "el", "ck", "tz", "ft", "cht", "ll", "nn", "rn"};
static const char *DE_ABBR[] = {"z.B.", "bzw.", "Dr.", "Prof.", "usw.", "ca.", "Nr.", "Abs.", "Art.", "Str.",
                                "etc.", "ggf.", "inkl.", "evtl.", "Tel.", "Mio.", "Mrd.", "Jh.", "Hrsg.", "vgl.",
                                "u.a.", "d.h.", "Abk.", "Bd.", "Aufl.", "Fr.", "Hr.", "St.", "allg.", "bes."};
static const char *EN_COMMON[] = {
  "the", "of", "and", "to", "a", "in", "is", "that", "it", "was", "for", "on", "are", "as", "with", "his", "they",
  "at", "be", "this", "from", "I", "have", "or", "by", "one", "had", "not", "but", "what", "all", "were", "when",
  "we", "there", "can", "an", "your", "which", "their", "said", "if", "do", "will", "each", "about", "how", "up",
  "out", "them", "then", "she", "many", "some", "so", "these", "would", "other", "into", "has", "more", "her",
  "two", "like", "him", "see", "time", "could", "no", "make", "than", "first", "been", "its", "who", "now",
  "people", "my", "made", "over", "did", "down", "only", "way", "find", "use", "may", "water", "long", "little",
  "very", "after", "words", "called", "just", "where", "most", "know", "government", "company", "history",
  "development", "possibility", "society", "citizen", "school", "house", "world", "life", "work",
};
static const char *EN_ONSET[] = {"b", "d", "f", "g", "h", "k", "l", "m", "n", "p", "r", "s", "t", "w", "sh", "st",
                                 "sp", "tr", "br", "gr", "cl", "fr", "bl", "cr", "th", "ch", "v", "j", ""};
static const char *EN_NUCLEUS[] = {"a", "e", "i", "o", "u", "e", "e", "a", "i", "ea", "ou", "ee", "oo", "ai", "o", "a"};
static const char *EN_CODA[] = {"", "", "", "n", "r", "t", "s", "l", "m", "ch", "ng", "nd", "rt", "st", "ck",
                                "ed", "er", "ly", "ing", "tion", "ll", "ss"};
static const char *EN_ABBR[] = {"Dr.", "Prof.", "Mr.", "Mrs.", "Ms.", "St.", "approx.", "Sept.", "Assoc.", "No.",
                                "pp.", "etc.", "e.g.", "i.e.", "vs.", "Inc.", "Ltd.", "Jan.", "Feb.", "Oct."};
static const char *EN_CLITIC[] = {"'ll", "'ve", "n't", "'s", "'re", "'d", "'m"};
static const char *XML_TAGS[] = {"<b>", "</b>", "<i>", "</i>", "<br />", "<p class=\"text\">", "</p>",
                                 "<a href=\"http://www.beispiel.de/seite\">", "</a>", "<x  y=\"alte zeit\">",
                                 "<!-- hm hm -->", "<?robot xgh ?>", "&amp;", "&quot;", "&nbsp;", "&lt;", "&gt;"};
static const char *TYPO[] = {"\xe2\x80\x9e", "\xe2\x80\x9c", "\xe2\x80\x9d", "\xc2\xbb", "\xc2\xab", "\xe2\x80\x93",
                             "\xe2\x80\xa6"};
static const char *EMOTICONS[] = {":-)", ";)", ":))", ":*(", "^___^", "T__T", "^^;", "-_-;;;", ":-*", "->", "<-"};
static const char *TLD[] = {"de", "com", "org", "net", "info", "eu"};

/* Abbreviation lists of the reference's grammar (src/de/abbrv.txt, src/en/abbrv.txt: one form per line,
 * without the final dot), handed in by the caller from fixture copies under testdata/; without them the
 * short built-in lists above are used.  [0] German, [1] English. */
static const char *g_abbr_text[2];
static size_t g_abbr_len[2];
static uint32_t *g_abbr_off[2];
static uint32_t g_abbr_n[2];
#include <stdlib.h>
void datok_corpus_set_abbreviations(int english, const char *text, size_t len) {
  const int k = english ? 1 : 0;
  free(g_abbr_off[k]);
  g_abbr_off[k] = NULL; g_abbr_n[k] = 0; g_abbr_text[k] = text; g_abbr_len[k] = len;
  if (!text || !len) return;
  uint32_t lines = 0;
  for (size_t i = 0; i < len; i++) lines += text[i] == '\n';
  g_abbr_off[k] = (uint32_t *)malloc((size_t)(lines + 2) * sizeof(uint32_t));
  size_t st = 0;
  for (size_t i = 0; i <= len; i++) {
    if (i == len || text[i] == '\n') {
      /* plain forms only: no blanks (the generator puts one abbreviation into one word slot) */
      int ok = i > st && i - st < 40;
      for (size_t j = st; j < i && ok; j++) ok = (unsigned char)text[j] > ' ';
      if (ok) g_abbr_off[k][g_abbr_n[k]++] = (uint32_t)st;
      st = i + 1;
    }
  }
}
static void put_abbreviation(gen_t *g, int en) {
  const int k = en ? 1 : 0;
  if (g_abbr_n[k] == 0) { puts_(g, en ? PICK(g, EN_ABBR) : PICK(g, DE_ABBR)); return; }
  const char *p = g_abbr_text[k] + g_abbr_off[k][rndn(g, g_abbr_n[k])];
  size_t l = 0;
  while (p + l < g_abbr_text[k] + g_abbr_len[k] && p[l] != '\n') l++;
  put(g, p, l);
  putc_(g, '.');
}

static void synth_word(gen_t *g, int en, int capital) {
  int syl = 1 + (int)rndn(g, 100) / 45; /* 1..3, mean ~1.8 */
  if (chance(g, 600)) syl += 2;         /* occasional compound */
  size_t start = g->n;
  for (int i = 0; i < syl; i++) {
    if (en) { puts_(g, PICK(g, EN_ONSET)); puts_(g, PICK(g, EN_NUCLEUS)); puts_(g, PICK(g, EN_CODA)); }
    else { puts_(g, PICK(g, DE_ONSET)); puts_(g, PICK(g, DE_NUCLEUS)); puts_(g, PICK(g, DE_CODA)); }
  }
  if (capital && start < g->n) {
    uint8_t c = g->out[start];
    if (c >= 'a' && c <= 'z') g->out[start] = (uint8_t)(c - 32);
    else if (c == 0xC3 && start + 1 < g->n && g->out[start + 1] >= 0xA0) g->out[start + 1] -= 0x20; /* ä->Ä */
  }
}

static void number_like(gen_t *g) {
  char b[64];
  int n = 0;
  switch (rndn(g, 7)) {
    case 0: n = __builtin_snprintf(b, sizeof b, "%u.", 1 + rndn(g, 30)); break;                       /* ordinal */
    case 1: n = __builtin_snprintf(b, sizeof b, "%u.%u.%u", 1 + rndn(g, 28), 1 + rndn(g, 12), 1900 + rndn(g, 125)); break;
    case 2: n = __builtin_snprintf(b, sizeof b, "%u:%02u", rndn(g, 24), rndn(g, 60)); break;
    case 3: n = __builtin_snprintf(b, sizeof b, "%u.%u.%u.%u", rndn(g, 256), rndn(g, 256), rndn(g, 256), rndn(g, 256)); break;
    case 4: n = __builtin_snprintf(b, sizeof b, "%u,%u%%", rndn(g, 100), rndn(g, 10)); break;
    case 5: n = __builtin_snprintf(b, sizeof b, "%u", rndn(g, 100000)); break;
    default: n = __builtin_snprintf(b, sizeof b, "%u/%u/%u", 1 + rndn(g, 28), 1 + rndn(g, 12), 1900 + rndn(g, 125)); break;
  }
  put(g, b, (size_t)n);
}

static void url_like(gen_t *g, int en) {
  switch (rndn(g, 3)) {
    case 0:
      puts_(g, chance(g, 5000) ? "https://www." : "http://");
      synth_word(g, en, 0); putc_(g, '.'); puts_(g, PICK(g, TLD)); putc_(g, '/'); synth_word(g, en, 0);
      if (chance(g, 3000)) { puts_(g, "?q="); synth_word(g, en, 0); }
      break;
    case 1:
      synth_word(g, en, 0); putc_(g, '@'); synth_word(g, en, 0); putc_(g, '.'); puts_(g, PICK(g, TLD));
      break;
    default:
      synth_word(g, en, 0); puts_(g, chance(g, 5000) ? ".txt" : ".jpeg");
      break;
  }
}

/* one word slot (possibly special), without the following separator */
static void word_slot(gen_t *g, int first) {
  int en = (g->kind == 3);
  int heavy = (g->kind == 4);
  uint32_t r = rndn(g, 10000);
  uint32_t p_abbr = heavy ? 1500 : 200, p_num = 100, p_url = 50, p_xml = heavy ? 500 : 50, p_typo = 30;
  uint32_t p_clitic = en ? 300 : 0, p_emot = heavy ? 100 : 0, p_hyph = heavy ? 100 : 0;
  uint32_t acc = 0;
  if (g->open_quote && g->n >= g->open_quote) {
    g->open_quote = 0;
    /* an unterminated attribute value makes the automaton read ahead until it meets
     * '<' (or the quote); bound that lookahead well below the 1024-rune buffer */
    puts_(g, chance(g, 5000) ? "<b>" : "\"> <i>");
    return;
  }
  if (r < (acc += p_abbr)) { put_abbreviation(g, en); return; }
  if (r < (acc += p_num)) { number_like(g); return; }
  if (r < (acc += p_url)) { url_like(g, en); return; }
  if (r < (acc += p_xml)) {
    if (heavy && chance(g, 2000)) { /* unclosed tag / attribute with spaces */
      if (g->open_quote == 0 && chance(g, 5000)) { puts_(g, "<x y=\"alte zeit"); g->open_quote = g->n + 20 + rndn(g, 400); }
      else puts_(g, "<br class=\"a b c\" ");
    } else {
      puts_(g, PICK(g, XML_TAGS));
    }
    return;
  }
  if (r < (acc += p_typo)) { puts_(g, PICK(g, TYPO)); return; }
  if (r < (acc += p_clitic)) { puts_(g, PICK(g, EN_COMMON)); puts_(g, PICK(g, EN_CLITIC)); return; }
  if (r < (acc += p_emot)) { puts_(g, PICK(g, EMOTICONS)); return; }
  if (r < (acc += p_hyph)) { /* long hyphen compound, up to ~500 runes */
    int parts = 2 + (int)rndn(g, (heavy && !g->open_quote) ? 60 : 4);
    size_t st = g->n;
    for (int i = 0; i < parts && g->n - st < 440; i++) { if (i) putc_(g, '-'); synth_word(g, en, 1); }
    return;
  }
  if (en && chance(g, 100)) { puts_(g, "I."); return; }
  int capital = first || (!en && chance(g, 2500));
  if (chance(g, 5500)) {
    const char *w = en ? PICK(g, EN_COMMON) : PICK(g, DE_COMMON);
    size_t st = g->n;
    puts_(g, w);
    if (first && st < g->n && g->out[st] >= 'a' && g->out[st] <= 'z') g->out[st] -= 32;
  } else {
    synth_word(g, en, capital);
  }
}

static void sentence(gen_t *g) {
  int heavy = (g->kind == 4);
  int words = 4 + (int)rndn(g, 27);
  int quoted = chance(g, 400);
  if (quoted) puts_(g, g->kind == 3 ? "\"" : (chance(g, 5000) ? "\xe2\x80\x9e" : "\xc2\xbb"));
  for (int i = 0; i < words; i++) {
    word_slot(g, i == 0);
    if (i + 1 < words) {
      if (chance(g, 800)) putc_(g, ',');
      if (chance(g, 100)) puts_(g, " -");
      if (chance(g, 30)) { puts_(g, " ("); word_slot(g, 0); putc_(g, ')'); }
      if (chance(g, 800)) putc_(g, '\n'); else if (chance(g, 100)) puts_(g, "  "); else putc_(g, ' ');
    }
  }
  uint32_t r = rndn(g, 100);
  if (r < 83) putc_(g, '.');
  else if (r < 90) putc_(g, '?');
  else if (r < 96) putc_(g, '!');
  else if (r < 98) puts_(g, heavy && chance(g, 5000) ? " ... " : "...");
  else puts_(g, "?!");
  if (quoted) puts_(g, g->kind == 3 ? "\"" : (chance(g, 5000) ? "\xe2\x80\x9c" : "\xc2\xab"));
}

static void document_body(gen_t *g, size_t target_end) {
  while (g->n + 400 < target_end) {
    sentence(g);
    if (chance(g, 1500)) puts_(g, "\n\n");
    else if (chance(g, 1000)) putc_(g, '\n');
    else putc_(g, ' ');
  }
}

static void simple_doc(gen_t *g) { /* C1 */
  static const char *SEP[] = {" ", "\t", "\n", "  "};
  while (g->n + 64 < g->cap) {
    int len = 1 + (int)rndn(g, 12);
    for (int i = 0; i < len; i++) {
      if (chance(g, 1000)) { putc_(g, (char)0xC3); putc_(g, (char)(0xA0 + rndn(g, 32))); } /* à-ÿ */
      else putc_(g, (char)('a' + rndn(g, 26)));
    }
    if (chance(g, 1200)) {
      static const char *P[] = {".", "?", "!"};
      puts_(g, PICK(g, P));
      if (chance(g, 1700)) puts_(g, chance(g, 5000) ? "?!" : "..");
    }
    puts_(g, PICK(g, SEP));
  }
  while (g->n < g->cap) putc_(g, g->n + 1 == g->cap ? 'a' : ' ');
}

/* Fills out[0..nbytes) exactly.  Returns the number of documents written. */
size_t datok_corpus_generate(int kind, uint64_t seed, uint8_t *out, size_t nbytes) {
  gen_t g;
  g.s = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  if (g.s == 0) g.s = 88172645463325252ull;
  g.out = out; g.n = 0; g.cap = nbytes; g.kind = kind; g.open_quote = 0;
  for (int i = 0; i < 8; i++) rnd(&g);
  if (nbytes == 0) return 0;
  if (kind == 1) { simple_doc(&g); return 1; }
  size_t docs = 0;
  if (kind == 4) {
    /* single document, no EOT: reserve the tail for a clean sentence end */
    g.cap = nbytes;
    document_body(&g, nbytes > 64 ? nbytes - 64 : 0);
    while (g.n + 2 < nbytes) { putc_(&g, 'a' + (char)rndn(&g, 26)); if (chance(&g, 1500)) putc_(&g, ' '); }
    while (g.n < nbytes) putc_(&g, g.n + 1 == nbytes ? '.' : 'e');
    return 1;
  }
  while (g.n < nbytes) {
    size_t remaining = nbytes - g.n;
    size_t dl = 6144 + rndn(&g, 8192);
    if (remaining < dl + 6144) dl = remaining; /* last document takes the rest */
    size_t end = g.n + dl;
    size_t tail = (dl == remaining || !chance(&g, 5000)) ? 1 : 2; /* EOT (+ "\n" with p=.5) */
    g.cap = end - tail;
    if (g.cap < g.n) g.cap = g.n;
    document_body(&g, g.cap);
    /* pad to the exact document length with short words, always ending in a token */
    while (g.n + 1 < g.cap) { putc_(&g, 'a' + (char)rndn(&g, 26)); if (chance(&g, 1500) && g.n + 2 < g.cap) putc_(&g, ' '); }
    while (g.n < g.cap) putc_(&g, '.');
    g.cap = nbytes;
    if (g.n < nbytes) putc_(&g, 4);
    if (tail == 2 && g.n < nbytes) putc_(&g, '\n');
    docs++;
  }
  return docs;
}
