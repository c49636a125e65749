// model.cpp -- .matok loader + GPU re-layout (host side).  See model.hpp.
#include "model.hpp"

#include <zlib.h>

#include <algorithm>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "../../include/datok_b200.h"

namespace datok {

// Go unicode/utf8 semantics: any malformed, overlong, surrogate or out-of-range
// sequence decodes to U+FFFD consuming exactly one byte.
int32_t decode_rune(const uint8_t* p, size_t n, int* width) {
  *width = n ? 1 : 0;
  if (!n) return 0xFFFD;
  uint32_t b0 = p[0];
  if (b0 < 0x80) return (int32_t)b0;
  int need;
  uint32_t lo = 0x80, hi = 0xBF, cp;
  if (b0 >= 0xC2 && b0 <= 0xDF) { need = 2; cp = b0 & 0x1F; }
  else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 3; cp = b0 & 0x0F; if (b0 == 0xE0) lo = 0xA0; if (b0 == 0xED) hi = 0x9F; }
  else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 4; cp = b0 & 0x07; if (b0 == 0xF0) lo = 0x90; if (b0 == 0xF4) hi = 0x8F; }
  else return 0xFFFD;
  if (n < (size_t)need) return 0xFFFD;
  for (int i = 1; i < need; i++) {
    uint32_t b = p[i];
    if (b < lo || b > hi) return 0xFFFD;
    cp = (cp << 6) | (b & 0x3F);
    lo = 0x80; hi = 0xBF;
  }
  *width = need;
  return (int32_t)cp;
}

static uint32_t rd16(const uint8_t* p) { return p[0] | (p[1] << 8); }
static uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

// ParseDatok (datok.go:621-729) + conversion of the double array into the matrix's cell layout, so that the
// rest of the loader and the kernels see one kind of model.  A double-array transition (datok.go:888-902,
// 1058-1066): u = base(t) + a is valid iff u <= check(array[1]) and check(array[u]) == t; array[u]'s check
// word carries the non-token flag (bit 31); the state that follows is base(array[u]) if array[u] is
// "separate" (bit 31 of its base word: it stands for a representative state), else u itself.  Reachable
// states are numbered in breadth-first order from state 1, which stays 1.
static int parse_datok_image(const uint8_t* d, size_t n, HostModel& m, std::string& why) {
  if (n < 21) { why = "Not a datok file"; return DATOK_ERR_FORMAT; }
  const uint8_t* h = d + 5;
  if (rd16(h) != 1) { why = "Version not compatible"; return DATOK_ERR_FORMAT; }  // datok.go:667-672
  m.epsilon = (int)rd16(h + 2);
  m.unknown = (int)rd16(h + 4);
  m.identity = (int)rd16(h + 6);
  m.sigmaCount = (int)rd16(h + 10);                 // (h + 8: the final symbol, not used by the walk)
  const size_t da_size = (size_t)rd32(h + 12) / 2;  // datok.go:681
  size_t p = 21;
  for (int i = 0; i < 256; i++) m.sigmaASCII[i] = m.identity;  // datok.go:686-691
  m.sigma.clear();
  for (int x = 0; x < m.sigmaCount; x++) {  // datok.go:693-701
    int w;
    int32_t sym = decode_rune(d + p, n - p, &w);
    if (w == 0) continue;
    p += (size_t)w;
    if (sym != 0) {
      if (sym < 256) m.sigmaASCII[sym] = x;
      m.sigma.emplace_back(sym, x);
    }
  }
  if (p >= n || d[p] != 'T') { why = "Not a datok file"; return DATOK_ERR_FORMAT; }  // datok.go:703-713
  p++;
  if (da_size < 2 || n - p < da_size * 8) { why = "Not enough bytes read"; return DATOK_ERR_FORMAT; }  // datok.go:722-725
  const uint8_t* da = d + p;
  constexpr uint32_t REST = 0x3FFFFFFFu, FIRST = 0x80000000u;  // datok.go:43-45
  auto base = [&](size_t x) { return rd32(da + 8 * x); };
  auto check = [&](size_t x) { return rd32(da + 8 * x + 4); };
  const uint32_t max_index = check(1) & REST;  // datok.go:891
  // follows (t, a): 0 = no transition, else the state that follows | FIRST if the target is non-token
  auto step = [&](uint32_t t, int a) -> uint32_t {
    const uint64_t u = (uint64_t)(base(t) & REST) + (uint64_t)a;
    if (u > max_index || u >= da_size || (check(u) & REST) != t) return 0;
    uint32_t nx = (uint32_t)u;
    if (base(u) & FIRST) nx = base(u) & REST;
    return nx | ((check(u) & FIRST) ? FIRST : 0u);
  };
  std::vector<uint32_t> id(da_size, 0), order;
  id[1] = 1;
  order.push_back(1);
  for (size_t k = 0; k < order.size(); k++) {
    for (int a = 1; a < m.sigmaCount; a++) {
      const uint32_t nx = step(order[k], a) & ~FIRST;
      if (!nx) continue;
      if (nx >= da_size) { why = "double array: target out of range"; return DATOK_ERR_FORMAT; }
      if (!id[nx]) { order.push_back(nx); id[nx] = (uint32_t)order.size(); }
    }
  }
  m.stateCount = (int)order.size();
  const size_t S = order.size();
  if (S + 1 >= 32768) { why = "state count not in 1..32766"; return DATOK_ERR_UNSUPPORTED_MODEL; }  // (before the matrix is allocated)
  m.array.assign((S + 1) * (size_t)m.sigmaCount, 0);  // cell (a, t) at (a - 1) * S + t, as in a .matok image
  for (size_t k = 0; k < S; k++)
    for (int a = 1; a < m.sigmaCount; a++) {
      const uint32_t e = step(order[k], a);
      if (e) m.array[(size_t)(a - 1) * S + (k + 1)] = id[e & ~FIRST] | (e & FIRST);
    }
  m.eot_rewind = false;
  return DATOK_OK;
}

int parse_matok_image(const uint8_t* d, size_t n, HostModel& m, std::string& why) {
  // magic + 14-byte little-endian header (matrix.go:246-285)
  if (n >= 5 && std::memcmp(d, "DATOK", 5) == 0) {  // fomafile.go:476-480
    return parse_datok_image(d, n, m, why);
  }
  if (n < 19 || std::memcmp(d, "MATOK", 5) != 0) { why = "Not a matok file"; return DATOK_ERR_FORMAT; }
  const uint8_t* h = d + 5;
  if (rd16(h) != 1) { why = "Version not compatible"; return DATOK_ERR_FORMAT; }
  m.epsilon = (int)rd16(h + 2);
  m.unknown = (int)rd16(h + 4);
  m.identity = (int)rd16(h + 6);
  m.stateCount = (int)rd32(h + 8);
  m.sigmaCount = (int)rd16(h + 12);
  size_t p = 19;
  for (int i = 0; i < 256; i++) m.sigmaASCII[i] = m.identity;  // matrix.go:289-293
  m.sigma.clear();
  for (int x = 0; x < m.sigmaCount; x++) {  // matrix.go:295-303
    int w;
    int32_t sym = decode_rune(d + p, n - p, &w);
    if (w == 0) continue;
    p += (size_t)w;
    if (sym != 0) {
      if (sym < 256) m.sigmaASCII[sym] = x;
      m.sigma.emplace_back(sym, x);
    }
  }
  if (p >= n || d[p] != 'M') { why = "Not a matok file"; return DATOK_ERR_FORMAT; }  // matrix.go:305-315
  p++;
  size_t cells = ((size_t)m.stateCount + 1) * (size_t)m.sigmaCount;  // matrix.go:286
  if (n - p < cells * 4) { why = "Not enough bytes read"; return DATOK_ERR_FORMAT; }  // matrix.go:327
  m.array.resize(cells);
  for (size_t x = 0; x < cells; x++) m.array[x] = rd32(d + p + 4 * x);
  return DATOK_OK;
}

static int gunzip_file(const char* path, std::vector<uint8_t>& img, std::string& why) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { why = std::string("cannot open ") + path; return DATOK_ERR_IO; }
  unsigned char mg[2] = {0, 0};
  size_t got = std::fread(mg, 1, 2, f);
  std::fclose(f);
  if (got != 2 || mg[0] != 0x1f || mg[1] != 0x8b) { why = "gzip: invalid header"; return DATOK_ERR_IO; }  // fomafile.go:64-68
  gzFile gz = gzopen(path, "rb");
  if (!gz) { why = "gzopen failed"; return DATOK_ERR_IO; }
  std::vector<uint8_t> chunk(1 << 20);
  for (;;) {
    int r = gzread(gz, chunk.data(), (unsigned)chunk.size());
    if (r < 0) { gzclose(gz); why = "gzip: read error"; return DATOK_ERR_IO; }
    if (r == 0) break;
    img.insert(img.end(), chunk.begin(), chunk.begin() + r);
  }
  gzclose(gz);
  return DATOK_OK;
}

int load_matok_file(const char* path, HostModel& m, std::string& why) {
  std::vector<uint8_t> img;
  int rc = gunzip_file(path, img, why);
  if (rc) return rc;
  rc = parse_matok_image(img.data(), img.size(), m, why);
  if (rc) return rc;
  return build_layout(m, why);
}

// ---------------------------------------------------------------------------------------------------
// The compile path: LoadFomaFile (fomafile.go:56-72) + ParseFoma (fomafile.go:77-450) + ToMatrix
// (matrix.go:30-99), i.e. what `datok convert` runs before Save (cmd/datok.go:63).  The foma text format:
//   ##props##   one line; field 6 "is_deterministic" and field 9 "is_epsilon_free" have to be 1, field 2 is the
//               state count
//   ##sigma##   "<number> <symbol>" per line; the symbol is one rune, one of the @_..._@ specials, or a
//               multi-character symbol (not supported by the tokenizer: arcs on it are dropped)
//   ##states##  arcs, 5 / 4 / 3 / 2 numbers per line (state in out target final | state in target final |
//               in out target | in target); "-1 ..." ends the list
// Symbol numbers and states are shifted by one (0 = "no symbol" / failure state).  Only three kinds of arc are
// accepted: x:x, x:epsilon (a non-token arc, FIRSTBIT) and epsilon:@_TOKEN_BOUND_@ (stored in the epsilon row).
namespace {
struct Arc { int sym; int target; bool nontoken; };

bool atoi_go(const std::string& s, int& out) {  // strconv.Atoi
  size_t i = (s.size() && (s[0] == '-' || s[0] == '+')) ? 1 : 0;
  if (i == s.size()) return false;
  long long v = 0;
  for (; i < s.size(); i++) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
    if (v > 0x7fffffffLL) return false;
  }
  out = (int)(s[0] == '-' ? -v : v);
  return true;
}
std::vector<std::string> split_blank(const std::string& s) {  // strings.Split(s, " ")
  std::vector<std::string> out(1);
  for (char c : s) { if (c == ' ') out.emplace_back(); else out.back().push_back(c); }
  return out;
}
}  // namespace

int compile_foma_image(const uint8_t* d, size_t n, HostModel& m, std::string& why) {
  int epsilon = -1, unknown = -1, identity = -1, final_sym = -1, tokenend = -1;
  int sigma_count = 0, state_count = -1;
  std::map<int, int32_t> rune_of_sym;      // Automaton.sigmaRev
  std::map<int, bool> multi_char;          // Automaton.sigmaMCS
  std::vector<std::vector<Arc>> arcs;      // Automaton.transitions (a later arc of the same symbol replaces the earlier)
  enum Mode { START, PROPS, SIGMA, STATES, DONE } mode = START;
  int cur_state = 0, in_sym = 0, out_sym = 0, target = 0, is_final = 0;  // the parser keeps these across lines
  auto add_arc = [&](int state, int sym, int tgt, bool nt) -> bool {
    if (state < 0 || state > state_count) return false;  // transitions[state+1]: index out of range in the reference
    for (Arc& a : arcs[(size_t)state])
      if (a.sym == sym) { a.target = tgt; a.nontoken = nt; return true; }
    arcs[(size_t)state].push_back({sym, tgt, nt});
    return true;
  };
  size_t pos = 0;
  auto next_line = [&](std::string& line) -> bool {  // bufio ReadString('\n'); an unterminated last line is dropped
    const void* q = pos < n ? std::memchr(d + pos, '\n', n - pos) : nullptr;
    if (!q) return false;
    const size_t e = (size_t)((const uint8_t*)q - d);
    line.assign((const char*)d + pos, e - pos);
    pos = e + 1;
    return true;
  };
  std::string line;
  while (next_line(line)) {
    if (line.compare(0, 2, "##") == 0) {
      if (line.compare(0, 9, "##props##") == 0) mode = PROPS;
      else if (line.compare(0, 10, "##states##") == 0) { mode = STATES; final_sym = ++sigma_count; }  // fomafile.go:118-123
      else if (line.compare(0, 9, "##sigma##") == 0) mode = SIGMA;
      else if (line.compare(0, 7, "##end##") == 0) mode = DONE;
      else if (line.compare(0, 10, "##foma-net") != 0) break;  // "Unknown input line" ends the parse
      continue;
    }
    if (mode == PROPS) {
      const auto f = split_blank(line);
      if (f.size() < 13) { why = "foma: short ##props## line"; return DATOK_ERR_FORMAT; }
      if (f[6] != "1") { why = "The FST needs to be deterministic"; return DATOK_ERR_FORMAT; }       // fomafile.go:159
      if (f[9] != "1") { why = "The FST needs to be epsilon free"; return DATOK_ERR_FORMAT; }        // fomafile.go:164
      int v;
      if (!atoi_go(f[1], v)) { why = "Can't read arccount"; return DATOK_ERR_FORMAT; }
      if (!atoi_go(f[2], v) || v < 0) { why = "Can't read statecount"; return DATOK_ERR_FORMAT; }
      state_count = v;
      arcs.assign((size_t)v + 1, {});
    } else if (mode == SIGMA) {
      const size_t sp = line.find(' ');
      if (sp == std::string::npos) { why = "foma: sigma line without a symbol"; return DATOK_ERR_FORMAT; }
      int number;
      if (!atoi_go(line.substr(0, sp), number) || number < -1) { why = "foma: bad symbol number"; return DATOK_ERR_FORMAT; }
      number++;  // fomafile.go:382
      sigma_count = number;
      const std::string sym = line.substr(sp + 1);
      size_t n_runes = 0;
      int32_t first_rune = 0;
      for (size_t q = 0; q < sym.size();) {
        int w;
        const int32_t r = decode_rune((const uint8_t*)sym.data() + q, sym.size() - q, &w);
        if (n_runes++ == 0) first_rune = r;
        q += (size_t)w;
      }
      if (n_runes == 1) {
        rune_of_sym[number] = first_rune;
      } else if (n_runes > 1) {
        if (sym == "@_EPSILON_SYMBOL_@") epsilon = number;
        else if (sym == "@_UNKNOWN_SYMBOL_@") unknown = number;
        else if (sym == "@_IDENTITY_SYMBOL_@") identity = number;
        else if (sym == "@_TOKEN_SYMBOL_@" || sym == "@_TOKEN_BOUND_@") tokenend = number;
        else multi_char[number] = true;
      } else {  // "<number> " followed by an empty line: the symbol is '\n' (fomafile.go:425-439)
        std::string rest;
        if (!next_line(rest)) { why = "foma: unexpected end in ##sigma##"; return DATOK_ERR_FORMAT; }
        if (!rest.empty()) multi_char[number] = true; else rune_of_sym[number] = '\n';
      }
    } else if (mode == STATES) {
      const auto f = split_blank(line);
      if (f[0] == "-1") continue;
      int e[5] = {0, 0, 0, 0, 0};
      bool numeric = true;
      for (size_t k = 0; k < f.size() && k < 5; k++) numeric = numeric && atoi_go(f[k], e[k]);
      if (!numeric) continue;  // "Unable to translate": the line is skipped
      if (state_count < 0) { why = "foma: ##states## before ##props##"; return DATOK_ERR_FORMAT; }
      if (f.size() == 5) { cur_state = e[0]; in_sym = e[1]; out_sym = e[2]; target = e[3]; is_final = e[4]; }
      else if (f.size() == 4 && e[1] == -1) {  // a state without outgoing arcs
        cur_state = e[0]; is_final = e[3];
        if (is_final == 1 && !add_arc(cur_state + 1, final_sym, 0, false)) { why = "foma: state out of range"; return DATOK_ERR_FORMAT; }
        continue;
      }
      else if (f.size() == 4) { cur_state = e[0]; in_sym = out_sym = e[1]; target = e[2]; is_final = e[3]; }
      else if (f.size() == 3) { in_sym = e[0]; out_sym = e[1]; target = e[2]; }
      else if (f.size() == 2) { in_sym = out_sym = e[0]; target = e[1]; }
      in_sym++; out_sym++;  // fomafile.go:287-288
      bool nontoken = false;
      if (in_sym != out_sym) {
        if (out_sym == tokenend && in_sym == epsilon) { /* token boundary: lives in the epsilon row */ }
        else if (out_sym == epsilon) nontoken = true;
        else { why = "Unsupported transition: " + std::to_string(cur_state) + " -> " + std::to_string(target); return DATOK_ERR_FORMAT; }
      } else if (in_sym == tokenend) {
        continue;  // tokenend accepting arcs are ignored
      } else if (in_sym == epsilon) {
        why = "General epsilon transitions are not supported"; return DATOK_ERR_FORMAT;
      } else if (multi_char.count(in_sym)) {
        continue;  // arcs on multi-character symbols are ignored
      }
      if (cur_state + 1 < 0 || cur_state + 1 > state_count) { why = "foma: state out of range"; return DATOK_ERR_FORMAT; }
      if (in_sym >= 0) add_arc(cur_state + 1, in_sym, target + 1, nontoken);
      if (is_final == 1) add_arc(cur_state + 1, final_sym, 0, false);  // the '#' arc of Mizobuchi et al.: target 0
    }
  }
  if (state_count < 0) { why = "foma: no ##props## section"; return DATOK_ERR_FORMAT; }

  // ToMatrix (matrix.go:30-99)
  m = HostModel();
  m.epsilon = epsilon; m.unknown = unknown; m.identity = identity; m.stateCount = state_count;
  int max_sym = 0;
  for (int i = 0; i < 256; i++) m.sigmaASCII[i] = identity != -1 ? identity : 0;
  if (identity != -1) max_sym = identity;
  for (auto& kv : rune_of_sym) {
    if (kv.second < 256) m.sigmaASCII[kv.second] = kv.first;
    m.sigma.emplace_back(kv.second, kv.first);
    max_sym = std::max(max_sym, kv.first);
  }
  m.sigmaCount = max_sym + 1;
  const size_t S = (size_t)state_count;
  m.array.assign((S + 1) * (size_t)(max_sym + 1), 0);
  std::vector<char> seen(S + 2, 0);
  std::vector<int> todo;
  if (S >= 1) { todo.push_back(1); seen[1] = 1; }
  while (!todo.empty()) {  // only what is reachable from state 1 is stored (matrix.go:76-96)
    const int t = todo.back();
    todo.pop_back();
    for (const Arc& a : arcs[(size_t)t]) {
      const long long cell = ((long long)a.sym - 1) * (long long)S + t;
      if (cell < 0 || (size_t)cell >= m.array.size()) { why = "foma: symbol outside the matrix"; return DATOK_ERR_FORMAT; }
      m.array[(size_t)cell] = (uint32_t)a.target | (a.nontoken ? 0x80000000u : 0u);
      if (a.target > state_count) { why = "stateCount is smaller"; return DATOK_ERR_FORMAT; }  // matrix.go:78
      if (a.target >= 1 && !seen[(size_t)a.target]) { seen[(size_t)a.target] = 1; todo.push_back(a.target); }
    }
  }
  m.eot_rewind = true;
  return DATOK_OK;
}

int load_foma_file(const char* path, HostModel& m, std::string& why) {
  std::vector<uint8_t> img;
  int rc = gunzip_file(path, img, why);
  if (rc) return rc;
  rc = compile_foma_image(img.data(), img.size(), m, why);
  if (rc) return rc;
  return build_layout(m, why);
}

// WriteTo (matrix.go:126-210): magic, 14-byte header, the runes of sigma by symbol number (0 where a number
// has no rune), 'M', the cells
static void put_rune(std::vector<uint8_t>& o, int32_t r) {  // bufio.Writer.WriteRune
  if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = 0xFFFD;
  if (r < 0x80) o.push_back((uint8_t)r);
  else if (r < 0x800) { o.push_back((uint8_t)(0xC0 | (r >> 6))); o.push_back((uint8_t)(0x80 | (r & 0x3F))); }
  else if (r < 0x10000) { o.push_back((uint8_t)(0xE0 | (r >> 12))); o.push_back((uint8_t)(0x80 | ((r >> 6) & 0x3F))); o.push_back((uint8_t)(0x80 | (r & 0x3F))); }
  else { o.push_back((uint8_t)(0xF0 | (r >> 18))); o.push_back((uint8_t)(0x80 | ((r >> 12) & 0x3F))); o.push_back((uint8_t)(0x80 | ((r >> 6) & 0x3F))); o.push_back((uint8_t)(0x80 | (r & 0x3F))); }
}
int write_matok_image(const HostModel& m, std::vector<uint8_t>& out, std::string& why) {
  if (!m.eot_rewind) { why = "not a matrix model"; return DATOK_ERR_INVALID_ARG; }
  std::map<int32_t, int32_t> sym_of_rune;
  for (auto& kv : m.sigma) sym_of_rune[kv.first] = kv.second;
  int top = 0;
  for (auto& kv : sym_of_rune) top = std::max(top, kv.second);
  std::vector<int32_t> runes((size_t)top + 1, 0);
  for (auto& kv : sym_of_rune) runes[(size_t)kv.second] = kv.first;
  out.clear();
  out.insert(out.end(), {'M', 'A', 'T', 'O', 'K'});
  auto put16 = [&](uint32_t v) { out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); };
  put16(1); put16((uint32_t)m.epsilon); put16((uint32_t)m.unknown); put16((uint32_t)m.identity);
  put16((uint32_t)m.stateCount); put16((uint32_t)m.stateCount >> 16);
  put16((uint32_t)runes.size());
  for (int32_t r : runes) put_rune(out, r);
  out.push_back('M');
  out.reserve(out.size() + 4 * m.array.size());
  for (uint32_t v : m.array) { put16(v); put16(v >> 16); }
  return DATOK_OK;
}
// Save (matrix.go:107-123)
int save_matok_file(const HostModel& m, const char* path, std::string& why) {
  std::vector<uint8_t> img;
  int rc = write_matok_image(m, img, why);
  if (rc) return rc;
  gzFile gz = gzopen(path, "wb");
  if (!gz) { why = std::string("cannot create ") + path; return DATOK_ERR_IO; }
  size_t off = 0;
  while (off < img.size()) {
    const unsigned step = (unsigned)std::min<size_t>(img.size() - off, 1u << 30);
    if (gzwrite(gz, img.data() + off, step) != (int)step) { gzclose(gz); why = "gzip: write error"; return DATOK_ERR_IO; }
    off += step;
  }
  if (gzclose(gz) != Z_OK) { why = "gzip: close error"; return DATOK_ERR_IO; }
  return DATOK_OK;
}

int build_layout(HostModel& m, std::string& why, const uint64_t* hist, const uint64_t* cls_hist) {
  const int S = m.stateCount, K = m.sigmaCount, eps = m.epsilon;
  if (S < 1 || S + 1 >= 32768) { why = "state count not in 1..32766"; return DATOK_ERR_UNSUPPORTED_MODEL; }
  if (eps < 1 || eps >= K) { why = "no epsilon symbol"; return DATOK_ERR_UNSUPPORTED_MODEL; }
  // identity == -1: a model compiled in memory from a foma file without @_IDENTITY_SYMBOL_@ (matrix.go:43,430,459):
  // runes outside sigma have no symbol and every lookup on them fails.  (A *loaded* file never has -1: the u16
  // header field reads back as 65535, for which the reference indexes out of range.)
  if (m.identity != -1 && (m.identity < 1 || m.identity >= K)) { why = "no identity symbol"; return DATOK_ERR_UNSUPPORTED_MODEL; }
  auto cell = [&](int a, int t) -> uint32_t {  // matrix.go:463; a==0 never matches (matrix.go:459)
    if (a < 1 || a >= K) return 0;
    return m.array[(size_t)(a - 1) * S + t];
  };
  // The reference retries a failed identity transition with the `unknown` symbol
  // (matrix.go:478-485).  With an empty unknown column that retry can never
  // succeed, so it is unobservable; that is what the kernels assume.
  if (m.unknown >= 1 && m.unknown < K)
    for (int t = 1; t <= S; t++)
      if (cell(m.unknown, t) != 0) { why = "model has transitions on the unknown symbol"; return DATOK_ERR_UNSUPPORTED_MODEL; }

  // final rune -> symbol map (later sigma entries win, as Go's map assignment does)
  std::map<int32_t, int32_t> sym_of_rune;
  for (auto& kv : m.sigma) sym_of_rune[kv.first] = kv.second;
  for (auto& kv : sym_of_rune)
    if (kv.second == eps) { why = "a rune maps to the epsilon symbol"; return DATOK_ERR_UNSUPPORTED_MODEL; }

  // --- state renumbering: hottest states first ---
  // rank = measured visit count if given, else breadth-first distance from the root
  std::vector<uint32_t> bfs_rank(S + 1, 0xFFFFFFFFu);
  {
    std::vector<int> queue;
    queue.push_back(1);
    bfs_rank[1] = 0;
    for (size_t qi = 0; qi < queue.size(); qi++) {
      const int t = queue[qi];
      for (int a = 1; a < K; a++) {
        const uint32_t tgt = cell(a, t) & 0x7FFFFFFFu;
        if (tgt >= 1 && tgt <= (uint32_t)S && bfs_rank[tgt] == 0xFFFFFFFFu) {
          bfs_rank[tgt] = (uint32_t)queue.size();
          queue.push_back((int)tgt);
        }
      }
    }
  }
  std::vector<int> order(S);
  for (int t = 1; t <= S; t++) order[t - 1] = t;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    const uint64_t hx = hist ? hist[x] : 0, hy = hist ? hist[y] : 0;
    if (hx != hy) return hx > hy;
    return bfs_rank[x] < bfs_rank[y];
  });
  m.new_of_old.assign(S + 1, 0);
  m.old_of_new.assign(S + 1, 0);
  for (int i = 0; i < S; i++) { m.new_of_old[order[i]] = (uint16_t)(i + 1); m.old_of_new[i + 1] = (uint16_t)order[i]; }
  m.start = m.new_of_old[1];
  for (int t = 1; t <= S; t++) {
    uint32_t c = cell(eps, t), tgt = c & 0x7FFFFFFFu;
    if (tgt > (uint32_t)S) { why = "transition target out of range"; return DATOK_ERR_FORMAT; }
  }
  // longest chain of consecutive epsilon transitions (Token, SentenceEnd, ...)
  m.max_eps_chain = 0;
  for (int t = 1; t <= S; t++) {
    uint32_t len = 0;
    int cur = t;
    while (len <= 4) {
      uint32_t c = cell(eps, cur) & 0x7FFFFFFFu;
      if (!c) break;
      len++;
      cur = (int)c;
    }
    m.max_eps_chain = std::max(m.max_eps_chain, len);
  }

  // --- classes: merge symbols with identical columns ---
  std::map<std::vector<uint32_t>, uint32_t> class_of_column;
  std::vector<std::vector<uint32_t>> columns;  // per class id >= CLS_FIRST
  auto column_of = [&](int a) {
    std::vector<uint32_t> col(S + 1, 0);
    for (int t = 1; t <= S; t++) col[t] = cell(a, t);
    return col;
  };
  std::map<int, uint32_t> class_of_sym;
  auto cls_of_sym = [&](int a) -> uint32_t {
    auto it = class_of_sym.find(a);
    if (it != class_of_sym.end()) return it->second;
    auto col = column_of(a);
    auto jt = class_of_column.find(col);
    uint32_t c;
    if (jt == class_of_column.end()) {
      c = CLS_FIRST + (uint32_t)columns.size();
      class_of_column.emplace(col, c);
      columns.push_back(std::move(col));
    } else {
      c = jt->second;
    }
    class_of_sym[a] = c;
    return c;
  };
  uint32_t raw_ascii[256];
  for (int r = 0; r < 256; r++) raw_ascii[r] = cls_of_sym(m.sigmaASCII[r]);
  uint32_t ident = cls_of_sym(m.identity);
  std::vector<std::pair<uint32_t, uint32_t>> hi;  // runes >= 256
  for (auto& kv : sym_of_rune)
    if (kv.first >= 256) hi.emplace_back((uint32_t)kv.first, cls_of_sym(kv.second));
  m.n_classes = CLS_FIRST + (uint32_t)columns.size();
  if (m.n_classes > 256) { why = "more than 253 symbol classes"; return DATOK_ERR_UNSUPPORTED_MODEL; }
  m.row_shift = m.n_classes <= 128 ? 7 : 8;
  // class ids in order of measured frequency (the ids above are "base" ids: order of first use, the same
  // for every layout of this model); epsilon, continuation and EOT keep their fixed ids
  std::vector<uint32_t> perm(m.n_classes);  // base id -> class id
  {
    std::vector<uint32_t> by_freq;
    for (uint32_t c = CLS_FIRST; c < m.n_classes; c++) by_freq.push_back(c);
    if (cls_hist)
      std::stable_sort(by_freq.begin(), by_freq.end(), [&](uint32_t x, uint32_t y) { return cls_hist[x] > cls_hist[y]; });
    for (uint32_t c = 0; c < CLS_FIRST; c++) perm[c] = c;
    for (size_t i = 0; i < by_freq.size(); i++) perm[by_freq[i]] = CLS_FIRST + (uint32_t)i;
    m.cls_base.assign(m.n_classes, 0);
    for (uint32_t c = 0; c < m.n_classes; c++) m.cls_base[perm[c]] = (uint8_t)c;
  }
  for (int r = 0; r < 128; r++) m.ascii_cls[r] = (uint8_t)perm[raw_ascii[r]];
  for (int r = 0; r < 128; r++) m.latin1_cls[r] = (uint8_t)perm[raw_ascii[128 + r]];
  m.ascii_cls[4] = (uint8_t)CLS_EOT;
  m.identity_cls = (uint8_t)perm[ident];
  std::sort(hi.begin(), hi.end());
  m.rune_key.clear(); m.rune_cls.clear();
  for (auto& kv : hi) { m.rune_key.push_back(kv.first); m.rune_cls.push_back((uint8_t)perm[kv.second]); }
  // Columns of the compact rows.  Fewer columns mean more resident rows (fewer steps in a state without a row)
  // but more bytes of a class without a column; both kinds of step go through the full table.  With both
  // histograms and the kernel's shared-memory budget the count that minimises their sum is taken (a rare class
  // costs about twice a cold state: its class has to be decoded from the raw bytes again); with the class
  // histogram alone, the classes that cover all but ~0.05 % of the measured bytes.
  m.hot_cols = m.n_classes;
  if (cls_hist && hist && m.row_budget_bytes) {
    std::vector<uint64_t> visits;  // visits per state, hottest first
    uint64_t vis_total = 0;
    for (int t = 1; t <= S; t++) { visits.push_back(hist[t]); vis_total += hist[t]; }
    std::sort(visits.begin(), visits.end(), std::greater<uint64_t>());
    std::vector<uint64_t> vis_prefix(visits.size() + 1, 0);
    for (size_t i = 0; i < visits.size(); i++) vis_prefix[i + 1] = vis_prefix[i] + visits[i];
    uint64_t cls_total = 0;
    for (uint32_t c = 0; c < m.n_classes; c++) cls_total += cls_hist[c];
    std::vector<uint64_t> cls_prefix(m.n_classes + 1, 0);  // by class id (frequency order behind the three fixed ones)
    for (uint32_t c = 0; c < m.n_classes; c++) cls_prefix[c + 1] = cls_prefix[c] + cls_hist[m.cls_base[c]];
    double best = 1e300;
    for (uint32_t w = m.n_classes; w >= std::min<uint32_t>(m.n_classes, 24); w--) {
      const uint32_t stride = ((w + 2) / 2 | 1u) * 2;
      size_t rows = m.row_budget_bytes / ((size_t)stride * 2);
      if (rows > visits.size()) rows = visits.size();
      if (rows > H16_MAX_ROWS - 1) rows = H16_MAX_ROWS - 1;
      const double cold = vis_total ? (double)(vis_total - vis_prefix[rows]) / (double)vis_total : 0.0;
      const double rare = cls_total ? (double)(cls_total - cls_prefix[w]) / (double)cls_total : 0.0;
      const double cost = cold + 2.0 * rare;
      if (cost < best - 1e-12) { best = cost; m.hot_cols = w; }
    }
  } else if (cls_hist) {
    uint64_t total = 0, cum = 0;
    for (uint32_t c = 0; c < m.n_classes; c++) total += cls_hist[c];
    cum = cls_hist[CLS_EPS] + cls_hist[CLS_CONT] + cls_hist[CLS_EOT];
    uint32_t w = CLS_FIRST;
    while (w < m.n_classes && (total - cum) * 2000 > total) { cum += cls_hist[m.cls_base[w]]; w++; }
    m.hot_cols = std::max<uint32_t>(w, std::min<uint32_t>(m.n_classes, 24));
  }
  if (m.force_hot_cols) m.hot_cols = std::min<uint32_t>(m.n_classes, std::max<uint32_t>(m.force_hot_cols, CLS_FIRST));

  // --- table ---
  const size_t R = (size_t)1 << m.row_shift;
  m.table.assign(((size_t)S + 1) * R, 0);
  auto conv = [&](uint32_t c) -> uint16_t {
    uint32_t tgt = c & 0x7FFFFFFFu;
    if (!tgt) return 0;
    return (uint16_t)(m.new_of_old[tgt] | ((c & 0x80000000u) ? NT_BIT : 0));
  };
  const std::vector<uint32_t> eot_col = column_of(m.sigmaASCII[4]);
  for (int t = 1; t <= S; t++) {
    uint16_t* row = &m.table[(size_t)m.new_of_old[t] * R];
    for (int a = 1; a < K; a++) {
      uint32_t tgt = cell(a, t) & 0x7FFFFFFFu;
      if (tgt > (uint32_t)S) { why = "transition target out of range"; return DATOK_ERR_FORMAT; }
    }
    row[CLS_EPS] = conv(cell(eps, t));
    row[CLS_CONT] = (uint16_t)(m.new_of_old[t] | NT_BIT);
    row[CLS_EOT] = conv(eot_col[t]);
    for (size_t c = 0; c < columns.size(); c++) row[perm[CLS_FIRST + c]] = conv(columns[c][t]);
  }
  std::memset(m.sync_mask, 0, sizeof m.sync_mask);
  std::memset(m.sync_ascii, 0, sizeof m.sync_ascii);
  const uint16_t* srow = &m.table[(size_t)m.start * R];
  for (uint32_t c = CLS_EOT; c < m.n_classes; c++)
    if (srow[c] == (uint16_t)(m.start | NT_BIT)) m.sync_mask[c >> 5] |= 1u << (c & 31);
  for (uint32_t bch = 0; bch < 128; bch++) {
    const uint32_t c = m.ascii_cls[bch];
    if ((m.sync_mask[c >> 5] >> (c & 31)) & 1u) m.sync_ascii[bch >> 5] |= 1u << (bch & 31);
  }

  // --- fused table T3 ---
  // At the loop top in state t reading class c the reference (matrix.go:437-497):
  //   records the epsilon point (t, here) if t has an epsilon transition, looks up
  //   (t, c); on failure with that point recorded it backtracks by 0 runes, takes the
  //   epsilon transition (Token or SentenceEnd), and re-reads c from the new state.
  // T3[t][c] is the transition that finally consumes c, with the number of epsilon
  // steps before it.  0: failure in a state without epsilon transition (backtrack to an
  // older point or hard fail).  T3_SLOW: anything else the fast path leaves to walk_run().
  m.stride2 = m.n_classes | 1u;
  m.table2.assign(((size_t)S + 1) * m.stride2, 0);
  // An in-place backtrack stacks the epsilon steps of two lookups and its own at one position; with at
  // most two chained epsilon transitions in the model that never exceeds the two boundaries a position
  // can hold.  The fast path keeps 2 * class in one byte.
  m.fast_ok = m.max_eps_chain <= 2 && m.n_classes <= 128;
  for (int t = 1; t <= S; t++) {
    uint32_t* row2 = &m.table2[(size_t)t * m.stride2];
    const uint16_t* row = &m.table[(size_t)t * R];
    row2[CLS_EPS] = row[CLS_EPS] & 0x7FFFu;
    for (uint32_t c = CLS_CONT; c < m.n_classes; c++) {
      uint32_t cur = (uint32_t)t, k = 0, e = 0;
      for (;;) {
        const uint16_t* r = &m.table[(size_t)cur * R];
        if (r[c] != 0) {
          e = r[c] | (k << T3_K_SHIFT) | (r[CLS_EPS] != 0 ? T3_EPSBIT : 0u);  // r[c] carries NT_BIT == T3_NTBIT
          break;
        }
        if (r[CLS_EPS] == 0) {
          // failure in a state without epsilon transition.  k == 0: the walk backtracks to an
          // OLDER point or fails hard (entry 0).  k > 0: hard fail right after the epsilon steps.
          if (k) e = T3_SLOW;
          break;
        }
        if (k == 2) { e = T3_SLOW; break; }  // more than two epsilon steps at one position
        cur = r[CLS_EPS] & 0x7FFFu;
        k++;
      }
      // double-array walk: an EOT does not rewind the buffer (datok.go:1019-1030).  Its steps are left to the exact
      // walker, which knows that rule; the fast path then never sees a consumed EOT of such a model.
      if (c == CLS_EOT && !m.eot_rewind) e = T3_SLOW;
      row2[c] = m.fast_ok ? e : T3_SLOW;
    }
  }
  // --- compact rows of the hottest states ---
  // (column hot_cols and the padding behind it stay zero)
  m.stride16 = ((m.hot_cols + 2) / 2 | 1u) * 2;
  m.hot16_rows = std::min<uint32_t>((uint32_t)S + 1, H16_MAX_ROWS);
  m.hot16.assign((size_t)m.hot16_rows * m.stride16, 0);
  for (uint32_t t = 1; t < m.hot16_rows && m.fast_ok; t++) {
    const uint32_t* row2 = &m.table2[(size_t)t * m.stride2];
    uint16_t* h = &m.hot16[(size_t)t * m.stride16];
    for (uint32_t c = 0; c < m.hot_cols; c++) {
      const uint32_t e = row2[c], tgt = e & T3_TGT, k = (e >> T3_K_SHIFT) & 3u;
      if (e == 0 && c != CLS_EPS) { h[c] = (uint16_t)H16_FAIL; continue; }
      if (e == 0 || (e & T3_SLOW) || tgt >= H16_MAX_ROWS) continue;
      h[c] = (uint16_t)(tgt | ((e & T3_NTBIT) ? H16_NTBIT : 0u) | ((e & T3_EPSBIT) ? H16_EPSBIT : 0u) |
                        (k >= 1 ? H16_KANYBIT : 0u) | (k >= 2 ? H16_K2BIT : 0u));
    }
  }
  return DATOK_OK;
}

}  // namespace datok
