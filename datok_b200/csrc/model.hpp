// model.hpp -- host-side .matok loader and GPU re-layout (no CUDA in this file).
//
// Replaces, for the MATOK magic, LoadTokenizerFile (fomafile.go:452-484),
// LoadMatrixFile (matrix.go:214-231) and ParseMatrix (matrix.go:235-337).
// The reference keeps a symbol-major uint32 matrix `array[(a-1)*S + t]`
// (matrix.go:85,442,463).  The GPU layout built here is different by design:
//   * runes are mapped to *classes* (symbols with identical columns merged;
//     EOT and UTF-8 continuation bytes get private classes),
//   * the table is state-major u16: table[t << row_shift | cls], bit 15 = the
//     reference's FIRSTBIT "non-token" flag (datok.go:43), bits 0..14 = target,
//   * states are renumbered by expected visit frequency, so that the hottest
//     rows are the ones with the smallest ids (they are kept in shared memory),
//   * a second, "fused" table T3 (u32) folds the reference's fail -> backtrack to
//     the epsilon recorded at the same position -> take epsilon -> re-read
//     sequence (matrix.go:472-497,563-576) into one lookup per input byte.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace datok {

constexpr uint32_t CLS_EPS = 0;   // column of the epsilon symbol (@_TOKEN_BOUND_@)
constexpr uint32_t CLS_CONT = 1;  // UTF-8 continuation byte of a valid sequence: no-op step
constexpr uint32_t CLS_EOT = 2;   // the byte 0x04 (matrix.go:13,422)
constexpr uint32_t CLS_FIRST = 3; // first ordinary class
constexpr uint16_t NT_BIT = 0x8000;
// fused table T3, entry (u32): [14:0] target state, [15] non-token transition, [17:16] number of
// epsilon transitions taken before the byte is consumed, [18] the state the byte is finally consumed
// from has an epsilon transition itself.  0 = not decidable locally (older epsilon point or hard
// fail); T3_SLOW = leave this (state, class) to the exact walker.  Column CLS_EPS holds the target of
// the state's epsilon transition (0: none).
constexpr uint32_t T3_TGT = 0x7FFFu;
constexpr uint32_t T3_NTBIT = 1u << 15;
constexpr uint32_t T3_K_SHIFT = 16;
constexpr uint32_t T3_EPSBIT = 1u << 18;
constexpr uint32_t T3_SLOW = 1u << 31;
// compact copy of the rows of the hottest states for shared memory, entry (u16): [10:0] target
// state, [11] non-token, [12] consuming state has an epsilon transition, [13] at least one epsilon step,
// [14] two epsilon steps -- one flag per bit of the high byte, below bit 7 of that byte: the walk moves all
// four into predicates with one R2P instruction.
// 0 = look the entry up in T3 (marked, or a target that is not among the hot rows); H16_FAIL = T3 holds 0.
// A compact row holds the hot_cols most frequent classes only (class ids are ordered by measured
// frequency); column hot_cols is all zero: rarer classes -- and the walk's end-of-range sentinel -- are
// mapped there and take the path through T3.
constexpr uint32_t H16_TGT = 0x07FFu;
constexpr uint32_t H16_NTBIT = 1u << 11;
constexpr uint32_t H16_EPSBIT = 1u << 12;
constexpr uint32_t H16_KANYBIT = 1u << 13;
constexpr uint32_t H16_K2BIT = 1u << 14;
constexpr uint32_t H16_MAX_ROWS = 2048;
constexpr uint32_t H16_FAIL = H16_NTBIT;  // target 0 with this flag: T3 holds 0 (failure without epsilon transition)

struct HostModel {
  // --- reference view (ParseMatrix) ---
  int epsilon = 0, unknown = 0, identity = 0, stateCount = 0, sigmaCount = 0;
  int32_t sigmaASCII[256];
  std::vector<std::pair<int32_t, int32_t>> sigma;  // rune -> symbol id, file order, later wins
  std::vector<uint32_t> array;
  // false: the model came from a double-array file (.datok).  Its walk (datok.go:781-1135) is the matrix
  // walk except that an EOT does not rewind the buffer; until the kernels have that variant, inputs that
  // hold an EOT are refused for such a model (api.cu do_count).
  bool eot_rewind = true;

  // --- GPU layout ---
  uint32_t n_classes = 0;
  uint32_t row_shift = 7;
  uint8_t ascii_cls[128];   // class of runes 0x00..0x7F
  uint8_t latin1_cls[128];  // class of runes 0x80..0xFF
  std::vector<uint32_t> rune_key;  // runes >= 0x100 present in sigma, ascending
  std::vector<uint8_t> rune_cls;
  uint8_t identity_cls = 0;        // class of every rune not in sigma
  std::vector<uint16_t> table;     // (S+1) << row_shift entries
  std::vector<uint16_t> new_of_old, old_of_new;
  uint16_t start = 0;   // new id of the reference's initial state 1 (matrix.go:351)
  uint32_t sync_mask[8];  // class c is a sync class iff table[start][c] == (start | NT_BIT)
  uint32_t sync_ascii[4]; // ASCII byte b is a sync byte iff its class is a sync class
  uint32_t max_eps_chain = 0;
  std::vector<uint32_t> table2;    // fused table T3, (S+1) * stride2 entries
  uint32_t stride2 = 0;            // >= n_classes
  std::vector<uint16_t> hot16;     // compact rows of states 0..hot16_rows-1, stride16 entries each
  uint32_t stride16 = 0;           // entries per compact row (> hot_cols); stride16 / 2 is odd (shared-memory bank spread)
  uint32_t hot16_rows = 0;
  uint32_t hot_cols = 0;           // classes 0..hot_cols-1 have a column in the compact rows
  uint32_t force_hot_cols = 0;     // != 0: that many columns whatever the histogram says (DATOK_HOT_COLS, tests)
  uint32_t row_budget_bytes = 0;   // shared memory the kernel has for compact rows (0: unknown): with both histograms the
                                   // column count is the one that leaves the fewest steps to the full table
  std::vector<uint8_t> cls_base;   // class id -> layout-independent id (order of first use), for calibration histograms
  bool fast_ok = false;            // the fused tables are usable (else every entry is marked T3_SLOW / 0)
};

// error codes are the DATOK_ERR_* values of include/datok_b200.h
int load_matok_file(const char* path, HostModel& m, std::string& why);
int parse_matok_image(const uint8_t* d, size_t n, HostModel& m, std::string& why);
// the compile path: LoadFomaFile + ParseFoma (fomafile.go:56-450) + ToMatrix (matrix.go:30-99) over a gzipped
// foma file / its gunzipped text; WriteTo / Save (matrix.go:107-210) of a matrix model
int load_foma_file(const char* path, HostModel& m, std::string& why);
int compile_foma_image(const uint8_t* d, size_t n, HostModel& m, std::string& why);
int write_matok_image(const HostModel& m, std::vector<uint8_t>& out, std::string& why);
int save_matok_file(const HostModel& m, const char* path, std::string& why);
// hist (optional): visits per reference state id (stateCount+1 entries) measured on
// representative text; without it states are ordered breadth-first from the root.
// cls_hist (optional): occurrences per layout-independent class id (256 entries, see cls_base): class ids
// are then ordered by frequency and the compact rows keep the frequent ones only.
int build_layout(HostModel& m, std::string& why, const uint64_t* hist = nullptr, const uint64_t* cls_hist = nullptr);

// Go unicode/utf8.DecodeRune (what bufio.Reader.ReadRune yields, matrix.go:392)
int32_t decode_rune(const uint8_t* p, size_t n, int* width);

}  // namespace datok
