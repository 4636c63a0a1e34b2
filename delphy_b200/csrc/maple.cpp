// maple.cpp -- MAPLE alignments straight to the flat arrays the SPR studies take (host side).
//
// A MAPLE file is a reference sequence followed, per sample, by its differences from it: "<letter> <site> [<length>]" lines --
// substitutions, and runs of missing / ambiguous sites.  The reference reads it (read_maple, core/io.cpp:98-254) with one getline +
// istringstream per line into a vector of Tip_desc {name, date range, vector<Seq_delta>, Missation_map}, which build_usher_like_tree
// then feeds, tip by tip, to an SPR study of a sequence that is not in the tree yet (core/phylo_tree.cpp:918-932).  Those studies take
// exactly "deltas from the reference sequence + missing intervals" (dphy_spr_request, DPHY_SPR_X_REL_REF), so this parser goes from
// the text to CSR arrays of that shape in one pass over the buffer -- no per-line stream objects, no per-tip containers -- and a
// request's x_delta_* / x_missing_* pointers are slices of them.  Host code: there is nothing here for a GPU to do.
//
// The behaviour is the reference's, including what it tolerates and what it drops (checked line for line against the compiled
// reference in tests/test_maple.py):
//   * ambiguous letters in the reference sequence become A, and a sample with a substitution at such a site is dropped (:128-150, :232);
//   * a sample whose id carries no valid date (..|YYYY-MM-DD, ..|YYYY-MM, ..|YYYY, ..|YYYY-MM-DD/YYYY-MM-DD; '-' also separates) is
//     dropped (core/sequence_utils.cpp:98-215), as is one with any malformed line; every such line counts as a warning;
//   * "T -> T" lines are skipped (MAPLE's spurious T->U), other letter == reference lines drop the sample (:229-237);
//   * a missing run without a length covers one site; runs are merged when they overlap or touch (Interval_set::insert,
//     core/interval_set.h:96-124);
//   * the number fields follow operator>>(int): leading blanks, an optional sign, digits; a field that is absent leaves the default,
//     one that is malformed reads as 0.
#include "delphy_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

struct dphy_maple {
  std::vector<uint8_t> ref;
  std::vector<double> t_min, t_max;
  std::vector<int64_t> name_off{0};
  std::string names;
  std::vector<int32_t> delta_off{0}, delta_site;
  std::vector<uint8_t> delta_from, delta_to;
  std::vector<int32_t> miss_off{0}, miss_start, miss_end;
  int64_t num_warnings = 0;
};

namespace {

thread_local std::string g_error;

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 'a' + 'A') : c; }

// Seq_letter bit sets (core/sequence.h:17-25, 115-152): A 1, C 2, G 4, T 8; 0 == not a sequence letter
int letter_bits(char c) {
  switch (upper(c)) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8;
    case 'R': return 1 | 4; case 'Y': return 2 | 8; case 'S': return 4 | 2; case 'W': return 1 | 8; case 'K': return 4 | 8; case 'M': return 1 | 2;
    case 'B': return 2 | 4 | 8; case 'D': return 1 | 4 | 8; case 'H': return 1 | 2 | 8; case 'V': return 1 | 2 | 4;
    case 'N': case '-': case '?': case '.': return 15;
    default: return 0;
  }
}
int real_letter(int bits) { return bits == 1 ? 0 : bits == 2 ? 1 : bits == 4 ? 2 : 3; }

// ---- the few pieces of std::istream behaviour the format's number fields rest on ----------------------------------------------------------
struct Field {          // an istringstream over one line: position + the two state bits that matter
  const char* p; const char* e;
  bool fail = false, eof = false;
  bool good() const { return !fail && !eof; }
  // the sentry of a formatted extraction: refuses a stream that is not good(); skips blanks; end of input sets eof | fail
  bool sentry() {
    if (!good()) { fail = true; return false; }
    while (p < e && is_space(*p)) ++p;
    if (p == e) { eof = true; fail = true; return false; }
    return true;
  }
  bool get_char(char& c) { if (!sentry()) return false; c = *p++; return true; }
  // operator>>(int&): untouched when the sentry fails; 0 + fail without digits; INT_MAX / INT_MIN + fail on overflow; eof when the
  // digits run to the end of the line
  void get_int(int& v) {
    if (!sentry()) return;
    const char* q = p;
    bool neg = false;
    if (q < e && (*q == '+' || *q == '-')) { neg = *q == '-'; ++q; }
    const char* d0 = q;
    long long acc = 0;
    bool over = false;
    while (q < e && *q >= '0' && *q <= '9') {
      if (!over) { acc = acc * 10 + (*q - '0'); if (acc > (long long)INT_MAX + 1) over = true; }
      ++q;
    }
    if (q == e) eof = true;
    if (q == d0) { p = q; v = 0; fail = true; return; }
    p = q;
    const long long val = neg ? -acc : acc;
    if (over || val > INT_MAX || val < INT_MIN) { v = neg ? INT_MIN : INT_MAX; fail = true; return; }
    v = (int)val;
  }
};

// ---- dates (core/dates.cpp:12-47 over absl::CivilDay; core/sequence_utils.cpp:63-215) ------------------------------------------------------
long long days_from_civil(long long y, int m, int d) {       // days since 1970-01-01 of a proleptic Gregorian date
  y -= m <= 2;
  const long long era = (y >= 0 ? y : y - 399) / 400;
  const long long yoe = y - era * 400;
  const long long doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const long long doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}
const long long kEpoch = 18262;                               // 2020-01-01 (core/dates.cpp:13)
bool leap(int y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }
int days_in_month(int y, int m) { static const int dm[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31}; return m == 2 && leap(y) ? 29 : dm[m - 1]; }
bool all_digits(const char* s, int n) { for (int i = 0; i < n; ++i) if (s[i] < '0' || s[i] > '9') return false; return true; }
int num(const char* s, int n) { int v = 0; for (int i = 0; i < n; ++i) v = v * 10 + (s[i] - '0'); return v; }
bool date_shape(const char* s) { return all_digits(s, 4) && s[4] == '-' && all_digits(s + 5, 2) && s[7] == '-' && all_digits(s + 8, 2); }
// parse_iso_date: false where absl::ParseCivilTime refuses the (already digit-shaped) string: month or day out of range
bool parse_date(const char* s, double& t) {
  const int y = num(s, 4), m = num(s + 5, 2), d = num(s + 8, 2);
  if (m < 1 || m > 12 || d < 1 || d > days_in_month(y, m)) return false;
  t = (double)(days_from_civil(y, m, d) - kEpoch);
  return true;
}
// extract_date_range_from_sequence_id: the first shape that fits the END of the id decides; a fitting shape that does not parse
// means "no date" (the reference's try block swallows the exception and gives up)
bool date_range_of_id(const std::string& id, double& t_min, double& t_max) {
  const size_t n = id.size();
  auto sep = [&](size_t tail) { return n >= tail + 1 && (id[n - tail - 1] == '|' || id[n - tail - 1] == '-'); };
  if (sep(21)) {
    const char* s = id.data() + n - 21;
    if (date_shape(s) && s[10] == '/' && date_shape(s + 11)) return parse_date(s, t_min) && parse_date(s + 11, t_max);
  }
  if (sep(10)) {
    const char* s = id.data() + n - 10;
    if (date_shape(s)) { if (!parse_date(s, t_min)) return false; t_max = t_min; return true; }
  }
  if (sep(7)) {
    const char* s = id.data() + n - 7;
    if (all_digits(s, 4) && s[4] == '-' && all_digits(s + 5, 2)) {
      const int y = num(s, 4), m = num(s + 5, 2);
      if (m < 1 || m > 12) return false;
      t_min = (double)(days_from_civil(y, m, 1) - kEpoch);
      t_max = (double)(days_from_civil(m == 12 ? y + 1 : y, m == 12 ? 1 : m + 1, 1) - kEpoch);
      return true;
    }
  }
  if (sep(4)) {
    const char* s = id.data() + n - 4;
    if (all_digits(s, 4)) {
      const int y = num(s, 4);
      t_min = (double)(days_from_civil(y, 1, 1) - kEpoch);
      t_max = (double)(days_from_civil(y + 1, 1, 1) - kEpoch);
      return true;
    }
  }
  return false;
}

// std::getline over the buffer: a line ends at '\n' or at the end of the text; no line once the text is used up
struct Lines {
  const char* p; const char* e;
  bool next(const char*& b, const char*& end) {
    if (p >= e) return false;
    b = p;
    const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(e - p)));
    if (nl) { end = nl; p = nl + 1; } else { end = e; p = e; }
    return true;
  }
};

int fail_with(const char* msg) { g_error = msg; return DPHY_ERR_INVALID_ARGUMENT; }

int parse(const char* text, size_t len, dphy_maple& M) {
  Lines in{text, text + len};
  const char *b, *e;
  if (!in.next(b, e)) return fail_with("Unexpected EOF while reading MAPLE file reference sequence");
  if (b == e) return fail_with("Unexpected empty reference id line in MAPLE file");
  if (*b != '>') return fail_with("Expected reference sequence id line to start with '>'");
  if (!in.next(b, e)) return fail_with("Unexpected EOF while reading MAPLE file reference sequence");
  // ---- the reference sequence: every non-blank letter of the lines up to the first sample id ----------------------------------------------
  std::vector<char> ambiguous;          // [L] 1 where the reference letter was ambiguous (stored as A)
  bool have_line = true;
  for (;;) {
    if (b != e) {
      if (*b == '>') break;
      for (const char* c = b; c != e; ++c) {
        if (is_space(*c)) continue;
        const int bits = letter_bits(*c);
        if (bits == 0) return fail_with("Reference sequence has an invalid state");
        const bool real = (bits & (bits - 1)) == 0;
        M.ref.push_back((uint8_t)(real ? real_letter(bits) : 0));
        ambiguous.push_back(real ? 0 : 1);
      }
    }
    if (!in.next(b, e)) { have_line = false; break; }
  }
  const int L = (int)M.ref.size();
  // ---- the samples -----------------------------------------------------------------------------------------------------------------------
  std::vector<std::pair<int, int>> runs;
  while (have_line) {
    if (b == e || *b != '>') return fail_with("Expected sequence id line to start with '>'");
    const char* ib = b + 1;
    const char* ie = e;
    while (ib < ie && is_space(*ib)) ++ib;
    while (ib < ie && is_space(*(ie - 1))) --ie;
    const std::string name(ib, ie);
    bool ignore = false;
    double t0 = 0.0, t1 = 0.0;
    if (!date_range_of_id(name, t0, t1)) { ++M.num_warnings; ignore = true; }
    const size_t d_mark = M.delta_site.size();
    runs.clear();
    have_line = false;
    while (in.next(b, e)) {
      if (b == e) continue;
      if (*b == '>') { have_line = true; break; }
      Field f{b, e};
      char c = 0;
      f.get_char(c);
      const int bits = f.fail ? 0 : letter_bits(c);
      if (bits == 0) { ignore = true; ++M.num_warnings; continue; }
      if ((bits & (bits - 1)) != 0) {
        // a run of missing / ambiguous sites: 1-based start, optional length (default 1)
        int start = 0; f.get_int(start); --start;
        const bool valid_start = !f.fail;
        int run = 1; f.get_int(run);
        const long long end = (long long)start + run;
        if (valid_start && 0 <= start && start < L && 0 < end && end <= L && start < end) runs.emplace_back(start, (int)end);
        else { ignore = true; ++M.num_warnings; }
      } else {
        int l = 0; f.get_int(l); --l;
        const int to = real_letter(bits);
        const int from = (0 <= l && l < L) ? M.ref[l] : 0;
        if (!f.fail && 0 <= l && l < L && (from != to || from == 3) && !ambiguous[l]) {
          if (!(from == 3 && to == 3)) { M.delta_site.push_back(l); M.delta_from.push_back((uint8_t)from); M.delta_to.push_back((uint8_t)to); }
        } else { ignore = true; ++M.num_warnings; }
      }
    }
    if (ignore) {
      M.delta_site.resize(d_mark); M.delta_from.resize(d_mark); M.delta_to.resize(d_mark);
    } else {
      M.t_min.push_back((double)(float)t0); M.t_max.push_back((double)(float)t1);       // Tip_desc keeps them as float (core/phylo_tree.h:139-140)
      M.names += name; M.name_off.push_back((int64_t)M.names.size());
      M.delta_off.push_back((int32_t)M.delta_site.size());
      std::sort(runs.begin(), runs.end());
      for (size_t i = 0; i < runs.size();) {                                           // union; runs that touch are one run
        int s = runs[i].first, t = runs[i].second;
        size_t j = i + 1;
        while (j < runs.size() && runs[j].first <= t) { t = std::max(t, runs[j].second); ++j; }
        M.miss_start.push_back(s); M.miss_end.push_back(t);
        i = j;
      }
      M.miss_off.push_back((int32_t)M.miss_start.size());
    }
  }
  return DPHY_OK;
}

}  // namespace

extern "C" {

int dphy_maple_parse(const char* text, size_t len, dphy_maple** out) {
  if (!out || (!text && len > 0)) return DPHY_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  auto* m = new (std::nothrow) dphy_maple();
  if (!m) return DPHY_ERR_OUT_OF_MEMORY;
  const int st = parse(text ? text : "", len, *m);
  if (st != DPHY_OK) { delete m; return st; }
  *out = m;
  return DPHY_OK;
}

const char* dphy_maple_last_error(void) { return g_error.c_str(); }

int dphy_maple_get(const dphy_maple* m, dphy_maple_view* v) {
  if (!m || !v) return DPHY_ERR_INVALID_ARGUMENT;
  v->num_sites = (int32_t)m->ref.size(); v->num_tips = (int32_t)m->t_min.size();
  v->num_warnings = m->num_warnings;
  v->ref = m->ref.data(); v->t_min = m->t_min.data(); v->t_max = m->t_max.data();
  v->name_off = m->name_off.data(); v->names = m->names.data();
  v->delta_off = m->delta_off.data(); v->delta_site = m->delta_site.data(); v->delta_from = m->delta_from.data(); v->delta_to = m->delta_to.data();
  v->miss_off = m->miss_off.data(); v->miss_start = m->miss_start.data(); v->miss_end = m->miss_end.data();
  return DPHY_OK;
}

void dphy_maple_free(dphy_maple* m) { delete m; }

}  // extern "C"
