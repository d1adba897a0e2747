// K0: occupation strings and single-excitation link tables (host side).
//
// Replaces pyscf.fci.cistring.{make_strings,str2addr,gen_linkstr_index}, reached
// by the reference through cisolver.trans_rdm12 (evcont/FCI_EVCont.py:121).
// Layout: strings ascending as integers; table rows
// [cre a, des i, address, sign], diagonal rows first, then (occupied i outer,
// virtual a inner).  The table is built once per (norb, nocc) and cached by the
// Python side; PySCF rebuilds it on every trans_rdm12 call.
#include <cstdint>
#include <vector>

#include "evcont_b200.h"

void evc_set_error(const char* fmt, ...);

namespace {

// Pascal triangle up to 64 choose k (fits in int64 for the sizes FCI can reach).
struct Binomials {
  int64_t c[65][65];
  Binomials() {
    for (int n = 0; n <= 64; ++n) {
      c[n][0] = 1;
      for (int k = 1; k <= 64; ++k) {
        if (k > n) { c[n][k] = 0; continue; }
        if (k == n) { c[n][k] = 1; continue; }
        __int128 v = static_cast<__int128>(c[n - 1][k - 1]) + c[n - 1][k];
        c[n][k] = v > INT64_MAX ? INT64_MAX : static_cast<int64_t>(v);
      }
    }
  }
};
const Binomials kBinom;

inline bool valid(int norb, int nocc) { return norb >= 0 && norb <= 62 && nocc >= 0 && nocc <= norb; }

inline int64_t rank_of(int norb, int64_t s) {
  int64_t addr = 0;
  int j = 0;
  for (int o = 0; o < norb; ++o)
    if ((s >> o) & 1) { ++j; addr += kBinom.c[o][j]; }
  return addr;
}

}  // namespace

extern "C" {

int64_t evc_num_strings(int norb, int nocc) { return valid(norb, nocc) ? kBinom.c[norb][nocc] : -1; }

int evc_num_links(int norb, int nocc) { return valid(norb, nocc) ? nocc + nocc * (norb - nocc) : -1; }

int64_t evc_str2addr(int norb, int nocc, int64_t string) {
  if (!valid(norb, nocc) || __builtin_popcountll(string) != nocc || (string >> norb) != 0) return -1;
  return rank_of(norb, string);
}

int64_t evc_addr2str(int norb, int nocc, int64_t addr) {
  if (!valid(norb, nocc) || addr < 0 || addr >= kBinom.c[norb][nocc]) return -1;
  int64_t s = 0;
  int k = nocc;
  for (int o = norb - 1; o >= 0 && k > 0; --o) {
    int64_t c = kBinom.c[o][k];
    if (addr >= c) { s |= int64_t(1) << o; addr -= c; --k; }
  }
  return s;
}

int evc_make_strings_host(int norb, int nocc, int64_t* out) {
  if (!valid(norb, nocc) || out == nullptr) { evc_set_error("evc_make_strings_host: bad arguments norb=%d nocc=%d", norb, nocc); return -1; }
  if (nocc == 0) { out[0] = 0; return 0; }
  int64_t s = (int64_t(1) << nocc) - 1;
  const int64_t limit = int64_t(1) << norb;
  int64_t k = 0;
  while (s < limit) {
    out[k++] = s;
    int64_t c = s & -s;
    int64_t r = s + c;
    s = (((r ^ s) >> 2) / c) | r;
  }
  return 0;
}

int evc_linkindex_build_host(int norb, int nocc, int32_t* out) {
  if (!valid(norb, nocc) || out == nullptr) { evc_set_error("evc_linkindex_build_host: bad arguments norb=%d nocc=%d", norb, nocc); return -1; }
  const int64_t nstr = kBinom.c[norb][nocc];
  if (nstr > INT32_MAX) { evc_set_error("evc_linkindex_build_host: %lld strings overflow int32 addresses", (long long)nstr); return -1; }
  const int nlink = nocc + nocc * (norb - nocc);
  std::vector<int64_t> strs(static_cast<size_t>(nstr));
  evc_make_strings_host(norb, nocc, strs.data());
  for (int64_t k = 0; k < nstr; ++k) {
    const int64_t s0 = strs[k];
    int32_t* row = out + k * nlink * 4;
    for (int o = 0; o < norb; ++o)
      if ((s0 >> o) & 1) { row[0] = o; row[1] = o; row[2] = static_cast<int32_t>(k); row[3] = 1; row += 4; }
    for (int i = 0; i < norb; ++i) {
      if (!((s0 >> i) & 1)) continue;
      for (int a = 0; a < norb; ++a) {
        if ((s0 >> a) & 1) continue;
        const int64_t s1 = (s0 ^ (int64_t(1) << i)) | (int64_t(1) << a);
        const int lo = i < a ? i : a, hi = i < a ? a : i;
        const int64_t between = s0 & ((int64_t(1) << hi) - (int64_t(1) << (lo + 1)));
        row[0] = a; row[1] = i; row[2] = static_cast<int32_t>(rank_of(norb, s1));
        row[3] = (__builtin_popcountll(between) & 1) ? -1 : 1;
        row += 4;
      }
    }
  }
  return 0;
}

int evc_linkindex_pack_host(int64_t nstr, int nlink, const int32_t* tab, int link_major, uint64_t* packed) {
  if (tab == nullptr || packed == nullptr || nstr < 0 || nlink < 0) { evc_set_error("evc_linkindex_pack_host: bad arguments"); return -1; }
  for (int64_t k = 0; k < nstr; ++k) {
    for (int l = 0; l < nlink; ++l) {
      const int32_t* r = tab + 4 * (k * nlink + l);
      const uint64_t rec = static_cast<uint64_t>(static_cast<uint32_t>(r[2])) |
                           (static_cast<uint64_t>(static_cast<uint8_t>(r[0])) << 32) |
                           (static_cast<uint64_t>(static_cast<uint8_t>(r[1])) << 40) |
                           (static_cast<uint64_t>(static_cast<uint8_t>(static_cast<int8_t>(r[3]))) << 48);
      packed[link_major ? (static_cast<int64_t>(l) * nstr + k) : (k * nlink + l)] = rec;
    }
  }
  return 0;
}

}  // extern "C"
