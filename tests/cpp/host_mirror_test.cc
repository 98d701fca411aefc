// C++ host mirror (include/dtb200.hpp) exercised the way the reference's group() is used
// (src/core/expr/eval_context.cc:249-288): built by __graft_entry__.build(), run by
// tests/test_gpu_cpp.py on the GPU box.  Prints "OK" on success.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include "dtb200.hpp"

int main() {
  const size_t n = 200000;
  std::vector<int32_t> k(n); std::vector<double> v(n);
  uint64_t x = 88172645463325252ull;
  for (size_t i = 0; i < n; i++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    k[i] = (x % 97 == 0) ? INT32_MIN : int32_t(x % 1000) - 500;      // NA keys form the first group
    v[i] = double((x >> 20) % 1000) / 8.0;
  }
  try {
    dtb::Column kc(k.data(), dtb::SType::INT32, n), vc(v.data(), dtb::SType::FLOAT64, n);
    dtb::RiGb rg = dtb::group({kc}, {dtb::SortFlag::NONE});
    const dtb::RowIndex& ri = rg.first; const dtb::Groupby& gb = rg.second;
    if (!gb || ri.size() != n) { printf("FAIL: no groupby\n"); return 1; }
    // reference semantics: NA first, ascending keys, stable inside groups
    std::map<int64_t, std::pair<double, int64_t>> want;             // key -> (sum, count); NA as INT64_MIN
    for (size_t i = 0; i < n; i++) { auto& w = want[k[i] == INT32_MIN ? INT64_MIN : k[i]]; w.first += v[i]; w.second++; }
    if (gb.size() != want.size()) { printf("FAIL: ngroups %zu vs %zu\n", gb.size(), want.size()); return 1; }
    std::vector<double> sums = dtb::reduce<double>(DTB_OP_SUM, vc, ri, gb);
    std::vector<int64_t> cnts = dtb::reduce<int64_t>(DTB_OP_NROWS, vc, ri, gb);
    size_t g = 0;
    for (auto& kv : want) {
      size_t i0, i1; gb.get_group(g, &i0, &i1);
      for (size_t p = i0; p < i1; p++) {
        const int32_t kk = k[size_t(ri[p])];
        if ((kk == INT32_MIN ? INT64_MIN : int64_t(kk)) != kv.first) { printf("FAIL: key order at group %zu\n", g); return 1; }
        if (p > i0 && ri[p] <= ri[p - 1]) { printf("FAIL: stability at group %zu\n", g); return 1; }
      }
      if (cnts[g] != kv.second.second || std::fabs(sums[g] - kv.second.first) > 1e-9 * std::fabs(kv.second.first) + 1e-9) {
        printf("FAIL: reducer at group %zu\n", g); return 1;
      }
      g++;
    }
    // median per group: sort_grouped + DTB_OP_MEDIAN (head_reduce_unary.cc:421-468), checked against std::sort
    dtb::RowIndex ri2 = dtb::sort_grouped(vc, ri, gb);
    std::vector<double> med = dtb::reduce<double>(DTB_OP_MEDIAN, vc, ri2, gb);
    for (size_t gg = 0; gg < gb.size(); gg += 97) {
      size_t i0, i1; gb.get_group(gg, &i0, &i1);
      std::vector<double> w;
      for (size_t p = i0; p < i1; p++) w.push_back(v[size_t(ri[p])]);
      std::sort(w.begin(), w.end());
      const size_t m = w.size();
      const double want_med = (m & 1) ? w[m / 2] : (w[m / 2] + w[m / 2 - 1]) / 2;
      if (med[gg] != want_med) { printf("FAIL: median of group %zu\n", gg); return 1; }
    }
    // natural join of the key column against its own sorted unique values (frame/join.cc:392-470)
    std::vector<int32_t> uniq;
    for (size_t gg = 0; gg < gb.size(); gg++) { size_t i0, i1; gb.get_group(gg, &i0, &i1); uniq.push_back(k[size_t(ri[i0])]); }
    dtb::Column jc(uniq.data(), dtb::SType::INT32, uniq.size());
    dtb::RowIndex jr = dtb::natural_join({kc}, {jc});
    for (size_t i = 0; i < n; i += 1013)
      if (jr[i] < 0 || uniq[size_t(jr[i])] != k[i]) { printf("FAIL: join at row %zu\n", i); return 1; }
    // DT[-2:, :, by(k)]: the last two rows of every group (expr/fexpr_literal_sliceint.cc:82-170)
    dtb::RiGb sl = dtb::slice_groups(gb, -2, DTB_SLICE_NA, DTB_SLICE_NA);
    if (sl.second.size() != gb.size()) { printf("FAIL: slice_groups dropped a group\n"); return 1; }
    for (size_t gg = 0; gg < gb.size(); gg += 89) {
      size_t i0, i1, j0, j1; gb.get_group(gg, &i0, &i1); sl.second.get_group(gg, &j0, &j1);
      const size_t want_n = (i1 - i0) < 2 ? (i1 - i0) : 2;
      if (j1 - j0 != want_n) { printf("FAIL: slice_groups size of group %zu\n", gg); return 1; }
      for (size_t q = 0; q < want_n; q++)
        if (size_t(sl.first[j0 + q]) != i1 - want_n + q) { printf("FAIL: slice_groups rows of group %zu\n", gg); return 1; }
    }
    // unsupported stype -> NotImplError, like sort.cc:673
    bool threw = false;
    try { dtb::Column sc(k.data(), static_cast<dtb::SType>(11), n); dtb::group({sc}, {dtb::SortFlag::NONE}); }
    catch (const dtb::NotImplError&) { threw = true; }
    if (!threw) { printf("FAIL: str32 key did not raise NotImplError\n"); return 1; }
  } catch (const dtb::Error& e) { printf("FAIL: %s\n", e.what()); return 1; }
  printf("OK\n");
  return 0;
}
