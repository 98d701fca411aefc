// dtb200.hpp -- header-only C++ mirror of the reference's own object layout for the
// DT[i, j, by(), sort()] path, on top of the C-ABI (dtb200.h).  This is what the in-extension
// hook of INTEGRATION.md (B) would use: same names, argument meaning and error behaviour as
//
//     RiGb group(const std::vector<Column>&, const std::vector<SortFlag>&, NaPosition)
//                                                        src/core/sort.h:36-58, sort.cc:1411-1495
//     class Groupby  { offsets_, ngroups_; get_group(i,&i0,&i1) }      src/core/groupby.h:54-89
//     class RowIndex { ARR32 indices, size(), max() }                  src/core/rowindex.h:46-196
//     reducer columns materialised over (Column, Groupby)              src/core/column/reduce_unary.h:30-68
//
// Buffers are host memory (std::vector) here, exactly like the reference's Buffers; the engine
// stages them through HBM.  Errors are thrown as dtb::Error (cf. dt::Error, utils/exceptions.h:43).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "dtb200.h"

namespace dtb {

struct Error : std::runtime_error { int code; Error(int c, const std::string& m) : std::runtime_error(m), code(c) {} };
struct NotImplError : Error { using Error::Error; };        // NotImplError, sort.cc:673
struct ValueError : Error { using Error::Error; };

inline void check(int rc) {
  if (rc == DTB_OK) return;
  const std::string msg = dtb_last_error();
  if (rc == DTB_ENOTIMPL) throw NotImplError(rc, msg);
  if (rc == DTB_EINVAL || rc == DTB_ENOSPACE) throw ValueError(rc, msg);
  throw Error(rc, msg);
}

enum class SType : int { BOOL = DTB_STYPE_BOOL, INT8 = DTB_STYPE_INT8, INT16 = DTB_STYPE_INT16,
                         INT32 = DTB_STYPE_INT32, INT64 = DTB_STYPE_INT64,
                         FLOAT32 = DTB_STYPE_FLOAT32, FLOAT64 = DTB_STYPE_FLOAT64 };

enum SortFlag : int { NONE = 0, DESCENDING = DTB_FLAG_DESCENDING, SORT_ONLY = DTB_FLAG_SORT_ONLY };
inline SortFlag operator|(SortFlag a, SortFlag b) { return static_cast<SortFlag>(int(a) | int(b)); }
enum NaPosition : int { FIRST = DTB_NA_FIRST, LAST = DTB_NA_LAST, REMOVE = DTB_NA_REMOVE };

// A material fixed-width column (SentinelFw_ColumnImpl): non-owning view of a typed buffer.
class Column {
  const void* data_; SType stype_; size_t nrows_;
 public:
  Column(const void* data, SType st, size_t nrows) : data_(data), stype_(st), nrows_(nrows) {}
  const void* get_data_readonly() const { return data_; }
  SType stype() const { return stype_; }
  size_t nrows() const { return nrows_; }
};

// ARR32 RowIndex: the ordering produced by group().
class RowIndex {
  std::vector<int32_t> ind_;
 public:
  RowIndex() {}
  explicit RowIndex(std::vector<int32_t>&& v) : ind_(std::move(v)) {}
  size_t size() const { return ind_.size(); }
  const int32_t* indices32() const { return ind_.data(); }
  int32_t operator[](size_t i) const { return ind_[i]; }
};

class Groupby {
  std::vector<int32_t> offsets_; size_t ngroups_ = 0; bool valid_ = false;
 public:
  Groupby() {}
  Groupby(size_t ng, std::vector<int32_t>&& offs) : offsets_(std::move(offs)), ngroups_(ng), valid_(true) {}
  explicit operator bool() const { return valid_; }
  size_t size() const { return ngroups_; }
  const int32_t* offsets_r() const { return offsets_.data(); }
  void get_group(size_t i, size_t* i0, size_t* i1) const {         // groupby.cc:92-97
    *i0 = size_t(offsets_[i]); *i1 = size_t(offsets_[i + 1]);
  }
};

using RiGb = std::pair<RowIndex, Groupby>;

// group(): sort.cc:1411-1495.
inline RiGb group(const std::vector<Column>& columns, const std::vector<SortFlag>& flags,
                  NaPosition na_pos = NaPosition::FIRST, dtb_stream stream = nullptr)
{
  if (columns.empty() || columns.size() != flags.size()) throw ValueError(DTB_EINVAL, "columns/flags mismatch");
  const size_t n = columns[0].nrows();
  std::vector<dtb_col> keys; std::vector<int> fl;
  for (size_t j = 0; j < columns.size(); ++j) {
    keys.push_back(dtb_col{columns[j].get_data_readonly(), int(columns[j].stype()), 0});
    fl.push_back(int(flags[j]));
  }
  std::vector<int32_t> order(n), offsets(n + 1);
  int64_t ng = -1, norder = 0;
  check(dtb_group(keys.data(), int(keys.size()), fl.data(), int(na_pos), int64_t(n), stream,
                  order.data(), offsets.data(), int64_t(n + 1), &ng, &norder));
  order.resize(size_t(norder));
  RiGb res;
  res.first = RowIndex(std::move(order));
  if (ng >= 0) { offsets.resize(size_t(ng) + 1); res.second = Groupby(size_t(ng), std::move(offsets)); }
  return res;
}

// Per-group reducer over `col` viewed through `ri` (empty RowIndex = identity), materialised.
template <typename TOut>
inline std::vector<TOut> reduce(int op, const Column& col, const RowIndex& ri, const Groupby& gby,
                                dtb_stream stream = nullptr)
{
  const int out_st = (op == DTB_OP_NROWS) ? DTB_STYPE_INT64 : dtb_reduce_out_stype(op, int(col.stype()));
  if (!out_st || size_t(dtb_stype_size(out_st)) != sizeof(TOut))
    throw ValueError(DTB_EINVAL, "reducer output type mismatch");
  std::vector<TOut> out(gby.size());
  check(dtb_reduce(op, dtb_col{col.get_data_readonly(), int(col.stype()), 0}, int64_t(col.nrows()),
                   ri.size() ? ri.indices32() : nullptr, 0, gby.offsets_r(), int64_t(gby.size()), stream, out.data()));
  return out;
}

// Column::sort_grouped (sort.cc:1499-1530): the rows of every group reordered by `col` (NA first), as a new
// RowIndex; Median_ColumnImpl / op_nunique read it (reduce<...>(DTB_OP_MEDIAN / DTB_OP_NUNIQUE, col, ri2, gby)).
inline RowIndex sort_grouped(const Column& col, const RowIndex& ri, const Groupby& gby, dtb_stream stream = nullptr)
{
  size_t i0 = 0, n = 0;
  if (gby.size()) gby.get_group(gby.size() - 1, &i0, &n);
  std::vector<int32_t> out(n);
  check(dtb_sort_grouped(dtb_col{col.get_data_readonly(), int(col.stype()), 0}, int64_t(col.nrows()),
                         ri.size() ? ri.indices32() : nullptr, gby.offsets_r(), int64_t(gby.size()), stream, out.data()));
  return RowIndex(std::move(out));
}

// natural_join (frame/join.cc:392-470): for every row of the X key columns the matching row of the keyed
// (sorted, unique) J key columns, or the NA index INT32_MIN.
inline RowIndex natural_join(const std::vector<Column>& xkeys, const std::vector<Column>& jkeys, dtb_stream stream = nullptr)
{
  if (xkeys.empty() || xkeys.size() != jkeys.size()) throw ValueError(DTB_EINVAL, "key columns mismatch");
  std::vector<dtb_col> xs, js;
  for (size_t c = 0; c < xkeys.size(); ++c) {
    xs.push_back(dtb_col{xkeys[c].get_data_readonly(), int(xkeys[c].stype()), 0});
    js.push_back(dtb_col{jkeys[c].get_data_readonly(), int(jkeys[c].stype()), 0});
  }
  std::vector<int32_t> out(xkeys[0].nrows());
  check(dtb_join(xs.data(), js.data(), int(xs.size()), int64_t(xkeys[0].nrows()), int64_t(jkeys[0].nrows()), stream, out.data()));
  return RowIndex(std::move(out));
}

// The `i` node under by() / sort() for an integer slice (expr/fexpr_literal_sliceint.cc:82-170): the slice applied
// inside every group; returns positions into the RowIndex of group() and the Groupby of the groups that remain.
// Missing slice members: DTB_SLICE_NA.
inline RiGb slice_groups(const Groupby& gby, int64_t start, int64_t stop, int64_t step, dtb_stream stream = nullptr)
{
  size_t i0 = 0, nrows = 0;
  if (gby.size()) gby.get_group(gby.size() - 1, &i0, &nrows);
  const int64_t cap = step == 0 ? int64_t(gby.size()) * (stop > 0 ? stop : 0) : int64_t(nrows);
  std::vector<int32_t> rows(size_t(cap > 0 ? cap : 0)), offs(gby.size() + 1);
  int64_t ngo = 0, nro = 0;
  check(dtb_slice_groups(gby.offsets_r(), int64_t(gby.size()), start, stop, step, stream, rows.data(), cap, offs.data(), &ngo, &nro));
  rows.resize(size_t(nro)); offs.resize(size_t(ngo) + 1);
  return RiGb(RowIndex(std::move(rows)), Groupby(size_t(ngo), std::move(offs)));
}

}  // namespace dtb
