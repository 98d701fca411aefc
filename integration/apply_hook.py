#!/usr/bin/env python3
"""Applies the in-extension hook of INTEGRATION.md (B) to a SCRATCH COPY of the reference.

    cp -r /root/reference /tmp/dthook
    python integration/apply_hook.py /tmp/dthook            # edits src/core/sort.cc and ci/ext.py
    (cd /tmp/dthook && python ci/ext.py build)              # the reference's own backend, ~2 min
    PYTHONPATH=/tmp/dthook/src python integration/check_hook.py

The script only INSERTS our own code next to short anchor strings; it carries no reference source.
What it adds:
  * src/core/sort.cc  -- `#include <dtb200.h>`, option `sort.b200` (bool, default False) and, inside
    `group()` just before the CPU SortContext is built, a call to `dtb_group` for material
    bool/int/float key columns whose outputs are wrapped exactly like get_result_rowindex() /
    extract_groups() do (sort.cc:596-616).  DTB_ENOTIMPL falls through to the CPU path (outside the
    named path); any other error is raised as a dt RuntimeError with dtb_last_error().
  * ci/ext.py         -- include path of include/dtb200.h, link + rpath of datatable_b200/lib/libdtb200.so.
/root/reference itself is never touched.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INCLUDE_ANCHOR = '#include "utils/misc.h"\n'
INCLUDE_ADD = '#include <dtb200.h>      // B200 engine C-ABI (drop-in for group()); -I<repo>/include, see ci/ext.py\n'

OPTION_ANCHOR = 'static bool sort_new = false;\n'
OPTION_ADD = 'static bool sort_b200 = false;   // option sort.b200: route group() through libdtb200\n'

REGISTER_ANCHOR = '  dt::register_option(\n    "sort.new",'
REGISTER_ADD = '''  dt::register_option(
    "sort.b200",
    []{ return py::obool(sort_b200); },
    [](const py::Arg& value) {
      sort_b200 = value.to_bool_strict();
    },
    nullptr);

'''

HOOK_ANCHOR = '  bool do_groups = n > 1 || !(flags[0] & SortFlag::SORT_ONLY);\n  SortContext sc('
HOOK_ADD = '''  // ---- dtb200: the whole ordering + grouping on the GPU ---------------------------------
  if (sort_b200 && nrows <= 0x7FFFFFFFu) {
    bool eligible = true;
    std::vector<dtb_col> b200_keys(n);
    std::vector<int> b200_flags(n);
    for (size_t j = 0; j < n; ++j) {
      const Column& cj = columns[j];
      switch (cj.stype()) {
        case dt::SType::BOOL: case dt::SType::INT8: case dt::SType::INT16: case dt::SType::INT32:
        case dt::SType::INT64: case dt::SType::FLOAT32: case dt::SType::FLOAT64: break;
        default: eligible = false;
      }
      if (!eligible) break;
      b200_keys[j].data = cj.get_data_readonly();                 // material after the loop above
      b200_keys[j].stype = static_cast<int>(cj.stype());          // SType values == DtStype_* codes
      b200_keys[j].reserved = 0;
      b200_flags[j] = static_cast<int>(flags[j]) & (DTB_FLAG_DESCENDING | DTB_FLAG_SORT_ONLY);  // same bit values
    }
    if (eligible) {
      Buffer b200_order = Buffer::mem(nrows * sizeof(int32_t));
      Buffer b200_offsets = Buffer::mem((nrows + 1) * sizeof(int32_t));
      int64_t b200_ng = -1, b200_norder = 0;
      int rc = dtb_group(b200_keys.data(), static_cast<int>(n), b200_flags.data(),
                         static_cast<int>(na_pos),                // NaPosition FIRST/LAST/REMOVE = 1/2/3 = DTB_NA_*
                         static_cast<int64_t>(nrows), nullptr,
                         b200_order.xptr(), b200_offsets.xptr(), static_cast<int64_t>(nrows + 1),
                         &b200_ng, &b200_norder);
      if (rc == DTB_OK) {
        b200_order.resize(static_cast<size_t>(b200_norder) * sizeof(int32_t));
        result.first = RowIndex(std::move(b200_order), RowIndex::ARR32);
        if (b200_ng >= 0) {
          b200_offsets.resize(static_cast<size_t>(b200_ng + 1) * sizeof(int32_t));
          result.second = Groupby(static_cast<size_t>(b200_ng), std::move(b200_offsets));
        }
        return result;
      }
      if (rc != DTB_ENOTIMPL) {
        throw RuntimeError() << "dtb200: " << dtb_last_error();
      }
    }
  }

'''

EXT_ANCHOR = '            ext.compiler.add_linker_flag("-lstdc++")\n'
EXT_ADD = '''            # dtb200: B200 engine C-ABI
            ext.compiler.add_compiler_flag("-I{inc}")
            ext.compiler.add_linker_flag("-L{lib}", "-ldtb200", "-Wl,-rpath,{lib}")
'''


def insert(text, anchor, add, before, what):
    if add in text:
        return text                                          # already applied
    i = text.find(anchor)
    if i < 0:
        sys.exit(f"apply_hook: anchor for {what} not found -- is this the expected reference revision?")
    if text.find(anchor, i + 1) >= 0:
        sys.exit(f"apply_hook: anchor for {what} is not unique")
    at = i if before else i + len(anchor)
    return text[:at] + add + text[at:]


def main():
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    tree = os.path.abspath(sys.argv[1])
    if tree.startswith("/root/reference"):
        sys.exit("apply_hook: refusing to modify /root/reference; work on a scratch copy")
    sort_cc = os.path.join(tree, "src", "core", "sort.cc")
    s = open(sort_cc).read()
    s = insert(s, INCLUDE_ANCHOR, INCLUDE_ADD, before=False, what="include")
    s = insert(s, OPTION_ANCHOR, OPTION_ADD, before=False, what="option flag")
    s = insert(s, REGISTER_ANCHOR, REGISTER_ADD, before=True, what="option registration")
    s = insert(s, HOOK_ANCHOR, HOOK_ADD, before=True, what="group() hook")
    open(sort_cc, "w").write(s)
    ext_py = os.path.join(tree, "ci", "ext.py")
    e = open(ext_py).read()
    add = EXT_ADD.format(inc=os.path.join(ROOT, "include"), lib=os.path.join(ROOT, "datatable_b200", "lib"))
    e = insert(e, EXT_ANCHOR, add, before=False, what="build flags")
    open(ext_py, "w").write(e)
    print("apply_hook: patched", sort_cc, "and", ext_py)


if __name__ == "__main__":
    main()
