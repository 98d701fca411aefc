"""
Host-side mirror of the reference's Python surface for the DT[i, j, by(), sort()] path:
Frame, f, by(), sort(), sum/mean/min/max/count.  Only what the hot path needs -- no fread,
no general expression engine (out of scope).

    reference                                   here
    ---------                                   ----
    dt.Frame                 src/datatable/frame.py:23 / src/core/frame/      Frame
    f.A, f["A"], -f.A        src/datatable/expr/                               f / ColRef
    by(...), sort(...)       src/core/expr/py_by.cc, py_sort.cc:40-110         by / sort
    dt.sum/mean/min/max/count  src/datatable/expr/reduce.py:49-153             sum_/mean/min_/max_/count
    DT[i, j, by, sort]       src/core/frame/__getitem__.cc:47-194,
                             src/core/expr/eval_context.cc:144-288, 473-520    Frame.__getitem__

Evaluation follows EvalContext: group() on the by/sort columns -> (RowIndex, Groupby);
reducers are evaluated over (value column, RowIndex, Groupby); plain columns are gathered
through the RowIndex; group keys are the first row of every group
(eval_context.cc:473-485).  All of it runs in libdtb200.so on the GPU.
"""
import numpy as np

from . import _lib, engine
from ._lib import (BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64, FLAG_DESCENDING,
                   FLAG_SORT_ONLY, NA_FIRST, NA_LAST, NA_REMOVE)

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_NA_POS = {"first": NA_FIRST, "last": NA_LAST, "remove": NA_REMOVE}
_NA_VALUE = {INT8: -2**7, INT16: -2**15, INT32: -2**31, INT64: -2**63, BOOL: -128}


# ---------------------------------------------------------------------------
# f-expressions (only column references, their negation, and reducers)
# ---------------------------------------------------------------------------
class ColRef:
    def __init__(self, name, negated=False):
        self.name = name
        self.negated = negated

    def __neg__(self):            # sort(-f.A) / by(-f.A): DESCENDING flag, not arithmetic (fexpr_list.cc:346-358)
        return ColRef(self.name, not self.negated)

    def __repr__(self):
        return f"{'-' if self.negated else ''}f.{self.name}"


class _FNamespace:
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return ColRef(name)

    def __getitem__(self, name):
        return ColRef(name)


f = _FNamespace()


class Reducer:
    def __init__(self, op, name, arg):
        self.op, self.opname, self.arg = op, name, arg


def sum(x): return Reducer(_lib.OP_SUM, "sum", x)          # noqa: A001  (mirrors dt.sum)
def mean(x): return Reducer(_lib.OP_MEAN, "mean", x)
def min(x): return Reducer(_lib.OP_MIN, "min", x)          # noqa: A001
def max(x): return Reducer(_lib.OP_MAX, "max", x)          # noqa: A001
def countna(x): return Reducer(_lib.OP_COUNTNA, "countna", x)


def count(x=None):
    return Reducer(_lib.OP_NROWS if x is None else _lib.OP_COUNT, "count", x)


class by:
    def __init__(self, *cols):
        self.cols = [_as_ref(c) for c in _flatten(cols)]


class sort:
    """sort(*cols, reverse=False, na_position="first") -- py_sort.cc:40-110."""

    def __init__(self, *cols, reverse=False, na_position="first"):
        self.cols = [_as_ref(c) for c in _flatten(cols)]
        n = len(self.cols)
        if isinstance(reverse, (list, tuple)):
            if len(reverse) != n:
                raise ValueError(f"number of elements (nflags={len(reverse)}) in the reverse flag list "
                                 f"does not match the number of sort columns ({n})")
            self.reverse = [bool(r) for r in reverse]
        elif isinstance(reverse, bool):
            self.reverse = [reverse] * n
        else:
            raise TypeError("reverse should be a boolean or a list of booleans")
        if na_position not in _NA_POS:
            raise ValueError(f"na position value `{na_position}` is not supported")
        self.na_position = na_position


def _flatten(cols):
    out = []
    for c in cols:
        if isinstance(c, (list, tuple)):
            out.extend(_flatten(c))
        else:
            out.append(c)
    return out


def _as_ref(c):
    if isinstance(c, ColRef):
        return c
    if isinstance(c, str):
        return ColRef(c)
    raise TypeError(f"Unsupported key expression {c!r}: only column references are on the GPU path")


# ---------------------------------------------------------------------------
# Frame
# ---------------------------------------------------------------------------
class Frame:
    """Column store: every column is a numpy array (host) or a torch CUDA tensor (HBM) plus an stype.
    Bool columns with NAs are int8 with -128 (the reference's bool8 layout)."""

    def __init__(self, data=None, stypes=None, **kwargs):
        self._cols = {}
        self._stypes = {}
        if data is None:
            data = kwargs
        if isinstance(data, Frame):
            self._cols, self._stypes = dict(data._cols), dict(data._stypes)
            return
        if not isinstance(data, dict):
            data = {"C0": data}
        n = None
        for name, col in data.items():
            st = None if stypes is None else stypes.get(name)
            if isinstance(col, (list, tuple)):
                col, st2 = _from_list(col)
                st = st or st2
            c = engine.Col(col, st)
            if n is None:
                n = c.nrows
            elif c.nrows != n:
                raise ValueError("columns have different numbers of rows")
            self._cols[name] = c.data
            self._stypes[name] = c.stype
        self._nrows = n or 0

    # -- metadata ---------------------------------------------------------------
    @property
    def names(self): return tuple(self._cols.keys())
    @property
    def nrows(self): return self._nrows
    @property
    def ncols(self): return len(self._cols)
    @property
    def shape(self): return (self.nrows, self.ncols)
    @property
    def stypes(self): return tuple(self._stypes[n] for n in self._cols)

    def _col(self, name):
        if name not in self._cols:
            raise KeyError(f"Column `{name}` does not exist in the Frame")
        return engine.Col(self._cols[name], self._stypes[name])

    def column(self, name):
        """Raw storage of a column (numpy array or CUDA tensor)."""
        return self._cols[name]

    def to_numpy(self, name=None):
        if name is None:
            return {n: self.to_numpy(n) for n in self._cols}
        c = self._cols[name]
        return c.cpu().numpy() if engine.is_tensor(c) else c

    def to_list(self):
        out = []
        for n in self._cols:
            a = self.to_numpy(n)
            st = self._stypes[n]
            if st in (FLOAT32, FLOAT64):
                out.append([None if np.isnan(x) else float(x) for x in a.tolist()])
            elif st == BOOL:
                out.append([None if x == -128 else bool(x) for x in a.tolist()])
            else:
                na = _NA_VALUE[st]
                out.append([None if x == na else int(x) for x in a.tolist()])
        return out

    def to_dict(self):
        return dict(zip(self.names, self.to_list()))

    def to_device(self):
        """Copy every column into HBM (the analogue of a device-backed Buffer)."""
        fr = Frame()
        for n, c in self._cols.items():
            fr._cols[n] = c if engine.is_tensor(c) else torch.from_numpy(np.ascontiguousarray(c)).cuda()
            fr._stypes[n] = self._stypes[n]
        fr._nrows = self._nrows
        return fr

    # -- DT.sort(cols) (sort.cc:1544-1574) ------------------------------------------
    def sort(self, *cols):
        return self[:, :, sort(*cols)]

    # -- DT[i, j, by, sort] -------------------------------------------------------------
    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = (slice(None), item)
        if len(item) < 2:
            raise ValueError("Frame[...] needs at least i and j")
        i, j = item[0], item[1]
        by_, sort_ = None, None
        for m in item[2:]:
            if isinstance(m, by):
                by_ = m
            elif isinstance(m, sort):
                sort_ = m
            else:
                raise TypeError(f"Unsupported modifier {m!r}")
        if not (isinstance(i, slice) and i == slice(None)):
            raise NotImplementedError("row filters are outside the GPU hot path (use i = :)")
        return _evaluate(self, j, by_, sort_)


def _from_list(lst):
    """Python list -> (array, stype) with None as NA (bool8 / int32 / int64 / float64 like the reference)."""
    vals = [x for x in lst if x is not None]
    if vals and all(isinstance(x, bool) for x in vals):
        return np.array([-128 if x is None else int(x) for x in lst], dtype=np.int8), BOOL
    if all(isinstance(x, int) for x in vals):
        big = any(abs(x) > 2**31 - 1 for x in vals)
        dt_, na = (np.int64, -2**63) if big else (np.int32, -2**31)
        return np.array([na if x is None else x for x in lst], dtype=dt_), (INT64 if big else INT32)
    return np.array([np.nan if x is None else float(x) for x in lst], dtype=np.float64), FLOAT64


def _evaluate(DT, j, by_, sort_):
    """EvalContext::evaluate (eval_context.cc:144-172) for the hot-path shapes."""
    # Host columns are uploaded once (pinned memory -> DMA), the whole query then runs on
    # HBM-resident buffers, and only the result frame travels back.
    if torch is None or not torch.cuda.is_available():
        raise _lib.DtbCudaError("no usable CUDA device: datatable_b200 has no CPU fallback")
    host_frame = not any(engine.is_tensor(c) and c.is_cuda for c in DT._cols.values())
    cache = {}
    names, exprs = _resolve_j(DT, j)

    # Host columns: start every upload the query needs on a copy stream, key columns first, so
    # that the PCIe transfer of the value columns overlaps the sort of the keys; each column is
    # awaited (stream event) only where it is first used.
    needed = []
    for m in (by_, sort_):
        if m is not None:
            needed += [r.name for r in m.cols]
    for e in exprs:
        nm = e.arg.name if isinstance(e, Reducer) and e.arg is not None else getattr(e, "name", None)
        if nm is not None:
            needed.append(nm)
    copy_stream = None
    pending = {}
    for nm in dict.fromkeys(needed):
        c = DT._col(nm)
        t = c.data if engine.is_tensor(c.data) else torch.from_numpy(c.data)
        if t.is_cuda:
            continue
        if copy_stream is None:
            copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(copy_stream):
            d = t.cuda(non_blocking=True)                           # pinned host memory -> async DMA
            ev = torch.cuda.Event(); ev.record(copy_stream)
        pending[nm] = (engine.Col(d, c.stype), ev)

    def dcol(name):
        if name not in cache:
            if name in pending:
                c, ev = pending[name]
                torch.cuda.current_stream().wait_event(ev)
                c.data.record_stream(torch.cuda.current_stream())
            else:
                c = DT._col(name)
                t = c.data if engine.is_tensor(c.data) else torch.from_numpy(c.data)
                if not t.is_cuda:
                    c = engine.Col(t.cuda(non_blocking=True), c.stype)
            cache[name] = c
        return cache[name]

    # ---- compute_groupby_and_sort (eval_context.cc:249-288) ----
    keycols, flags = [], []
    na_pos = NA_FIRST
    if by_ is not None:
        for ref in by_.cols:
            keycols.append(dcol(ref.name))
            flags.append(FLAG_DESCENDING if ref.negated else 0)
    if sort_ is not None:
        na_pos = _NA_POS[sort_.na_position]
        for ref, rev in zip(sort_.cols, sort_.reverse):
            keycols.append(dcol(ref.name))
            desc = (not rev) if ref.negated else rev                 # fexpr_list.cc:346-358
            flags.append((FLAG_DESCENDING if desc else 0) | FLAG_SORT_ONLY)
    order = offsets = None
    ngroups = None
    gb = None
    has_reducer = any(isinstance(e, Reducer) for e in exprs)
    if keycols:
        if by_ is not None and has_reducer:
            # RowIndex + Groupby stay in HBM behind a handle; reducers go through it
            # the reducers of j are known before group() runs: hand them over so that the engine can
            # overlap them with the sort (dtb_groupby_create_reduce)
            reds = [(e.op, None if e.arg is None else dcol(e.arg.name)) for e in exprs if isinstance(e, Reducer)]
            gb = engine.Groupby(keycols, flags, na_pos, reducers=reds)
            ngroups = gb.ngroups
        else:
            order, offsets, ngroups = engine.group(keycols, flags, na_pos)

    # ---- j ----
    out = Frame()

    def add(name, data, st):
        base, k = name, 0
        while name in out._cols:
            k += 1
            name = f"{base}.{k - 1}"
        if host_frame and engine.is_tensor(data):
            data = data.cpu().numpy()
        out._cols[name] = data
        out._stypes[name] = st

    if by_ is not None:
        if has_reducer:
            # group keys = first row of every group (get_group_rowindex, eval_context.cc:124-135)
            first = gb.first_rows()
            for ref in by_.cols:
                c = dcol(ref.name)
                add(ref.name, engine.gather(c, first), c.stype)
            ired = 0
            for name, e in zip(names, exprs):
                if isinstance(e, Reducer):
                    add(name, gb.reduced(ired), _red_stype(dcol, e))
                    ired += 1
                else:
                    raise NotImplementedError("mixing reducers and plain columns under by() is outside the hot path")
            for n_ in out._cols:
                if out._stypes[n_] is None:
                    out._stypes[n_] = engine.Col(out._cols[n_]).stype
            out._nrows = ngroups
            gb.close()
            return out
        # by() without reducers: every row, grouped order, key columns first
        bynames = [r.name for r in by_.cols]
        for ref in by_.cols:
            c = dcol(ref.name)
            add(ref.name, engine.gather(c, order), c.stype)
        for name, e in zip(names, exprs):
            if e.name in bynames and j_is_all(j):
                continue
            c = dcol(e.name)
            add(name, engine.gather(c, order), c.stype)
        out._nrows = len(order)
        return out

    if has_reducer:
        # reducers without by(): one group over all rows (Groupby::single_group, groupby.cc:60-68)
        nrows = DT.nrows if order is None else len(order)
        offs = torch.tensor([0, nrows], dtype=torch.int32, device="cuda")
        for name, e in zip(names, exprs):
            if not isinstance(e, Reducer):
                raise NotImplementedError("mixing reducers and plain columns is outside the hot path")
            add(name, _reduce(dcol, e, order, offs), _red_stype(dcol, e))
        out._nrows = 1
        return out

    for name, e in zip(names, exprs):
        if order is None:
            c = DT._col(e.name)
            out._cols[name] = c.data; out._stypes[name] = c.stype
        else:
            c = dcol(e.name)
            add(name, engine.gather(c, order), c.stype)
    out._nrows = DT.nrows if order is None else len(order)
    return out


def j_is_all(j):
    return isinstance(j, slice) and j == slice(None)


def _resolve_j(DT, j):
    if j_is_all(j):
        return list(DT.names), [ColRef(n) for n in DT.names]
    if isinstance(j, dict):
        return list(j.keys()), [_as_expr(v) for v in j.values()]
    if isinstance(j, (list, tuple)):
        es = [_as_expr(v) for v in j]
    else:
        es = [_as_expr(j)]
    names = []
    for e in es:
        if isinstance(e, Reducer):
            names.append("count" if e.arg is None else e.arg.name)     # reducers keep the column's name
        else:
            names.append(e.name)
    return names, es


def _as_expr(v):
    if isinstance(v, (Reducer, ColRef)):
        return v
    if isinstance(v, str):
        return ColRef(v)
    raise TypeError(f"Unsupported j expression {v!r}")


def _index_through(order, pos):
    """order[pos] -- composition of RowIndexes (rowindex.cc:246-250) done as a gather."""
    return engine.gather(engine.Col(order, INT32), pos)


def _red_stype(dcol, e):
    """Output stype of a reducer column (bool8 min/max stay bool8, fexpr_minmax.cc:50-72)."""
    if e.op == _lib.OP_NROWS or e.arg is None:
        return INT64
    return engine.reduce_out_stype(e.op, dcol(e.arg.name).stype)


def _reduce(dcol, e, order, offsets):
    if e.op == _lib.OP_NROWS:
        return engine.reduce(e.op, None, order, offsets)
    return engine.reduce(e.op, dcol(e.arg.name), order, offsets)


# ---------------------------------------------------------------------------
# First consumers of group() beyond DT[i, j, by, sort]
# ---------------------------------------------------------------------------
def unique(frame):
    """dt.unique(frame): the sorted unique values (NA first) of a single-column frame --
    group() + first row of every group, as src/core/set_funcs.cc:100-140 does."""
    if frame.ncols != 1:
        raise NotImplementedError("unique() of a multi-column frame (set union) is outside the GPU hot path")
    name = frame.names[0]
    c = frame._col(name)
    host = not (engine.is_tensor(c.data) and c.data.is_cuda)
    t = c.data if engine.is_tensor(c.data) else torch.from_numpy(c.data)
    cd = engine.Col(t if t.is_cuda else t.cuda(), c.stype)
    gb = engine.Groupby([cd], [0], NA_FIRST)
    vals = engine.gather(cd, gb.first_rows())
    gb.close()
    out = Frame()
    out._cols[name] = vals.cpu().numpy() if host else vals
    out._stypes[name] = c.stype
    out._nrows = int(vals.shape[0])
    return out


def nunique(frame):
    """Frame.nunique(): number of distinct non-NA values per column (stats.cc:949-1010 via group())."""
    out = Frame()
    for name in frame.names:
        c = frame._col(name)
        t = c.data if engine.is_tensor(c.data) else torch.from_numpy(c.data)
        cd = engine.Col(t if t.is_cuda else t.cuda(), c.stype)
        gb = engine.Groupby([cd], [0], NA_FIRST, reducers=[(_lib.OP_COUNT, cd)])
        valid = gb.reduced(0)                      # groups whose key is NA have count(key) == 0
        n = int((valid > 0).sum().item()) if gb.ngroups > 0 else 0
        gb.close()
        out._cols[name] = np.array([n], dtype=np.int64)
        out._stypes[name] = INT64
    out._nrows = 1
    return out
