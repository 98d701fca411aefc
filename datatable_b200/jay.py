"""Jay ingest: the reference's binary frame format (src/core/jay/README.md, jay.fbs; reader open_jay.cc:60-330)
read straight into the buffers the engine consumes -- fixed-width columns only (bool8 / int8..64 / float32/64 /
date32 / time64), whose Jay data buffers ARE the reference's NA-sentinel arrays, so a column goes from the mapped
file to HBM in one copy.  The meta section is FlatBuffers; its few tables are walked by hand
here (no flatbuffers dependency): Frame{nrows, ncols, nkeys, columns}, Column{stype, data, strdata, name, nullcount,
stats, type, nrows, buffers, children}, Type{stype, extra}, Buffer{offset, length}.
Host side only (plumbing): no kernels; the upload is torch's pinned/async copy like every other host column."""
import mmap
import struct

import numpy as np

from . import _lib
from ._lib import BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64, DATE32, TIME64

# jay::SType (jay.fbs) -> (engine stype, numpy dtype)
_JAY_STYPE = {0: (BOOL, np.int8), 1: (INT8, np.int8), 2: (INT16, np.int16), 3: (INT32, np.int32), 4: (INT64, np.int64),
              5: (FLOAT32, np.float32), 6: (FLOAT64, np.float64), 9: (DATE32, np.int32), 10: (TIME64, np.int64)}
_JAY_NAMES = {7: "str32", 8: "str64", 11: "void", 12: "arr32", 13: "arr64"}


class JayError(_lib.DtbValueError):
    pass


class _FB:
    """Minimal FlatBuffers reader over a bytes-like object (little-endian, offsets as in the FlatBuffers spec)."""

    def __init__(self, buf, base):
        self.b, self.base = buf, base                       # base: file offset of the meta section

    def u8(self, p): return self.b[self.base + p]
    def u16(self, p): return struct.unpack_from("<H", self.b, self.base + p)[0]
    def i32(self, p): return struct.unpack_from("<i", self.b, self.base + p)[0]
    def u32(self, p): return struct.unpack_from("<I", self.b, self.base + p)[0]
    def u64(self, p): return struct.unpack_from("<Q", self.b, self.base + p)[0]

    def field(self, table, idx):
        """position of field `idx` of the table at `table`, or 0 when absent (default value)"""
        vt = table - self.i32(table)
        if 4 + 2 * idx >= self.u16(vt):
            return 0
        off = self.u16(vt + 4 + 2 * idx)
        return table + off if off else 0

    def indirect(self, p): return p + self.u32(p)

    def string(self, p):
        s = self.indirect(p)
        n = self.u32(s)
        return bytes(self.b[self.base + s + 4: self.base + s + 4 + n]).decode("utf-8")

    def vector(self, p):
        v = self.indirect(p)
        return v + 4, self.u32(v)                           # (position of element 0, count)


def read_meta(buf):
    """Parses the meta section of a Jay file held in `buf` (bytes / mmap).  Returns
    {"nrows", "ncols", "nkeys", "columns": [{"name", "jay_stype", "offset", "length", "nullcount", "nrows"}]};
    offsets are FILE offsets of the column's data buffer."""
    n = len(buf)
    if n < 24 or n % 8 or bytes(buf[:3]) != b"JAY" or bytes(buf[n - 3:n]) != b"JAY":
        raise JayError("not a Jay file (signature / size)")
    if bytes(buf[:8]) != b"JAY1\0\0\0\0" or bytes(buf[n - 8:n]) != b"\0\0\0\0" + b"1JAY":
        raise JayError("unsupported Jay version")
    meta_size = struct.unpack_from("<q", buf, n - 16)[0]
    if meta_size % 8 or meta_size <= 0 or meta_size > n - 24:            # README.md: such a file is invalid
        raise JayError("invalid Jay meta size")
    base = n - 16 - meta_size
    fb = _FB(buf, base)
    frame = fb.indirect(0)
    f_nrows, f_ncols, f_nkeys, f_cols = (fb.field(frame, i) for i in range(4))
    nrows = fb.u64(f_nrows) if f_nrows else 0
    out = {"nrows": nrows, "ncols": fb.u64(f_ncols) if f_ncols else 0, "nkeys": fb.i32(f_nkeys) if f_nkeys else 0,
           "columns": []}
    if not f_cols:
        return out
    first, count = fb.vector(f_cols)
    for i in range(count):
        col = fb.indirect(first + 4 * i)
        f_stype, f_data, f_name, f_null, f_type, f_nr, f_bufs = (fb.field(col, k) for k in (0, 1, 3, 4, 7, 8, 9))
        name = fb.string(f_name) if f_name else f"C{i}"
        if f_type:                                                      # current layout: type + buffers (open_jay.cc:254-330)
            ty = fb.indirect(f_type)
            f_st = fb.field(ty, 0)
            jst = fb.u8(f_st) if f_st else 0
            if not f_bufs:
                raise JayError(f"column `{name}` has no buffers")
            b0, nb = fb.vector(f_bufs)
            if jst in _JAY_STYPE and nb != 2:
                raise JayError(f"column `{name}`: a fixed-width column has a validity and a data buffer, found {nb}")
            data_at = b0 + 16 * (nb - 1 if jst in _JAY_STYPE else 0)    # fixed-width: [validity (may be empty), data]
            off, length = fb.u64(data_at), fb.u64(data_at + 8)
            if jst in _JAY_STYPE and nb == 2 and fb.u64(b0 + 8) != 0:
                raise JayError(f"column `{name}` carries a validity bitmap: outside the GPU hot path")
        else:                                                           # legacy layout: stype + data (open_jay.cc:176-216)
            jst = fb.u8(f_stype) if f_stype else 0
            if not f_data:
                raise JayError(f"column `{name}` has no data buffer")
            off, length = fb.u64(f_data), fb.u64(f_data + 8)
        cn = fb.u64(f_nr) if f_nr else nrows
        out["columns"].append({"name": name, "jay_stype": jst, "offset": 8 + off, "length": length,
                               "nullcount": fb.u64(f_null) if f_null else 0, "nrows": cn or nrows})
    return out


def open_jay(path, columns=None, device=True):
    """Jay file -> Frame.  device=True: every column goes from the mapped file into HBM (one pinned staging copy
    per column, asynchronous upload); device=False: numpy views of a private copy.  `columns`: names to read.
    String / void / array columns raise DtbNotImplError (outside the path) unless left out with `columns`."""
    from .frame import Frame
    with open(path, "rb") as fh:
        mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
    try:
        meta = read_meta(mm)
        cols, sts = {}, {}
        for c in meta["columns"]:
            if columns is not None and c["name"] not in columns:
                continue
            if c["jay_stype"] not in _JAY_STYPE:
                raise _lib.DtbNotImplError(f"Jay column `{c['name']}` of type {_JAY_NAMES.get(c['jay_stype'], c['jay_stype'])} "
                                           "is outside the GPU hot path")
            st, npdt = _JAY_STYPE[c["jay_stype"]]
            nr = c["nrows"]
            if c["length"] != nr * np.dtype(npdt).itemsize or c["offset"] + c["length"] > len(mm):
                raise JayError(f"column `{c['name']}`: buffer of {c['length']} bytes does not hold {nr} rows")
            view = np.frombuffer(mm, dtype=npdt, count=nr, offset=c["offset"])
            if device:
                import torch
                host = torch.empty(nr, dtype=getattr(torch, np.dtype(npdt).name), pin_memory=torch.cuda.is_available())
                host.numpy()[:] = view                                   # file pages -> pinned staging buffer
                cols[c["name"]] = host.cuda(non_blocking=True)
            else:
                cols[c["name"]] = np.array(view)                         # private copy: the map is closed below
            sts[c["name"]] = st
            del view
        if device:
            import torch
            torch.cuda.synchronize()
        fr = Frame(cols, stypes=sts)
        fr._jay_nkeys = meta["nkeys"]
        nk = meta["nkeys"]
        if nk > 0 and all(c["name"] in cols for c in meta["columns"][:nk]):
            # "the Frame is sorted by the first nkeys columns, and those columns have unique values" (README.md):
            # the key is taken over as stored, like open_jay.cc:112 (dt->set_nkeys_unsafe)
            fr._key = tuple(c["name"] for c in meta["columns"][:nk])
        return fr
    finally:
        try:
            mm.close()
        except BufferError:                                              # a numpy view is still alive: let the GC close it
            pass


# ---------------------------------------------------------------------------------------------------------------
# writer (Frame.to_jay of the reference, src/core/jay/save_jay.cc): the result frame of a query, in the format the
# reference opens with dt.fread / dt.open.  Fixed-width columns; the meta section is laid out front to back (every
# FlatBuffers reference points forward) with the alignment the reference's flatbuffers::Verifier checks.
# ---------------------------------------------------------------------------------------------------------------
_TO_JAY_STYPE = {st: j for j, (st, _) in _JAY_STYPE.items()}


class _Meta:
    def __init__(self):
        self.b = bytearray()
        self.patches = []                                    # (position of a uoffset32, its target's name)
        self.at = {}

    def pad(self, align, plus=0):
        while (len(self.b) + plus) % align:
            self.b += b"\0"

    def ref(self, target):                                   # placeholder for a forward uoffset32
        self.patches.append((len(self.b), target))
        self.b += b"\0\0\0\0"

    def finish(self):
        for pos, target in self.patches:
            struct.pack_into("<I", self.b, pos, self.at[target] - pos)
        self.pad(8)
        return bytes(self.b)


def save_jay(frame, path):
    """Frame -> Jay file (fixed-width columns; the key, if any, must be the leading columns like the reference's)."""
    names = list(frame.names)
    nk = len(frame.key)
    if tuple(names[:nk]) != tuple(frame.key):
        raise JayError("a keyed frame is stored with its key columns first")
    data = bytearray(b"JAY1\0\0\0\0")
    cols = []
    for nm in names:
        st = frame._stypes[nm]
        if st not in _TO_JAY_STYPE:
            raise _lib.DtbNotImplError(f"column `{nm}` of stype {st} cannot be written to Jay by this engine")
        a = np.ascontiguousarray(frame.to_numpy(nm))
        raw = a.tobytes()
        if a.dtype.kind == "f":
            nulls = int(np.isnan(a).sum())
        else:
            nulls = int((a == np.iinfo(a.dtype).min).sum())
        cols.append((nm, _TO_JAY_STYPE[st], len(data) - 8, len(raw), nulls))
        data += raw
        data += b"\0" * (-len(data) % 8)
    m = _Meta()
    m.ref("frame")                                           # root uoffset
    # Frame: vtable [12, 28, nrows@8, ncols@16, nkeys@4, columns@24], table 8-aligned
    m.pad(8, plus=12)
    vt = len(m.b)
    m.b += struct.pack("<6H", 12, 28, 8, 16, 4, 24)
    m.at["frame"] = len(m.b)
    m.b += struct.pack("<iiQQ", len(m.b) - vt, nk, frame.nrows, len(names))
    m.ref("columns")
    m.pad(4)
    m.at["columns"] = len(m.b)
    m.b += struct.pack("<I", len(cols))
    for i in range(len(cols)):
        m.ref(f"col{i}")
    for i, (nm, jst, off, length, nulls) in enumerate(cols):
        # Column: ids 3 name@4, 4 nullcount@8, 7 type@24, 8 nrows@16, 9 buffers@28; table size 32, 8-aligned
        m.pad(8, plus=26 + 2)                                # vtable of 26 bytes + 2 bytes of padding before the table
        vt = len(m.b)
        m.b += struct.pack("<13H", 26, 32, 0, 0, 0, 4, 8, 0, 0, 24, 16, 28, 0)
        m.b += b"\0\0"
        m.at[f"col{i}"] = len(m.b)
        m.b += struct.pack("<i", len(m.b) - vt)
        m.ref(f"name{i}")
        m.b += struct.pack("<QQ", nulls, frame.nrows)
        m.ref(f"type{i}")
        m.ref(f"bufs{i}")
        # Type: id 0 stype@4
        m.pad(4, plus=6 + 2)
        vt = len(m.b)
        m.b += struct.pack("<3H", 6, 8, 4) + b"\0\0"
        m.at[f"type{i}"] = len(m.b)
        m.b += struct.pack("<iB3x", len(m.b) - vt, jst)
        # buffers: [validity (empty), data]; the structs are 8-aligned, their count sits right before them
        m.pad(8, plus=4)
        m.at[f"bufs{i}"] = len(m.b)
        m.b += struct.pack("<IQQQQ", 2, 0, 0, off, length)
        enc = nm.encode("utf-8")
        m.pad(4)
        m.at[f"name{i}"] = len(m.b)
        m.b += struct.pack("<I", len(enc)) + enc + b"\0"
        m.pad(4)
    meta = m.finish()
    with open(path, "wb") as fh:
        fh.write(bytes(data) + meta + struct.pack("<q", len(meta)) + b"\0\0\0\0" + b"1JAY")
