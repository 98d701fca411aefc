#!/usr/bin/env python
"""
Writes tests/golden/jay_v1.jay / jay_keyed.jay with the *reference itself* (Frame.to_jay, src/core/jay/save_jay.cc)
and the columns' values as NA-sentinel arrays in jay_expected.npz:

    PYTHONPATH=oracle/_ref python tests/golden/make_golden_jay.py

datatable_b200/jay.py (the Jay ingest) must read exactly these buffers.
"""
import datetime
import os

import numpy as np

import datatable as dt

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(3)
n = 1000
NA = {"i8": -2**7, "i16": -2**15, "i32": -2**31, "i64": -2**63}
cols = {
    "b": [None if i % 17 == 0 else bool(i % 3 == 0) for i in range(n)],
    "i8": [None if i % 13 == 0 else (i % 200) - 100 for i in range(n)],
    "i16": [None if i % 11 == 0 else (i * 37) % 30000 - 15000 for i in range(n)],
    "i32": [None if i % 7 == 0 else int(x) for i, x in enumerate(rng.integers(-10**9, 10**9, n))],
    "i64": [None if i % 5 == 0 else int(x) for i, x in enumerate(rng.integers(-10**17, 10**17, n))],
    "f32": [None if i % 19 == 0 else float(np.float32(x)) for i, x in enumerate(rng.standard_normal(n))],
    "f64": [None if i % 23 == 0 else float(x) for i, x in enumerate(rng.standard_normal(n))],
    "d32": [None if i % 31 == 0 else datetime.date(2000, 1, 1) + datetime.timedelta(days=int(x)) for i, x in enumerate(rng.integers(-5000, 9000, n))],
    "s": [None if i % 29 == 0 else "s%d" % (i % 50) for i in range(n)],
}
DT = dt.Frame(cols, stypes={"b": dt.bool8, "i8": dt.int8, "i16": dt.int16, "i32": dt.int32, "i64": dt.int64,
                            "f32": dt.float32, "f64": dt.float64})
DT.to_jay(os.path.join(HERE, "jay_v1.jay"))
exp = {}
exp["b"] = np.array([-128 if x is None else int(x) for x in cols["b"]], dtype=np.int8)
for nm, npdt in (("i8", np.int8), ("i16", np.int16), ("i32", np.int32), ("i64", np.int64)):
    exp[nm] = np.array([NA[nm] if x is None else x for x in cols[nm]], dtype=npdt)
exp["f32"] = np.array([np.nan if x is None else x for x in cols["f32"]], dtype=np.float32)
exp["f64"] = np.array([np.nan if x is None else x for x in cols["f64"]], dtype=np.float64)
epoch = datetime.date(1970, 1, 1)
exp["d32"] = np.array([-2**31 if x is None else (x - epoch).days for x in cols["d32"]], dtype=np.int32)
exp["stypes"] = np.array([str(s) for s in DT.stypes])

K = dt.Frame(k=list(range(0, 600, 3)), v=[float(i) * 0.5 for i in range(200)], stypes={"k": dt.int32})
K.key = "k"
K.to_jay(os.path.join(HERE, "jay_keyed.jay"))
exp["keyed.k"] = np.arange(0, 600, 3, dtype=np.int32)
exp["keyed.v"] = np.arange(200, dtype=np.float64) * 0.5
# a 0-row frame and a time64 column (expected values are written in tests/test_jay.py)
dt.Frame(a=[], b=[], stypes={"a": dt.int32, "b": dt.float64}).to_jay(os.path.join(HERE, "jay_empty.jay"))
dt.Frame(t=[datetime.datetime(2020, 1, 1, 12, 0, 0), None, datetime.datetime(1969, 12, 31, 23, 59, 59)],
         x=[1, 2, 3]).to_jay(os.path.join(HERE, "jay_time.jay"))
np.savez_compressed(os.path.join(HERE, "jay_expected.npz"), **exp)
print("wrote", os.path.getsize(os.path.join(HERE, "jay_v1.jay")), os.path.getsize(os.path.join(HERE, "jay_keyed.jay")), "bytes;", DT.stypes)
