"""CPU: host-side logic of the Frame mirror that needs no GPU -- Arrow ingest/export with NA sentinels,
list ingest, key validation messages."""
import numpy as np
import pytest


def test_arrow_roundtrip_na_sentinels():
    pa = pytest.importorskip("pyarrow")
    import datatable_b200 as dtb
    from datatable_b200._lib import BOOL, INT8, INT32, INT64, FLOAT32, FLOAT64
    t = pa.table({"b": pa.array([True, None, False]), "i8": pa.array([1, None, -3], type=pa.int8()),
                  "i": pa.array([5, 6, None], type=pa.int32()), "l": pa.array([None, 2**40, 7], type=pa.int64()),
                  "f": pa.array([1.5, None, 2.5], type=pa.float32()), "d": pa.array([None, 0.25, -1.0])})
    F = dtb.Frame.from_arrow(t)
    assert F.names == ("b", "i8", "i", "l", "f", "d") and F.stypes == (BOOL, INT8, INT32, INT64, FLOAT32, FLOAT64)
    assert F.to_list() == [[True, None, False], [1, None, -3], [5, 6, None], [None, 2**40, 7],
                           [1.5, None, 2.5], [None, 0.25, -1.0]]
    assert F.to_numpy("b").tolist() == [1, -128, 0] and F.to_numpy("i").tolist() == [5, 6, -2**31]
    back = F.to_arrow()
    assert back.to_pydict() == t.to_pydict()
    with pytest.raises(NotImplementedError):
        dtb.Frame.from_arrow(pa.table({"s": pa.array(["a", "b"])}))


def test_key_and_join_argument_errors():
    import datatable_b200 as dtb
    F = dtb.Frame({"k": np.array([1, 2, 3], np.int32), "v": np.array([1.0, 2.0, 3.0])})
    with pytest.raises(KeyError):
        F.key = "nope"
    with pytest.raises(ValueError, match="multiple times"):
        F.key = ["k", "k"]
    with pytest.raises(TypeError):
        dtb.join(3)
    with pytest.raises(ValueError, match="not keyed"):
        dtb.join(F)
    assert F.key == ()
