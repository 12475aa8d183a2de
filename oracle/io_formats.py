"""TEST INFRASTRUCTURE — binary formats shared by oracle/ref_harness.cc and the python tests.

.fmat       : int32 rows, int32 cols, float32[rows*cols] row-major
record file : repeated { int32 name_len, name, int32 dtype (0=i32, 1=f32, 2=i64, 3=f64), int64 count, data }
"""
import struct

import numpy as np

_DT = {0: np.int32, 1: np.float32, 2: np.int64, 3: np.float64}


def write_fmat(path, a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", a.shape[0], a.shape[1]))
        f.write(a.tobytes())


def read_fmat(path):
    with open(path, "rb") as f:
        r, c = struct.unpack("<ii", f.read(8))
        return np.frombuffer(f.read(r * c * 4), dtype=np.float32).reshape(r, c).copy()


def read_records(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    while o < len(buf):
        (nl,) = struct.unpack_from("<i", buf, o); o += 4
        name = buf[o:o + nl].decode(); o += nl
        dt, cnt = struct.unpack_from("<iq", buf, o); o += 12
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(buf, dtype=dtype, count=cnt, offset=o).copy(); o += cnt * dtype.itemsize
    return out


def tree_from_records(rec):
    """Reshape the flat arrays of a dumped tree (ref_harness.cc dump_tree) into a dict."""
    dim, n, nn, L, depth, thr = [int(x) for x in rec["meta"]]
    t = dict(rec)
    t.update(dim=dim, n_points=n, n_nodes=nn, n_leaves=L, depth=depth, leaf_threshold=thr)
    t["sorted_points"] = rec["sorted_points"].reshape(n, dim)
    for k in ("node_left_center", "node_right_center", "node_normal"):
        t[k] = rec[k].reshape(nn, dim)
    del t["meta"]
    return t
