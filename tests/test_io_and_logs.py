"""CPU-only: data-set loaders (Code/data_processor.h:5-99) and the experiment-log text blocks (f4,
Code/Convex_Tree.h:964-991, Code/GPU_Convex_Tree.h:973-1005) of the C ABI."""
import os
import struct

import numpy as np
import pytest


def ref_parse(text, delims=" "):
    """The reference's contract restated in Python: one row per gets() line, dim = tokens of the first line,
    value = (float)atof(token)."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    rows = []
    dim = None
    for ln in lines:
        toks = [t for t in ln.replace("\r", " ").split(delims) if t != ""] if delims == " " else [t for t in ln.split(delims) if t.strip("\r") != ""]
        if dim is None:
            dim = len(toks)
        vals = [np.float32(float(t)) for t in toks[:dim]]
        vals += [np.float32(0)] * (dim - len(vals))
        rows.append(vals)
    return np.array(rows, dtype=np.float32).reshape(len(rows), dim)


def test_read_text_matches_reference_contract(pkg, tmp_path):
    rng = np.random.default_rng(0)
    data = (rng.standard_normal((1000, 7)) * 1e3).astype(np.float32)
    # mixed formatting like real dumps: fixed, scientific, integers, repeated separators, a trailing blank
    lines = []
    for r, row in enumerate(data):
        toks = []
        for j, v in enumerate(row):
            toks.append(("%.6f" % v) if (r + j) % 3 == 0 else ("%.8e" % v) if (r + j) % 3 == 1 else repr(float(v)))
        lines.append(("  " if r % 5 == 0 else " ").join(toks) + (" " if r % 7 == 0 else ""))
    text = "\n".join(lines) + "\n"
    p = tmp_path / "d.txt"
    p.write_text(text)
    got = pkg.read_text(str(p))
    want = ref_parse(text)
    assert got.shape == (1000, 7)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


REF_AUX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_aux")
needs_ref_aux = pytest.mark.skipif(not os.access(REF_AUX, os.X_OK), reason="oracle/_ref/ref_aux not built (reference sources absent)")


def read_fmat(path):
    raw = open(path, "rb").read()
    rows, cols = struct.unpack("<ii", raw[:8])
    return np.frombuffer(raw, dtype=np.float32, offset=8).reshape(rows, cols)


@needs_ref_aux
@pytest.mark.parametrize("delims", [" ", ",", "\t"])
def test_read_text_vs_reference_read_data(pkg, tmp_path, delims):
    """scht_read_text against the reference's own read_data (Code/data_processor.h:75-99, compiled unmodified into
    oracle/_ref/ref_aux) on well-formed files: same shape, same float bits"""
    import subprocess
    rng = np.random.default_rng(7)
    data = np.concatenate([(rng.standard_normal((700, 9)) * 10.0 ** rng.integers(-6, 7, (700, 1))).astype(np.float32),
                           rng.integers(-300, 300, (300, 9)).astype(np.float32)])
    lines = []
    for r, row in enumerate(data):
        toks = [("%.6f" % v) if (r + j) % 4 == 0 else ("%.8e" % v) if (r + j) % 4 == 1 else ("%d" % v) if float(v).is_integer() else repr(float(v))
                for j, v in enumerate(row)]
        lines.append(delims.join(toks))
    p = tmp_path / "d.txt"
    p.write_text("\n".join(lines) + "\n")
    out = tmp_path / "ref.fmat"
    subprocess.run([REF_AUX, "readdata", str(p), delims, str(out)], check=True, capture_output=True)
    want = read_fmat(out)
    got = pkg.read_text(str(p), delims)
    assert got.shape == want.shape == (1000, 9)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and a file written by scht_write_text is read back by the reference with the original bits
    q = tmp_path / "w.txt"
    pkg.write_text(str(q), data)
    subprocess.run([REF_AUX, "readdata", str(q), " ", str(out)], check=True, capture_output=True)
    assert np.array_equal(read_fmat(out).view(np.uint32), data.view(np.uint32))


@needs_ref_aux
@pytest.mark.parametrize("n,dim,pct,seed", [(2000, 4, 0.01, 1), (5000, 16, 0.004, 3), (777, 3, 0.1, 1)])
def test_tree_info_vs_reference_save_tree_info(pkg, tmp_path, n, dim, pct, seed):
    """scht_format_tree_info against the text the reference's own save_tree_info writes (Code/Convex_Tree.h:964-991) for the
    same data and seed; only the build time differs"""
    import subprocess
    sys_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    import sys
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    from io_formats import write_fmat
    data = np.random.default_rng(n).random((n, dim), dtype=np.float32)
    dp, out = tmp_path / "d.fmat", tmp_path / "info.txt"
    write_fmat(str(dp), data)
    subprocess.run([REF_AUX, "treeinfo", str(dp), repr(pct), str(seed), str(out)], check=True, capture_output=True)
    want = out.read_text()
    got = pkg.format_tree_info(pkg.HostTree(data, pct, seed=seed), pct)
    marker = "tree building time:  "
    assert marker in want and want.endswith("s\n")
    assert got[:got.index(marker) + len(marker)] == want[:want.index(marker) + len(marker)]
    assert got.endswith("s\n") and float(got[got.index(marker) + len(marker):-2]) >= 0.0


def test_read_text_edge_cases(pkg, tmp_path):
    p = tmp_path / "e.txt"
    # no trailing newline, CRLF, a short line (zero-filled), a long line (extra tokens ignored), an empty line (a row)
    p.write_bytes(b"1 2 3\r\n4 5\n\n6 7 8 9\n10 11 12")
    got = pkg.read_text(str(p))
    want = np.array([[1, 2, 3], [4, 5, 0], [0, 0, 0], [6, 7, 8], [10, 11, 12]], dtype=np.float32)
    assert np.array_equal(got, want)
    q = tmp_path / "c.csv"
    q.write_text("1.5,2.5\n3.5,4.5\n")
    assert np.array_equal(pkg.read_text(str(q), ","), np.array([[1.5, 2.5], [3.5, 4.5]], dtype=np.float32))
    with pytest.raises(pkg.SchtError):
        pkg.read_text(str(tmp_path / "missing.txt"))
    z = tmp_path / "z.txt"
    z.write_text("")
    with pytest.raises(pkg.SchtError):
        pkg.read_text(str(z))


def test_text_round_trip_is_exact_and_parallel(pkg, tmp_path):
    rng = np.random.default_rng(1)
    data = rng.random((20000, 16), dtype=np.float32) * np.float32(1e5)  # > 4096 rows: the multi-threaded path
    p = tmp_path / "big.txt"
    pkg.write_text(str(p), data)
    got = pkg.read_text(str(p))
    assert np.array_equal(got.view(np.uint32), data.view(np.uint32))


def test_bin_format(pkg, tmp_path):
    rng = np.random.default_rng(2)
    data = rng.standard_normal((777, 5)).astype(np.float32)
    p = tmp_path / "d.bin"
    pkg.write_bin(str(p), data)
    raw = p.read_bytes()
    assert raw[:8] == b"SCHTBIN1" and struct.unpack("<iiqq", raw[8:32]) == (5, 0, 777, 0) and len(raw) == 32 + 777 * 5 * 4
    assert np.array_equal(pkg.read_bin(str(p)).view(np.uint32), data.view(np.uint32))
    (tmp_path / "bad.bin").write_bytes(raw[:-4])
    with pytest.raises(pkg.SchtError):
        pkg.read_bin(str(tmp_path / "bad.bin"))
    (tmp_path / "bad2.bin").write_bytes(b"NOTMAGIC" + raw[8:])
    with pytest.raises(pkg.SchtError):
        pkg.read_bin(str(tmp_path / "bad2.bin"))


def test_fvecs(pkg, tmp_path):
    rng = np.random.default_rng(3)
    data = rng.integers(0, 219, (50, 128)).astype(np.float32)
    p = tmp_path / "s.fvecs"
    with open(p, "wb") as f:
        for row in data:
            f.write(struct.pack("<i", 128))
            f.write(row.tobytes())
    assert np.array_equal(pkg.read_fvecs(str(p)), data)
    with open(p, "ab") as f:
        f.write(b"\x00\x00")
    with pytest.raises(pkg.SchtError):
        pkg.read_fvecs(str(p))


def test_loaded_data_builds_the_same_tree(pkg, tmp_path):
    rng = np.random.default_rng(4)
    data = rng.random((3000, 6), dtype=np.float32)
    p = tmp_path / "t.txt"
    pkg.write_text(str(p), data)
    ta, tb = pkg.HostTree(pkg.read_text(str(p)), 0.01, seed=1), pkg.HostTree(data, 0.01, seed=1)  # (arrays() are views: keep the trees)
    a, b = ta.arrays(), tb.arrays()
    assert np.array_equal(a["sorted_to_orig"], b["sorted_to_orig"]) and np.array_equal(a["node_beta"].view(np.uint32), b["node_beta"].view(np.uint32))


def test_tree_info_text_follows_the_reference_layout(pkg):
    data = np.random.default_rng(5).random((2000, 4), dtype=np.float32)
    ht = pkg.HostTree(data, 0.01, seed=1)
    a = ht.arrays()
    text = pkg.format_tree_info(ht, 0.01)
    # Code/Convex_Tree.h:964-991, field by field (the first fragment is the reference's own stray strcat)
    want = ("\n   data dim=Tree info: \n      data size= 2000, dim=4\n      depth= %d\n      nodes num= %d\n      leaf nodes num=%d"
            "\n      percent threshold to be a leaf=0.010000\n      number threshold to be a leaf=%d\n      tree building time:  "
            % (a["depth"], a["n_nodes"], a["n_leaves"], a["leaf_threshold"]))
    assert text.startswith(want)
    tail = text[len(want):]
    assert tail.endswith("s\n") and float(tail[:-2]) >= 0.0


def test_knn_log_text_follows_the_reference_layout(pkg):
    st = dict(n_queries=1000, dist_computations=123456789012, hyperplane_tests=4242, descent_dists=28000, batches=3, ms_total=1500.0,
              ms_device=1250.0)
    text = pkg.format_knn_log(st, 1000, 10, 400)
    want = ("\n*-----------------------------------------KNN QUERY RESULT ----------------------------------------\n"
            "*     Query points number:1000, and K=10"
            "\n*     GPU Alorithm running time:1.500000"
            "\n*          Where:"
            "\n*          (1) Init query points: 0.250000"
            "\n*              where Approximate searching: 0.000000"
            "\n*          (2) Quadratic programming Calls in devices=4242"
            "\n*              where inner productions (alpha'*q) in quadprog =4242"
            "\n*          (3) Distance computations in devices=123456789012"
            "\n*          (4) Distance computations in approximate searching=28000"
            "\n*          (5) Overall distance computations in host and devices=(2)+ (3)+(4)=123456821254"
            "\n*          (6) BATCH_SIZE=400, batch iteration times=3"
            "\n*-----------------------------------------KNN QUERY RESULT----------------------------------------")
    assert text == want
    # small buffers are refused, never overrun
    import ctypes as C
    buf = C.create_string_buffer(16)
    stc = pkg.Stats()
    assert pkg.lib().scht_format_knn_log(C.byref(stc), 1, 1, 1, buf, 16, None) == -1
