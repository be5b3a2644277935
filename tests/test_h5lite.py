"""h5lite -- the from-scratch HDF5 subset reader/writer of the host `corr` (the image has no
libhdf5): round trips of every supported type, the reference's overwrite-in-place result
semantics, large groups (several symbol nodes), and reading the one genuine libhdf5-written
file available in the image (a MATLAB 7.4 file with a 512-byte user block)."""
import os

import numpy as np
import pytest

SAMPLE = "/opt/prime-rl/.venv/lib/python3.12/site-packages/scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat"


def test_roundtrip_all_types(pkg, tmp_path):
    H = pkg.h5lite
    rng = np.random.default_rng(0)
    f = H.File()
    vals = {}
    for t in ("int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64"):
        a = (rng.standard_normal((3, 5)) * 100).astype(t)
        vals["/types/" + t] = a
        f.put("/types/" + t, a)
    vals["/xpcs/dqmap"] = rng.integers(0, 37, (64, 48)).astype(np.int32)
    f.put("/xpcs/dqmap", vals["/xpcs/dqmap"])
    f.put("/xpcs/compression", "ENABLED")
    f.put("/measurement/instrument/detector/efficiency", np.array([[0.5]], np.float32))
    f.put("/a/b/c/d/deep", np.arange(7, dtype=np.int64))
    f.put("/exchange/cube", np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    p = str(tmp_path / "t.h5")
    f.save(p)
    f.close()
    assert open(p, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    g = H.File(p)
    for k, v in vals.items():
        got = g.get(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
    assert g.get("/xpcs/compression") == "ENABLED"
    assert g.get("/measurement/instrument/detector/efficiency")[0, 0] == np.float32(0.5)
    assert np.array_equal(g.get("/a/b/c/d/deep"), np.arange(7))
    assert g.get("/exchange/cube").shape == (2, 3, 4)
    assert sorted(g.list("/")) == ["a", "exchange", "measurement", "types", "xpcs"]
    assert g.kind("/xpcs") == "group" and g.kind("/xpcs/dqmap") == "dataset" and g.kind("/nope") == "absent"
    g.close()


def test_results_are_added_and_overwritten_in_place(pkg, tmp_path):
    """H5Result semantics (h5_result.cpp:67-103): reopen the config file, create the group and
    the dataset if missing, overwrite the values if they exist; everything else is preserved."""
    H = pkg.h5lite
    p = str(tmp_path / "cfg.h5")
    f = H.File()
    f.put("/xpcs/delays_per_level", np.array([[8]], np.int32))
    f.save(p)
    f.close()
    f = H.File(p)
    f.put("/exchange/norm-0-g2", np.ones((5, 3), np.float32))
    f.save(p)
    f.close()
    f = H.File(p)
    f.put("/exchange/norm-0-g2", np.full((5, 3), 2.0, np.float32))
    f.put("/exchange/C2T_all/g2_00001", np.eye(4, dtype=np.float32))
    f.save(p)
    f.close()
    f = H.File(p)
    assert f.get("/xpcs/delays_per_level")[0, 0] == 8
    assert (f.get("/exchange/norm-0-g2") == 2.0).all()
    assert np.array_equal(f.get("/exchange/C2T_all/g2_00001"), np.eye(4, dtype=np.float32))
    f.close()


def test_large_group_spans_several_symbol_nodes(pkg, tmp_path):
    H = pkg.h5lite
    f = H.File()
    for i in range(300):
        f.put("/many/ds_%04d" % i, np.array([i], np.int64))
    p = str(tmp_path / "many.h5")
    f.save(p)
    f.close()
    g = H.File(p)
    assert len(g.list("/many")) == 300
    for i in (0, 63, 64, 128, 299):
        assert g.get("/many/ds_%04d" % i)[0] == i
    g.close()


@pytest.mark.skipif(not os.path.exists(SAMPLE), reason="genuine HDF5 sample not in this image")
def test_reads_a_genuine_libhdf5_file(pkg):
    g = pkg.h5lite.File(SAMPLE)
    assert g.list("/") == ["testdouble"]
    a = g.get("/testdouble")
    assert a.dtype == np.float64 and a.shape == (9, 1)
    assert np.allclose(a.ravel(), np.arange(9) * np.pi / 4)
    g.close()


def test_rejects_non_hdf5(pkg, tmp_path):
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(pkg.h5lite.H5Error):
        pkg.h5lite.File(str(p))


def test_corr_fails_loudly_without_gpu(pkg, tmp_path):
    """The host program has no CPU path: on a box without a usable GPU it exits non-zero with the
    library's message instead of computing anything."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import refdrv
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    corr = os.path.join(os.path.dirname(pkg.cabi.LIB_PATH), "corr")
    dq, sq = pkg.synth.annular_qmaps(16, 16, n_dynamic=2, static_per_dynamic=2, r_min=1.0)
    off, idx, val = pkg.synth.sparse_frames(256, 50, 0.05, seed=1)
    imm = str(tmp_path / "d.imm")
    pkg.synth.write_imm_sparse(imm, 16, 16, off, idx, val)
    cfg = str(tmp_path / "c.h5")
    f = pkg.h5lite.File()
    for path, value in refdrv.config_items(dq, sq, 50, imm)[0]:
        f.put(path, value)
    f.save(cfg)
    f.close()
    p = subprocess.run([corr, cfg], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 3 and "no CPU path" in p.stderr
    assert subprocess.run([corr], stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode == 1


@pytest.mark.parametrize("deflate,shuffle", [(0, False), (6, False), (6, True)])
def test_chunked_filtered_datasets_roundtrip(pkg, tmp_path, deflate, shuffle):
    """Chunked storage with the deflate / shuffle filters (H5Pset_chunk / H5Pset_deflate / H5Pset_shuffle): edge chunks,
    a 3-d frame stack chunked per frame with more than 64 chunks (B-tree with an internal level), a single-chunk
    matrix (how the reference stores C2T_all/g2_*).  Written and read back by h5lite; the reader's chunk path is the
    one that decodes libhdf5-written files."""
    H = pkg.h5lite
    rng = np.random.default_rng(3)
    stack = (rng.random((70, 24, 32)) < 0.03).astype(np.uint16) * rng.integers(1, 5, (70, 24, 32)).astype(np.uint16)
    mat = rng.standard_normal((37, 53)).astype(np.float32)
    vol = rng.integers(-1000, 1000, (5, 7, 9)).astype(np.int32)
    f = H.File()
    f.put("/entry/data/data", stack)
    f.set_storage("/entry/data/data", chunk=(1, 24, 32), deflate=deflate, shuffle=shuffle)
    f.put("/exchange/C2T_all/g2_00001", mat)
    f.set_storage("/exchange/C2T_all/g2_00001", chunk=mat.shape, deflate=deflate, shuffle=shuffle)
    f.put("/odd/vol", vol)
    f.set_storage("/odd/vol", chunk=(2, 4, 4), deflate=deflate, shuffle=shuffle)   # edge chunks on every axis
    f.put("/plain", np.arange(10, dtype=np.float64))
    p = str(tmp_path / "c.h5")
    f.save(p)
    f.close()
    g = H.File(p)
    assert np.array_equal(g.get("/entry/data/data"), stack)
    assert np.array_equal(g.get("/exchange/C2T_all/g2_00001"), mat)
    assert np.array_equal(g.get("/odd/vol"), vol)
    assert np.array_equal(g.get("/plain"), np.arange(10, dtype=np.float64))
    notes = g.report("notes")
    assert any("/entry/data/data" in n and "chunked" in n for n in notes)
    g.close()
    if deflate:
        assert os.path.getsize(p) < stack.nbytes   # the sparse stack compresses


def test_attributes_of_a_genuine_file_survive_a_rewrite(pkg, tmp_path):
    """The reference adds datasets to the user's file in place; h5lite rewrites the file, so the attribute messages of
    every object must be carried over byte for byte (MATLAB_class on the dataset of the genuine sample file)."""
    H = pkg.h5lite
    f = H.File(SAMPLE)
    names = f.list("/")
    before = {n: f.extra("/" + n) for n in names}
    assert any(len(v) > 0 and any(b"MATLAB_class" in body for _, body in v) for v in before.values())
    assert f.report("lossy") == []
    f.put("/exchange/new", np.arange(4, dtype=np.float32))
    p = str(tmp_path / "rw.h5")
    f.save(p)
    f.close()
    g = H.File(p)
    for n in names:
        assert g.extra("/" + n) == before[n], n
    assert np.array_equal(g.get("/exchange/new"), np.arange(4, dtype=np.float32))
    g.close()


def test_independent_parser_reads_what_h5lite_writes(pkg, tmp_path):
    """tests/h5check.py walks the file format on its own (pure Python, no code shared with h5lite) and is first pointed
    at the genuine libhdf5-written sample; it must then find every group, dataset, attribute and value in a file h5lite
    wrote -- contiguous, chunked (with and without filters, edge chunks, > 64 chunks) and a rewritten genuine file."""
    from h5check import H5Check
    ref = H5Check(SAMPLE)                                  # the parser itself, on genuine libhdf5 output
    assert list(ref.datasets) == ["/testdouble"] and ref.datasets["/testdouble"].dtype == np.float64
    assert any(b"MATLAB_class" in a for a in ref.attrs["/testdouble"])
    H = pkg.h5lite
    g = H.File(SAMPLE)
    assert np.array_equal(g.get("/testdouble"), ref.datasets["/testdouble"])   # both readers agree on the genuine file
    rng = np.random.default_rng(5)
    vals = {"/xpcs/dqmap": rng.integers(0, 37, (64, 48)).astype(np.int32),
            "/xpcs/delays_per_level": np.array([[8]], np.int32),
            "/measurement/instrument/detector/efficiency": np.array([[0.5]], np.float32),
            "/exchange/norm-0-g2": rng.standard_normal((88, 36)).astype(np.float32),
            "/exchange/timestamp_clock": rng.standard_normal((2, 500)),
            "/exchange/frames_out": rng.standard_normal((4, 6, 3)).astype(np.float32),
            "/exchange/C2T_all/g2_00001": np.triu(rng.standard_normal((150, 150))).astype(np.float32),
            "/entry/data/data": (rng.random((70, 12, 16)) < 0.05).astype(np.uint16)}
    for k, v in vals.items():
        g.put(k, v)
    for i in range(150):                                   # a group that needs several symbol nodes
        g.put("/many/d%03d" % i, np.array([i], np.int64))
    g.put("/xpcs/compression", "ENABLED")
    g.set_storage("/exchange/C2T_all/g2_00001", chunk=(150, 150), deflate=6)
    g.set_storage("/entry/data/data", chunk=(1, 12, 16), deflate=6, shuffle=True)
    g.set_storage("/exchange/frames_out", chunk=(3, 4, 2))
    p = str(tmp_path / "w.h5")
    g.save(p)
    g.close()
    chk = H5Check(p)
    for k, v in vals.items():
        assert k in chk.datasets, k
        assert chk.datasets[k].dtype == v.dtype and chk.datasets[k].shape == v.shape, k
        assert np.array_equal(chk.datasets[k], v), k
    assert np.array_equal(chk.datasets["/testdouble"], ref.datasets["/testdouble"])
    assert chk.attrs["/testdouble"] == ref.attrs["/testdouble"]            # attribute messages byte for byte
    assert chk.datasets["/xpcs/compression"].tobytes().rstrip(b"\0") == b"ENABLED"
    assert all(int(chk.datasets["/many/d%03d" % i][0]) == i for i in range(150))
    assert {"/", "/xpcs", "/exchange", "/exchange/C2T_all", "/many", "/entry/data"} <= set(chk.groups)
