"""The loader of tests/test_reference_golden.py is itself tested (it normally skips: no Julia here): a synthetic fixture directory in the
dumper's format -- raw column-major binaries + manifest -- must come back with the axes the comparisons assume."""
import os

import numpy as np

import test_reference_golden as trg


def test_loader_reads_column_major_dump(tmp_path, monkeypatch):
    K = np.arange(2 * 3 * 4, dtype=np.float64).reshape(2, 3, 4)          # numpy [rows][nu][NK]
    A = np.arange(9, dtype=np.float64).reshape(3, 3)
    dfs = np.array([5, 2, 4, 1, 3], dtype=np.int64)
    # what Julia's write() puts on disk for arrays of shape (NK, nu, rows) / (n, n): column-major = numpy's C-order of the reversed axes
    (tmp_path / "x__K.f64").write_bytes(K.tobytes())
    (tmp_path / "x__A.f64").write_bytes(np.asfortranarray(A).T.tobytes())   # Julia A[i,j] column-major
    (tmp_path / "x__dfs_list.i64").write_bytes(dfs.tobytes())
    (tmp_path / "manifest.txt").write_text("x K f64 4x3x2\nx A f64 3x3\nx dfs_list i64 5\ny B f64 1\n")
    monkeypatch.setattr(trg, "JDIR", str(tmp_path))
    monkeypatch.setattr(trg, "MANIFEST", str(tmp_path / "manifest.txt"))
    g = trg.load_case("x")
    assert set(g) == {"K", "A", "dfs_list"}
    assert np.array_equal(g["K"], K)
    assert np.array_equal(g["A"].T, A)           # the reversed-shape view of a column-major matrix is its transpose
    assert np.array_equal(g["dfs_list"], dfs)
    assert trg.parse_manifest(open(trg.MANIFEST).read()) == ["x", "y"]


def test_dumper_and_loader_agree_on_case_names():
    """every case the loader looks for is written by tools/dump_golden.jl (and vice versa)"""
    src = open(os.path.join(os.path.dirname(trg.HERE), "tools", "dump_golden.jl")).read()
    wanted = trg.CASES + ["c5_pid_doublependulum", "tracking_triple_cartpole"]
    for c in wanted:
        assert f'"{c}"' in src, c
    for name in ("A", "Bu", "Bl", "G", "K", "storage", "state0", "dfs_list", "body_ids", "eqc_ids", "storage0", "storage_tracked", "state0_tracked"):
        assert f'"{name}"' in src, name
