"""CPU: DGL 0.4.x save_graphs (.bin) reader (gcc_b200/datasets/dgl_bin.py; reference call sites
gcc/datasets/graph_dataset.py:26-28,58-60 and gcc/utils/x2dgl.py:129-131).  DGL is absent, so these tests pin the
reader against this repo's own writer of the restated layout AND against header variants the reader must tolerate;
parity with a file written by DGL itself stays unpinned (said so in the module and in DESIGN.md)."""
import struct

import numpy as np
import pytest

from gcc_b200.datasets import dgl_bin, synthetic


def _graphs():
    return [synthetic.erdos_renyi(60, 150, seed=1), synthetic.star_graph(12), synthetic.chung_lu(300, 900, seed=2)]


def test_round_trip_graphs_and_labels(tmp_path):
    gs = _graphs()
    path = dgl_bin.write_dgl_bin(str(tmp_path / "small.bin"), gs)
    got, labels = dgl_bin.read_dgl_bin(path)
    assert labels["graph_sizes"].tolist() == [g.num_nodes for g in gs]
    assert dgl_bin.read_labels(path)["graph_sizes"].dtype == np.int64
    for a, b in zip(got, gs):
        assert a.num_nodes == b.num_nodes
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    sub, _ = dgl_bin.read_dgl_bin(path, idx_list=[2, 0])            # load_graphs(file, jobs[worker_id])
    assert [g.num_nodes for g in sub] == [gs[2].num_nodes, gs[0].num_nodes]


def test_reader_tolerates_header_variants(tmp_path):
    """No graph-type word after the version / an extra word: the offset table is still found."""
    gs = _graphs()[:2]
    path = dgl_bin.write_dgl_bin(str(tmp_path / "a.bin"), gs)
    raw = open(path, "rb").read()
    for variant, delta in (("no_type", -8), ("extra", +8)):
        head = raw[:16] + (b"" if delta < 0 else raw[16:24] + struct.pack("<Q", 7))
        ng = struct.unpack_from("<Q", raw, 24)[0]
        offs = np.frombuffer(raw, dtype=np.uint64, count=ng, offset=40) + np.uint64(delta if delta > 0 else 0) - np.uint64(-delta if delta < 0 else 0)
        body = raw[24:32] + struct.pack("<Q", ng) + offs.tobytes() + raw[40 + 8 * ng:]
        p2 = str(tmp_path / (variant + ".bin"))
        open(p2, "wb").write(head + body)
        got, labels = dgl_bin.read_dgl_bin(p2)
        assert [g.num_nodes for g in got] == [g.num_nodes for g in gs], variant
        assert np.array_equal(got[1].indices, gs[1].indices)
        assert labels["graph_sizes"].tolist() == [g.num_nodes for g in gs]


def test_rejects_foreign_and_asymmetric_files(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"\0" * 64)
    with pytest.raises(dgl_bin.DglBinError):
        dgl_bin.read_dgl_bin(str(p))
    g = synthetic.CSRGraph(np.array([0, 1, 1], np.int64), np.array([1], np.int32), 2, "one-way")
    path = dgl_bin.write_dgl_bin(str(tmp_path / "asym.bin"), [g])
    with pytest.raises(dgl_bin.DglBinError):
        dgl_bin.read_dgl_bin(path)


def test_load_graphs_accepts_bin(tmp_path):
    from gcc_b200.datasets.graph_dataset import load_graphs
    gs = _graphs()
    path = dgl_bin.write_dgl_bin(str(tmp_path / "small.bin"), gs)
    union, sizes = load_graphs(path)
    assert sizes == [g.num_nodes for g in gs] and union.num_nodes == sum(sizes)
    assert len(union.indices) == sum(len(g.indices) for g in gs)
