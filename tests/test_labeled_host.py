"""CPU: host logic of the finetune datasets (gcc_b200/datasets/labeled.py) -- the file parsers of the reference's
Edgelist / TU inputs, the seed-first relabelling of whole graphs and the option plumbing of train.py --finetune."""
import os

import numpy as np
import torch

from gcc_b200.datasets import labeled


def test_edgelist_parser_matches_reference_rules(tmp_path):
    # data_util.py:61-113: ids in order of first appearance, both directions, labels in order of first appearance
    (tmp_path / "toy.edgelist").write_text("10 20\n20 30\n30 10\n40 10\n")
    (tmp_path / "toy.nodelabel").write_text("10 7\n20 5\n30 7\n40 9\n")
    e = labeled.Edgelist(str(tmp_path), "toy")
    assert e.node2id == {10: 0, 20: 1, 30: 2, 40: 3}
    assert e.data.edge_index.tolist() == [[0, 1, 1, 2, 2, 0, 3, 0], [1, 0, 2, 1, 0, 2, 0, 3]]
    assert e.data.y.argmax(dim=1).tolist() == [0, 1, 0, 2] and tuple(e.data.y.shape) == (4, 3)
    g = labeled.graph_from_edge_index(e.data.edge_index.numpy())
    assert g.num_nodes == 4 and np.diff(g.indptr).tolist() == [3, 2, 2, 1]
    assert g.indices.tolist() == [1, 2, 3, 0, 2, 0, 1, 0]
    # h-index sets: label = value > median, num_labels = number of distinct raw values (data_util.py:98-107)
    (tmp_path / "x_hindex.edgelist").write_text("1 2\n2 3\n")
    (tmp_path / "x_hindex.nodelabel").write_text("1 4\n2 9\n3 1\n")
    h = labeled.Edgelist(str(tmp_path), "x_hindex")
    assert h.data.y.argmax(dim=1).tolist() == [0, 1, 0] and h.data.y.shape[1] == 3


def test_tu_reader(tmp_path):
    d = tmp_path / "TOY"
    d.mkdir()
    (d / "TOY_A.txt").write_text("1, 2\n2, 1\n2, 3\n3, 2\n4, 5\n5, 4\n")
    (d / "TOY_graph_indicator.txt").write_text("1\n1\n1\n2\n2\n2\n")
    (d / "TOY_graph_labels.txt").write_text("-1\n1\n")
    graphs, labels = labeled.read_tu_dataset(str(tmp_path), "TOY")
    assert labels.tolist() == [0, 1]
    assert [g.num_nodes for g in graphs] == [3, 3]
    assert graphs[0].indptr.tolist() == [0, 1, 3, 4] and graphs[0].indices.tolist() == [1, 0, 2, 1]
    assert graphs[1].indptr.tolist() == [0, 1, 2, 2]                 # node 6 is isolated and kept


def test_seed_first_is_an_isomorphism():
    rng = np.random.RandomState(0)
    for trial in range(20):
        n = int(rng.randint(2, 30))
        src, dst = rng.randint(0, n, 3 * n), rng.randint(0, n, 3 * n)
        g = labeled._simple_csr(src, dst, n, "t")
        seed = int(np.argmax(np.diff(g.indptr)))
        ip, ix, perm = labeled.seed_first(g.indptr, g.indices, seed)
        assert perm[0] == seed and sorted(perm.tolist()) == list(range(n))
        assert perm[1:].tolist() == sorted(perm[1:].tolist())          # the others keep their order
        old = {(u, int(v)) for u in range(n) for v in g.indices[g.indptr[u]:g.indptr[u + 1]]}
        new = {(int(perm[u]), int(perm[v])) for u in range(n) for v in ix[ip[u]:ip[u + 1]]}
        assert old == new
        for u in range(n):
            row = ix[ip[u]:ip[u + 1]]
            assert np.all(np.diff(row) > 0)
        assert ip[1] - ip[0] == np.diff(g.indptr).max()


def test_finetune_options_and_names(tmp_path):
    import train
    a = train.parse_option(["--finetune", "--fold-idx", "3", "--dataset", "usa_airport", "--model-path",
                            str(tmp_path / "m"), "--tb-path", str(tmp_path / "t")])
    assert a.finetune and a.fold_idx == 3 and not a.cv
    a = train.option_update(a)
    assert "_ft_True_" in a.model_name and os.path.isdir(a.model_folder)
    assert labeled.GRAPH_CLASSIFICATION_DSETS == ["collab", "imdb-binary", "imdb-multi", "rdt-b", "rdt-5k"]
    # the reference's split: StratifiedKFold(10, shuffle=True, random_state=seed) (train.py:536-545)
    from sklearn.model_selection import StratifiedKFold
    y = np.arange(100) % 4
    folds = list(StratifiedKFold(n_splits=10, shuffle=True, random_state=0).split(np.zeros(100), y))
    assert len(folds) == 10 and all(len(te) == 10 for _, te in folds)
    assert torch.is_tensor(labeled.Data(None, None, torch.zeros(1)).y)
