#!/usr/bin/env python
"""generate.py -- the reference's inference export (generate.py:33-134) on the B200 kernels.

Loads a pretraining checkpoint (`--load-path`, written by this repo's train.py or by the reference:
same `model` state_dict keys), walks two ego-nets from EVERY node of the target graph in order
(NodeClassificationDataset, graph_dataset.py:279-309), encodes both with the model in eval mode
(BatchNorm running statistics, no dropout) and saves (f(q) + f(k)) / 2 as `<model_folder>/<name>.npy`
(generate.py:48-53,133-134).  The reference batches the whole dataset into ONE batch (:90); here
the nodes go through in `--batch-size` chunks -- eval-mode encoding is per-graph, so the result is
the same.  `--dataset` is an .npz CSR (indptr, indices) or `synthetic-<kind>`: the reference's named
datasets need the network / DGL.
"""
import argparse
import os

import numpy as np
import torch

from gcc_b200.datasets import synthetic
from gcc_b200.datasets.graph_dataset import NodeClassificationDataset
from gcc_b200.models import GraphEncoder


def test_moco(train_loader, model, opt):
    """generate.py:33-53."""
    model.eval()
    emb_list = []
    for graph_q, graph_k, count in train_loader:
        bsz = graph_q.batch_size
        with torch.no_grad():
            feat_q = model(graph_q)
            feat_k = model(graph_k)
        assert feat_q.shape == (bsz, opt.hidden_size)
        emb_list.append(((feat_q + feat_k) / 2)[:count].detach().cpu())
    return torch.cat(emb_list)


def build_graph(name, nodes, edges):
    if name.endswith(".npz"):
        return name
    kind = name.split("-", 1)[1] if "-" in name else "er"
    if kind == "chunglu":
        return synthetic.chung_lu(nodes, edges, 0.5, seed=0)
    return synthetic.erdos_renyi(nodes, edges, seed=0)


def main(args_test):
    if not os.path.isfile(args_test.load_path):
        raise SystemExit("=> no checkpoint found at '{}'".format(args_test.load_path))
    print("=> loading checkpoint '{}'".format(args_test.load_path))
    checkpoint = torch.load(args_test.load_path, map_location="cpu", weights_only=False)
    print("=> loaded successfully '{}' (epoch {})".format(args_test.load_path, checkpoint["epoch"]))
    args = checkpoint["opt"]
    if not torch.cuda.is_available():
        raise SystemExit("generate.py needs a CUDA device (sm_100a); there is no CPU path")
    torch.cuda.set_device(args_test.gpu or 0)
    args.device = torch.device("cuda", args_test.gpu or 0)
    train_dataset = NodeClassificationDataset(
        dataset=build_graph(args_test.dataset, args_test.graph_nodes, args_test.graph_edges),
        rw_hops=args.rw_hops, subgraph_size=args.subgraph_size, restart_prob=args.restart_prob,
        positional_embedding_size=args.positional_embedding_size, device=args.device,
        seed=getattr(args, "seed", 0), batch_size=args_test.batch_size)
    model = GraphEncoder(
        positional_embedding_size=args.positional_embedding_size, max_node_freq=args.max_node_freq,
        max_edge_freq=args.max_edge_freq, max_degree=args.max_degree,
        freq_embedding_size=args.freq_embedding_size, degree_embedding_size=args.degree_embedding_size,
        output_dim=args.hidden_size, node_hidden_dim=args.hidden_size, edge_hidden_dim=args.hidden_size,
        num_layers=args.num_layer, num_step_set2set=args.set2set_iter,
        num_layer_set2set=args.set2set_lstm_layer, gnn_model=args.model, norm=args.norm,
        degree_input=True)
    model.load_state_dict(checkpoint["model"])
    model = model.to(args.device)
    del checkpoint
    emb = test_moco(train_dataset, model, args)
    out = os.path.join(getattr(args, "model_folder", "."), os.path.basename(args_test.dataset).replace(".npz", ""))
    np.save(out, emb.numpy())
    print("saved {} embeddings of dim {} to {}.npy".format(emb.shape[0], emb.shape[1], out))
    return emb


if __name__ == "__main__":
    parser = argparse.ArgumentParser("argument for generation")
    parser.add_argument("--load-path", type=str, required=True, help="path to load model")
    parser.add_argument("--dataset", type=str, default="synthetic-er",
                        help=".npz CSR file (indptr, indices) or synthetic-er / synthetic-chunglu")
    parser.add_argument("--graph-nodes", type=int, default=2000)
    parser.add_argument("--graph-edges", type=int, default=10000)
    parser.add_argument("--batch-size", type=int, default=256)
    parser.add_argument("--gpu", default=None, type=int, help="GPU id to use.")
    main(parser.parse_args())
