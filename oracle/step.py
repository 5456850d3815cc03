"""Oracle (test infrastructure): one full pretraining step on CPU.

Restates train.py:378-434 (train_moco body) with the pieces in oracle/model.py.
State is a plain dict so tests can seed it from golden fixtures:
  state = dict(params={key: tensor}, ema={key: tensor}, memory=tensor[K,d], index=int,
               adam_m={key: tensor}, adam_v={key: tensor}, adam_t=int)
`params`/`ema` hold every state_dict entry of GraphEncoder -- the
GIN-path ones are used, the rest (set2set.*, lin_readout.*) only ride along in
the EMA (train.py:169-172 averages model.parameters(), used or not).
"""
import numpy as np
import torch

from . import model as om
from . import rwr as orwr


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked", ".eps"))


def _apply_bn_stats(sd, stats, momentum=0.1):
    for key, (mean, var_unb) in stats.items():
        sd[key + "running_mean"] = (1 - momentum) * sd[key + "running_mean"] + momentum * mean
        sd[key + "running_var"] = (1 - momentum) * sd[key + "running_var"] + momentum * var_unb
        sd[key + "num_batches_tracked"] = sd[key + "num_batches_tracked"] + 1


def train_step(state, batch_q, batch_k, *, num_layers, moco=True, T=0.07, lr=0.005,
               alpha=0.999, clip_norm=1.0, weight_decay=1e-5, beta1=0.9, beta2=0.999,
               dropout_key=None, step_index=0, max_degree=512, hidden=None):
    """batch_* = dict(indptr, indices, pos, seed, sub_deg, node_off).
    Returns dict(loss, grad_norm, feat_q, feat_k, out, grads)."""
    params = state["params"]
    dt = next(v for k, v in params.items() if k.endswith("linears.0.weight")).dtype
    live = [k for k, v in params.items() if not is_buffer(k)]
    for k in live:
        params[k] = params[k].detach().clone().requires_grad_(True)
    B = len(batch_q["node_off"]) - 1
    out_dim = params["gnn.linears_prediction.0.weight"].shape[0]

    def masks(view):
        if dropout_key is None:
            return None
        return [orwr.dropout_mask(dropout_key, step_index, view * num_layers + i,
                                  B * out_dim, 0.5).reshape(B, out_dim)
                for i in range(num_layers)]

    def enc(p, b, keep):
        return om.gin_encoder_forward(
            p, b["indptr"], b["indices"], torch.as_tensor(b["pos"]).to(dt), b["seed"],
            b["sub_deg"], b["node_off"], num_layers=num_layers, max_degree=max_degree,
            norm=True, bn_train=True, dropout_keep=keep)

    feat_q, _, stats_q = enc(params, batch_q, masks(0))
    _apply_bn_stats(params, stats_q)
    if moco:
        with torch.no_grad():       # train.py:390-391; BN of model_ema runs in train mode (:360-365)
            feat_k, _, stats_k = enc(state["ema"], batch_k, None)
        _apply_bn_stats(state["ema"], stats_k)
        out = om.moco_logits(feat_q, feat_k, state["memory"], T)     # memory_moco.py:33-44
        loss = om.nce_softmax_loss(out)                               # criterions.py:12-17
    else:
        feat_k, _, stats_k = enc(params, batch_k, masks(1))
        _apply_bn_stats(params, stats_k)
        out = feat_k @ feat_q.t() / T                                 # train.py:400
        loss = om.nce_softmax_loss_ns(out)                            # criterions.py:27-33
    grads = torch.autograd.grad(loss, [params[k] for k in live], allow_unused=True)
    gdict = {k: g for k, g in zip(live, grads) if g is not None}
    # clip (train.py:409), Adam (train.py:417), on the parameters that received grads
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in gdict.values()))
    coef = clip_norm / (float(total) + 1e-6)
    state["adam_t"] = state.get("adam_t", 0) + 1
    t = state["adam_t"]
    with torch.no_grad():
        for k, g in gdict.items():
            if coef < 1.0:
                g = g * coef
            p = params[k]
            g = g + weight_decay * p
            m = state["adam_m"].setdefault(k, torch.zeros_like(p))
            v = state["adam_v"].setdefault(k, torch.zeros_like(p))
            m.mul_(beta1).add_(g, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = v.sqrt() / np.sqrt(1 - beta2 ** t) + 1e-8
            params[k] = (p - (lr / (1 - beta1 ** t)) * m / denom).detach()
        for k in live:
            params[k] = params[k].detach()
        if moco:
            for k in live:                                            # train.py:169-172
                state["ema"][k] = state["ema"][k] * alpha + (1 - alpha) * params[k]
            state["index"] = om.moco_enqueue(state["memory"], feat_k, state["index"])
    return dict(loss=float(loss), grad_norm=float(total), feat_q=feat_q.detach(),
                feat_k=feat_k.detach(), out=out.detach(), grads=gdict)


def _adam(state, name, p, g, lr, t, weight_decay, beta1, beta2):
    g = g + weight_decay * p
    m = state["adam_m"].setdefault(name, torch.zeros_like(p))
    v = state["adam_v"].setdefault(name, torch.zeros_like(p))
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = v.sqrt() / np.sqrt(1 - beta2 ** t) + 1e-8
    return (p - (lr / (1 - beta1 ** t)) * m / denom).detach()


def finetune_step(state, batch, y, *, num_layers, lr=0.005, weight_decay=1e-5, beta1=0.9, beta2=0.999,
                  dropout_key=None, step_index=0, max_degree=512):
    """One step of train.py:train_finetune (:213-238): encoder forward (train mode), Linear head
    (state["out_w"], state["out_b"]), cross entropy, clip_grad_value_(1) on both parameter sets, Adam on both
    (two optimisers with the same hyper-parameters and step count).  Returns dict(loss, logits, grads)."""
    params = state["params"]
    dt = next(v for k, v in params.items() if k.endswith("linears.0.weight")).dtype
    live = [k for k, v in params.items() if not is_buffer(k)]
    for k in live:
        params[k] = params[k].detach().clone().requires_grad_(True)
    w = state["out_w"].detach().clone().requires_grad_(True)
    b = state["out_b"].detach().clone().requires_grad_(True)
    B = len(batch["node_off"]) - 1
    out_dim = params["gnn.linears_prediction.0.weight"].shape[0]
    keep = None
    if dropout_key is not None:
        keep = [orwr.dropout_mask(dropout_key, step_index, i, B * out_dim, 0.5).reshape(B, out_dim)
                for i in range(num_layers)]
    feat, _, stats = om.gin_encoder_forward(
        params, batch["indptr"], batch["indices"], torch.as_tensor(batch["pos"]).to(dt), batch["seed"],
        batch["sub_deg"], batch["node_off"], num_layers=num_layers, max_degree=max_degree, norm=True,
        bn_train=True, dropout_keep=keep)
    _apply_bn_stats(params, stats)
    logits = feat @ w.t() + b
    loss = torch.nn.functional.cross_entropy(logits, torch.as_tensor(y).long())
    grads = torch.autograd.grad(loss, [params[k] for k in live] + [w, b], allow_unused=True)
    gdict = {k: g.clamp(-1.0, 1.0) for k, g in zip(live + ["out_w", "out_b"], grads) if g is not None}
    state["adam_t"] = state.get("adam_t", 0) + 1
    t = state["adam_t"]
    with torch.no_grad():
        for k, g in gdict.items():
            if k == "out_w":
                state["out_w"] = _adam(state, k, w, g, lr, t, weight_decay, beta1, beta2)
            elif k == "out_b":
                state["out_b"] = _adam(state, k, b, g, lr, t, weight_decay, beta1, beta2)
            else:
                params[k] = _adam(state, k, params[k], g, lr, t, weight_decay, beta1, beta2)
        for k in live:
            params[k] = params[k].detach()
    return dict(loss=float(loss.detach()), logits=logits.detach(), grads=gdict)


def finetune_eval(state, batch, y, *, num_layers, max_degree=512):
    """train.py:test_finetune (:300-337) on one batch: eval-mode encoder, Linear head, cross entropy."""
    params = state["params"]
    dt = next(v for k, v in params.items() if k.endswith("linears.0.weight")).dtype
    with torch.no_grad():
        feat, _, _ = om.gin_encoder_forward(
            params, batch["indptr"], batch["indices"], torch.as_tensor(batch["pos"]).to(dt), batch["seed"],
            batch["sub_deg"], batch["node_off"], num_layers=num_layers, max_degree=max_degree, norm=True,
            bn_train=False, dropout_keep=None)
        logits = feat @ state["out_w"].t() + state["out_b"]
        loss = torch.nn.functional.cross_entropy(logits, torch.as_tensor(y).long())
    return dict(loss=float(loss), logits=logits)
