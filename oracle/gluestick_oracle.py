"""CPU oracle for the GlueStick matcher (TEST INFRASTRUCTURE ONLY -- see oracle/lightglue_oracle.py for the rules).

A plain-PyTorch restatement of gluefactory/models/matchers/gluestick.py written from the math (SURVEY.md Appendix
A.7), token-major ([B, N, C], channels last: a Conv1d with kernel size 1 is a Linear over channels), pinned against
the UNMODIFIED reference through tests/golden/gluestick_*.npz (oracle/make_golden.py `gluestick`): predictions, all
loss entries, every parameter gradient and the updated BatchNorm running statistics.

`w` is a flat dict under the reference's state_dict names (Conv1d weights [out, in, 1]).  BatchNorm runs in training
mode (batch statistics over all tokens of the call, biased variance, eps 1e-5), exactly as `model.train()` does in
train.py:465; `bn_stats`, when given, collects the batch mean / unbiased variance per BatchNorm call so that the
running-statistics update (momentum 0.1, two calls per layer: image 0 then image 1) can be checked too.
"""
import torch
import torch.nn.functional as F

from .lightglue_oracle import filter_matches, log_double_softmax  # gluestick.py:772-783, :318-331 (same rule)

EPS_BN = 1e-5


def normalize_keypoints(kpts, size):
    """gluestick.py:478-490: centre, divide by 0.7 * the longer side."""
    size = size.to(kpts)
    c = size / 2
    f = size.max(1, keepdim=True).values * 0.7
    return (kpts - c[:, None, :]) / f[:, None, :]


def _conv(x, w, name):
    """Conv1d(kernel 1) on token-major x [..., C_in]."""
    return x @ w[name + ".weight"][:, :, 0].t() + w[name + ".bias"]


def _bn(x, w, name, bn_stats=None):
    """BatchNorm1d in training mode over all leading dims of x [..., C] (gluestick.py:465-475 MLP)."""
    flat = x.reshape(-1, x.shape[-1])
    mean = flat.mean(0)
    var = flat.var(0, unbiased=False)
    if bn_stats is not None:
        n = flat.shape[0]
        bn_stats.setdefault(name, []).append((mean.detach(), (var * n / max(n - 1, 1)).detach()))
    return (x - mean) / torch.sqrt(var + EPS_BN) * w[name + ".weight"] + w[name + ".bias"]


def _mlp(x, w, pre, n_layers, bn_stats=None):
    """MLP(channels, do_bn=True): conv, [BN, ReLU, conv]*  -- module indices 0, (1, 2, 3), (4, 5, 6) ..."""
    for i in range(n_layers):
        x = _conv(x, w, f"{pre}.{3 * i}")
        if i < n_layers - 1:
            x = torch.relu(_bn(x, w, f"{pre}.{3 * i + 1}", bn_stats))
    return x


def keypoint_encoder(kpts, scores, w, bn_stats=None):
    """KeypointEncoder (gluestick.py:493-501): MLP([3, 32, 64, 128, 256, D]) of (x, y, score)."""
    return _mlp(torch.cat([kpts, scores[..., None]], -1), w, "kenc.encoder", 5, bn_stats)


def endpoint_encoder(lines, scores, w, bn_stats=None):
    """EndPtEncoder (gluestick.py:504-523): per endpoint (x, y, offset to the other endpoint, line score)."""
    B, L = lines.shape[:2]
    off = lines[:, :, 1] - lines[:, :, 0]
    off = torch.stack([off, -off], 2).reshape(B, 2 * L, 2)
    # NB the reference tiles the line scores (`scores.repeat(1, 2)`, gluestick.py:520) while the endpoints are interleaved:
    # endpoint p gets the score of line p % L -- reproduced as is
    inp = torch.cat([lines.reshape(B, 2 * L, 2), off, scores.repeat(1, 2)[..., None]], -1)
    return _mlp(inp, w, "lenc.encoder", 5, bn_stats)


def mha(x, src, w, pre, H=4):
    """MultiHeadedAttention (gluestick.py:532-551): heads split as channel = d * H + h (view(b, dim, h, n))."""
    B, N, D = x.shape
    dh = D // H
    q = _conv(x, w, pre + ".proj.0").view(B, N, dh, H)
    k = _conv(src, w, pre + ".proj.1").view(B, -1, dh, H)
    v = _conv(src, w, pre + ".proj.2").view(B, -1, dh, H)
    s = torch.einsum("bndh,bmdh->bhnm", q, k) / dh**0.5
    o = torch.einsum("bhnm,bmdh->bndh", torch.softmax(s, -1), v).reshape(B, N, D)
    return _conv(o, w, pre + ".merge")


def gnn_layer(d0, d1, w, i, kind, bn_stats=None):
    """GNNLayer + AttentionalPropagation (gluestick.py:553-586): x + MLP([x ; MHA(x, src)])."""
    pre = f"gnn.layers.{i}.update"
    s0, s1 = (d1, d0) if kind == "cross" else (d0, d1)
    out = []
    for x, src in ((d0, s0), (d1, s1)):
        msg = mha(x, src, w, pre + ".attn")
        out.append(x + _mlp(torch.cat([x, msg], -1), w, pre + ".mlp", 2, bn_stats))
    return out[0], out[1]


def line_layer(d0, d1, enc0, enc1, idx0, idx1, w, j, bn_stats=None):
    """LineLayer without line attention (gluestick.py:589-691): per line endpoint a message MLP([this endpoint's
    junction ; the other endpoint's junction ; endpoint encoding]), averaged over the endpoints that share a junction
    (scatter_reduce mean, include_self=False) and added to the junction's descriptor."""
    pre = f"gnn.line_layers.{j}.mlp"
    out = []
    for d, enc, idx in ((d0, enc0, idx0), (d1, enc1, idx1)):
        B, n2 = idx.shape
        D = d.shape[-1]
        ld = torch.gather(d, 1, idx[..., None].expand(B, n2, D))
        other = ld.view(B, n2 // 2, 2, D).flip(2).reshape(B, n2, D)
        upd = _mlp(torch.cat([ld, other, enc], -1), w, pre, 2, bn_stats)
        acc = torch.zeros_like(d).scatter_reduce(1, idx[..., None].expand(B, n2, D), upd, reduce="mean", include_self=False)
        out.append(d + acc)
    return out[0], out[1]


def line_scores_from_junctions(ld0, ld1, idx0, idx1, w, proj, D):
    """_get_line_matches (gluestick.py:333-377) up to the raw line scores: junction similarity gathered per endpoint
    pair, best of the two endpoint orderings."""
    m0, m1 = _conv(ld0, w, proj), _conv(ld1, w, proj)
    s = m0 @ m1.transpose(1, 2) / D**0.5
    B, n20 = idx0.shape
    n21 = idx1.shape[1]
    s = torch.gather(s, 2, idx1[:, None, :].expand(B, s.shape[1], n21))
    s = torch.gather(s, 1, idx0[:, :, None].expand(B, n20, n21))
    s = s.reshape(B, n20 // 2, 2, n21 // 2, 2)
    return 0.5 * torch.maximum(s[:, :, 0, :, 0] + s[:, :, 1, :, 1], s[:, :, 0, :, 1] + s[:, :, 1, :, 0])


def gluestick_forward(w, data, conf, bn_stats=None):
    """GlueStick._forward (gluestick.py:143-316), lines present, no inter-supervision, no line attention."""
    layers = conf["GNN_layers"]
    D = conf.get("descriptor_dim", 256)
    th = conf.get("filter_threshold", 0.2)
    k0 = normalize_keypoints(data["keypoints0"], data["view0"]["image_size"])
    k1 = normalize_keypoints(data["keypoints1"], data["view1"]["image_size"])
    d0, d1 = data["descriptors0"], data["descriptors1"]
    if "input_proj.weight" in w:
        d0, d1 = _conv(d0, w, "input_proj"), _conv(d1, w, "input_proj")
    d0 = d0 + keypoint_encoder(k0, data["keypoint_scores0"], w, bn_stats)
    d1 = d1 + keypoint_encoder(k1, data["keypoint_scores1"], w, bn_stats)
    B, L0 = data["lines0"].shape[:2]
    L1 = data["lines1"].shape[1]
    l0 = normalize_keypoints(data["lines0"].reshape(B, 2 * L0, 2), data["view0"]["image_size"]).reshape(B, L0, 2, 2)
    l1 = normalize_keypoints(data["lines1"].reshape(B, 2 * L1, 2), data["view1"]["image_size"]).reshape(B, L1, 2, 2)
    enc0 = endpoint_encoder(l0, data["line_scores0"], w, bn_stats)
    enc1 = endpoint_encoder(l1, data["line_scores1"], w, bn_stats)
    idx0, idx1 = data["lines_junc_idx0"].reshape(B, -1), data["lines_junc_idx1"].reshape(B, -1)
    for i, kind in enumerate(layers):
        d0, d1 = gnn_layer(d0, d1, w, i, kind, bn_stats)
        if kind == "self":
            d0, d1 = line_layer(d0, d1, enc0, enc1, idx0, idx1, w, i // 2, bn_stats)
    md0, md1 = _conv(d0, w, "final_proj"), _conv(d1, w, "final_proj")
    scores = log_double_softmax(md0 @ md1.transpose(1, 2) / D**0.5, w["bin_score"])
    m0, m1, ms0, ms1 = filter_matches(scores, th)
    raw = line_scores_from_junctions(d0[:, :2 * L0], d1[:, :2 * L1], idx0, idx1, w, "final_line_proj", D)
    lscores = log_double_softmax(raw, w["line_bin_score"])
    lm0, lm1, lms0, lms1 = filter_matches(lscores, th)
    return {"log_assignment": scores, "matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1,
            "line_log_assignment": lscores, "line_matches0": lm0, "line_matches1": lm1, "line_matching_scores0": lms0,
            "line_matching_scores1": lms1, "raw_line_scores": raw}


def _sub_loss(la, asg, m0, m1, bal):
    """GlueStick.sub_loss (gluestick.py:379-415)."""
    pos = asg.to(la.dtype)
    num_pos = pos.sum((1, 2)).clamp(min=1.0)
    neg0, neg1 = (m0 == -1).to(la.dtype), (m1 == -1).to(la.dtype)
    num_neg = (neg0.sum(1) + neg1.sum(1)).clamp(min=1.0)
    nll_pos = -(la[:, :-1, :-1] * pos).sum((1, 2)) / num_pos
    nll_neg = (-(la[:, :-1, -1] * neg0).sum(1) - (la[:, -1, :-1] * neg1).sum(1)) / num_neg
    return bal * nll_pos + (1 - bal) * nll_neg, num_pos, num_neg


def gluestick_loss(w, pred, data, conf):
    """GlueStick.loss (gluestick.py:417-462), training branch (no metrics)."""
    lc = conf.get("loss", {})
    bal, wgt = lc.get("nll_balancing", 0.5), lc.get("nll_weight", 1.0)
    losses = {"total": 0}
    for prefix, bin_name in (("", "bin_score"), ("line_", "line_bin_score")):
        la = pred[prefix + "log_assignment"]
        nll, num_pos, num_neg = _sub_loss(la, data["gt_" + prefix + "assignment"], data["gt_" + prefix + "matches0"],
                                          data["gt_" + prefix + "matches1"], bal)
        losses[prefix + "assignment_nll"] = nll
        losses["total"] = losses["total"] + nll * wgt
        losses[prefix + "num_matchable"] = num_pos
        losses[prefix + "num_unmatchable"] = num_neg
        losses[prefix + "sinkhorn_norm"] = la.exp()[:, :-1].sum(2).mean(1)
        losses[prefix + "bin_score"] = w[bin_name][None]
    return losses
