"""B200-native GlueStick matcher (points + lines) -- drop-in for gluefactory's in-tree `matchers.gluestick`
(BASELINE.json configs[4]: SP + LSD + GlueStick).

Select it with `model.matcher.name: gluefactory_b200.matchers.gluestick`.  Same constructor, `forward(data) -> pred`,
`loss(pred, data) -> (losses, metrics)`, `required_data_keys`, and the same module tree, so parameter / buffer names
and shapes equal the reference's state_dict (Conv1d weights [out, in, 1], BatchNorm running statistics) and
checkpoints are interchangeable (gluestick.py:68-141).

Layout and kernels.  The reference keeps channels first ([B, D, N], Conv1d with kernel 1, gluestick.py:158); here the
node features are token-major [B*N, D] like in the LightGlue plugin, so
  * every Conv1d(k=1) with a tensor-core-sized input is the library's tcgen05 GEMM (`ops.LinearFn`: bias in the
    epilogue, split-K weight gradient); the tiny encoder layers (3 / 5 / 32 input channels) are plain fp32 linears;
  * the attention core (gluestick.py:524-529) is the same tcgen05 flash-attention pair as LightGlue's.  GlueStick splits
    heads as channel = d * H + h; instead of permuting activations, the rows of the q / k / v projection weights and the
    columns of the merge weight are permuted once per call so the projections emit head-major [B, N, H, 64] directly;
  * the assignment is `log_double_softmax` with a learnt bin (gluestick.py:772-783) on the similarity produced by the
    batched tcgen05 GEMM: forward kernels + closed-form backward (`ops.log_double_softmax`, heads_grad.py);
  * match filtering (`_get_matches`, gluestick.py:318-331) is `filter_matches_kernel`;
  * BatchNorm1d (training-mode batch statistics, running-stat update), the LineLayer's gather / scatter-mean over
    junctions and the O(lines^2) line-score gathers are torch ops.
`precision: "bf16"` (default) rounds GEMM / attention operands to bf16 (the reference forces its attention to fp32,
gluestick.py:17-22; the fp32 mode reproduces that); `precision: "fp32"` is the parity path (fp32 cuBLAS linears,
CUDA-core attention).  No CPU fallback.

Not built (raise NotImplementedError): `line_attention`, `inter_supervision`, inputs without lines.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib, ops
from .lightglue import _Conf, _merge, _to_plain, matcher_metrics


def MLP(channels, do_bn=True):
    """gluestick.py:465-475 (same module indices: conv, BN, ReLU, conv, ...)."""
    layers = []
    n = len(channels)
    for i in range(1, n):
        layers.append(nn.Conv1d(channels[i - 1], channels[i], kernel_size=1, bias=True))
        if i < n - 1:
            if do_bn:
                layers.append(nn.BatchNorm1d(channels[i]))
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class KeypointEncoder(nn.Module):
    def __init__(self, feature_dim, layers):
        super().__init__()
        self.encoder = MLP([3] + list(layers) + [feature_dim])
        nn.init.constant_(self.encoder[-1].bias, 0.0)


class EndPtEncoder(nn.Module):
    def __init__(self, feature_dim, layers):
        super().__init__()
        self.encoder = MLP([5] + list(layers) + [feature_dim])
        nn.init.constant_(self.encoder[-1].bias, 0.0)


class MultiHeadedAttention(nn.Module):
    def __init__(self, h, d_model):
        super().__init__()
        self.dim, self.h = d_model // h, h
        self.merge = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.proj = nn.ModuleList([nn.Conv1d(d_model, d_model, kernel_size=1) for _ in range(3)])


class AttentionalPropagation(nn.Module):
    def __init__(self, num_dim, num_heads):
        super().__init__()
        self.attn = MultiHeadedAttention(num_heads, num_dim)
        self.mlp = MLP([num_dim * 2, num_dim * 2, num_dim])
        nn.init.constant_(self.mlp[-1].bias, 0.0)


class GNNLayer(nn.Module):
    def __init__(self, feature_dim, layer_type):
        super().__init__()
        assert layer_type in ("cross", "self")
        self.type = layer_type
        self.update = AttentionalPropagation(feature_dim, 4)


class LineLayer(nn.Module):
    def __init__(self, feature_dim):
        super().__init__()
        self.mlp = MLP([feature_dim * 3, feature_dim * 2, feature_dim])


class AttentionalGNN(nn.Module):
    def __init__(self, feature_dim, layer_types):
        super().__init__()
        self.layers = nn.ModuleList([GNNLayer(feature_dim, t) for t in layer_types])
        self.line_layers = nn.ModuleList([LineLayer(feature_dim) for _ in range(len(layer_types) // 2)])


class _BmmNT(torch.autograd.Function):
    """sim [B,M,N] fp32 = alpha * a [B,M,D] . b [B,N,D]^T on the batched tcgen05 GEMM, with its two backward
    contractions (gluestick.py:248-249 `einsum("bdn,bdm->bnm")`)."""

    @staticmethod
    def forward(ctx, a, b, alpha):
        ctx.save_for_backward(a, b)
        ctx.alpha = alpha
        return ops.gemm_bf16(a, b, alpha=alpha)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g16 = g.to(torch.bfloat16).contiguous()
        da = ops.gemm_bf16(g16, b, a_mn_major=False, b_mn_major=True, out_dtype=torch.bfloat16, alpha=ctx.alpha)
        db = ops.gemm_bf16(g16, a, a_mn_major=True, b_mn_major=True, out_dtype=torch.bfloat16, alpha=ctx.alpha)
        return da, db, None


class GlueStick(nn.Module):
    default_conf = {
        "name": "gluestick",
        "input_dim": 256,
        "descriptor_dim": 256,
        "weights": None,
        "version": "v0.1_arxiv",
        "keypoint_encoder": [32, 64, 128, 256],
        "GNN_layers": ["self", "cross"] * 9,
        "num_line_iterations": 1,
        "line_attention": False,
        "filter_threshold": 0.2,
        "checkpointed": False,  # accepted; the fused attention keeps no N x N activations to checkpoint
        "skip_init": False,
        "inter_supervision": None,
        "loss": {"nll_weight": 1.0, "nll_balancing": 0.5, "inter_supervision": [0.3, 0.6]},
        "precision": "bf16",
    }
    required_data_keys = ["view0", "view1", "keypoints0", "keypoints1", "descriptors0", "descriptors1",
                          "keypoint_scores0", "keypoint_scores1", "lines0", "lines1", "lines_junc_idx0",
                          "lines_junc_idx1", "line_scores0", "line_scores1"]

    def __init__(self, conf=None):
        super().__init__()
        self.conf = conf = _Conf(_merge(self.default_conf, _to_plain(conf)))
        assert conf.precision in ("bf16", "fp32"), conf.precision
        if conf.line_attention or conf.inter_supervision or conf.skip_init:
            raise NotImplementedError("gluefactory_b200 gluestick: line_attention / inter_supervision / skip_init are not built")
        D = conf.descriptor_dim
        assert D == 256, "the lgb200 attention kernels are built for 4 heads of 64 channels"
        if conf.input_dim != D:
            self.input_proj = nn.Conv1d(conf.input_dim, D, kernel_size=1)
            nn.init.constant_(self.input_proj.bias, 0.0)
        # registration order = the reference's (gluestick.py:69-113): state_dict keys come out identical
        bin_score = nn.Parameter(torch.tensor(1.0))
        line_bin_score = nn.Parameter(torch.tensor(1.0))
        self.kenc = KeypointEncoder(D, conf.keypoint_encoder)
        self.lenc = EndPtEncoder(D, conf.keypoint_encoder)
        self.gnn = AttentionalGNN(D, list(conf.GNN_layers))
        self.final_proj = nn.Conv1d(D, D, kernel_size=1)
        nn.init.constant_(self.final_proj.bias, 0.0)
        nn.init.orthogonal_(self.final_proj.weight, gain=1)
        self.final_line_proj = nn.Conv1d(D, D, kernel_size=1)
        nn.init.constant_(self.final_line_proj.bias, 0.0)
        nn.init.orthogonal_(self.final_line_proj.weight, gain=1)
        self.register_parameter("bin_score", bin_score)
        self.register_parameter("line_bin_score", line_bin_score)
        if conf.weights:
            sd = torch.load(conf.weights, map_location="cpu")
            if "model" in sd:  # full glue-factory checkpoint (gluestick.py:131-139)
                sd = {k.replace("matcher.", "").replace("module.", ""): v for k, v in sd["model"].items() if "matcher." in k}
            self.load_state_dict(sd, strict=False)
        # head permutation: reference channel c = d * H + h  <->  head-major channel h * 64 + d
        H, dh = 4, D // 4
        c = torch.arange(D)
        self.register_buffer("_perm", ((c % dh) * H + c // dh), persistent=False)

    # ------------------------------------------------------------------------------------------ building blocks
    @property
    def _bf16(self):
        return self.conf.precision == "bf16"

    def _conv(self, x, conv, w=None, b=None, keep_bf16=False):
        """Conv1d(kernel 1) on token-major x [T, C_in]; w / b override the module's parameters (permuted views).
        bf16 mode: x may already be bf16 (no second cast), and keep_bf16 leaves the GEMM's bf16 output as it is for a
        consumer that rounds to bf16 anyway (the next GEMM, BatchNorm + ReLU in front of one, the attention kernels)."""
        w = conv.weight[:, :, 0] if w is None else w
        b = conv.bias if b is None else b
        if self._bf16 and w.shape[1] % 64 == 0 and w.shape[0] % 8 == 0:
            y = ops.LinearFn.apply(x.to(torch.bfloat16).contiguous(), w, b)
            return y if keep_bf16 else y.float()
        return F.linear(x.float(), w, b)

    def _mlp(self, x, seq):
        mods = list(seq)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Conv1d):
                # an inner layer's output only feeds BatchNorm / ReLU and the next GEMM, which rounds it to bf16 at the
                # same point either way: keep it bf16 (BatchNorm still accumulates its statistics in fp32)
                x = self._conv(x, m, keep_bf16=self._bf16 and i < len(mods) - 1)
            elif isinstance(m, nn.BatchNorm1d):
                x = m(x if x.dtype == torch.bfloat16 else x.float())  # [T, C]: statistics over all tokens of this call
            else:
                x = torch.relu(x)
        return x

    def _attn_weights(self, attn):
        """The projection / merge weights of one MultiHeadedAttention with the head permutation applied (rows of q, k, v,
        columns of merge); computed once per layer and shared by both images."""
        perm = self._perm
        return ([(p.weight[:, :, 0][perm], p.bias[perm]) for p in attn.proj], attn.merge.weight[:, :, 0][:, perm])

    def _attention(self, x, src, attn, B, wts=None):
        """MultiHeadedAttention (gluestick.py:532-551): x [B*N, D] attends to src [B*M, D]."""
        D = x.shape[1]
        proj_w, merge_w = wts if wts is not None else self._attn_weights(attn)
        if self._bf16:  # one cast per input, projections stay bf16 into the attention kernels and out of them
            x, src = x.to(torch.bfloat16), (src.to(torch.bfloat16) if src is not x else None)
            src = x if src is None else src
        q, k, v = (self._conv(t, p, w, b, keep_bf16=True) for p, (w, b), t in zip(attn.proj, proj_w, (x, src, src)))
        cdt = torch.bfloat16 if self._bf16 else torch.float32
        shp = lambda t: t.to(cdt).view(B, -1, 4, D // 4)  # noqa: E731  head-major thanks to the permuted rows
        o = ops.Attention.apply(shp(q), shp(k), shp(v), 0, (D // 4) ** -0.5)
        return self._conv(o.reshape(-1, D), attn.merge, merge_w, attn.merge.bias)

    def _gnn_layer(self, d0, d1, layer, B):
        s0, s1 = (d1, d0) if layer.type == "cross" else (d0, d1)
        wts = self._attn_weights(layer.update.attn)
        out = []
        for x, src in ((d0, s0), (d1, s1)):
            msg = self._attention(x, src, layer.update.attn, B, wts)
            out.append(x + self._mlp(torch.cat([x, msg], 1), layer.update.mlp))
        return out[0], out[1]

    def _line_layer(self, d0, d1, enc0, enc1, idx0, idx1, layer, B):
        """LineLayer (gluestick.py:589-691, line_attention = False)."""
        out = []
        for d, enc, idx in ((d0, enc0, idx0), (d1, enc1, idx1)):
            D = d.shape[1]
            n = d.shape[0] // B
            n2 = idx.shape[1]
            flat = (idx + torch.arange(B, device=idx.device)[:, None] * n).reshape(-1)  # rows of the [B*n, D] matrix
            ld = d.index_select(0, flat)
            other = ld.view(B, n2 // 2, 2, D).flip(2).reshape(B * n2, D)
            upd = self._mlp(torch.cat([ld, other, enc], 1), layer.mlp)
            acc = torch.zeros_like(d).scatter_reduce(0, flat[:, None].expand(-1, D), upd.to(d.dtype), reduce="mean",
                                                     include_self=False)
            out.append(d + acc)
        return out[0], out[1]

    @staticmethod
    def _norm_kpts(kpts, size):
        """gluestick.py:478-490."""
        size = size.to(kpts)
        return (kpts - (size / 2)[:, None, :]) / (size.max(1, keepdim=True).values * 0.7)[:, None, :]

    def _matches(self, scores):
        """_get_matches (gluestick.py:318-331) on the filter kernel."""
        inner = scores[:, :-1, :-1]
        mx0, mx1 = inner.max(2), inner.max(1)
        return ops.filter_matches(mx0.values.contiguous(), mx0.indices.to(torch.int32).contiguous(),
                                  mx1.indices.to(torch.int32).contiguous(), self.conf.filter_threshold)

    def _similarity(self, m0, m1, B):
        D = m0.shape[1]
        a, b = m0.view(B, -1, D), m1.view(B, -1, D)
        if self._bf16 and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0:
            return _BmmNT.apply(a.to(torch.bfloat16).contiguous(), b.to(torch.bfloat16).contiguous(), D ** -0.5)
        return torch.bmm(a, b.transpose(1, 2)) / D ** 0.5

    # ------------------------------------------------------------------------------------------ forward / loss
    def forward(self, data):
        with torch.autocast(device_type="cuda", enabled=False):
            return self._forward(data)

    def _forward(self, data):
        for key in self.required_data_keys:
            assert key in data, f"Missing key {key} in data"
        _lib.load()  # raises unless the CUDA library is built and the device is a B200 (no CPU fallback)
        kp0, kp1 = data["keypoints0"].float(), data["keypoints1"].float()
        B, n0 = kp0.shape[:2]
        n1 = kp1.shape[1]
        L0, L1 = data["lines0"].shape[1], data["lines1"].shape[1]
        if n0 == 0 or n1 == 0 or L0 == 0 or L1 == 0:
            raise NotImplementedError("gluefactory_b200 gluestick: empty keypoint / line sets are not supported")
        size0, size1 = data["view0"]["image_size"], data["view1"]["image_size"]
        D = self.conf.descriptor_dim
        d0 = data["descriptors0"].float().reshape(B * n0, -1)
        d1 = data["descriptors1"].float().reshape(B * n1, -1)
        if self.conf.input_dim != D:
            d0, d1 = self._conv(d0, self.input_proj), self._conv(d1, self.input_proj)
        enc_in = lambda k, s: torch.cat([k, s.float()[..., None]], -1).reshape(-1, 3)  # noqa: E731
        d0 = d0 + self._mlp(enc_in(self._norm_kpts(kp0, size0), data["keypoint_scores0"]), self.kenc.encoder)
        d1 = d1 + self._mlp(enc_in(self._norm_kpts(kp1, size1), data["keypoint_scores1"]), self.kenc.encoder)

        def line_enc(lines, scores, size):
            Ln = lines.shape[1]
            ln = self._norm_kpts(lines.float().reshape(B, 2 * Ln, 2), size).reshape(B, Ln, 2, 2)
            off = ln[:, :, 1] - ln[:, :, 0]
            off = torch.stack([off, -off], 2).reshape(B, 2 * Ln, 2)
            # the reference tiles the line scores over the interleaved endpoints (`scores.repeat(1, 2)`, gluestick.py:520)
            inp = torch.cat([ln.reshape(B, 2 * Ln, 2), off, scores.float().repeat(1, 2)[..., None]], -1)
            return self._mlp(inp.reshape(-1, 5), self.lenc.encoder)

        enc0, enc1 = line_enc(data["lines0"], data["line_scores0"], size0), line_enc(data["lines1"], data["line_scores1"], size1)
        idx0, idx1 = data["lines_junc_idx0"].reshape(B, -1), data["lines_junc_idx1"].reshape(B, -1)
        for i, layer in enumerate(self.gnn.layers):
            d0, d1 = self._gnn_layer(d0, d1, layer, B)
            if layer.type == "self":
                for _ in range(self.conf.num_line_iterations):
                    d0, d1 = self._line_layer(d0, d1, enc0, enc1, idx0, idx1, self.gnn.line_layers[i // 2], B)
        # points (keypoints and line junctions together)
        scores = ops.log_double_softmax(self._similarity(self._conv(d0, self.final_proj), self._conv(d1, self.final_proj), B),
                                        self.bin_score)
        m0, m1, ms0, ms1 = self._matches(scores.detach())
        pred = {"log_assignment": scores, "matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1}
        # lines: junction similarity gathered per endpoint pair, best of the two endpoint orderings (gluestick.py:333-377)
        ld0 = d0.view(B, n0, D)[:, :2 * L0].reshape(-1, D)
        ld1 = d1.view(B, n1, D)[:, :2 * L1].reshape(-1, D)
        s = self._similarity(self._conv(ld0, self.final_line_proj), self._conv(ld1, self.final_line_proj), B)
        n20, n21 = idx0.shape[1], idx1.shape[1]
        s = torch.gather(s, 2, idx1[:, None, :].expand(B, s.shape[1], n21))
        s = torch.gather(s, 1, idx0[:, :, None].expand(B, n20, n21)).reshape(B, n20 // 2, 2, n21 // 2, 2)
        raw = 0.5 * torch.maximum(s[:, :, 0, :, 0] + s[:, :, 1, :, 1], s[:, :, 0, :, 1] + s[:, :, 1, :, 0])
        lscores = ops.log_double_softmax(raw.contiguous(), self.line_bin_score)
        lm0, lm1, lms0, lms1 = self._matches(lscores.detach())
        pred.update({"line_log_assignment": lscores, "line_matches0": lm0, "line_matches1": lm1,
                     "line_matching_scores0": lms0, "line_matching_scores1": lms1, "raw_line_scores": raw})
        return pred

    def _sub_loss(self, pred, data, losses, bin_score, prefix):
        """gluestick.py:379-415."""
        positive = data["gt_" + prefix + "assignment"].float()
        num_pos = positive.sum((1, 2)).clamp(min=1.0)
        neg0 = (data["gt_" + prefix + "matches0"] == -1).float()
        neg1 = (data["gt_" + prefix + "matches1"] == -1).float()
        num_neg = (neg0.sum(1) + neg1.sum(1)).clamp(min=1.0)
        la = pred[prefix + "log_assignment"]
        nll_pos = -(la[:, :-1, :-1] * positive).sum((1, 2)) / num_pos
        nll_neg = (-(la[:, :-1, -1] * neg0).sum(1) - (la[:, -1, :-1] * neg1).sum(1)) / num_neg
        bal = self.conf.loss.nll_balancing
        nll = bal * nll_pos + (1 - bal) * nll_neg
        losses[prefix + "assignment_nll"] = nll
        if self.conf.loss.nll_weight > 0:
            losses["total"] = losses["total"] + nll * self.conf.loss.nll_weight
        losses[prefix + "num_matchable"] = num_pos
        losses[prefix + "num_unmatchable"] = num_neg
        losses[prefix + "sinkhorn_norm"] = la.exp()[:, :-1].sum(2).mean(1)
        losses[prefix + "bin_score"] = bin_score[None]
        return losses

    def loss(self, pred, data):
        """gluestick.py:417-462."""
        with torch.autocast(device_type="cuda", enabled=False):
            losses = {"total": 0}
            losses = self._sub_loss(pred, data, losses, self.bin_score, "")
            losses = self._sub_loss(pred, data, losses, self.line_bin_score, "line_")
            metrics = {}
            if not self.training:
                metrics = {**matcher_metrics(pred, data),
                           **{"line_" + k: v for k, v in matcher_metrics(
                               {"matches0": pred["line_matches0"], "matching_scores0": pred["line_matching_scores0"]},
                               {"gt_matches0": data["gt_line_matches0"]}).items()}}
            return losses, metrics


__main_model__ = GlueStick
