"""Hand-scheduled forward/backward of the matcher's two repeating units.

`LayerFn`  : one LightGlue transformer layer (self block on both images + bidirectional cross block,
             lightglue.py:224-245) as ONE autograd node.
`HeadFn`   : one deep-supervision head (MatchAssignment + NLL terms + argmax, lightglue.py:271-290,
             losses.py:6-73) as ONE autograd node.

Compared with composing the same kernels through torch autograd op by op (the `engine: autograd`
path kept in matchers/lightglue.py as the cross-check), the explicit schedule removes the per-op
dtype casts, the cat/slice copies and ~2/3 of the launches: weights are read from a compute-dtype
shadow refreshed once per step, weight gradients are produced in fp32 directly by the GEMM
(`out_dtype`), the residual update emits the next GEMM's operand in the same pass, and nothing
N x N is ever kept besides `sim`.

All dense projections are plain cuBLAS GEMMs (`torch.mm/addmm`); everything else is an lgb200 kernel.
"""
import torch
import torch.nn.functional as F

from . import ops

_OUT_DTYPE_OK = True
_ADDMM_DTYPE_OK = True
_head_counters = {}


def _wgrad(dy, a):
    """dW [out, in] fp32 = dy^T a for compute-dtype dy [T, out], a [T, in]."""
    global _OUT_DTYPE_OK
    if dy.dtype == torch.float32:
        return torch.mm(dy.t(), a)
    # the split-K tcgen05 GEMM beats cuBLAS only while one 128x128 tile grid leaves most SMs to the K splits
    # (256x256: 38 vs 46 us at T = 131072; wider outputs re-read the operands through L2 and tie at ~50 us)
    if (dy.dtype == torch.bfloat16 and dy.shape[1] * a.shape[1] <= 256 * 256 and dy.stride(1) == 1 and a.stride(1) == 1
            and dy.stride(0) % 8 == 0 and a.stride(0) % 8 == 0 and dy.data_ptr() % 16 == 0 and a.data_ptr() % 16 == 0):
        return ops.wgrad_bf16(dy, a)
    if _OUT_DTYPE_OK:
        try:
            return torch.mm(dy.t(), a, out_dtype=torch.float32)
        except (TypeError, RuntimeError):
            _OUT_DTYPE_OK = False
    return torch.mm(dy.t(), a).float()


def _bgrad(dy):
    return ops.colsum(dy)


def _dgrad_acc(acc, dy, w):
    """acc (fp32 [T, in]) += dy (compute dtype [T, out]) @ w ([out, in]), IN PLACE, in the GEMM epilogue when
    torch exposes addmm(out_dtype=..., out=acc) for bf16 operands, else GEMM + mixed-dtype add.  Every `acc`
    handed in is a gradient buffer this engine owns (HeadFn's fresh dx or autograd's accumulation result), so
    overwriting it saves the 134 MB copy an out-of-place addmm starts with."""
    global _ADDMM_DTYPE_OK
    if dy.dtype == torch.float32:
        return acc.addmm_(dy, w)
    if _ADDMM_DTYPE_OK:
        try:
            return torch.addmm(acc, dy, w, out_dtype=torch.float32, out=acc)
        except (TypeError, RuntimeError):
            _ADDMM_DTYPE_OK = False
    return acc.add_(torch.mm(dy, w))


def _attend_fwd(q, k, v, sizes, H, cross):
    """q,k,v [T, H*64] over tokens [image0 (B*M); image1 (B*N)] -> (out [T, H*64], lse tuple)."""
    B, M, N = sizes
    scale = 0.125
    if M == N:
        shp = (2 * B, M, H, 64)
        out, lse = ops.attn_fwd(q.view(shp), k.view(shp), v.view(shp), B if cross else 0, scale)
        return out.view(q.shape), (lse,)
    t0 = B * M
    q0, q1 = q[:t0].view(B, M, H, 64), q[t0:].view(B, N, H, 64)
    k0, k1 = k[:t0].view(B, M, H, 64), k[t0:].view(B, N, H, 64)
    v0, v1 = v[:t0].view(B, M, H, 64), v[t0:].view(B, N, H, 64)
    if cross:
        o0, l0 = ops.attn_fwd(q0, k1, v1, 0, scale)
        o1, l1 = ops.attn_fwd(q1, k0, v0, 0, scale)
    else:
        o0, l0 = ops.attn_fwd(q0, k0, v0, 0, scale)
        o1, l1 = ops.attn_fwd(q1, k1, v1, 0, scale)
    return torch.cat([o0.reshape(t0, -1), o1.reshape(B * N, -1)], 0), (l0, l1)


def _attend_bwd(q, k, v, out, lses, dout, sizes, H, cross):
    B, M, N = sizes
    scale = 0.125
    if M == N:
        shp = (2 * B, M, H, 64)
        dq, dk, dv = ops.attn_bwd(q.view(shp), k.view(shp), v.view(shp), out.view(shp), lses[0], dout.view(shp),
                                  B if cross else 0, scale)
        return dq.view(q.shape), dk.view(q.shape), dv.view(q.shape)
    t0 = B * M
    sp = lambda t: (t[:t0].view(B, M, H, 64), t[t0:].view(B, N, H, 64))  # noqa: E731
    (q0, q1), (k0, k1), (v0, v1), (o0, o1), (g0, g1) = sp(q), sp(k), sp(v), sp(out), sp(dout.contiguous())
    if cross:
        dq0, dk1, dv1 = ops.attn_bwd(q0, k1, v1, o0, lses[0], g0, 0, scale)
        dq1, dk0, dv0 = ops.attn_bwd(q1, k0, v0, o1, lses[1], g1, 0, scale)
    else:
        dq0, dk0, dv0 = ops.attn_bwd(q0, k0, v0, o0, lses[0], g0, 0, scale)
        dq1, dk1, dv1 = ops.attn_bwd(q1, k1, v1, o1, lses[1], g1, 0, scale)
    cat = lambda a, b: torch.cat([a.reshape(t0, -1), b.reshape(B * N, -1)], 0)  # noqa: E731
    return cat(dq0, dq1), cat(dk0, dk1), cat(dv0, dv1)


# order of the per-layer tensors (reference names, lightglue.py:139-148, 174-183)
LAYER_PARAMS = [
    "self_attn.Wqkv.weight", "self_attn.Wqkv.bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
    "self_attn.ffn.0.weight", "self_attn.ffn.0.bias", "self_attn.ffn.1.weight", "self_attn.ffn.1.bias",
    "self_attn.ffn.3.weight", "self_attn.ffn.3.bias",
    "cross_attn.to_qk.weight", "cross_attn.to_qk.bias", "cross_attn.to_v.weight", "cross_attn.to_v.bias",
    "cross_attn.to_out.weight", "cross_attn.to_out.bias", "cross_attn.ffn.0.weight", "cross_attn.ffn.0.bias",
    "cross_attn.ffn.1.weight", "cross_attn.ffn.1.bias", "cross_attn.ffn.3.weight", "cross_attn.ffn.3.bias",
]
_LN_SLOTS = (6, 7, 18, 19)  # LayerNorm affine stays fp32


class LayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, sizes, H, cdt, eps, w, *params):
        (Wqkv, bqkv, Wo, bo, W0, b0, g1, be1, W3, b3,
         Wqk, bqk, Wv, bv, Wout, bout, W0c, b0c, g2, be2, W3c, b3c) = w
        D = x.shape[1]
        x = x.contiguous()
        # The FFN input cat([x, msg], -1) (lightglue.py:162, 219) is never concatenated: the compute-dtype copy of x
        # and the out-projection's result are written straight into the two halves of one [T, 2D] buffer, which
        # then feeds ffn.0 as ONE K=2D GEMM (and, in backward, one [2D x 2D] weight-gradient GEMM).
        T = x.shape[0]
        # ---- self block (lightglue.py:150-163)
        cat1 = torch.empty(T, 2 * D, device=x.device, dtype=cdt)
        x16, msg = cat1[:, :D], cat1[:, D:]
        ops.residual_add_cast(x, None, cdt, out_cast=x16)
        qkv = torch.addmm(bqkv, x16, Wqkv.t())
        q, k, v = ops.rope_fwd(qkv, theta, H)
        del qkv
        att, lse1 = _attend_fwd(q, k, v, sizes, H, cross=False)
        torch.addmm(bo, att, Wo.t(), out=msg)
        h = torch.addmm(b0, cat1, W0.t())
        g, mean1, rstd1 = ops.ln_gelu_fwd(h, g1, be1, eps)
        y = torch.addmm(b3, g, W3.t())
        cat2 = torch.empty(T, 2 * D, device=x.device, dtype=cdt)
        x1_16, msg2 = cat2[:, :D], cat2[:, D:]
        x1, _ = ops.residual_add_cast(x, y, cdt, out_cast=x1_16)
        del y
        # ---- cross block (lightglue.py:195-221)
        qk = torch.addmm(bqk, x1_16, Wqk.t())
        vv = torch.addmm(bv, x1_16, Wv.t())
        m, lse2 = _attend_fwd(qk, qk, vv, sizes, H, cross=True)
        torch.addmm(bout, m, Wout.t(), out=msg2)
        h2 = torch.addmm(b0c, cat2, W0c.t())
        gg, mean2, rstd2 = ops.ln_gelu_fwd(h2, g2, be2, eps)
        y2 = torch.addmm(b3c, gg, W3c.t())
        x2, _ = ops.residual_add_cast(x1, y2, None)
        ctx.save_for_backward(theta, cat1, q, k, v, att, h, mean1, rstd1, g, cat2, qk, vv, m, h2, mean2,
                              rstd2, gg, *lse1, *lse2, *w)
        ctx.meta = (sizes, H, cdt, len(lse1), len(lse2), D)
        return x2

    @staticmethod
    def backward(ctx, dx):
        sizes, H, cdt, n1, n2, D = ctx.meta
        sv = ctx.saved_tensors
        (theta, cat1, q, k, v, att, h, mean1, rstd1, g, cat2, qk, vv, m, h2, mean2, rstd2, gg) = sv[:18]
        x16, x1_16 = cat1[:, :D], cat2[:, :D]
        lse1, lse2 = sv[18:18 + n1], sv[18 + n1:18 + n1 + n2]
        (Wqkv, bqkv, Wo, bo, W0, b0, g1, be1, W3, b3,
         Wqk, bqk, Wv, bv, Wout, bout, W0c, b0c, g2, be2, W3c, b3c) = sv[18 + n1 + n2:]
        dx = dx.contiguous()
        # ---- cross block
        dy2 = dx.to(cdt)
        dW3c, db3c = _wgrad(dy2, gg), _bgrad(dy2)
        dgg = torch.mm(dy2, W3c)
        dh2, dg2, dbe2, db0c = ops.ln_gelu_bwd(dgg, h2, g2, be2, mean2, rstd2, want_dxsum=True)
        del dgg
        dW0c = _wgrad(dh2, cat2)
        dmsg2 = torch.mm(dh2, W0c[:, D:])
        dx1 = _dgrad_acc(dx, dh2, W0c[:, :D])  # fp32 accumulation of the residual-stream gradient
        del dh2
        dWout, dbout = _wgrad(dmsg2, m), _bgrad(dmsg2)
        dm = torch.mm(dmsg2, Wout)
        dq_, dk_, dvv = _attend_bwd(qk, qk, vv, m, lse2, dm, sizes, H, cross=True)
        dqk = dq_.add_(dk_)  # the shared to_qk projection is query in one direction and key in the other
        dWqk, dbqk = _wgrad(dqk, x1_16), _bgrad(dqk)
        dWv, dbv = _wgrad(dvv, x1_16), _bgrad(dvv)
        dx1 = _dgrad_acc(dx1, dqk, Wqk)
        dx1 = _dgrad_acc(dx1, dvv, Wv)
        # ---- self block
        dy = dx1.to(cdt)
        dW3, db3 = _wgrad(dy, g), _bgrad(dy)
        dg = torch.mm(dy, W3)
        dh, dg1, dbe1, db0 = ops.ln_gelu_bwd(dg, h, g1, be1, mean1, rstd1, want_dxsum=True)
        del dg
        dW0 = _wgrad(dh, cat1)
        dmsg = torch.mm(dh, W0[:, D:])
        dx0 = _dgrad_acc(dx1, dh, W0[:, :D])
        del dh
        dWo, dbo = _wgrad(dmsg, att), _bgrad(dmsg)
        datt = torch.mm(dmsg, Wo)
        dq, dk, dv = _attend_bwd(q, k, v, att, lse1, datt, sizes, H, cross=False)
        dqkv, dtheta = ops.rope_bwd(dq, dk, dv, q, k, theta, H)
        dWqkv, dbqkv = _wgrad(dqkv, x16), _bgrad(dqkv)
        dx0 = _dgrad_acc(dx0, dqkv, Wqkv)
        grads = (dWqkv, dbqkv, dWo, dbo, dW0, db0, dg1, dbe1, dW3, db3,
                 dWqk, dbqk, dWv, dbv, dWout, dbout, dW0c, db0c, dg2, dbe2, dW3c, db3c)
        return (dx0, dtheta, None, None, None, None, None) + grads


class HeadFn(torch.autograd.Function):
    """x [T, D] fp32 (tokens of both images) -> (nll [B], conf [B]) of this layer: the NLL of its assignment
    and the token-confidence BCE against the final layer's argmax (`fin`; None for the last layer),
    plus the detached nll_pos / nll_neg for logging."""

    @staticmethod
    def forward(ctx, x, sizes, cdt, gt, bal, fin, wfp, bfp, Wfp_p, bfp_p, wm, bm, wt, bt):
        B, M, N = sizes
        D = x.shape[1]
        t0 = B * M
        dev = x.device
        x = x.contiguous()
        has_tok = wt is not None
        # compute-dtype x for final_proj + [matchability logit, token-confidence logit] per token, one pass over x
        x16, zt, ls, du = ops.head_token_fwd(x, wm.view(-1), bm, wt.view(-1) if has_tok else None,
                                             bt if has_tok else None, cdt)
        md = torch.addmm(bfp, x16, wfp.t())  # final_proj, un-scaled; d^-1/2 is folded into sim
        md0, md1 = md[:t0].view(B, M, D), md[t0:].view(B, N, D)
        alpha = float(D) ** -0.5
        if cdt == torch.bfloat16:
            sim = ops.gemm_bf16(md0, md1, alpha=alpha)
        else:
            sim = torch.bmm(md0, md1.transpose(1, 2)).mul_(alpha)
        ls0, ls1, du0, du1 = ls[:t0].view(B, M), ls[t0:].view(B, N), du[:t0].view(B, M), du[t0:].view(B, N)
        st = ops.assign_stats(sim, ls0, ls1, du0, du1, gt_u8=gt["u8"], dense=False)
        out = torch.empty(4, B, device=dev, dtype=torch.float32)
        f0, f1 = (fin if (fin is not None and has_tok) else (None, None))
        hws = torch.empty(4 * B * ((M + N + 255) // 256), device=dev, dtype=torch.float32)
        cnt = _head_counters.get(dev)
        if cnt is None:
            cnt = _head_counters[dev] = torch.zeros(4096, device=dev, dtype=torch.int32)
        assert B <= 4096
        ops.call("lgb200_head_terms_fwd", ops.ptr(zt), ops.ptr(st["pos_row_sum"]), ops.ptr(gt["rowcnt"]),
                 ops.ptr(gt["colcnt"]), ops.ptr(gt["neg0"]), ops.ptr(gt["neg1"]), ops.ptr(st["rowmax"]),
                 ops.ptr(st["rowarg"]), ops.ptr(st["colmax"]), ops.ptr(st["colarg"]), ops.ptr(f0), ops.ptr(f1),
                 ops.ptr(gt["num_pos"]), ops.ptr(gt["num_neg"]), float(bal), ops.ptr(out[0]), ops.ptr(out[1]),
                 ops.ptr(out[2]), ops.ptr(out[3]), ops.ptr(hws), ops.ptr(cnt), B, M, N, ops.stream_ptr())
        nll, nll_pos, nll_neg, conf = out[0], out[1], out[2], out[3]
        saved = [x, x16, md, sim, st["lse_row"], st["lse_col"], zt, st["rowmax"], st["rowarg"], st["colmax"],
                 st["colarg"], wfp, wm, gt["u8"], gt["rowcnt"], gt["colcnt"], gt["neg0"], gt["neg1"], gt["num_pos"],
                 gt["num_neg"]]
        if f0 is not None:
            saved += [f0, f1]
        ctx.save_for_backward(*saved)
        ctx.meta = (sizes, cdt, bal, alpha, has_tok, f0 is not None)
        ctx.mark_non_differentiable(nll_pos, nll_neg)
        return nll, conf, nll_pos, nll_neg

    @staticmethod
    def backward(ctx, g_nll, g_conf, *_):
        sv = ctx.saved_tensors
        (x, x16, md, sim, lse_row, lse_col, zt, rowmax, rowarg, colmax, colarg, wfp, wm, gt_u8, rowcnt, colcnt, neg0,
         neg1, num_pos, num_neg) = sv[:20]
        (B, M, N), cdt, bal, alpha, has_tok, has_fin = ctx.meta
        f0, f1 = (sv[20], sv[21]) if has_fin else (None, None)
        D = x.shape[1]
        t0 = B * M
        g_nll = g_nll.float().contiguous()
        g_conf = g_conf.float().contiguous() if g_conf is not None else torch.zeros_like(g_nll)
        dzt = torch.empty_like(zt)
        ops.call("lgb200_head_terms_bwd", ops.ptr(zt), ops.ptr(rowcnt), ops.ptr(colcnt), ops.ptr(neg0), ops.ptr(neg1),
                 ops.ptr(rowmax), ops.ptr(rowarg), ops.ptr(colmax), ops.ptr(colarg), ops.ptr(f0), ops.ptr(f1),
                 ops.ptr(num_pos), ops.ptr(num_neg), float(bal), ops.ptr(g_nll), ops.ptr(g_conf), ops.ptr(dzt), B, M, N,
                 ops.stream_ptr())
        # similarity: dsim = gc (2 gt - softmax_row * rowcnt - softmax_col * colcnt), gc includes d^-1/2
        gc = (g_nll * (-bal * alpha) / num_pos).contiguous()
        tc = cdt == torch.bfloat16 and N % 8 == 0 and M % 8 == 0
        dsim = torch.empty(B, M, N, device=x.device, dtype=torch.bfloat16 if tc else torch.float32)
        ops.call("lgb200_assign_bwd", ops.ptr(sim), ops.ptr(lse_row), ops.ptr(lse_col), ops.ptr(gt_u8), ops.ptr(gc),
                 ops.ptr(rowcnt), ops.ptr(colcnt), ops.ptr(dsim), ops._code(dsim.dtype), B, M, N, ops.stream_ptr())
        md0, md1 = md[:t0].view(B, M, D), md[t0:].view(B, N, D)
        if tc:
            dmd0 = ops.gemm_bf16(dsim, md1, a_mn_major=False, b_mn_major=True, out_dtype=cdt)  # dsim   md1
            dmd1 = ops.gemm_bf16(dsim, md0, a_mn_major=True, b_mn_major=True, out_dtype=cdt)   # dsim^T md0
        else:
            dmd0 = torch.bmm(dsim, md1.float()).to(cdt)
            dmd1 = torch.bmm(dsim.transpose(1, 2), md0.float()).to(cdt)
        dmd = torch.cat([dmd0.reshape(t0, D), dmd1.reshape(B * N, D)], 0)
        dWfp, dbfp = _wgrad(dmd, x16), _bgrad(dmd)
        # dx = dmd W_fp + dzt[:,0] wm (the token-confidence head reads a detached x, lightglue.py:82-83), dW2, db2
        dx, dW2, db2 = ops.head_token_bwd(x, torch.mm(dmd, wfp), dzt, wm.view(-1))
        dwt, dbt = (dW2[1:2], db2[1:2]) if has_tok else (None, None)
        return dx, None, None, None, None, None, None, None, dWfp, dbfp, dW2[0:1], db2[0:1], dwt, dbt
