"""autograd glue for the M-A causal-branch kernels (csrc/ma_tail.cu; ABI in include/cvad_b200.h)."""
from __future__ import annotations

import torch

from .ops import _call, _cuda, _f32c, _ptr, _st, grad_buffer, _wants_grad

MAXDET, NF, HID = 5, 6, 64


class _DetDecode(torch.autograd.Function):
    """raw (B,T,5,4) -> box (B,T,5,4); cnt (B,T) int32 and src (B,T,5) int32 are returned non-differentiable."""

    @staticmethod
    def forward(ctx, raw, flag):
        _cuda(raw)
        raw = _f32c(raw)
        B, T = raw.shape[:2]
        R = B * T
        box = torch.empty((B, T, MAXDET, 4), device=raw.device, dtype=torch.float32)
        cnt = torch.empty((B, T), device=raw.device, dtype=torch.int32)
        src = torch.empty((B, T, MAXDET), device=raw.device, dtype=torch.int32)
        _call("cvad_det_decode_f32", _ptr(raw), R, _ptr(box), _ptr(cnt), _ptr(src), _ptr(flag), _st())
        ctx.save_for_backward(raw, src)
        ctx.mark_non_differentiable(cnt, src)
        return box, cnt, src

    @staticmethod
    def backward(ctx, dbox, _c, _s):
        raw, src = ctx.saved_tensors
        dbox = _f32c(dbox)
        draw = torch.empty_like(raw)
        _call("cvad_det_decode_bwd_f32", _ptr(raw), _ptr(dbox), _ptr(src), src.numel() // MAXDET, _ptr(draw), _st())
        return draw, None


def det_decode(raw, flag=None):
    return _DetDecode.apply(raw, flag)


class _TrajAssemble(torch.autograd.Function):
    """box (B,T,5,4), reid (B,T,5,D), cnt (B,T) -> traj (B,5,T,4+D), ntr (B) int32."""

    @staticmethod
    def forward(ctx, box, reid, cnt, flag):
        _cuda(box, reid, cnt)
        box, reid = _f32c(box), _f32c(reid)
        B, T = box.shape[:2]
        D = reid.shape[-1]
        traj = torch.empty((B, MAXDET, T, 4 + D), device=box.device, dtype=torch.float32)
        ntr = torch.empty((B,), device=box.device, dtype=torch.int32)
        _call("cvad_traj_assemble_f32", _ptr(box), _ptr(reid), _ptr(cnt), B, T, D, _ptr(traj), _ptr(ntr), _ptr(flag), _st())
        ctx.save_for_backward(cnt)
        ctx.meta = (B, T, D)
        ctx.mark_non_differentiable(ntr)
        return traj, ntr

    @staticmethod
    def backward(ctx, dtraj, _n):
        (cnt,) = ctx.saved_tensors
        B, T, D = ctx.meta
        dtraj = _f32c(dtraj)
        dbox = torch.empty((B, T, MAXDET, 4), device=dtraj.device, dtype=torch.float32)
        dreid = torch.empty((B, T, MAXDET, D), device=dtraj.device, dtype=torch.float32)
        _call("cvad_traj_assemble_bwd_f32", _ptr(dtraj), _ptr(cnt), B, T, D, _ptr(dbox), _ptr(dreid), _st())
        return dbox, dreid, None, None


def traj_assemble(box, reid, cnt, flag=None):
    return _TrajAssemble.apply(box, reid, cnt, flag)


class _GruLast(torch.autograd.Function):
    """gi (B*5,T,192) precomputed input projection -> last hidden state (B*5,64)."""

    @staticmethod
    def forward(ctx, gi, w_hh, b_hh, ntr):
        _cuda(gi, w_hh, b_hh, ntr)
        gi = _f32c(gi)
        N, T, _ = gi.shape
        B = N // MAXDET
        hT = torch.empty((N, HID), device=gi.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        saved = torch.empty((N, T, 5, HID), device=gi.device, dtype=torch.float32) if need else None
        _call("cvad_gru_fwd_f32", _ptr(gi), _ptr(w_hh), _ptr(b_hh), _ptr(ntr), B, T, _ptr(hT), _ptr(saved), _st())
        ctx.save_for_backward(saved, ntr)
        ctx.w_hh, ctx.b_hh = w_hh, b_hh
        ctx.meta = (B, T)
        return hT

    @staticmethod
    def backward(ctx, dhT):
        saved, ntr = ctx.saved_tensors
        B, T = ctx.meta
        w_hh, b_hh = ctx.w_hh, ctx.b_hh
        dhT = _f32c(dhT)
        dgi = torch.empty((B * MAXDET, T, 3 * HID), device=dhT.device, dtype=torch.float32)
        dw = grad_buffer(w_hh) if _wants_grad(w_hh) else None
        db = grad_buffer(b_hh) if _wants_grad(b_hh) else None
        _call("cvad_gru_bwd_f32", _ptr(dhT), _ptr(saved), _ptr(w_hh), _ptr(ntr), B, T, _ptr(dgi), _ptr(dw), _ptr(db), _st())
        return dgi, None, None, None


def gru_last(gi, w_hh, b_hh, ntr):
    return _GruLast.apply(gi, w_hh, b_hh, ntr)


class _ReparamKl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, lv, eps, ntr):
        _cuda(mu, lv, eps, ntr)
        mu, lv, eps = _f32c(mu), _f32c(lv), _f32c(eps)
        B = ntr.shape[0]
        z = torch.empty((B, MAXDET, NF), device=mu.device, dtype=torch.float32)
        kl = torch.empty((B,), device=mu.device, dtype=torch.float32)
        _call("cvad_reparam_kl_f32", _ptr(mu), _ptr(lv), _ptr(eps), _ptr(ntr), B, _ptr(z), _ptr(kl), _st())
        ctx.save_for_backward(mu, lv, eps, ntr)
        return z, kl

    @staticmethod
    def backward(ctx, dz, dkl):
        mu, lv, eps, ntr = ctx.saved_tensors
        B = ntr.shape[0]
        dz, dkl = _f32c(dz), _f32c(dkl)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(lv)
        _call("cvad_reparam_kl_bwd_f32", _ptr(mu), _ptr(lv), _ptr(eps), _ptr(ntr), B, _ptr(dz), _ptr(dkl), _ptr(dmu), _ptr(dlv), _st())
        return dmu, dlv, None, None


def reparam_kl(mu, lv, eps, ntr):
    return _ReparamKl.apply(mu, lv, eps, ntr)


class _PairConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node):
        _cuda(node)
        node = _f32c(node)
        B, _, Hn = node.shape
        pair = torch.empty((B, MAXDET, MAXDET, 2 * Hn), device=node.device, dtype=torch.float32)
        _call("cvad_pair_concat_f32", _ptr(node), B, Hn, _ptr(pair), _st())
        ctx.meta = (B, Hn)
        return pair

    @staticmethod
    def backward(ctx, dpair):
        B, Hn = ctx.meta
        dpair = _f32c(dpair)
        dnode = torch.empty((B, MAXDET, Hn), device=dpair.device, dtype=torch.float32)
        _call("cvad_pair_concat_bwd_f32", _ptr(dpair), B, Hn, _ptr(dnode), _st())
        return dnode


def pair_concat(node):
    return _PairConcat.apply(node)


class _AdjAssemble(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, ntr):
        _cuda(e, ntr)
        e = _f32c(e)
        B = ntr.shape[0]
        adj = torch.empty((B, NF, NF), device=e.device, dtype=torch.float32)
        _call("cvad_adj_assemble_f32", _ptr(e), _ptr(ntr), B, _ptr(adj), 0, _st())
        ctx.save_for_backward(ntr)
        ctx.eshape = e.shape
        return adj

    @staticmethod
    def backward(ctx, dadj):
        (ntr,) = ctx.saved_tensors
        B = ntr.shape[0]
        dadj = _f32c(dadj)
        de = torch.empty(ctx.eshape, device=dadj.device, dtype=torch.float32)
        _call("cvad_adj_assemble_f32", _ptr(dadj), _ptr(ntr), B, _ptr(de), 1, _st())
        return de, None


def adj_assemble(e, ntr):
    return _AdjAssemble.apply(e, ntr)


class _Structured(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adj, z):
        _cuda(adj, z)
        adj, z = _f32c(adj), _f32c(z)
        B = adj.shape[0]
        out = torch.empty_like(z)
        _call("cvad_structured_f32", _ptr(adj), _ptr(z), B, _ptr(out), _st())
        ctx.save_for_backward(adj, z)
        return out

    @staticmethod
    def backward(ctx, dout):
        adj, z = ctx.saved_tensors
        dout = _f32c(dout)
        dadj, dz = torch.empty_like(adj), torch.empty_like(z)
        _call("cvad_structured_bwd_f32", _ptr(adj), _ptr(z), _ptr(dout), adj.shape[0], _ptr(dadj), _ptr(dz), _st())
        return dadj, dz


def structured(adj, z):
    return _Structured.apply(adj, z)


class _ScorerInputs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, pred, ntr):
        _cuda(z, pred, ntr)
        z, pred = _f32c(z), _f32c(pred)
        B = ntr.shape[0]
        cin = torch.empty((B, 18), device=z.device, dtype=torch.float32)
        mn = torch.empty((B, 12), device=z.device, dtype=torch.float32)
        tin = torch.empty((B, 6), device=z.device, dtype=torch.float32)
        _call("cvad_scorer_inputs_f32", _ptr(z), _ptr(pred), _ptr(ntr), B, _ptr(cin), _ptr(mn), _ptr(tin), _st())
        ctx.save_for_backward(cin, ntr)
        return cin, mn, tin

    @staticmethod
    def backward(ctx, dcin, dmn, dtin):
        cin, ntr = ctx.saved_tensors
        B = ntr.shape[0]
        dcin, dmn, dtin = _f32c(dcin), _f32c(dmn), _f32c(dtin)
        dz = torch.empty((B, MAXDET, NF), device=cin.device, dtype=torch.float32)
        dpred = torch.empty_like(dz)
        _call("cvad_scorer_inputs_bwd_f32", _ptr(cin), _ptr(ntr), B, _ptr(dcin), _ptr(dmn), _ptr(dtin), _ptr(dz), _ptr(dpred), _st())
        return dz, dpred, None


def scorer_inputs(z, pred, ntr):
    return _ScorerInputs.apply(z, pred, ntr)


class _LinComb3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, y, b, z, c):
        _cuda(x, y, z)
        x, y, z = _f32c(x), _f32c(y), _f32c(z)
        out = torch.empty_like(x)
        _call("cvad_lincomb3_f32", _ptr(out), _ptr(x), float(a), _ptr(y), float(b), _ptr(z), float(c), x.numel(), _st())
        ctx.coef = (a, b, c)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        outs = []
        for coef in ctx.coef:
            d = torch.empty_like(g)
            _call("cvad_lincomb3_f32", _ptr(d), _ptr(g), float(coef), _ptr(g), 0.0, _ptr(g), 0.0, g.numel(), _st())
            outs.append(d)
        return outs[0], None, outs[1], None, outs[2], None


def lincomb3(x, a, y, b, z, c):
    return _LinComb3.apply(x, a, y, b, z, c)


class _SoftmaxRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _cuda(x)
        x = _f32c(x)
        y = torch.empty_like(x)
        _call("cvad_softmax_rows_f32", _ptr(x), x.numel() // x.shape[-1], x.shape[-1], _ptr(y), _st())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(y)
        _call("cvad_softmax_rows_bwd_f32", _ptr(y), _ptr(dy), y.numel() // y.shape[-1], y.shape[-1], _ptr(dx), _st())
        return dx


def softmax_rows(x):
    return _SoftmaxRows.apply(x)
