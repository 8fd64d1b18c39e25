"""bf16 tensor-core execution of the M-A backbone (cad:141-158): one autograd node for the whole conv/BN/ReLU stack.

Data flow per step (activations NHWC bf16 in HBM, statistics / gradients of parameters fp32):

  x fp32 (B*T,1,240,360) --conv 7x7 s2 (fp32 FFMA, frozen stem)--> y1 fp32 NCHW
      --bn1 batch stats--> fused BN+ReLU+MaxPool(3,2,1) --> a0 bf16 NHWC (.,60,90,32)
  for the 8 layers: raw_i = conv3x3_tcgen05(a_{i-1}) ; (mean, invstd) = stats(raw_i) ; a_i = relu(bn(raw_i))
  features = AdaptiveAvgPool(4,6)(a_8) -> fp32 (B*T, 6144) in the reference's (c,h,w) order

Backward walks the same list in reverse: fused ReLU+BN backward (bf16), tcgen05 weight-gradient (fp32 atomics into the
gradient arena) and tcgen05 data-gradient.  Convolution biases that feed a BatchNorm have an analytically zero gradient
(the reference only accumulates round-off there); they receive exactly zero here.
"""
from __future__ import annotations

import ctypes

import torch

from . import ops
from .ops import _call, _ptr, _st, grad_buffer, _wants_grad

BF16 = torch.bfloat16


def _layers(bb):
    out = []
    for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
        for ci, bi in ((0, 1), (3, 4)):
            out.append((layer[ci], layer[bi]))
    return out


class _BackboneBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bb, *params):
        dev = x.device
        N, _, H, W = x.shape
        st = _st()
        training = bb.training
        # ---- stem (frozen in the reference's training recipe, cad:596-598): fp32 conv + bn1 statistics
        y1 = ops.conv_act(x, bb.conv1.weight.detach(), bb.conv1.bias.detach(), 2, 3, ops.ACT_NONE)
        C1 = y1.shape[1]
        H1, W1 = y1.shape[2], y1.shape[3]
        mean = torch.empty(C1, device=dev, dtype=torch.float32)
        invstd = torch.empty_like(mean)
        bn1 = bb.bn1
        if bn1.training:
            _call("cvad_bn_train_stats_f32", _ptr(y1), N, C1, H1 * W1, _ptr(ops.bn_workspace(dev, C1)), float(bn1.eps), float(bn1.momentum),
                  _ptr(mean), _ptr(invstd), _ptr(bn1.running_mean), _ptr(bn1.running_var), _ptr(bn1.num_batches_tracked), st)
        else:
            _call("cvad_bn_eval_prepare_f32", C1, float(bn1.eps), _ptr(bn1.running_mean), _ptr(bn1.running_var), _ptr(mean), _ptr(invstd), st)
        PH, PW = (H1 - 1) // 2 + 1, (W1 - 1) // 2 + 1
        a = torch.empty((N, PH, PW, C1), device=dev, dtype=BF16)
        _call("cvad_stem_bn_relu_maxpool_bf16", _ptr(y1), N, C1, H1, W1, _ptr(mean), _ptr(invstd), _ptr(bn1.weight), _ptr(bn1.bias), _ptr(a), st)
        del y1
        need_bwd = any(ctx.needs_input_grad)
        saved = []
        h, w, cin = PH, PW, C1
        for conv, bn in _layers(bb):
            cout, stride = conv.out_channels, conv.stride[0]
            wf = torch.empty((cout, 9 * cin), device=dev, dtype=BF16)
            wd = torch.empty((cin, 9 * cout), device=dev, dtype=BF16) if need_bwd else None
            _call("cvad_pack_w3x3_bf16", _ptr(conv.weight), cout, cin, _ptr(wf), _ptr(wd), st)
            ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
            raw = torch.empty((N, ho, wo, cout), device=dev, dtype=BF16)
            _call("cvad_conv3x3_fwd_bf16", _ptr(a), _ptr(wf), _ptr(conv.bias), _ptr(raw), N, h, w, cin, cout, stride, st)
            P = N * ho * wo
            mean = torch.empty(cout, device=dev, dtype=torch.float32)
            invstd = torch.empty_like(mean)
            if bn.training:
                _call("cvad_bn_stats_nhwc_bf16", _ptr(raw), P, cout, _ptr(ops.bn_workspace(dev, cout)), float(bn.eps), float(bn.momentum),
                      _ptr(mean), _ptr(invstd), _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(bn.num_batches_tracked), st)
            else:
                _call("cvad_bn_eval_prepare_f32", cout, float(bn.eps), _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(mean), _ptr(invstd), st)
            act = torch.empty_like(raw)
            _call("cvad_bn_apply_relu_nhwc_bf16", _ptr(raw), _ptr(act), P, cout, _ptr(mean), _ptr(invstd), _ptr(bn.weight), _ptr(bn.bias), st)
            if need_bwd:
                saved.append((a, raw, mean, invstd, wd, (h, w, cin, cout, stride, ho, wo), bn.training))
            a, h, w, cin = act, ho, wo, cout
        feats = torch.empty((N, cin, 4, 6), device=dev, dtype=torch.float32)
        _call("cvad_avgpool_nhwc_bf16_fwd", _ptr(a), N, h, w, cin, 4, 6, _ptr(feats), st)
        ctx.bb, ctx.saved, ctx.last = bb, saved, (N, h, w, cin)
        return feats.reshape(N, -1)

    @staticmethod
    def backward(ctx, dfeat):
        bb, saved = ctx.bb, ctx.saved
        N, h, w, c = ctx.last
        st = _st()
        dev = dfeat.device
        dfeat = ops._f32c(dfeat)
        dact = torch.empty((N, h, w, c), device=dev, dtype=BF16)
        _call("cvad_avgpool_nhwc_bf16_bwd", _ptr(dfeat), N, h, w, c, 4, 6, _ptr(dact), st)
        layers = _layers(bb)
        for idx in range(len(layers) - 1, -1, -1):
            conv, bn = layers[idx]
            a_in, raw, mean, invstd, wd, (hi, wi, cin, cout, stride, ho, wo), bn_training = saved[idx]
            P = N * ho * wo
            draw = torch.empty_like(raw)
            dg = grad_buffer(bn.weight) if _wants_grad(bn.weight) else None
            db = grad_buffer(bn.bias) if _wants_grad(bn.bias) else None
            _call("cvad_bn_relu_bwd_nhwc_bf16", _ptr(raw), _ptr(dact), _ptr(draw), P, cout, _ptr(mean), _ptr(invstd), _ptr(bn.weight),
                  _ptr(bn.bias), int(bn_training), _ptr(ops.bn_workspace(dev, cout)), _ptr(dg), _ptr(db), st)
            if _wants_grad(conv.weight):
                _call("cvad_conv3x3_wgrad_bf16", _ptr(a_in), _ptr(draw), _ptr(grad_buffer(conv.weight)), N, hi, wi, cin, cout, stride, st)
            if _wants_grad(conv.bias):
                grad_buffer(conv.bias)          # analytically zero (BatchNorm removes the mean); keep the tensor "with grad"
            if idx > 0:
                dact = torch.empty((N, hi, wi, cin), device=dev, dtype=BF16)
                _call("cvad_conv3x3_dgrad_bf16", _ptr(draw), _ptr(wd), _ptr(dact), N, hi, wi, cin, cout, stride, st)
            saved[idx] = None
        return (None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def backbone_forward_bf16(bb, x):
    """x (B*T, 1, H, W) fp32 -> features (B*T, 6144) fp32."""
    ops._cuda(x)
    params = [p for p in bb.parameters() if p.requires_grad]
    return _BackboneBF16.apply(x.float().contiguous(), bb, *params)
