import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.nn.functional as F
import cvad_b200
from cvad_b200 import tc
from cvad_b200.ops import _call, _ptr, _st
dev = torch.device("cuda:0")
def g(s): return torch.Generator().manual_seed(s)
for (N, H, W, Ci, Co, s) in [(1, 8, 12, 256, 256, 1), (4, 8, 12, 256, 256, 1)]:
    x = torch.randn(N, Ci, H, W, generator=g(1)).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(Co, Ci, 3, 3, generator=g(2)) / (3.0 * Ci ** 0.5)).to(dev)
    b = torch.randn(Co, generator=g(3)).to(dev)
    ref = F.conv2d(x, w.to(torch.bfloat16).float(), b, stride=s, padding=1)
    wf = torch.empty(9 * Co, Ci, device=dev, dtype=torch.bfloat16)
    wd = torch.empty(9 * Ci, Co, device=dev, dtype=torch.bfloat16)
    _call("cvad_flat_pack_w3x3_bf16", _ptr(w), Co, Ci, s, _ptr(wf), _ptr(wd), _st())
    xin = tc.to_padded(x)
    y = torch.full((N, H + 2, W + 2, Co), 7.0, device=dev, dtype=torch.bfloat16)
    _call("cvad_flat_conv3x3_fwd_bf16", _ptr(xin), _ptr(wf), _ptr(b), _ptr(y), N, H, W, Ci, Co, s, _st())
    torch.cuda.synchronize()
    out = tc.from_padded(y, H, W)          # N,C,H,W
    err = (out - ref).abs()
    print("case", N, "max err", float(err.max()), "ref max", float(ref.abs().max()))
    print(" err by 32-channel block:", [round(float(err[:, c:c + 32].max()), 3) for c in range(0, Co, 32)])
    print(" err by row h:", [round(float(err[:, :, h].max()), 3) for h in range(H)])
    print(" fraction exactly 7:", float((y.float() == 7.0).float().mean()), " by channel block:",
          [round(float((y[..., c:c + 32].float() == 7.0).float().mean()), 3) for c in range(0, Co, 32)])
    yf = y.float().reshape(-1, Co)
    rows7 = (yf == 7.0).all(dim=1).nonzero().flatten().tolist()
    print(" flat rows entirely 7:", rows7[:40], len(rows7))
