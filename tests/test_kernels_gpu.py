"""Per-kernel parity on the B200: every C-ABI op (through the ctypes/autograd binding) against the plain PyTorch fp32
op it replaces, forward and backward, on seeded inputs incl. ragged / odd shapes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def setup_module(_m):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b, floor=1e-6):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), floor))


def _g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


CONV_CASES = [
    # N, Cin, D, H, W, Cout, k, stride, pad
    (2, 3, 8, 20, 24, 16, (3, 3, 3), (1, 2, 2), (1, 1, 1)),
    (3, 16, 8, 12, 12, 32, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    (2, 32, 5, 9, 7, 64, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    (2, 1, 6, 18, 18, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (1, 8, 4, 10, 10, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 1, 1, 37, 53, 32, (1, 7, 7), (1, 2, 2), (0, 3, 3)),
    (2, 32, 1, 15, 23, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (3, 64, 1, 9, 11, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (2, 1, 1, 16, 16, 32, (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    (2, 130, 1, 6, 5, 70, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("act", [0, 1])
def test_conv_fwd_bwd(dev, case, act):
    from cvad_b200 import ops
    N, Ci, D, H, W, Co, k, s, p = case
    x = torch.randn(N, Ci, D, H, W, generator=_g(1)).to(dev).requires_grad_(True)
    w = (torch.randn(Co, Ci, *k, generator=_g(2)) * 0.1).to(dev).requires_grad_(True)
    b = torch.randn(Co, generator=_g(3)).to(dev).requires_grad_(True)
    ref = F.conv3d(x, w, b, stride=s, padding=p)
    if act:
        ref = F.relu(ref)
    gy = torch.randn(ref.shape, generator=_g(4)).to(dev)
    gx_r, gw_r, gb_r = torch.autograd.grad(ref, (x, w, b), gy)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y = ops.conv_act(x2, w2, b2, s, p, act)
    assert y.shape == ref.shape
    assert rel(y, ref) < 2e-5
    y.backward(gy)
    assert rel(x2.grad, gx_r) < 5e-5
    assert rel(w2.grad, gw_r) < 5e-5
    assert rel(b2.grad, gb_r) < 5e-5


def test_conv2d_4d_input_and_frozen_weight(dev):
    from cvad_b200 import ops
    x = torch.randn(4, 1, 30, 44, generator=_g(5)).to(dev)
    w = torch.randn(32, 1, 7, 7, generator=_g(6)).to(dev)
    b = torch.randn(32, generator=_g(7)).to(dev)
    y = ops.conv_act(x, w, b, 2, 3, 0)
    assert rel(y, F.conv2d(x, w, b, stride=2, padding=3)) < 2e-5


LIN_CASES = [(4, 4096, 16), (32, 4096, 16), (7, 16, 32), (5, 32, 256), (33, 256, 128), (3, 80, 32), (9, 32, 1), (130, 300, 70),
             (512, 1024, 512)]


@pytest.mark.parametrize("M,K,O", LIN_CASES)
@pytest.mark.parametrize("act", [0, 1, 3])
@pytest.mark.parametrize("masked", [False, True])
def test_linear_fwd_bwd(dev, M, K, O, act, masked):
    from cvad_b200 import ops
    x = torch.randn(M, K, generator=_g(1)).to(dev).requires_grad_(True)
    w = (torch.randn(O, K, generator=_g(2)) / K ** 0.5).to(dev).requires_grad_(True)
    b = torch.randn(O, generator=_g(3)).to(dev).requires_grad_(True)
    mask = (torch.rand(M, O, generator=_g(4)) > 0.3).float().to(dev) if masked else None
    ref = F.linear(x, w, b)
    ref = F.relu(ref) if act == 1 else (torch.sigmoid(ref) if act == 3 else ref)
    if masked:
        ref = ref * mask / 0.7
    gy = torch.randn(M, O, generator=_g(5)).to(dev)
    gr = torch.autograd.grad(ref, (x, w, b), gy)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y = ops.linear_act(x2, w2, b2, act, mask, 0.3 if masked else 0.0)
    assert rel(y, ref) < 2e-5
    y.backward(gy)
    for got, want in zip((x2.grad, w2.grad, b2.grad), gr):
        assert rel(got, want) < 5e-5


@pytest.mark.parametrize("shape,nd", [((3, 8, 6, 10, 12), 3), ((5, 16, 1, 9, 7), 3), ((4, 32, 13, 17), 2)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_batchnorm(dev, shape, nd, training, act):
    from cvad_b200 import ops
    C = shape[1]
    x = (torch.randn(shape, generator=_g(1)) * 3 + 40).to(dev).requires_grad_(True)
    g = (torch.rand(C, generator=_g(2)) + 0.5).to(dev).requires_grad_(True)
    b = torch.randn(C, generator=_g(3)).to(dev).requires_grad_(True)
    rm = torch.randn(C, generator=_g(4)).to(dev) + 40
    rv = (torch.rand(C, generator=_g(5)) + 5).to(dev)
    rm2, rv2 = rm.clone(), rv.clone()
    nbt = torch.tensor(3, device=dev)
    ref = F.batch_norm(x, rm, rv, g, b, training, 0.1, 1e-5)
    ref = F.relu(ref) if act == 1 else (F.leaky_relu(ref, 0.1) if act == 2 else ref)
    gy = torch.randn(shape, generator=_g(6)).to(dev)
    gr = torch.autograd.grad(ref, (x, g, b), gy)
    x2, g2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, g, b))
    y = ops.batchnorm_act(x2, g2, b2, rm2, rv2, nbt, ops.bn_workspace(dev, C), training, act)
    assert rel(y, ref) < 3e-5
    y.backward(gy)
    assert rel(x2.grad, gr[0]) < 2e-4
    assert rel(g2.grad, gr[1]) < 2e-4 and rel(b2.grad, gr[2]) < 2e-4
    if training:
        assert rel(rm2, rm) < 1e-5 and rel(rv2, rv) < 1e-5 and int(nbt) == 4


@pytest.mark.parametrize("shape,k,s,p", [((2, 8, 6, 10, 12), (1, 2, 2), (1, 2, 2), 0), ((2, 16, 7, 9, 11), (2, 2, 2), (2, 2, 2), 0),
                                          ((3, 32, 21, 33), 3, 2, 1)])
def test_maxpool(dev, shape, k, s, p):
    from cvad_b200 import ops
    x = torch.randn(shape, generator=_g(1)).to(dev).requires_grad_(True)
    fn = F.max_pool3d if len(shape) == 5 else F.max_pool2d
    ref = fn(x, k, s, p)
    gy = torch.randn(ref.shape, generator=_g(2)).to(dev)
    (gr,) = torch.autograd.grad(ref, x, gy)
    x2 = x.detach().clone().requires_grad_(True)
    y = ops.maxpool(x2, k, s, p)
    assert torch.equal(y, ref)
    y.backward(gy)
    assert rel(x2.grad, gr) < 1e-6


@pytest.mark.parametrize("shape,out", [((2, 64, 2, 8, 8), (4, 4, 4)), ((2, 64, 3, 6, 6), (4, 4, 4)), ((3, 64, 16, 5, 7), (4, 4, 4)),
                                        ((2, 32, 4, 8, 8), (1, 1, 1)), ((3, 256, 8, 12), (4, 6)), ((2, 256, 4, 6), (4, 6)),
                                        ((2, 16, 7, 11), (4, 6))])
def test_adaptive_avgpool(dev, shape, out):
    from cvad_b200 import ops
    x = torch.randn(shape, generator=_g(1)).to(dev).requires_grad_(True)
    fn = F.adaptive_avg_pool3d if len(shape) == 5 else F.adaptive_avg_pool2d
    ref = fn(x, out)
    gy = torch.randn(ref.shape, generator=_g(2)).to(dev)
    (gr,) = torch.autograd.grad(ref, x, gy)
    x2 = x.detach().clone().requires_grad_(True)
    y = ops.adaptive_avgpool(x2, out)
    assert rel(y, ref) < 1e-5
    y.backward(gy)
    assert rel(x2.grad, gr) < 1e-5


def test_mean_mid_and_mask_scale_and_lincomb(dev):
    from cvad_b200 import ops
    x = torch.randn(3, 5, 40, generator=_g(1)).to(dev).requires_grad_(True)
    y = ops.mean_mid(x)
    assert rel(y, x.mean(1)) < 1e-6
    gy = torch.randn(3, 40, generator=_g(2)).to(dev)
    y.backward(gy)
    assert rel(x.grad, (gy / 5).unsqueeze(1).expand(3, 5, 40)) < 1e-6
    m = (torch.rand(6, 32, generator=_g(3)) > 0.5).float().to(dev)
    z = torch.randn(6, 32, generator=_g(4)).to(dev).requires_grad_(True)
    o = ops.mask_scale(z, m, 0.5)
    assert rel(o, z * m * 2) < 1e-6
    o.backward(torch.ones_like(o))
    assert rel(z.grad, m * 2) < 1e-6
    a = torch.randn(7, generator=_g(5)).to(dev).requires_grad_(True)
    pr = torch.rand(7, 2, generator=_g(6)).to(dev).requires_grad_(True)
    f = ops.lincomb2(a, 0.6, pr, 1, 0.4)
    assert rel(f, 0.6 * a + 0.4 * pr[:, 1]) < 1e-6
    f.backward(torch.ones(7, device=dev))
    assert rel(a.grad, torch.full((7,), 0.6, device=dev)) < 1e-6
    assert rel(pr.grad[:, 1], torch.full((7,), 0.4, device=dev)) < 1e-6 and float(pr.grad[:, 0].abs().max()) == 0.0


def test_mb_loss_kernel_vs_golden(dev, gold):
    from cvad_b200 import ops
    c = gold("mb.pt")["loss32"]
    sc = c["scores"].to(dev).requires_grad_(True)
    ad = c["adj"].to(dev).requires_grad_(True)
    loss, comp = ops.mb_loss(sc, ad, c["pseudo"].to(dev))
    loss.backward()
    assert abs(float(loss) - float(c["loss"])) < 1e-6 * max(1, abs(float(c["loss"])))
    keys = ("anomaly_loss", "acyclicity_loss", "sparsity_loss", "consistency_loss", "structure_loss", "edge_count", "sparsity_ratio")
    for i, k in enumerate(keys):
        assert abs(float(comp[i + 1]) - c["comps"][k]) <= 2e-5 * max(1.0, abs(c["comps"][k])), k
    assert rel(sc.grad.cpu(), c["dscores"]) < 1e-4
    assert rel(ad.grad.cpu(), c["dadj"]) < 1e-4


@pytest.mark.parametrize("B,nanom", [(1, 0), (2, 0), (5, 5), (8, 1), (64, 3), (300, 20)])
def test_mb_loss_kernel_vs_oracle(dev, B, nanom):
    from cvad_b200 import ops
    from oracle import mb as o_mb
    g = _g(B)
    sc = (torch.rand(B, 1, generator=g) * 0.9 + 0.05)
    ad = torch.rand(B, 16, 16, generator=g) * (1 - torch.eye(16))
    ps = torch.zeros(B)
    ps[:nanom] = 1
    s1, a1 = sc.clone().requires_grad_(True), ad.clone().requires_grad_(True)
    lo, co = o_mb.mb_loss(s1, a1, ps)
    lo.backward()
    s2, a2 = sc.to(dev).requires_grad_(True), ad.to(dev).requires_grad_(True)
    loss, comp = ops.mb_loss(s2, a2, ps.to(dev))
    loss.backward()
    assert abs(float(loss) - float(lo)) < 2e-5 * max(1, abs(float(lo)))
    assert rel(s2.grad.cpu(), s1.grad, 1e-9) < 1e-4
    assert rel(a2.grad.cpu(), a1.grad, 1e-9) < 1e-4


def test_bce_and_ma_loss(dev):
    from cvad_b200 import ops
    g = _g(3)
    s = torch.rand(9, generator=g).requires_grad_(True)
    y = (torch.rand(9, generator=g) > 0.5).float()
    ref = F.binary_cross_entropy(s, y)
    ref.backward()
    s2 = s.detach().to(dev).requires_grad_(True)
    l2 = ops.bce_loss(s2, y.to(dev))
    l2.backward()
    assert abs(float(l2) - float(ref)) < 1e-6 and rel(s2.grad.cpu(), s.grad) < 1e-5
    B = 6
    pr = torch.softmax(torch.randn(B, 2, generator=g), -1).requires_grad_(True)
    fi, ca, kl = (torch.rand(B, generator=g).requires_grad_(True) for _ in range(3))
    lab = (torch.rand(B, generator=g) > 0.5).long()
    ref = 0.4 * F.cross_entropy(pr, lab) + 0.3 * F.mse_loss(fi, lab.float()) + 0.2 * F.mse_loss(ca, lab.float()) + 0.1 * kl.sum() / B
    ref.backward()
    t = [v.detach().to(dev).requires_grad_(True) for v in (pr, fi, ca, kl)]
    loss, comp = ops.ma_loss(*t, lab.to(dev))
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    for a, b in zip(t, (pr, fi, ca, kl)):
        assert rel(a.grad.cpu(), b.grad) < 1e-5


@pytest.mark.parametrize("decoupled,clip_mode", [(True, 1), (False, 2), (True, 0)])
def test_fused_adam_matches_torch(dev, decoupled, clip_mode):
    from cvad_b200.arena import FusedAdam
    g = _g(11)
    shapes = [(16, 3, 3, 3, 3), (16,), (300, 70), (1,), (1500,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    ref_p = [torch.nn.Parameter(p.clone().to(dev)) for p in ps]
    my_p = [torch.nn.Parameter(p.clone().to(dev)) for p in ps]
    if decoupled:
        ropt = torch.optim.AdamW(ref_p, lr=5e-3, weight_decay=1e-2)
    else:
        ropt = torch.optim.Adam(ref_p, lr=5e-3, weight_decay=1e-2)
    mopt = FusedAdam(my_p, lr=5e-3, weight_decay=1e-2, decoupled=decoupled, clip_mode=clip_mode, max_norm=0.5 if clip_mode == 1 else 1.0,
                     clip_threshold=10.0, nan_mode=1)
    for it in range(4):
        grads = [torch.randn(s, generator=g) * (30.0 if it == 2 else 1.0) for s in shapes]
        mopt.zero_grad()
        for p, q, gr in zip(ref_p, my_p, grads):
            p.grad = gr.clone().to(dev)
            q.grad.copy_(gr.to(dev))
        if clip_mode == 1:
            torch.nn.utils.clip_grad_norm_(ref_p, 0.5)
        elif clip_mode == 2:
            n = sum(float(p.grad.norm()) ** 2 for p in ref_p) ** 0.5
            if n > 10.0:
                torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        ropt.step()
        mopt.step()
    for p, q in zip(ref_p, my_p):
        assert rel(q.data, p.data) < 2e-5
    sd = mopt.state_dict()
    assert float(sd["state"][0]["step"]) == 4.0
    # a NaN gradient skips the whole step
    before = [q.detach().clone() for q in my_p]
    mopt.zero_grad()
    my_p[2].grad[0, 0] = float("nan")
    mopt.step()
    for a, b in zip(before, my_p):
        assert torch.equal(a, b.data)
    assert mopt.skipped_steps() == 1


# dims, activations, which layers carry a dropout keep-mask -- the M-A stacks (cad:167-179, 240-246, 318-326, 361-367, 407-413, 435-461)
CHAIN_CASES = [
    ((256, 128, 64, 20), (1, 1, 0), (0,)),                    # detector behind its 6144 -> 512 -> 256 GEMMs (dropout mask on a chained layer)
    ((4, 32, 64, 64), (1, 1, 0), ()),                         # re-id
    ((68, 192), (0,), ()),                                    # GRU input projection
    ((64, 32, 1), (1, 3), ()),                                # edge predictor (sigmoid head)
    ((18, 64, 32, 1), (1, 1, 3), (0,)),                       # causal scorer
    ((6, 32, 32, 6), (1, 1, 0), ()),                          # dynamics
    ((256, 128, 64, 2), (1, 1, 0), (0,)),                     # direct classifier behind its first two GEMMs
    ((100, 70, 33, 130, 5, 77, 9, 3, 40), (1, 4, 3, 0, 1, 2, 1, 0), (1, 4)),      # eight ragged layers, every activation
]


@pytest.mark.parametrize("dims,acts,masked", CHAIN_CASES)
@pytest.mark.parametrize("rows", [1, 7, 8, 33, 160, 2560])
def test_mlp_chain_fwd_bwd(dev, dims, acts, masked, rows):
    """ops.mlp_chain (one launch forward, one for the data-gradient chain) against the same stack in plain PyTorch fp32: output,
    input gradient and every weight / bias gradient."""
    from cvad_b200 import ops
    actf = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.1), 3: torch.sigmoid, 4: torch.tanh}
    n = len(dims) - 1
    x = torch.randn(rows, dims[0], generator=_g(1)).to(dev).requires_grad_(True)
    Ws = [(torch.randn(dims[l + 1], dims[l], generator=_g(10 + l)) / dims[l] ** 0.5).to(dev).requires_grad_(True) for l in range(n)]
    bs = [torch.randn(dims[l + 1], generator=_g(30 + l)).mul(0.1).to(dev).requires_grad_(True) for l in range(n)]
    keeps = [(torch.rand(rows, dims[l + 1], generator=_g(50 + l)) >= 0.3).float().to(dev) if l in masked else None for l in range(n)]
    h = x
    for l in range(n):
        h = actf[acts[l]](F.linear(h, Ws[l], bs[l]))
        if keeps[l] is not None:
            h = h * keeps[l] / 0.7
    dy = torch.randn(rows, dims[-1], generator=_g(3)).to(dev)
    ref = torch.autograd.grad(h, [x] + Ws + bs, dy)
    x2 = x.detach().clone().requires_grad_(True)
    W2 = [w.detach().clone().requires_grad_(True) for w in Ws]
    b2 = [b.detach().clone().requires_grad_(True) for b in bs]
    y = ops.mlp_chain(x2, [(W2[l], b2[l], acts[l], keeps[l], 0.3 if keeps[l] is not None else 0.0) for l in range(n)])
    assert rel(y, h) < 2e-5
    with ops.param_grad_overlap():
        y.backward(dy)
    torch.cuda.synchronize()
    got = [x2.grad] + [w.grad for w in W2] + [b.grad for b in b2]
    for i, (a, r) in enumerate(zip(got, ref)):
        assert rel(a, r) < 5e-5, (i, rel(a, r))
    # inference: nothing is stored, same numbers
    with torch.no_grad():
        y2 = ops.mlp_chain(x.detach(), [(w.detach(), b.detach(), acts[l], keeps[l], 0.3 if keeps[l] is not None else 0.0) for l, (w, b) in enumerate(zip(Ws, bs))])
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("M,N,K", [(512, 512, 6144), (32, 512, 6144), (130, 200, 2048), (1, 128, 4096), (300, 512, 96)])
def test_linear_fwd_tf32x3_matches_fp32(dev, M, N, K):
    """tcgen05 kind::tf32 GEMM with the 3xTF32 operand split: fp32-level accuracy (a plain tf32 product would be ~1e-3 off), ragged M / N
    through TMA zero fill, split-K reduction.  Reference: fp64 matmul of the same fp32 operands."""
    from cvad_b200.ops import _call, _ptr, _st
    x = torch.randn(M, K, generator=_g(1)).to(dev)
    w = (torch.randn(N, K, generator=_g(2)) / K ** 0.5).to(dev)
    y = torch.zeros(M, N, device=dev)
    _call("cvad_linear_fwd_tf32x3", _ptr(x), _ptr(w), _ptr(y), M, N, K, _st())
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    e = rel(y, ref)
    e32 = rel(x @ w.t(), ref)
    print(f"[tf32x3] M{M} N{N} K{K}: rel err {e:.2e} (torch fp32 matmul: {e32:.2e})")
    assert e < 5e-6


@pytest.mark.parametrize("n", [1, 3, 4, 5, 17, 1023, 4096, 1000003])
@pytest.mark.parametrize("offset", [0, 1, 2, 3])
def test_fill_f32_any_alignment(dev, n, offset):
    """cvad_fill_f32 (zero-initialises the accumulating GEMMs' outputs and the gradient arena every step): 16-byte stores over the aligned
    body, scalars around it -- every element written, nothing outside the range touched."""
    from cvad_b200.ops import _call, _ptr, _st
    buf = torch.full((n + 8,), 7.0, device=dev)
    view = buf[offset + 1: offset + 1 + n]
    _call("cvad_fill_f32", view.data_ptr(), n, -2.5, _st())
    torch.cuda.synchronize()
    assert bool((view == -2.5).all())
    assert float(buf[: offset + 1].min()) == 7.0 and float(buf[offset + 1 + n:].min()) == 7.0
