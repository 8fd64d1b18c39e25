"""Device-side evaluation tail against the numpy / sklearn calls the reference makes on the host:
np.percentile s1:60, roc_auc_score mc3:388, the eight metrics of s2:286-295, np.convolve cad:1085-1087."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 5, 100, 2047, 2048, 2049, 5000, 70000]


def _scores(n, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, generator=g) * scale).float()


@pytest.mark.parametrize("n", SIZES)
def test_percentile_and_labels_bit_equal_numpy(dev, n):
    """Thresholds are compared by the reference with `>` (s1:61): a threshold one ulp off flips labels, so this is bit-exact."""
    from cvad_b200 import evaltail
    for seed, scale in ((1, 1.0), (2, 1e-3), (3, 40.0)):
        x = _scores(n, seed * 100 + n, scale)
        if n > 10:
            x[n // 3] = x[n // 2]                       # ties
        xd = x.to(dev)
        for q in (95, 50, 0, 100, 12.5, 99.9):
            want = np.percentile(x.numpy(), q)
            got = evaltail.percentile(xd, q).cpu().numpy()[0]
            assert got == want, (n, q, got, want)
        thr, labels = evaltail.percentile_labels(xd, 95)
        t = np.percentile(x.numpy(), 95)
        assert np.array_equal(labels.cpu().numpy(), (x.numpy() > t).astype(np.float32))


@pytest.mark.parametrize("n", SIZES)
def test_sort_matches_numpy_incl_nan_and_negative(dev, n):
    from cvad_b200 import evaltail
    x = _scores(n, 7 + n, 2.0) - 1.0
    if n > 4:
        x[1] = float("nan")
        x[2] = -0.0
        x[3] = 0.0
        x[4] = float("inf")
    srt, order = evaltail.sort_scores(x.to(dev))
    want = np.sort(x.numpy())
    assert np.array_equal(srt.cpu().numpy(), want, equal_nan=True)
    o = order.cpu().numpy()
    assert sorted(o.tolist()) == list(range(n))
    assert np.array_equal(x.numpy()[o], want, equal_nan=True)


@pytest.mark.parametrize("n", [2, 37, 1000, 4096, 30000])
def test_roc_auc_matches_sklearn(dev, n):
    from sklearn.metrics import roc_auc_score
    from cvad_b200 import evaltail
    from cvad_b200.mc import roc_auc as host_auc
    g = torch.Generator().manual_seed(n)
    for quant in (0, 16):                                # continuous scores, and heavily tied ones
        s = torch.rand(n, generator=g)
        if quant:
            s = (s * quant).round() / quant
        y = (torch.rand(n, generator=g) < 0.3).float()
        y[0], y[1] = 0.0, 1.0
        got = float(evaltail.roc_auc(s.to(dev), y.to(dev)))
        want = roc_auc_score(y.numpy(), s.numpy())
        assert abs(got - want) < 1e-12, (n, quant, got, want)
        assert abs(got - host_auc(y.numpy(), s.numpy())) < 1e-12
    one_class = torch.zeros(n)
    assert float(evaltail.roc_auc(torch.rand(n).to(dev), one_class.to(dev))) == 0.0


@pytest.mark.parametrize("n", [1, 9, 300, 5000])
def test_mb_eval_metrics_match_numpy(dev, n):
    """s2:286-295 incl. len(np.unique(graphs.reshape(N, -1), axis=0)) with repeated graphs and a -0.0 / +0.0 pair."""
    from cvad_b200 import evaltail
    g = torch.Generator().manual_seed(100 + n)
    preds = torch.rand(n, generator=g) * 0.3 + 0.1
    base = torch.rand(max(n // 3, 1), 16, 16, generator=g) * (1 - torch.eye(16))
    graphs = base[torch.randint(0, base.shape[0], (n,), generator=g)].clone()
    if n > 2:
        graphs[1] = graphs[0]
        graphs[1, 0, 0] = -0.0                           # equal to graphs[0] for np.unique
        graphs[2, 3, 4] += 1e-7
    got = evaltail.mb_eval_metrics(preds.to(dev), graphs.to(dev)).cpu().numpy()
    p, cg = preds.numpy(), graphs.numpy()
    e = np.sum(cg > 0.1, axis=(1, 2))
    want = [float(np.mean(p)), float(np.std(p)), float(np.min(p)), float(np.max(p)), float(np.max(p) - np.min(p)), float(np.mean(e)),
            float(np.mean(e / 256)), len(np.unique(cg.reshape(len(cg), -1), axis=0))]
    assert got[7] == want[7] and got[2] == want[2] and got[3] == want[3] and got[4] == want[4]
    assert abs(got[5] - want[5]) < 1e-9 and abs(got[6] - want[6]) < 1e-12
    assert abs(got[0] - want[0]) < 2e-6 * abs(want[0]) and abs(got[1] - want[1]) < 1e-5 * max(abs(want[1]), 1e-3)   # numpy sums in float32


@pytest.mark.parametrize("n,w", [(50, 10), (10, 10), (7, 10), (3000, 25)])
def test_moving_average_matches_np_convolve(dev, n, w):
    from cvad_b200 import evaltail
    x = _scores(n, n + w)
    got = evaltail.moving_average(x.to(dev), w).cpu().numpy()
    if n < w:
        assert got.size == 0
        return
    want = np.convolve(x.numpy(), np.ones(w) / w, mode="valid")
    assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=0)


def test_evaluate_improved_metrics_on_device_match_host_formulas(dev, gold):
    """evaluate_improved end to end on the shipped checkpoint: device metrics == the reference's numpy formulas on the returned arrays,
    and the metrics-only call returns no arrays."""
    import synth
    from cvad_b200.mb import ImprovedMiniCausalVAD
    tr = ImprovedMiniCausalVAD(device=dev, verbose=False)
    tr.model.load_state_dict(gold("best_improved_model.pth")["model_state_dict"], strict=True)
    loader = [(synth.mb_clips_bright(6, 8, 64, 64, 900 + i), torch.zeros(6)) for i in range(4)]
    p, cg, m = tr.evaluate_improved(loader)
    e = np.sum(cg > 0.1, axis=(1, 2))
    assert m["unique_graphs"] == len(np.unique(cg.reshape(len(cg), -1), axis=0))
    assert abs(m["avg_edges"] - float(np.mean(e))) < 1e-9 and m["min_score"] == float(np.min(p)) and m["max_score"] == float(np.max(p))
    assert abs(m["mean_score"] - float(np.mean(p))) < 2e-6 and abs(m["std_score"] - float(np.std(p))) < 1e-6
    p2, cg2, m2 = tr.evaluate_improved(loader, return_arrays=False)
    assert p2 is None and cg2 is None and set(m2) == set(m)
    for k in m:      # two passes over the same clips agree to round-off (split-K atomics order the fp32 sums differently), not bitwise
        assert abs(m2[k] - m[k]) <= 1e-5 * max(abs(m[k]), 1.0), k
