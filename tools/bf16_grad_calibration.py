"""How far are per-tensor gradient norms of STOCK torch bf16 autocast from fp32 for the M-A training step?  (calibrates the bound of
tests/test_models_gpu.py::_bf16_grad_check: the unmodified reference model, CPU, torch.autocast(bfloat16) vs fp32, same injected noise).

    python tools/bf16_grad_calibration.py [case index in tests/golden/ma.pt, default 3 = live_train]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from oracle.ref_harness import NoiseInjector, import_ref  # noqa: E402
from test_oracle_golden import ma_noise, ma_synth_state  # noqa: E402

torch.set_num_threads(8)
cad = import_ref("causal_anomaly_detection")
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
c = torch.load(os.path.join(ROOT, "tests/golden/ma.pt"), weights_only=False)["cases"][idx]
B, T = c["B"], c["T"]
x = synth.ma_clips(B, T, c["H"], c["W"], c["xseed"], c["wide"])
eps, keep = ma_noise(c)
res = {}
for mode in ("fp32", "bf16"):
    torch.manual_seed(0)
    model = cad.CausalAnomalyDetector()
    model.load_state_dict(ma_synth_state(c["seed"], c["live"]), strict=True)
    for n_, p in model.named_parameters():
        if "backbone.conv1" in n_ or "backbone.bn1" in n_:
            p.requires_grad = False
    model.train()
    with NoiseInjector() as inj:
        inj.randn = [eps[b, : int(c["n_tracks"][b])].clone() for b in range(B)]
        inj.dropout[id(model.detector.detector_net[2])] = [keep["det0"]]
        inj.dropout[id(model.detector.detector_net[5])] = [keep["det1"]]
        inj.dropout[id(model.anomaly_scorer.causal_scorer[2])] = [keep["scorer0"][b:b + 1] for b in range(B)]
        inj.dropout[id(model.direct_classifier[2])] = [keep["cls0"]]
        inj.dropout[id(model.direct_classifier[5])] = [keep["cls1"]]
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            out = model(x)
            y = c["labels"]
            ce = torch.nn.CrossEntropyLoss()(out["direct_predictions"].float(), y)
            an = torch.nn.MSELoss()(out["anomaly_scores"].float(), y.float())
            kl = sum(k for k in out["kl_losses"]) / len(out["kl_losses"])
            cs = torch.nn.MSELoss()(out["causal_anomaly_scores"].float(), y.float())
            total = 0.4 * ce + 0.3 * an + 0.2 * cs + 0.1 * kl
        total.backward()
    res[mode] = ({k: float(p.grad.double().norm()) for k, p in model.named_parameters() if p.grad is not None}, float(total))
g32, g16 = res["fp32"][0], res["bf16"][0]
print(f"case {c['name']}: loss fp32 {res['fp32'][1]:.6f} bf16-autocast {res['bf16'][1]:.6f}")
gmax = max(g32.values())
rows = sorted(((abs(g16[k] - g32[k]) / g32[k], k, g32[k]) for k in g32 if g32[k] > 1e-3 * gmax and not synth.is_bn_fed_conv_bias(k)), reverse=True)
for e, k, n in rows[:12]:
    print(f"   {k:42s} |g| {n:.3e}  stock bf16 autocast rel err {e:.3f}")
