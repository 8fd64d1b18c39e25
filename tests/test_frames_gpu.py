"""Device-side frame path against the host code of the reference's loaders: cv2.resize (bit-exact), the Avenue clip
layout (BGR->RGB, /255, (3,T,H,W)), sliding windows, and M-A's uint8 clips."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("sh,sw,dh,dw,C", [(360, 640, 240, 360, 1), (360, 640, 64, 64, 3), (240, 360, 240, 360, 1), (158, 238, 240, 360, 1),
                                             (480, 856, 64, 64, 3), (100, 77, 240, 360, 3), (33, 50, 64, 64, 1), (64, 64, 120, 200, 3)])
def test_resize_bit_exact_vs_cv2(dev, sh, sw, dh, dw, C):
    from cvad_b200.frames import DeviceFrames
    rng = np.random.default_rng(sh * 7 + dw)
    vid = rng.integers(0, 256, (5, sh, sw, C) if C > 1 else (5, sh, sw), dtype=np.uint8)
    got = DeviceFrames(vid, dev).resized((dw, dh)).data.cpu().numpy()
    want = np.stack([cv2.resize(f, (dw, dh)) for f in vid])
    if C == 1:
        want = want[..., None]
    assert got.shape == want.shape and np.array_equal(got, want)


def test_avenue_clip_layout_and_sliding_windows(dev):
    """What create_avenue_dataloaders' dataset does per clip on the host (resize, BGR->RGB, /255, permute), for all windows of a video."""
    from cvad_b200.frames import DeviceFrames, window_starts
    rng = np.random.default_rng(3)
    vid = rng.integers(0, 256, (41, 90, 160, 3), dtype=np.uint8)                 # BGR frames as cv2.imread returns them
    starts = window_starts(len(vid), 8, 4)
    assert starts == list(range(0, 41 - 8, 4))
    got = DeviceFrames(vid, dev).resized((64, 64)).clips_f32(starts, 8).cpu()
    want = []
    for s in starts:
        clip = np.stack([cv2.cvtColor(cv2.resize(f, (64, 64)), cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0 for f in vid[s:s + 8]])
        want.append(torch.from_numpy(clip).permute(3, 0, 1, 2))
    want = torch.stack(want)
    assert got.shape == want.shape == (len(starts), 3, 8, 64, 64)
    assert float((got - want).abs().max()) <= 1e-7                                # x * (1/255) vs x / 255: one rounding


def test_ma_uint8_clips_feed_the_model(dev):
    """Grayscale video -> resize (360,240) -> uint8 clips (B,T,1,H,W) -> M-A: the same scores as the reference's host path
    (cv2.resize + FloatTensor + Normalize(0.5, 0.5), cad:89-104, 1177-1179)."""
    import synth
    from test_oracle_golden import ma_synth_state
    from cvad_b200.frames import DeviceFrames
    from cvad_b200.ma import CausalAnomalyDetector
    from cvad_b200.noise import FixedNoise
    rng = np.random.default_rng(5)
    vid = rng.integers(0, 256, (12, 158, 238), dtype=np.uint8)                    # UCSD ped1-sized frames
    starts, T = [0, 4], 8
    clips_u8 = DeviceFrames(vid, dev).resized((360, 240)).clips_u8(starts, T)
    host = np.stack([np.stack([cv2.resize(f, (360, 240)) for f in vid[s:s + T]]) for s in starts])
    assert np.array_equal(clips_u8.cpu().numpy()[:, :, 0], host)
    xf = (torch.from_numpy(host).float().unsqueeze(2) - 0.5) / 0.5
    m = CausalAnomalyDetector()
    m.load_state_dict(ma_synth_state(4, True), strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    eps = torch.randn(2, 5, 6, generator=synth.gen(6))
    outs = []
    for x in (xf.to(dev), clips_u8):
        m.noise = FixedNoise({"eps": eps})
        with torch.no_grad():
            outs.append(m(x)["anomaly_scores"])
    assert torch.equal(outs[0], outs[1])
