"""Oracle (TEST INFRASTRUCTURE) for M-D: conv autoencoder + LSTM + memory bank, causal_anomaly_detection1.py (cad1).

Functional fp32 restatement in plain torch CPU ops of ``VideoAutoEncoder`` cad1:124-321 (encoder cad1:129-153 applied per time
step cad1:227-231, LSTM cad1:182-188/238-239, decoder cad1:156-179 applied T times to the same feature cad1:254-257, memory
score cad1:262-301), ``reconstruction_loss`` cad1:323-344 and the combined score cad1:545-552.  ``P`` is the model's
state_dict (fp32 tensors); BatchNorm running statistics are updated in place on ``P`` in training mode, exactly as the
reference's T sequential calls do.  tools/make_golden.py asserts it equal to the reference itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS, MOM = 1e-5, 0.1


def _bn(P, pre, x, train):
    out = F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"], P[pre + ".weight"], P[pre + ".bias"], train, MOM, EPS)
    if train:
        P[pre + ".num_batches_tracked"] += 1
    return out


def encode_frame(P, x, train):
    """cad1:129-153 on one time step (B,C,H,W) -> (B,64)."""
    h = x
    for ci in (0, 3, 6, 9):
        h = F.conv2d(h, P[f"encoder.{ci}.weight"], P[f"encoder.{ci}.bias"], stride=2, padding=1)
        h = F.leaky_relu(_bn(P, f"encoder.{ci + 1}", h, train), 0.1)
    return torch.tanh(F.linear(h.flatten(1), P["encoder.13.weight"], P["encoder.13.bias"]))


def lstm_last(P, seq):
    """nn.LSTM(64,64,batch_first) last hidden state; gate order i,f,g,o."""
    Wi, Wh = P["temporal_encoder.weight_ih_l0"], P["temporal_encoder.weight_hh_l0"]
    bi, bh = P["temporal_encoder.bias_ih_l0"], P["temporal_encoder.bias_hh_l0"]
    B, T, _ = seq.shape
    h = torch.zeros(B, 64, dtype=seq.dtype)
    c = torch.zeros(B, 64, dtype=seq.dtype)
    for t in range(T):
        g = F.linear(seq[:, t], Wi, bi) + F.linear(h, Wh, bh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def decode(P, z, train):
    """cad1:156-179 -> (B,C,64,64)."""
    h = F.leaky_relu(F.linear(z, P["decoder.0.weight"], P["decoder.0.bias"]), 0.1).reshape(-1, 128, 4, 4)
    for ci in (3, 6, 9):
        h = F.conv_transpose2d(h, P[f"decoder.{ci}.weight"], P[f"decoder.{ci}.bias"], stride=2, padding=1)
        h = F.leaky_relu(_bn(P, f"decoder.{ci + 1}", h, train), 0.1)
    return torch.sigmoid(F.conv_transpose2d(h, P["decoder.12.weight"], P["decoder.12.bias"], stride=2, padding=1))


def memory_score(P, z):
    """cad1:262-301."""
    filled = int(P["memory_ptr"][0])
    if filled < 10:
        return torch.zeros(z.shape[0])
    mem = P["normal_memory"][:filled]
    zn = z / z.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    mn = mem / mem.norm(dim=-1, keepdim=True).clamp(min=1e-8)
    sim = (zn @ mn.t()).clamp(-1, 1)
    return (1 - sim).min(dim=1)[0].clamp(0, 2) / 2.0


def md_forward(P, frames, train):
    """cad1:303-321 -> (reconstructed (B,T,C,H,W), sequence_feature (B,64), frame_features (B,T,64), anomaly_score (B,))."""
    B, T = frames.shape[:2]
    ff = torch.stack([encode_frame(P, frames[:, t], train) for t in range(T)], dim=1)
    z = lstm_last(P, ff)
    rec = torch.stack([decode(P, z, train) for _ in range(T)], dim=1)
    return rec, z, ff, memory_score(P, z.detach())


def recon_loss(frames, rec):
    return ((rec - frames) ** 2).mean()


def combined_scores(frames, rec, mem_score):
    """cad1:545-552."""
    return 0.7 * ((rec - frames) ** 2).flatten(1).mean(dim=1) + 0.3 * mem_score


def update_memory(P, feats):
    """cad1:201-219 (in place on P)."""
    bs, ptr, size = feats.shape[0], int(P["memory_ptr"][0]), P["normal_memory"].shape[0]
    feats = feats.detach()
    if ptr + bs <= size:
        P["normal_memory"][ptr:ptr + bs] = feats
        P["memory_ptr"][0] = (ptr + bs) % size
    else:
        rem = size - ptr
        P["normal_memory"][ptr:] = feats[:rem]
        P["normal_memory"][:bs - rem] = feats[rem:]
        P["memory_ptr"][0] = bs - rem
