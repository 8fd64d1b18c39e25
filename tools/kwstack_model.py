#!/usr/bin/env python
"""CPU model of the kw-stacked flat convolution planned for the 32/64-channel layers.

Today a stride-1 3x3 convolution over the padded-flat layout is nine row-shifted GEMMs, out[q] = sum_{kh,kw} in[q + (kh-1)Wp + (kw-1)] W[kh][kw],
i.e. nine M128 x N(=Cout) MMAs per K16 chunk, each streaming the 4 KB activation tile again.  Stacked form: per kernel row kh ONE MMA of
width N = 3*Cout against the weights [ci][(kw, co)],

        D[r][(kw, co)] = sum_kh sum_ci in[r + (kh-1)Wp][ci] * W[kh][kw][ci][co]            (three MMAs per K16 chunk, not nine)

and the epilogue adds the three column groups with a one-row shift:  out[q][co] = D[q-1][(0,co)] + D[q][(1,co)] + D[q+1][(2,co)].
A 128-row accumulator tile therefore yields 126 output rows (rows 1..126; one halo row on either side).  This script checks that algebra,
including the 126-row tiling, against torch's conv2d.  Pure torch on the CPU; nothing here is on the product path.
"""
import torch
import torch.nn.functional as F


def kwstack_conv(x_pad_flat, w, Wp, tile_rows=128):
    """x_pad_flat (R, Ci): padded-flat pixels of the whole batch; w (Co, Ci, 3, 3).  Returns out (R, Co) in the same flat index space
    (border rows hold junk-free but meaningless values, as in the kernel)."""
    R, Ci = x_pad_flat.shape
    Co = w.shape[0]
    # weights as the B operand of kernel row kh: [(kw, co)][ci]
    wk = [torch.cat([w[:, :, kh, kw] for kw in range(3)], dim=0) for kh in range(3)]          # 3 x (3*Co, Ci)
    out = torch.zeros(R, Co)
    useful = tile_rows - 2
    pad = Wp + 1
    xz = torch.cat([torch.zeros(pad + 1, Ci), x_pad_flat, torch.zeros(pad + tile_rows, Ci)])   # TMA zero-fill outside the tensor
    for q0 in range(0, R, useful):
        r0 = q0 - 1                                                                            # accumulator row i <-> flat pixel r0 + i
        D = torch.zeros(tile_rows, 3 * Co)
        for kh in range(3):
            a = xz[r0 + (kh - 1) * Wp + pad + 1: r0 + (kh - 1) * Wp + pad + 1 + tile_rows]    # row-shifted view of the segment
            D += a @ wk[kh].t()
        o = D[0:useful, 0:Co] + D[1:useful + 1, Co:2 * Co] + D[2:useful + 2, 2 * Co:3 * Co]    # the one-row shifts of the epilogue
        n = min(useful, R - q0)
        out[q0:q0 + n] = o[:n]
    return out


def main():
    torch.manual_seed(0)
    for (N, H, W, Ci, Co) in [(2, 5, 7, 4, 6), (3, 12, 17, 8, 8), (1, 60, 90, 4, 4)]:
        x = torch.randn(N, Ci, H, W)
        w = torch.randn(Co, Ci, 3, 3)
        ref = F.conv2d(x, w, padding=1)
        Wp = W + 2
        xp = torch.zeros(N, H + 2, Wp, Ci)
        xp[:, 1:H + 1, 1:W + 1] = x.permute(0, 2, 3, 1)
        out = kwstack_conv(xp.reshape(-1, Ci), w, Wp).reshape(N, H + 2, Wp, Co)
        got = out[:, 1:H + 1, 1:W + 1].permute(0, 3, 1, 2)
        err = float((got - ref).abs().max() / ref.abs().max())
        print(f"N={N} {H}x{W} {Ci}->{Co}: max rel err {err:.2e}")
        assert err < 1e-5
    print("kw-stacked algebra matches conv2d")


if __name__ == "__main__":
    main()
