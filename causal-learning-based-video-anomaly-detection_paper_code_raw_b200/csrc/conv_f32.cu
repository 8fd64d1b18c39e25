// fp32 implicit-GEMM convolution (1-D/2-D/3-D via D=1), forward / data-gradient / weight-gradient.
//
// This is the exact-precision (fp32 FFMA) path: it serves the small 3-D CNNs of M-B / M-C
// (avenue_training_script2.py:19-21, minicausal_vad_complete3.py:38-53), the M-D encoder/decoder
// (causal_anomaly_detection1.py:129-179; ConvTranspose2d == data-gradient used as a forward op)
// and the "fp32 mode" of the M-A backbone (causal_anomaly_detection.py:115-153).  The bf16
// tensor-core path for the M-A backbone lives in conv_tc.cu.
//
// All three problems are one tiled GEMM  C[i][j] = sum_r A(i,r) * B(r,j)  whose operands are
// gathered on the fly (no im2col buffer in HBM):
//   FWD    i = output position (n,od,oh,ow)   r = (ci,kd,kh,kw)      j = co
//   DGRAD  i = input  position (n,id,ih,iw)   r = (co,kd,kh,kw)      j = ci
//   WGRAD  i = (ci,kd,kh,kw)                  r = output position    j = co   (split over r, atomics)
// Tensors are addressed through explicit element strides (n,c,d,h,w), so NCDHW-contiguous torch
// tensors and channels-last views both work; weights are read/written in torch's OIDHW layout.
#include "common.cuh"
#include "cvad_b200.h"

namespace {

constexpr int BM = 128;   // tile rows (i)
constexpr int BK = 16;    // reduction slab (r)
constexpr int TM = 4;     // rows per thread
constexpr int NTHREADS = 256;

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct Geo {
  int N, Cin, Din, Hin, Win, Cout, Dout, Hout, Wout;
  int kD, kH, kW, sD, sH, sW, pD, pH, pW;
  long long xs[5], ys[5];
  int taps, K;            // taps = kD*kH*kW, K = Cin*taps
  long long Pout, Pin;    // N*Dout*Hout*Wout, N*Din*Hin*Win
  unsigned m_taps, s_taps, m_khw, s_khw, m_kw, s_kw;      // multiply-shift forms of / taps, / (kH*kW), / kW (the gathers run them per element)
  unsigned m_wo, s_wo, m_ho, s_ho, m_do, s_do;            // / Wout, / Hout, / Dout (position decode; used when Pout < 2^31)
  int small;                                               // Pout and Pin below 2^31: 32-bit position arithmetic
};

// Strided data-gradient, one launch per PARITY CLASS of input positions (blockIdx.z = class).  With stride s an input coordinate i only
// meets the kernel taps k with (i + p - k) % s == 0, so gathering all taps (the transposed-convolution view) multiplies 1 - 1/s^dims of
// the operands by structural zeros: 7/8 of the work of a stride-2 3-D convolution.  A class fixes (i % s) per dimension; its rows are the
// positions of that class, its reduction runs over (co, the taps that do meet them), and everything it touches is dense.
struct Phase {
  int pd, ph, pw;         // parity of (id, ih, iw) modulo the stride
  int Dc, Hc, Wc;         // positions of this class per dimension
  int kd0, kh0, kw0;      // first matching tap per dimension
  int nkd, nkh, nkw;      // matching taps per dimension
  long long rows;         // N*Dc*Hc*Wc
  unsigned m_pt, s_pt, m_hw, s_hw, m_w, s_w;              // / (nkd*nkh*nkw), / (nkh*nkw), / nkw
};

__host__ __device__ inline void fdiv_init(unsigned d, unsigned& mul, unsigned& shr) {
  if (d <= 1) { mul = 0; shr = 0; return; }
  unsigned lg = 0;
  while ((1u << lg) < d) ++lg;
  const unsigned p = 31 + lg;
  mul = (unsigned)(((1ull << p) + d - 1) / d);
  shr = p - 32;
}

__host__ __device__ inline Phase make_phase(const Geo& g, int cls) {
  Phase f;
  f.pw = cls % g.sW; cls /= g.sW;
  f.ph = cls % g.sH; cls /= g.sH;
  f.pd = cls;
  f.Dc = f.pd < g.Din ? (g.Din - f.pd + g.sD - 1) / g.sD : 0;
  f.Hc = f.ph < g.Hin ? (g.Hin - f.ph + g.sH - 1) / g.sH : 0;
  f.Wc = f.pw < g.Win ? (g.Win - f.pw + g.sW - 1) / g.sW : 0;
  f.kd0 = (f.pd + g.pD) % g.sD; f.kh0 = (f.ph + g.pH) % g.sH; f.kw0 = (f.pw + g.pW) % g.sW;
  f.nkd = f.kd0 < g.kD ? (g.kD - f.kd0 + g.sD - 1) / g.sD : 0;
  f.nkh = f.kh0 < g.kH ? (g.kH - f.kh0 + g.sH - 1) / g.sH : 0;
  f.nkw = f.kw0 < g.kW ? (g.kW - f.kw0 + g.sW - 1) / g.sW : 0;
  f.rows = (long long)g.N * f.Dc * f.Hc * f.Wc;
  fdiv_init((unsigned)(f.nkd * f.nkh * f.nkw), f.m_pt, f.s_pt);
  fdiv_init((unsigned)(f.nkh * f.nkw), f.m_hw, f.s_hw);
  fdiv_init((unsigned)f.nkw, f.m_w, f.s_w);
  return f;
}

// One decoded position: base element offset and the top-left coordinate of its receptive field.
struct PosInfo {
  long long base;
  int d0, h0, w0;
  bool ok;
};

__device__ __forceinline__ PosInfo decode_out_pos(const Geo& g, long long m) {
  // output position m -> input-window origin (FWD / WGRAD gather of x)
  PosInfo q;
  q.ok = m < g.Pout;
  long long mm = q.ok ? m : 0;
  int ow, oh, od, n;
  if (g.small) {
    int t = (int)mm, u = fast_div(t, g.m_wo, g.s_wo);
    ow = t - u * g.Wout;
    t = fast_div(u, g.m_ho, g.s_ho);
    oh = u - t * g.Hout;
    n = fast_div(t, g.m_do, g.s_do);
    od = t - n * g.Dout;
  } else {
    ow = (int)(mm % g.Wout); mm /= g.Wout;
    oh = (int)(mm % g.Hout); mm /= g.Hout;
    od = (int)(mm % g.Dout); mm /= g.Dout;
    n = (int)mm;
  }
  q.d0 = od * g.sD - g.pD; q.h0 = oh * g.sH - g.pH; q.w0 = ow * g.sW - g.pW;
  q.base = n * g.xs[0] + (long long)q.d0 * g.xs[2] + (long long)q.h0 * g.xs[3] + (long long)q.w0 * g.xs[4];
  return q;
}

__device__ __forceinline__ long long out_offset(const Geo& g, long long m) {
  if (g.small) {
    int t = (int)m, u = fast_div(t, g.m_wo, g.s_wo);
    const int ow = t - u * g.Wout;
    t = fast_div(u, g.m_ho, g.s_ho);
    const int oh = u - t * g.Hout;
    const int n = fast_div(t, g.m_do, g.s_do);
    return n * g.ys[0] + (t - n * g.Dout) * g.ys[2] + oh * g.ys[3] + ow * g.ys[4];
  }
  int ow = (int)(m % g.Wout); m /= g.Wout;
  int oh = (int)(m % g.Hout); m /= g.Hout;
  int od = (int)(m % g.Dout); m /= g.Dout;
  return m * g.ys[0] + od * g.ys[2] + oh * g.ys[3] + ow * g.ys[4];
}

__device__ __forceinline__ PosInfo decode_in_pos(const Geo& g, const Phase& f, long long m) {
  // row m of parity class f -> input position (DGRAD rows); d0/h0/w0 hold id+pD etc.
  PosInfo q;
  q.ok = m < f.rows;
  long long mm = q.ok ? m : 0;
  int iw = (int)(mm % f.Wc) * g.sW + f.pw; mm /= f.Wc;
  int ih = (int)(mm % f.Hc) * g.sH + f.ph; mm /= f.Hc;
  int id = (int)(mm % f.Dc) * g.sD + f.pd; mm /= f.Dc;
  int n = (int)mm;
  q.d0 = id + g.pD; q.h0 = ih + g.pH; q.w0 = iw + g.pW;
  q.base = n * g.ys[0];
  return q;
}

__device__ __forceinline__ long long in_offset(const Geo& g, const Phase& f, long long m) {
  int iw = (int)(m % f.Wc) * g.sW + f.pw; m /= f.Wc;
  int ih = (int)(m % f.Hc) * g.sH + f.ph; m /= f.Hc;
  int id = (int)(m % f.Dc) * g.sD + f.pd; m /= f.Dc;
  return m * g.xs[0] + id * g.xs[2] + ih * g.xs[3] + iw * g.xs[4];
}

// x[pos-window + (ci,kd,kh,kw)] with zero padding
__device__ __forceinline__ float gather_x(const Geo& g, const float* __restrict__ x, const PosInfo& q, int k) {
  if (!q.ok || k >= g.K) return 0.f;
  int ci = fast_div(k, g.m_taps, g.s_taps), tap = k - ci * g.taps;
  int kd = fast_div(tap, g.m_khw, g.s_khw); tap -= kd * g.kH * g.kW;
  int kh = fast_div(tap, g.m_kw, g.s_kw), kw = tap - kh * g.kW;
  int id = q.d0 + kd, ih = q.h0 + kh, iw = q.w0 + kw;
  if ((unsigned)id >= (unsigned)g.Din || (unsigned)ih >= (unsigned)g.Hin || (unsigned)iw >= (unsigned)g.Win) return 0.f;
  return __ldg(x + q.base + ci * g.xs[1] + (long long)kd * g.xs[2] + (long long)kh * g.xs[3] + (long long)kw * g.xs[4]);
}

// dy[(id+p-kd)/s ...][co] for DGRAD; r = (co, kd', kh', kw') over the taps kd = kd0 + kd'*sD ... of the row's parity class, for which the
// division by the stride is exact by construction
__device__ __forceinline__ float gather_dy(const Geo& g, const Phase& f, const float* __restrict__ dy, const PosInfo& q, int r, int ptaps) {
  if (!q.ok || r >= g.Cout * ptaps) return 0.f;
  int co = fast_div(r, f.m_pt, f.s_pt), tap = r - co * ptaps;
  int kd = fast_div(tap, f.m_hw, f.s_hw); tap -= kd * f.nkh * f.nkw;
  int kh = fast_div(tap, f.m_w, f.s_w), kw = tap - kh * f.nkw;
  int td = q.d0 - (f.kd0 + kd * g.sD), th = q.h0 - (f.kh0 + kh * g.sH), tw = q.w0 - (f.kw0 + kw * g.sW);
  if (td < 0 || th < 0 || tw < 0) return 0.f;
  int od = g.sD == 2 ? td >> 1 : td / g.sD, oh = g.sH == 2 ? th >> 1 : th / g.sH, ow = g.sW == 2 ? tw >> 1 : tw / g.sW;
  if (od >= g.Dout || oh >= g.Hout || ow >= g.Wout) return 0.f;
  return __ldg(dy + q.base + co * g.ys[1] + (long long)od * g.ys[2] + (long long)oh * g.ys[3] + (long long)ow * g.ys[4]);
}

template <int MODE, int BN, int TN>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_kernel(Geo g, const float* __restrict__ src, const float* __restrict__ wgt,
                                                             const float* __restrict__ aux, float* __restrict__ dst, int act,
                                                             long long r_per_split) {
  // src: x (FWD/WGRAD) or dy (DGRAD); aux: bias (FWD) or dy (WGRAD); dst: y / dx / dW
  static_assert((BM / TM) * (BN / TN) == NTHREADS, "tile/thread mismatch");
  constexpr int BNP = BN + 4;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BNP];
  constexpr int A_PER_THREAD = BM * BK / NTHREADS;                 // 8
  constexpr int B_PER_THREAD = (BN * BK + NTHREADS - 1) / NTHREADS;

  const int tid = threadIdx.x;
  const long long i0 = (long long)blockIdx.x * BM;
  const int j0 = blockIdx.y * BN;
  const int tm = tid & 31, tn = tid >> 5;

  long long R, r_begin, r_end;
  int NJ;
  Phase f = {};
  int ptaps = 0;
  if (MODE == MODE_FWD) { R = g.K; NJ = g.Cout; }
  else if (MODE == MODE_DGRAD) {
    f = make_phase(g, blockIdx.z);                 // blockIdx.z = parity class, never a reduction split
    ptaps = f.nkd * f.nkh * f.nkw;
    R = (long long)g.Cout * ptaps;
    NJ = g.Cin;
    if (i0 >= f.rows) return;
  } else { R = g.Pout; NJ = g.Cout; }
  r_begin = MODE == MODE_DGRAD ? 0 : (long long)blockIdx.z * r_per_split;
  r_end = MODE == MODE_DGRAD ? R : (r_begin + r_per_split < R ? r_begin + r_per_split : R);
  if (MODE != MODE_DGRAD && r_begin >= r_end) return;

  // ---- operand loaders -------------------------------------------------------------------
  // FWD / DGRAD: A rows are positions; each thread owns one row (tid & 127) and strides over r.
  // WGRAD: A rows are k=(ci,tap); the reduction index r is the position -> lanes run along r.
  PosInfo rowq;
  if (MODE == MODE_FWD) rowq = decode_out_pos(g, i0 + (tid & (BM - 1)));
  if (MODE == MODE_DGRAD) rowq = decode_in_pos(g, f, i0 + (tid & (BM - 1)));

  float ra[A_PER_THREAD], rb[B_PER_THREAD];

  auto load_tiles = [&](long long r0) {
    if (MODE == MODE_WGRAD) {
      // r fast: lane (tid & 15) walks BK positions, (tid >> 4) + 16*e walks the 128 k rows
      const int rr = tid & (BK - 1);
      PosInfo q = decode_out_pos(g, r0 + rr);
      q.ok = q.ok && (r0 + rr) < r_end;
#pragma unroll
      for (int e = 0; e < A_PER_THREAD; ++e) {
        int ii = (tid >> 4) + 16 * e;
        ra[e] = gather_x(g, src, q, (int)(i0 + ii));
      }
      // B(r, j) = dy[pos r][co j]
      long long m = r0 + rr;
      bool ok = m < r_end;
      long long off = ok ? out_offset(g, m) : 0;
#pragma unroll
      for (int e = 0; e < B_PER_THREAD; ++e) {
        int jj = (tid >> 4) + 16 * e;
        int co = j0 + jj;
        rb[e] = (ok && jj < BN && co < NJ) ? __ldg(aux + off + co * g.ys[1]) : 0.f;
      }
    } else {
      const int rbase = tid >> 7;   // 0..1
#pragma unroll
      for (int e = 0; e < A_PER_THREAD; ++e) {
        int rr = rbase + 2 * e;
        long long r = r0 + rr;
        ra[e] = (r < r_end) ? (MODE == MODE_FWD ? gather_x(g, src, rowq, (int)r) : gather_dy(g, f, src, rowq, (int)r, ptaps)) : 0.f;
      }
      // B(r, j): FWD w[co*K + k];  DGRAD w[(co*Cin + ci)*taps + tap]
      const int rr = tid & (BK - 1);
      long long r = r0 + rr;
#pragma unroll
      for (int e = 0; e < B_PER_THREAD; ++e) {
        int jj = (tid >> 4) + 16 * e;
        int j = j0 + jj;
        float v = 0.f;
        if (r < r_end && jj < BN && j < NJ) {
          if (MODE == MODE_FWD) v = __ldg(wgt + (long long)j * g.K + r);
          else {
            int co = fast_div((int)r, f.m_pt, f.s_pt), tap = (int)r - co * ptaps;
            int kd = fast_div(tap, f.m_hw, f.s_hw); tap -= kd * f.nkh * f.nkw;
            int kh = fast_div(tap, f.m_w, f.s_w), kw = tap - kh * f.nkw;
            const int full = ((f.kd0 + kd * g.sD) * g.kH + (f.kh0 + kh * g.sH)) * g.kW + (f.kw0 + kw * g.sW);
            v = __ldg(wgt + ((long long)co * g.Cin + j) * g.taps + full);
          }
        }
        rb[e] = v;
      }
    }
  };
  auto store_tiles = [&]() {
    if (MODE == MODE_WGRAD) {
      const int rr = tid & (BK - 1);
#pragma unroll
      for (int e = 0; e < A_PER_THREAD; ++e) As[rr][(tid >> 4) + 16 * e] = ra[e];
#pragma unroll
      for (int e = 0; e < B_PER_THREAD; ++e) {
        int jj = (tid >> 4) + 16 * e;
        if (jj < BN) Bs[rr][jj] = rb[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < A_PER_THREAD; ++e) As[(tid >> 7) + 2 * e][tid & (BM - 1)] = ra[e];
      const int rr = tid & (BK - 1);
#pragma unroll
      for (int e = 0; e < B_PER_THREAD; ++e) {
        int jj = (tid >> 4) + 16 * e;
        if (jj < BN) Bs[rr][jj] = rb[e];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  if (r_begin < r_end) {          // (a parity class that no tap reaches still writes its zeros below)
    load_tiles(r_begin);
    store_tiles();
  }
  __syncthreads();
  for (long long r0 = r_begin; r0 < r_end; r0 += BK) {
    const bool more = r0 + BK < r_end;
    if (more) load_tiles(r0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm * TM]);
      float bv[TN];
#pragma unroll
      for (int b = 0; b < TN; ++b) bv[b] = Bs[kk][tn * TN + b];
      const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(a4[a], bv[b], acc[a][b]);
    }
    __syncthreads();
    if (more) {
      store_tiles();
      __syncthreads();
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    long long i = i0 + tm * TM + a;
    if (MODE == MODE_FWD) {
      if (i >= g.Pout) continue;
      long long off = out_offset(g, i);
#pragma unroll
      for (int b = 0; b < TN; ++b) {
        int j = j0 + tn * TN + b;
        if (j < NJ) {
          float v = acc[a][b] + (aux ? __ldg(aux + j) : 0.f);
          dst[off + j * g.ys[1]] = cvad_act(v, act);
        }
      }
    } else if (MODE == MODE_DGRAD) {
      if (i >= f.rows) continue;
      long long off = in_offset(g, f, i);
#pragma unroll
      for (int b = 0; b < TN; ++b) {
        int j = j0 + tn * TN + b;
        if (j < NJ) {
          float v = acc[a][b];
          if (act) v += dst[off + j * g.xs[1]];     // act != 0 here means "accumulate into dx"
          dst[off + j * g.xs[1]] = v;
        }
      }
    } else {
      if (i >= g.K) continue;
#pragma unroll
      for (int b = 0; b < TN; ++b) {
        int j = j0 + tn * TN + b;
        if (j < NJ) atomicAdd(dst + (long long)j * g.K + i, acc[a][b]);
      }
    }
  }
}

Geo make_geo(const cvad_conv_desc* d) {
  Geo g;
  g.N = d->N; g.Cin = d->Cin; g.Din = d->Din; g.Hin = d->Hin; g.Win = d->Win;
  g.Cout = d->Cout; g.Dout = d->Dout; g.Hout = d->Hout; g.Wout = d->Wout;
  g.kD = d->kD; g.kH = d->kH; g.kW = d->kW; g.sD = d->sD; g.sH = d->sH; g.sW = d->sW;
  g.pD = d->pD; g.pH = d->pH; g.pW = d->pW;
  for (int i = 0; i < 5; ++i) { g.xs[i] = d->xs[i]; g.ys[i] = d->ys[i]; }
  g.taps = g.kD * g.kH * g.kW;
  g.K = g.Cin * g.taps;
  g.Pout = (long long)g.N * g.Dout * g.Hout * g.Wout;
  g.Pin = (long long)g.N * g.Din * g.Hin * g.Win;
  fdiv_init((unsigned)g.taps, g.m_taps, g.s_taps);
  fdiv_init((unsigned)(g.kH * g.kW), g.m_khw, g.s_khw);
  fdiv_init((unsigned)g.kW, g.m_kw, g.s_kw);
  fdiv_init((unsigned)g.Wout, g.m_wo, g.s_wo);
  fdiv_init((unsigned)g.Hout, g.m_ho, g.s_ho);
  fdiv_init((unsigned)g.Dout, g.m_do, g.s_do);
  g.small = g.Pout < (1LL << 31) - 4096 && g.Pin < (1LL << 31) - 4096;
  return g;
}

template <int MODE>
int launch(const Geo& g, const float* src, const float* wgt, const float* aux, float* dst, int act, cudaStream_t st) {
  long long rows = MODE == MODE_FWD ? g.Pout : (MODE == MODE_DGRAD ? g.Pin : g.K);
  int nj = MODE == MODE_DGRAD ? g.Cin : g.Cout;
  long long R = MODE == MODE_FWD ? g.K : (MODE == MODE_DGRAD ? (long long)g.Cout * g.taps : g.Pout);
  int splits = 1;
  if (MODE == MODE_DGRAD) {       // grid.x covers the largest parity class (class 0), grid.z the sD*sH*sW classes
    rows = make_phase(g, 0).rows;
    splits = g.sD * g.sH * g.sW;
  }
  long long gx = (rows + BM - 1) / BM;
  if (MODE == MODE_WGRAD) {
    // split the (huge) position reduction so that the grid fills the machine ~4x
    int bn = nj > 32 ? 64 : (nj > 16 ? 32 : (nj > 8 ? 16 : 8));
    long long tiles = gx * ((nj + bn - 1) / bn);
    long long want = (4LL * cvad_num_sms() + tiles - 1) / tiles;
    long long maxs = (R + 511) / 512;
    splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
    if (splits < 1) splits = 1;
  }
  long long rps = R;
  if (MODE != MODE_DGRAD) {
    rps = ((R + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (int)((R + rps - 1) / rps);
  }
  if (nj > 32) {
    dim3 grid((unsigned)gx, (nj + 63) / 64, splits);
    conv_gemm_kernel<MODE, 64, 8><<<grid, NTHREADS, 0, st>>>(g, src, wgt, aux, dst, act, rps);
  } else if (nj > 16) {
    dim3 grid((unsigned)gx, 1, splits);
    conv_gemm_kernel<MODE, 32, 4><<<grid, NTHREADS, 0, st>>>(g, src, wgt, aux, dst, act, rps);
  } else if (nj > 8) {
    dim3 grid((unsigned)gx, 1, splits);
    conv_gemm_kernel<MODE, 16, 2><<<grid, NTHREADS, 0, st>>>(g, src, wgt, aux, dst, act, rps);
  } else {
    dim3 grid((unsigned)gx, 1, splits);
    conv_gemm_kernel<MODE, 8, 1><<<grid, NTHREADS, 0, st>>>(g, src, wgt, aux, dst, act, rps);
  }
  CVAD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

CVAD_API int cvad_conv_fwd_f32(const cvad_conv_desc* d, const float* x, const float* w, const float* bias, float* y, int act,
                               void* stream) {
  Geo g = make_geo(d);
  return launch<MODE_FWD>(g, x, w, bias, y, act, (cudaStream_t)stream);
}

CVAD_API int cvad_conv_dgrad_f32(const cvad_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream) {
  Geo g = make_geo(d);
  return launch<MODE_DGRAD>(g, dy, w, nullptr, dx, accumulate, (cudaStream_t)stream);
}

CVAD_API int cvad_conv_wgrad_f32(const cvad_conv_desc* d, const float* x, const float* dy, float* dw, void* stream) {
  // dw (OIDHW, contiguous) must be zero-initialised (or hold a value to accumulate onto): partial sums are added atomically.
  Geo g = make_geo(d);
  return launch<MODE_WGRAD>(g, x, nullptr, dy, dw, 0, (cudaStream_t)stream);
}
