// Frame preparation on the device: what the reference's loaders do per frame on the host with OpenCV --
//   cv2.resize(frame, (360, 240)) on grayscale uint8 frames                     causal_anomaly_detection.py:89-96  (then Normalize, fused into the stem)
//   cv2.resize(img, frame_size) + BGR->RGB + /255 + (T,H,W,3)->(3,T,H,W)        avenue loaders (s1:19, 86-92; bbox:397-411)
// so that a video is uploaded ONCE as raw uint8 frames and every clip / sliding window is cut, resized and laid out where the model
// reads it, instead of crossing PCIe as fp32 per clip.
//
// cvad_resize_bilinear_u8 reproduces cv2.resize(..., interpolation=INTER_LINEAR) for 8-bit images BIT-EXACTLY: OpenCV's fixed-point
// scheme -- 11-bit coefficients saturate_cast<short>(w * 2048) from fx = (float)((dx + 0.5) * scale - 0.5), horizontal pass in int32,
// vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 -- including its asymmetric border rules (columns: the
// fraction is zeroed when the tap is clamped; rows: only the tap index is clamped).
#include "common.cuh"
#include "cvad_b200.h"

namespace {

struct Tap {
  int i0, i1, w0, w1;
};

__device__ __forceinline__ Tap make_tap(int d, int src, double scale, bool zero_frac_when_clamped) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (zero_frac_when_clamped) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.i0 = min(max(s, 0), src - 1);
  t.i1 = min(max(s + 1, 0), src - 1);
  t.w0 = __float2int_rn((1.f - f) * 2048.f);
  t.w1 = __float2int_rn(f * 2048.f);
  return t;
}

// src (N, sh, sw, C) uint8 interleaved -> dst (N, dh, dw, C) uint8
__global__ void resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, int N, int sh, int sw, int C, uint8_t* __restrict__ dst, int dh, int dw,
                                          double scale_x, double scale_y) {
  const long long total = (long long)N * dh * dw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(t % dw);
    const long long r = t / dw;
    const int dy = (int)(r % dh), n = (int)(r / dh);
    const Tap tx = make_tap(dx, sw, scale_x, true), ty = make_tap(dy, sh, scale_y, false);
    const uint8_t* f = src + (long long)n * sh * sw * C;
    const uint8_t* r0 = f + (long long)ty.i0 * sw * C;
    const uint8_t* r1 = f + (long long)ty.i1 * sw * C;
    uint8_t* o = dst + t * C;
    for (int c = 0; c < C; ++c) {
      const int s0 = (int)__ldg(r0 + tx.i0 * C + c) * tx.w0 + (int)__ldg(r0 + tx.i1 * C + c) * tx.w1;
      const int s1 = (int)__ldg(r1 + tx.i0 * C + c) * tx.w0 + (int)__ldg(r1 + tx.i1 * C + c) * tx.w1;
      const int v = (((ty.w0 * (s0 >> 4)) >> 16) + ((ty.w1 * (s1 >> 4)) >> 16) + 2) >> 2;
      o[c] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

// frames (F, H, W, C) uint8 (C = 1 or 3) + clip starts -> clips (B, C, T, H, W) fp32 = value * scale, channel order optionally reversed
// (BGR -> RGB): the Avenue loaders' cvtColor + /255 + permute, and the sliding-window gather of bbox:392-411, in one pass.
__global__ void clips_from_frames_kernel(const uint8_t* __restrict__ frames, int F, int H, int W, int C, const int* __restrict__ starts, int B, int T,
                                         int frame_stride, float scale, int reverse_channels, float* __restrict__ out) {
  const long long hw = (long long)H * W;
  const long long total = (long long)B * C * T * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long p = t % hw;
    long long r = t / hw;
    const int tt = (int)(r % T); r /= T;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    int fr = starts[b] + tt * frame_stride;
    fr = min(max(fr, 0), F - 1);
    const int cs = reverse_channels ? C - 1 - c : c;
    out[t] = (float)__ldg(frames + ((long long)fr * hw + p) * C + cs) * scale;
  }
}

// the same gather for the M-A input: clips (B, T, 1, H, W) uint8 from grayscale frames (F, H, W) (cad:57 sequence windows)
__global__ void clips_from_frames_u8_kernel(const uint8_t* __restrict__ frames, int F, long long hw, const int* __restrict__ starts, int B, int T,
                                            int frame_stride, uint8_t* __restrict__ out) {
  const long long total = (long long)B * T * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long p = t % hw;
    const long long r = t / hw;
    const int tt = (int)(r % T), b = (int)(r / T);
    int fr = starts[b] + tt * frame_stride;
    fr = min(max(fr, 0), F - 1);
    out[t] = __ldg(frames + (long long)fr * hw + p);
  }
}

inline unsigned frame_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 16LL * cvad_num_sms()) b = 16LL * cvad_num_sms();
  return (unsigned)b;
}

}  // namespace

CVAD_API int cvad_resize_bilinear_u8(const void* src, int N, int src_h, int src_w, int C, void* dst, int dst_h, int dst_w, void* stream) {
  if (N <= 0) return 0;
  if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || C < 1 || C > 4) return (int)cudaErrorInvalidValue;
  // OpenCV: inv_scale = dsize / ssize (double); scale = 1. / inv_scale
  const double scale_x = 1.0 / ((double)dst_w / src_w), scale_y = 1.0 / ((double)dst_h / src_h);
  resize_bilinear_u8_kernel<<<frame_blocks((long long)N * dst_h * dst_w), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, N, src_h, src_w, C,
                                                                                                         (uint8_t*)dst, dst_h, dst_w, scale_x, scale_y);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_clips_from_frames_f32(const void* frames, int F, int H, int W, int C, const int* starts, int B, int T, int frame_stride,
                                        float scale, int reverse_channels, float* out, void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if (F <= 0 || H <= 0 || W <= 0 || C < 1 || C > 4) return (int)cudaErrorInvalidValue;
  clips_from_frames_kernel<<<frame_blocks((long long)B * C * T * H * W), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames, F, H, W, C, starts, B, T,
                                                                                                       frame_stride, scale, reverse_channels, out);
  CVAD_LAUNCH_CHECK();
  return 0;
}

CVAD_API int cvad_clips_from_frames_u8(const void* frames, int F, int H, int W, const int* starts, int B, int T, int frame_stride, void* out,
                                       void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if (F <= 0 || H <= 0 || W <= 0) return (int)cudaErrorInvalidValue;
  clips_from_frames_u8_kernel<<<frame_blocks((long long)B * T * H * W), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames, F, (long long)H * W, starts,
                                                                                                      B, T, frame_stride, (uint8_t*)out);
  CVAD_LAUNCH_CHECK();
  return 0;
}
