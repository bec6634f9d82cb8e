// CUDA-core implicit-GEMM convolution (fp32 accumulate), NHWC.
//
// Role: (1) the "strict" fp32 engine used for parity gating (the tensor-core engine stores fp16 activations),
//       (2) the engine for the few layers the tcgen05 kernel does not take (C_in = 3 stems).
// Replaces nn.Conv2d + folded BatchNorm2d (+ residual add) (+ ReLU) as used by
//   /root/reference/celldetection/models/commons.py:120-149 (TwoConvNormRelu), :461-511 (ReadOut),
//   models/resnet.py:56-116,265-290 (stem, BasicBlock, Bottleneck incl. grouped 3x3), models/fpn.py:79-134.
//
// Tiling: one CTA = 64 output pixels x 64 output channels, K chunks of 16 input channels per filter tap, 256 threads,
// each thread a 4x4 register tile.  Weights are pre-packed on the host as float [R*S][kslab][cout]; grouped
// convolutions are expanded to block-diagonal 64-channel slabs (see cpn_op_t::kslab).
#include "common.cuh"

namespace cpn {

struct ConvSimtParams {
  const void* src;
  void* dst;
  const void* res;
  const float* wgt;
  const float* bias;
  int N, H, W, src_pitch;
  int Ho, Wo, dst_pitch, cout;
  int res_h, res_w, res_pitch;
  int R, S, stride, pad;
  int kslab, slab_mode;
  int act;
  long long M;
};

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int cbase = p.slab_mode ? (n0 / p.kslab) * p.kslab : 0;

  // A-load assignment: pixel la_m, channels la_c .. la_c+3 of the current chunk
  const int la_m = tid >> 2, la_c = (tid & 3) * 4;
  const long long m_ld = m0 + la_m;
  const bool m_ok = m_ld < p.M;
  int n_img = 0, oy = 0, ox = 0;
  if (m_ok) {
    n_img = (int)(m_ld / ((long long)p.Ho * p.Wo));
    int rem = (int)(m_ld - (long long)n_img * p.Ho * p.Wo);
    oy = rem / p.Wo;
    ox = rem - oy * p.Wo;
  }
  const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
  const T* src = reinterpret_cast<const T*>(p.src);
  const bool vec_ok = (p.kslab % 4 == 0) && (p.src_pitch % 4 == 0);

  // B-load assignment: k = lb_k, couts lb_n..+3
  const int lb_k = tid >> 4, lb_n = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < p.R * p.S; ++tap) {
    const int r = tap / p.S, s = tap - r * p.S;
    const int iy = iy0 + r, ix = ix0 + s;
    const bool pix_ok = m_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const T* sp = src + (((long long)n_img * p.H + iy) * p.W + ix) * p.src_pitch + cbase;
    const float* wp = p.wgt + (long long)tap * p.kslab * p.cout;
    for (int c0 = 0; c0 < p.kslab; c0 += BK) {
      // ---- stage A chunk (transposed into k-major) ----
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (pix_ok) {
        const int c = c0 + la_c;
        if (vec_ok && c + 3 < p.kslab) {
          if (sizeof(T) == 4) {
            float4 v = *reinterpret_cast<const float4*>(sp + c);
            av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
          } else {
            uint2 raw = *reinterpret_cast<const uint2*>(sp + c);
            __half2 h0 = *reinterpret_cast<__half2*>(&raw.x), h1 = *reinterpret_cast<__half2*>(&raw.y);
            av[0] = __low2float(h0); av[1] = __high2float(h0); av[2] = __low2float(h1); av[3] = __high2float(h1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < p.kslab) av[j] = to_f32<T>(sp[c + j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[la_c + j][la_m] = av[j];
      // ---- stage B chunk ----
      {
        const int k = c0 + lb_k;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.kslab) {
          const float* w = wp + (long long)k * p.cout + n0 + lb_n;
          if ((p.cout % 4 == 0) && n0 + lb_n + 3 < p.cout) {
            bv = *reinterpret_cast<const float4*>(w);
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n0 + lb_n + j < p.cout) t[j] = w[j];
            bv = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        *reinterpret_cast<float4*>(&Bs[lb_k][lb_n]) = bv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w};
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias, residual (optionally through nearest up-sampling), activation, store ----
  T* dst = reinterpret_cast<T*>(p.dst);
  const T* res = reinterpret_cast<const T*>(p.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int n_i = (int)(m / ((long long)p.Ho * p.Wo));
    const int rem = (int)(m - (long long)n_i * p.Ho * p.Wo);
    const int y = rem / p.Wo, x = rem - y * p.Wo;
    T* dp = dst + m * p.dst_pitch;
    const T* rp = nullptr;
    if (res) {
      const int ry = (p.res_h == p.Ho) ? y : (int)(((long long)y * p.res_h) / p.Ho);
      const int rx = (p.res_w == p.Wo) ? x : (int)(((long long)x * p.res_w) / p.Wo);
      rp = res + (((long long)n_i * p.res_h + ry) * p.res_w + rx) * p.res_pitch;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (rp) v += to_f32<T>(rp[n]);
      if (p.act == CPN_ACT_RELU) v = fmaxf(v, 0.f);
      dp[n] = from_f32<T>(v);
    }
  }
}

int conv_simt_launch(const cpn_op_t& op, const void* src, void* dst, const void* res, const void* wgt,
                     const float* bias, cudaStream_t st) {
  CPN_REQUIRE(op.src.dtype == op.dst.dtype && (op.src.dtype == CPN_DT_F32 || op.src.dtype == CPN_DT_F16),
              "conv_simt: unsupported dtypes %d -> %d", op.src.dtype, op.dst.dtype);
  CPN_REQUIRE(op.res.n == 0 || op.res.dtype == op.dst.dtype, "conv_simt: residual dtype mismatch");
  ConvSimtParams p;
  p.src = src; p.dst = dst; p.res = op.res.n ? res : nullptr;
  p.wgt = reinterpret_cast<const float*>(wgt); p.bias = bias;
  p.N = op.src.n; p.H = op.src.h; p.W = op.src.w; p.src_pitch = op.src.pitch;
  p.Ho = op.dst.h; p.Wo = op.dst.w; p.dst_pitch = op.dst.pitch; p.cout = op.dst.c;
  p.res_h = op.res.n ? op.res.h : 0; p.res_w = op.res.n ? op.res.w : 0; p.res_pitch = op.res.pitch;
  p.R = op.r; p.S = op.s; p.stride = op.stride; p.pad = op.pad;
  p.kslab = op.kslab; p.slab_mode = op.slab_mode; p.act = op.act;
  p.M = (long long)op.dst.n * op.dst.h * op.dst.w;
  const int eh = (op.src.h + 2 * op.pad - op.r) / op.stride + 1, ew = (op.src.w + 2 * op.pad - op.s) / op.stride + 1;
  CPN_REQUIRE(eh == op.dst.h && ew == op.dst.w && op.src.n == op.dst.n,
              "conv_simt: output shape mismatch (%dx%d expected %dx%d)", op.dst.h, op.dst.w, eh, ew);
  CPN_REQUIRE(!op.slab_mode || (op.kslab % 64 == 0), "conv_simt: grouped slab must be a multiple of 64");
  dim3 grid((unsigned)ceil_div64(p.M, BM), (unsigned)((p.cout + BN - 1) / BN));
  if (op.src.dtype == CPN_DT_F32)
    conv_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  else
    conv_simt_kernel<__half><<<grid, 256, 0, st>>>(p);
  CPN_CHECK_LAUNCH();
  return 0;
}

}  // namespace cpn
