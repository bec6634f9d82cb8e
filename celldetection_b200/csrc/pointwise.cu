// Memory-bound layer kernels of the plan (NHWC, fp32 or fp16 activations): input preparation, max-pooling, nearest and
// bilinear resampling and the ReadOut projection.  All are HBM-bound element-wise / gather work: coalesced along the
// channel axis, 8-byte or 16-byte vector accesses where the pitch allows, grid sized in multiples of the SM count.
#include "common.cuh"
#include <cstring>
#include <cstdlib>
#include <cmath>

namespace cpn {

static inline int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------------------------------
// PREP: network input -> NHWC activations.  Fuses Normalize's range assertion (commons.py:694-700; mean 0 / std 1 so the
// affine part is the identity), the uint8 -> float / 255 of lightning_base.py:774-780 and the NCHW -> NHWC change.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store(__half* o, int lo_delta, float v) {
  const __half h = __float2half_rn(v);
  o[0] = h;
  o[lo_delta] = __float2half_rn(v - __half2float(h));
}

// Second block of a pixel: none (lo_delta == 0), fp16 lo halves (CPN_DT_F16X2) or the e4m3 (lo8 | hi8) chunks of
// CPN_DT_F16F8 with their power-of-two scales.
struct LoFmt {
  int lo_delta;
  int f8;
  float lo_scale, hi8_scale, lo_inv;
};
static LoFmt lo_fmt(const cpn_view_t& v) {
  LoFmt f;
  f.lo_delta = dtype_has_lo(v.dtype) ? v.lo_delta : 0;
  f.f8 = v.dtype == CPN_DT_F16F8;
  f.lo_scale = ldexpf(1.f, 8 + v.fp8_exp); f.hi8_scale = ldexpf(1.f, -2 + v.fp8_exp); f.lo_inv = ldexpf(1.f, -(8 + v.fp8_exp));
  return f;
}
// 8 consecutive channels (c0 % 8 == 0) of a CPN_DT_F16F8 pixel: hi halves + 8 lo8 bytes + 8 hi8 bytes
__device__ __forceinline__ void f16f8_store8(__half* px, const LoFmt& F, int c0, const float (&v)[8]) {
  __align__(16) __half h[8];
  float lo[8], hf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = __float2half_rn(v[j]); hf[j] = __half2float(h[j]); lo[j] = (v[j] - hf[j]) * F.lo_scale; hf[j] *= F.hi8_scale; }
  *reinterpret_cast<uint4*>(px + c0) = *reinterpret_cast<const uint4*>(h);
  uint8_t* q = f8_block(px, F.lo_delta, c0);
  *reinterpret_cast<uint2*>(q) = make_uint2(pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]), pack_e4m3x4(lo[4], lo[5], lo[6], lo[7]));
  *reinterpret_cast<uint2*>(q + 32) = make_uint2(pack_e4m3x4(hf[0], hf[1], hf[2], hf[3]), pack_e4m3x4(hf[4], hf[5], hf[6], hf[7]));
}

template <typename T>
__global__ void prep_kernel(const void* __restrict__ in, int fmt, T* __restrict__ out, int N, int C, int H, int W,
                            int pitch, int32_t* __restrict__ flags, const LoFmt F) {
  const int lo_delta = F.lo_delta;
  const long long total = (long long)N * H * W;
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ((long long)H * W));
    const long long rem = i - (long long)n * H * W;
    T* o = out + i * pitch;
    for (int c = 0; c < C; ++c) {
      float v;
      if (fmt == CPN_IN_F32_NCHW) {
        v = reinterpret_cast<const float*>(in)[((long long)n * C + c) * H * W + rem];
        bad |= !(v >= 0.f && v <= 1.f);
      } else if (fmt == CPN_IN_U8_NCHW) {
        v = (float)reinterpret_cast<const uint8_t*>(in)[((long long)n * C + c) * H * W + rem] / 255.f;
      } else {
        v = (float)reinterpret_cast<const uint8_t*>(in)[i * C + c] / 255.f;
      }
      if (sizeof(T) == 2 && F.f8) { f16f8_store(reinterpret_cast<__half*>(o), lo_delta, c, v, F.lo_scale, F.hi8_scale); continue; }
      o[c] = from_f32<T>(v);
      if (sizeof(T) == 2 && lo_delta > 0) o[c + lo_delta] = from_f32<T>(v - to_f32<T>(o[c]));
    }
  }
  if (bad) atomicOr(flags, 1);
}

// PREP, im2col mode (op.r > 0): the network input is expanded to the K-major operand of the stem convolution,
//   dst[n, oy, ox, (r*k + s)*C + c] = in[n, c, oy*stride - pad + r, ox*stride - pad + s]   (zero outside, zero-padded
// to a multiple of 64 channels), so that the C_in = 3 stem (7x7 s2, resnet.py:273; 3x3 s1, unet.py:52) runs on the
// tensor cores as a 1x1 convolution with K = 64-padded k*k*C.  One thread writes 8 consecutive channels (16 bytes).
template <typename T>
__global__ void __launch_bounds__(256) prep_im2col_kernel(const void* __restrict__ in, int fmt, T* __restrict__ out, int N,
                                                          int C, int H, int W, int Ho, int Wo, int Kp, int pitch, int k,
                                                          int stride, int pad, int32_t* __restrict__ flags,
                                                          const LoFmt F) {
  const int lo_delta = F.lo_delta;
  // per-entry tables: e -> (dy, dx, c) and the NCHW / NHWC offsets relative to the window origin
  __shared__ int off_nchw[512], off_nhwc[512];
  __shared__ signed char dys[512], dxs[512];
  const int kk = k * k * C;
  for (int e = threadIdx.x; e < Kp && e < 512; e += blockDim.x) {
    const int tap = e / C, c = e - tap * C;
    const int r = tap / k, s_ = tap - r * k;
    dys[e] = (signed char)r; dxs[e] = (signed char)s_;
    off_nchw[e] = (c * H + r) * W + s_;
    off_nhwc[e] = (r * W + s_) * C + c;
  }
  __syncthreads();
  const int chunks = Kp / 8;
  const long long total = (long long)N * Ho * Wo * chunks;
  bool bad = false;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long pix = i / chunks;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
    const bool interior = iy0 >= 0 && ix0 >= 0 && iy0 + k <= H && ix0 + k <= W;
    const long long base_nchw = ((long long)n * C * H + iy0) * W + ix0;
    const long long base_nhwc = (((long long)n * H + iy0) * W + ix0) * C;
    __align__(16) T vals[8];
    __align__(16) T vlo[8];
    float vf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = ch * 8 + j;
      float v = 0.f;
      if (e < kk) {
        const bool ok = interior || (iy0 + dys[e] >= 0 && iy0 + dys[e] < H && ix0 + dxs[e] >= 0 && ix0 + dxs[e] < W);
        if (ok) {
          if (fmt == CPN_IN_F32_NCHW) {
            v = __ldg(reinterpret_cast<const float*>(in) + base_nchw + off_nchw[e]);
            bad |= !(v >= 0.f && v <= 1.f);
          } else if (fmt == CPN_IN_U8_NCHW) {
            v = (float)__ldg(reinterpret_cast<const uint8_t*>(in) + base_nchw + off_nchw[e]) / 255.f;
          } else {
            v = (float)__ldg(reinterpret_cast<const uint8_t*>(in) + base_nhwc + off_nhwc[e]) / 255.f;
          }
        }
      }
      vals[j] = from_f32<T>(v);
      vlo[j] = from_f32<T>(v - to_f32<T>(vals[j]));
      vf[j] = v;
    }
    T* o = out + (((long long)n * Ho + oy) * Wo + ox) * pitch + ch * 8;
    if (sizeof(T) == 2 && F.f8) {
      f16f8_store8(reinterpret_cast<__half*>(o) - ch * 8, F, ch * 8, vf);
    } else if (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(vals);
      if (lo_delta > 0) *reinterpret_cast<uint4*>(o + lo_delta) = *reinterpret_cast<const uint4*>(vlo);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = vals[j];
    }
  }
  if (bad) atomicOr(flags, 1);
}

// Tiled variant: a CTA stages the input window of a TY x TX tile of output pixels in shared memory with coalesced loads
// (each input value is needed by up to k*k / stride^2 output pixels: the direct kernel re-fetched it with scattered 4-byte
// loads and ran at 0.9 TB/s of output, ncu r01: 10.8 % DRAM) and then writes the K-major operand rows with 16-byte stores,
// consecutive threads on consecutive 8-channel chunks of one pixel.
constexpr int PREP_TY = 4, PREP_TX = 32;

template <typename T>
__global__ void __launch_bounds__(256) prep_im2col_tiled_kernel(const void* __restrict__ in, int fmt, T* __restrict__ out,
                                                                int N, int C, int H, int W, int Ho, int Wo, int Kp,
                                                                int pitch, int k, int stride, int pad,
                                                                int32_t* __restrict__ flags, const LoFmt F, int WH, int WW) {
  const int lo_delta = F.lo_delta;
  extern __shared__ float prep_sm[];                  // [C][WH][WW] window, then int off_sm[Kp]
  int* off_sm = reinterpret_cast<int*>(prep_sm + C * WH * WW);
  const int kk = k * k * C;
  for (int e = threadIdx.x; e < Kp; e += blockDim.x) {
    const int tap = e / C, c = e - tap * C;
    const int r = tap / k, s_ = tap - r * k;
    off_sm[e] = e < kk ? (c * WH + r) * WW + s_ : -1;
  }
  const int tiles_x = (Wo + PREP_TX - 1) / PREP_TX, tiles_y = (Ho + PREP_TY - 1) / PREP_TY;
  int b = blockIdx.x;
  const int tx = b % tiles_x; b /= tiles_x;
  const int ty = b % tiles_y;
  const int n = b / tiles_y;
  const int oy0 = ty * PREP_TY, ox0 = tx * PREP_TX;
  const int iy0 = oy0 * stride - pad, ix0 = ox0 * stride - pad;
  bool bad = false;
  for (int i = threadIdx.x; i < C * WH * WW; i += blockDim.x) {
    const int wx = i % WW, t = i / WW, wy = t % WH, c = t / WH;
    const int iy = iy0 + wy, ix = ix0 + wx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      if (fmt == CPN_IN_F32_NCHW) {
        v = __ldg(reinterpret_cast<const float*>(in) + (((long long)n * C + c) * H + iy) * W + ix);
        bad |= !(v >= 0.f && v <= 1.f);
      } else if (fmt == CPN_IN_U8_NCHW) {
        v = (float)__ldg(reinterpret_cast<const uint8_t*>(in) + (((long long)n * C + c) * H + iy) * W + ix) / 255.f;
      } else {
        v = (float)__ldg(reinterpret_cast<const uint8_t*>(in) + (((long long)n * H + iy) * W + ix) * C + c) / 255.f;
      }
    }
    prep_sm[i] = v;
  }
  __syncthreads();
  const int chunks = Kp / 8;
  for (int i = threadIdx.x; i < PREP_TY * PREP_TX * chunks; i += blockDim.x) {
    const int ch = i % chunks, pxl = i / chunks;
    const int px = pxl % PREP_TX, py = pxl / PREP_TX;
    const int oy = oy0 + py, ox = ox0 + px;
    if (oy >= Ho || ox >= Wo) continue;
    const float* win = prep_sm + (py * stride) * WW + px * stride;
    __align__(16) T vals[8];
    __align__(16) T vlo[8];
    float vf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = off_sm[ch * 8 + j];
      const float v = o >= 0 ? win[o] : 0.f;
      vals[j] = from_f32<T>(v);
      vlo[j] = from_f32<T>(v - to_f32<T>(vals[j]));
      vf[j] = v;
    }
    T* o = out + (((long long)n * Ho + oy) * Wo + ox) * pitch + ch * 8;
    if (sizeof(T) == 2 && F.f8) {
      f16f8_store8(reinterpret_cast<__half*>(o) - ch * 8, F, ch * 8, vf);
    } else if (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(vals);
      if (lo_delta > 0) *reinterpret_cast<uint4*>(o + lo_delta) = *reinterpret_cast<const uint4*>(vlo);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = vals[j];
    }
  }
  if (bad) atomicOr(flags, 1);
}

int prep_launch(const cpn_op_t& op, const void* input, int input_format, void* dst, int32_t* flags, cudaStream_t st) {
  CPN_REQUIRE(input_format >= 0 && input_format <= 2, "prep: bad input format %d", input_format);
  if (op.r > 0) {  // im2col mode: src view describes the logical input (n, h, w, c)
    CPN_REQUIRE(op.dst.dtype != CPN_DT_F16F8 || op.dst.lo_delta % 32 == 0, "prep: fp16+e4m3 lo_delta must be a multiple of 32");
    CPN_REQUIRE(op.dst.c % 8 == 0 && op.dst.pitch % 8 == 0 && op.dst.c >= op.r * op.r * op.src.c && op.dst.c <= 512,
                "prep(im2col): dst channels %d must be a multiple of 8, >= k*k*c and <= 512", op.dst.c);
    {
      // tiled path (CPN_PREP_TILED=0 selects the direct kernel): window of a PREP_TY x PREP_TX output tile in smem
      static int tiled_env = -1;
      if (tiled_env < 0) { const char* e = getenv("CPN_PREP_TILED"); tiled_env = (e && atoi(e) == 0) ? 0 : 1; }
      const int WH = (PREP_TY - 1) * op.stride + op.r, WW = (PREP_TX - 1) * op.stride + op.r;
      const size_t smem = ((size_t)op.src.c * WH * WW + op.dst.c) * 4;
      const long long blocks = (long long)op.dst.n * ((op.dst.h + PREP_TY - 1) / PREP_TY) * ((op.dst.w + PREP_TX - 1) / PREP_TX);
      if (tiled_env && smem <= 48 * 1024 && blocks < (1ll << 31)) {
        if (op.dst.dtype == CPN_DT_F32)
          prep_im2col_tiled_kernel<float><<<(int)blocks, 256, smem, st>>>(
              input, input_format, (float*)dst, op.src.n, op.src.c, op.src.h, op.src.w, op.dst.h, op.dst.w, op.dst.c,
              op.dst.pitch, op.r, op.stride, op.pad, flags, LoFmt{0, 0, 1.f, 1.f, 1.f}, WH, WW);
        else
          prep_im2col_tiled_kernel<__half><<<(int)blocks, 256, smem, st>>>(
              input, input_format, (__half*)dst, op.src.n, op.src.c, op.src.h, op.src.w, op.dst.h, op.dst.w, op.dst.c,
              op.dst.pitch, op.r, op.stride, op.pad, flags, lo_fmt(op.dst), WH, WW);
        CPN_CHECK_LAUNCH();
        return 0;
      }
    }
    const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.c / 8);
    const int grid = grid_for(total, 256);
    if (op.dst.dtype == CPN_DT_F32)
      prep_im2col_kernel<float><<<grid, 256, 0, st>>>(input, input_format, (float*)dst, op.src.n, op.src.c, op.src.h,
                                                      op.src.w, op.dst.h, op.dst.w, op.dst.c, op.dst.pitch, op.r,
                                                      op.stride, op.pad, flags, LoFmt{0, 0, 1.f, 1.f, 1.f});
    else
      prep_im2col_kernel<__half><<<grid, 256, 0, st>>>(input, input_format, (__half*)dst, op.src.n, op.src.c, op.src.h,
                                                       op.src.w, op.dst.h, op.dst.w, op.dst.c, op.dst.pitch, op.r,
                                                       op.stride, op.pad, flags, lo_fmt(op.dst));
    CPN_CHECK_LAUNCH();
    return 0;
  }
  const long long total = (long long)op.dst.n * op.dst.h * op.dst.w;
  const int grid = grid_for(total, 256);
  if (op.dst.dtype == CPN_DT_F32)
    prep_kernel<float><<<grid, 256, 0, st>>>(input, input_format, (float*)dst, op.dst.n, op.dst.c, op.dst.h, op.dst.w,
                                             op.dst.pitch, flags, LoFmt{0, 0, 1.f, 1.f, 1.f});
  else
    prep_kernel<__half><<<grid, 256, 0, st>>>(input, input_format, (__half*)dst, op.dst.n, op.dst.c, op.dst.h,
                                              op.dst.w, op.dst.pitch, flags, lo_fmt(op.dst));
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// vector helpers: a "pack" is 4 channels (float4 / 4 halves)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Pack4;
template <>
struct Pack4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void get(float (&f)[4]) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void set(const float (&f)[4]) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <>
struct Pack4<__half> {
  uint2 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint2*>(p) = v; }
  __device__ __forceinline__ void get(float (&f)[4]) const {
    const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  __device__ __forceinline__ void set(const float (&f)[4]) {
    __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
    v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// MAXPOOL (nn.MaxPool2d(k, stride, pad); padding never wins) -- resnet.py:279 (3,2,1), unet.py:56 (2,2,0)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int C, int sp,
                               int Ho, int Wo, int dp, int k, int stride, int pad) {
  const int c4 = C / 4;
  const long long total = (long long)N * Ho * Wo * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    long long pix = i / c4;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int dy = 0; dy < k; ++dy) {
      const int iy = oy * stride - pad + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int ix = ox * stride - pad + dx;
        if (ix < 0 || ix >= W) continue;
        Pack4<T> pk;
        pk.load(src + (((long long)n * H + iy) * W + ix) * sp + c);
        float f[4];
        pk.get(f);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = fmaxf(m[j], f[j]);
      }
    }
    Pack4<T> o;
    o.set(m);
    o.store(dst + (((long long)n * Ho + oy) * Wo + ox) * dp + c);
  }
}

// split fp16 pairs: the maximum is taken on the reconstructed value hi + lo and the winning PAIR is copied
__global__ void maxpool_split_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N, int H, int W, int C,
                                     int sp, int slo, int Ho, int Wo, int dp, int dlo, int k, int stride, int pad,
                                     int f8, float lo_inv) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long pix = i / C;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float best = -INFINITY;
    __half bh = __float2half(0.f), bl = __float2half(0.f);
    uint8_t b8lo = 0, b8hi = 0;
    for (int dy = 0; dy < k; ++dy) {
      const int iy = oy * stride - pad + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int ix = ox * stride - pad + dx;
        if (ix < 0 || ix >= W) continue;
        const __half* q = src + (((long long)n * H + iy) * W + ix) * sp + c;
        if (f8) {   // the winning (hi, lo8, hi8) triple is copied (source and destination share the scales)
          const uint8_t* q8 = f8_block(q - c, slo, c);
          const float v = __half2float(q[0]) + e4m3_to_f32(q8[0]) * lo_inv;
          if (v > best) { best = v; bh = q[0]; b8lo = q8[0]; b8hi = q8[32]; }
          continue;
        }
        const __half h = q[0], l = q[slo];
        const float v = __half2float(h) + __half2float(l);
        if (v > best) { best = v; bh = h; bl = l; }
      }
    }
    __half* o = dst + (((long long)n * Ho + oy) * Wo + ox) * dp + c;
    o[0] = bh;
    if (f8) {
      uint8_t* o8 = f8_block(o - c, dlo, c);
      o8[0] = b8lo; o8[32] = b8hi;
      continue;
    }
    o[dlo] = bl;
  }
}

// CPN_DT_F16F8, C % 8 == 0: one thread owns 8 consecutive channels of one output pixel -- per window pixel one 16-byte
// load of the hi halves and one 8-byte load of the lo8 bytes (the hi8 bytes are fetched only for the eight winners)
__global__ void __launch_bounds__(256) maxpool_f16f8_v8_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                               int N, int H, int W, int c8, int sp, int slo, int Ho, int Wo,
                                                               int dp, int dlo, int k, int stride, int pad, float lo_inv) {
  const long long total = (long long)N * Ho * Wo * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    long long pix = i / c8;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float best[8];
    uint16_t bh[8];
    uint8_t bl[8];
    int bpos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bh[j] = 0; bl[j] = 0; bpos[j] = -1; }
    for (int dy = 0; dy < k; ++dy) {
      const int iy = oy * stride - pad + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int ix = ox * stride - pad + dx;
        if (ix < 0 || ix >= W) continue;
        const __half* q = src + (((long long)n * H + iy) * W + ix) * sp;
        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(q + c));
        const uint2 lv = __ldg(reinterpret_cast<const uint2*>(f8_block(q, slo, c)));
        const uint16_t* hh = reinterpret_cast<const uint16_t*>(&hv);
        const uint8_t* ll = reinterpret_cast<const uint8_t*>(&lv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = __half2float(__ushort_as_half(hh[j])) + e4m3_to_f32(ll[j]) * lo_inv;
          if (v > best[j]) { best[j] = v; bh[j] = hh[j]; bl[j] = ll[j]; bpos[j] = iy * W + ix; }
        }
      }
    }
    __half* o = dst + (((long long)n * Ho + oy) * Wo + ox) * dp;
    __align__(16) uint16_t oh[8];
    __align__(8) uint8_t ol[8], o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      oh[j] = bh[j]; ol[j] = bl[j];
      o8[j] = bpos[j] >= 0 ? f8_block(src + ((long long)n * H * W + bpos[j]) * sp, slo, c + j)[32] : (uint8_t)0;
    }
    *reinterpret_cast<uint4*>(o + c) = *reinterpret_cast<const uint4*>(oh);
    uint8_t* q8 = f8_block(o, dlo, c);
    *reinterpret_cast<uint2*>(q8) = *reinterpret_cast<const uint2*>(ol);
    *reinterpret_cast<uint2*>(q8 + 32) = *reinterpret_cast<const uint2*>(o8);
  }
}

int maxpool_launch(const cpn_op_t& op, const void* src, void* dst, cudaStream_t st) {
  if (op.src.dtype == CPN_DT_F16F8 && op.dst.dtype == CPN_DT_F16F8 && op.src.c == op.dst.c && op.src.c % 8 == 0 &&
      op.src.fp8_exp == op.dst.fp8_exp && op.src.pitch % 8 == 0 && op.dst.pitch % 8 == 0 && op.src.lo_delta % 32 == 0 &&
      op.dst.lo_delta % 32 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 &&
      (long long)op.src.h * op.src.w < (1ll << 31)) {
    const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.c / 8);
    maxpool_f16f8_v8_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h,
                                                                 op.src.w, op.src.c / 8, op.src.pitch, op.src.lo_delta,
                                                                 op.dst.h, op.dst.w, op.dst.pitch, op.dst.lo_delta, op.r,
                                                                 op.stride, op.pad, lo_fmt(op.src).lo_inv);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  if (dtype_has_lo(op.src.dtype)) {
    CPN_REQUIRE(op.dst.dtype == op.src.dtype && op.src.c == op.dst.c && op.src.fp8_exp == op.dst.fp8_exp,
                "maxpool: split dtype/channel/scale mismatch");
    const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * op.dst.c;
    maxpool_split_kernel<<<grid_for(total, 256), 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h,
                                                              op.src.w, op.src.c, op.src.pitch, op.src.lo_delta, op.dst.h,
                                                              op.dst.w, op.dst.pitch, op.dst.lo_delta, op.r, op.stride,
                                                              op.pad, op.src.dtype == CPN_DT_F16F8, lo_fmt(op.src).lo_inv);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  CPN_REQUIRE(op.src.c % 4 == 0 && op.src.pitch % 4 == 0 && op.dst.pitch % 4 == 0 && op.src.c == op.dst.c,
              "maxpool: channels/pitch must be multiples of 4");
  CPN_REQUIRE(op.src.dtype == op.dst.dtype, "maxpool: dtype mismatch");
  const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.c / 4);
  const int grid = grid_for(total, 256);
  if (op.src.dtype == CPN_DT_F32)
    maxpool_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, op.src.n, op.src.h, op.src.w, op.src.c,
                                                op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch, op.r, op.stride, op.pad);
  else
    maxpool_kernel<__half><<<grid, 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h, op.src.w,
                                                 op.src.c, op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch, op.r,
                                                 op.stride, op.pad);
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// UPSAMPLE: F.interpolate(mode='nearest') -- src index = floor(dst * in / out) (unet.py:215-217, torchvision FPN)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void upsample_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int C, int sp,
                                int Ho, int Wo, int dp) {
  const int c4 = C / 4;
  const long long total = (long long)N * Ho * Wo * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    long long pix = i / c4;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const int iy = (int)(((long long)oy * H) / Ho), ix = (int)(((long long)ox * W) / Wo);
    Pack4<T> pk;
    pk.load(src + (((long long)n * H + iy) * W + ix) * sp + c);
    pk.store(dst + (((long long)n * Ho + oy) * Wo + ox) * dp + c);
  }
}

// 16-byte variant (fp16, C % 8 == 0): one thread = one source chunk of 8 channels of one SOURCE pixel, replicated to
// the (up to fy x fx) destination pixels that map onto it (exact integer factors only) -- 1 load, fy*fx stores.
__global__ void __launch_bounds__(256) upsample_int_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N,
                                                           int H, int W, int c8, int sp8, int dp8, int fy, int fx) {
  const int total = N * H * W * c8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % c8;
    int pix = i / c8;
    const int x = pix % W;
    pix /= W;
    const int y = pix % H;
    const int n = pix / H;
    const uint4 v = __ldg(src + (size_t)((n * H + y) * W + x) * sp8 + c);
    const int Wo = W * fx, Ho = H * fy;
    for (int dy = 0; dy < fy; ++dy)
      for (int dx = 0; dx < fx; ++dx)
        dst[(size_t)((n * Ho + y * fy + dy) * Wo + x * fx + dx) * dp8 + c] = v;
  }
}

int upsample_launch(const cpn_op_t& op_in, const void* src, void* dst, cudaStream_t st) {
  if (dtype_has_lo(op_in.src.dtype)) {   // nearest copy is exact per block: run the fp16 kernel on the hi and the lo / 8-bit
                                         // planes (the latter as opaque 2-byte slots)
    CPN_REQUIRE(op_in.dst.dtype == op_in.src.dtype && op_in.src.fp8_exp == op_in.dst.fp8_exp, "upsample: split dtype mismatch");
    cpn_op_t o = op_in;
    o.src.dtype = o.dst.dtype = CPN_DT_F16;
    if (upsample_launch(o, src, dst, st)) return 1;
    return upsample_launch(o, reinterpret_cast<const __half*>(src) + op_in.src.lo_delta,
                           reinterpret_cast<__half*>(dst) + op_in.dst.lo_delta, st);
  }
  const cpn_op_t& op = op_in;
  CPN_REQUIRE(op.src.c % 4 == 0 && op.src.pitch % 4 == 0 && op.dst.pitch % 4 == 0 && op.src.c == op.dst.c,
              "upsample: channels/pitch must be multiples of 4");
  CPN_REQUIRE(op.src.dtype == op.dst.dtype, "upsample: dtype mismatch");
  const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.c / 4);
  const int grid = grid_for(total, 256);
  const bool integer = op.dst.h % op.src.h == 0 && op.dst.w % op.src.w == 0;
  if (op.src.dtype == CPN_DT_F16 && integer && op.src.c % 8 == 0 && op.src.pitch % 8 == 0 && op.dst.pitch % 8 == 0 &&
      (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 &&
      (long long)op.src.n * op.src.h * op.src.w * (op.src.c / 8) < (1ll << 31) &&
      (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.pitch / 8) < (1ll << 31)) {
    const long long items = (long long)op.src.n * op.src.h * op.src.w * (op.src.c / 8);
    upsample_int_kernel<<<grid_for(items, 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, op.src.n, op.src.h, op.src.w,
                                                            op.src.c / 8, op.src.pitch / 8, op.dst.pitch / 8,
                                                            op.dst.h / op.src.h, op.dst.w / op.src.w);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  if (op.src.dtype == CPN_DT_F32)
    upsample_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, op.src.n, op.src.h, op.src.w,
                                                 op.src.c, op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch);
  else
    upsample_kernel<__half><<<grid, 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h, op.src.w,
                                                  op.src.c, op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch);
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// BILINEAR: F.interpolate(mode='bilinear', align_corners=False) (models/cpn.py:109-115): half-pixel centres,
// src = (dst + 0.5) * in/out - 0.5 clamped at 0, upper neighbour clamped to in-1 (ATen area_pixel_compute_source_index).
// ---------------------------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void bilinear_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int N, int H, int W, int C, int sp,
                                int Ho, int Wo, int dp) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long pix = i / C;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const TI* b = src + (long long)n * H * W * sp + c;
    const float v00 = to_f32<TI>(b[((long long)y0 * W + x0) * sp]), v01 = to_f32<TI>(b[((long long)y0 * W + x1) * sp]);
    const float v10 = to_f32<TI>(b[((long long)y1 * W + x0) * sp]), v11 = to_f32<TI>(b[((long long)y1 * W + x1) * sp]);
    const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    dst[(((long long)n * Ho + oy) * Wo + ox) * dp + c] = from_f32<TO>(v);
  }
}

// fp16 -> fp16, C % 8 == 0: one thread blends 8 channels (four 16-byte loads, one 16-byte store)
__global__ void __launch_bounds__(256) bilinear_h8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N,
                                                          int H, int W, int c8, int sp8, int Ho, int Wo, int dp8) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)N * Ho * Wo * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    long long pix = i / c8;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const uint4* b = src + (size_t)n * H * W * sp8 + c;
    const uint4 q00 = __ldg(b + ((size_t)y0 * W + x0) * sp8), q01 = __ldg(b + ((size_t)y0 * W + x1) * sp8);
    const uint4 q10 = __ldg(b + ((size_t)y1 * W + x0) * sp8), q11 = __ldg(b + ((size_t)y1 * W + x1) * sp8);
    const __half2* h00 = reinterpret_cast<const __half2*>(&q00);
    const __half2* h01 = reinterpret_cast<const __half2*>(&q01);
    const __half2* h10 = reinterpret_cast<const __half2*>(&q10);
    const __half2* h11 = reinterpret_cast<const __half2*>(&q11);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h00[j]), bb = __half22float2(h01[j]);
      const float2 cc = __half22float2(h10[j]), d = __half22float2(h11[j]);
      const float vx = hy * (hx * a.x + lx * bb.x) + ly * (hx * cc.x + lx * d.x);
      const float vy = hy * (hx * a.y + lx * bb.y) + ly * (hx * cc.y + lx * d.y);
      ho[j] = __floats2half2_rn(vx, vy);
    }
    dst[(((size_t)n * Ho + oy) * Wo + ox) * dp8 + c] = o;
  }
}

// split fp16 pairs: blend the reconstructed values in fp32, split the result again
__global__ void bilinear_split_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N, int H, int W, int C,
                                      int sp, int slo, int Ho, int Wo, int dp, int dlo, const LoFmt FS, const LoFmt FD) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long pix = i / C;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const __half* b = src + (long long)n * H * W * sp + c;
    auto val = [&](int y, int x) {
      const __half* q = b + ((long long)y * W + x) * sp;
      if (FS.f8) return f16f8_load(q - c, slo, c, FS.lo_inv);
      return __half2float(q[0]) + __half2float(q[slo]);
    };
    const float v = hy * (hx * val(y0, x0) + lx * val(y0, x1)) + ly * (hx * val(y1, x0) + lx * val(y1, x1));
    __half* o = dst + (((long long)n * Ho + oy) * Wo + ox) * dp + c;
    if (FD.f8) f16f8_store(o - c, dlo, c, v, FD.lo_scale, FD.hi8_scale);
    else split_store(o, dlo, v);
  }
}

// fp16 + e4m3 (CPN_DT_F16F8), C % 8 == 0: one thread blends 8 channels of one output pixel -- per tap one 16-byte load of the
// fp16 values and one 8-byte load of their lo8 residual bytes; one 16-byte + two 8-byte stores (f16f8_store8).  Same
// arithmetic and association as bilinear_split_kernel (reconstruct in fp32, blend, split again), 8 x fewer memory instructions.
__global__ void __launch_bounds__(256) bilinear_f8v8_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int N,
                                                            int H, int W, int C, int sp, int Ho, int Wo, int dp,
                                                            const LoFmt FS, const LoFmt FD) {
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const int c8 = C / 8;
  const long long total = (long long)N * Ho * Wo * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8) * 8;
    long long pix = i / c8;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const __half* b = src + (long long)n * H * W * sp;
    auto tap = [&](int y, int x, float (&o)[8]) {
      const __half* px = b + ((long long)y * W + x) * sp;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(px + c0));
      const uint2 l8 = __ldg(reinterpret_cast<const uint2*>(f8_block(px, FS.lo_delta, c0)));
      const __half2* h = reinterpret_cast<const __half2*>(&q);
      float lo[8];
      unpack_e4m3x4(l8.x, lo[0], lo[1], lo[2], lo[3]);
      unpack_e4m3x4(l8.y, lo[4], lo[5], lo[6], lo[7]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        o[2 * j] = f.x + lo[2 * j] * FS.lo_inv;
        o[2 * j + 1] = f.y + lo[2 * j + 1] * FS.lo_inv;
      }
    };
    float a[8], bb[8], v[8];
    tap(y0, x0, a); tap(y0, x1, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = hy * (hx * a[j] + lx * bb[j]);
    tap(y1, x0, a); tap(y1, x1, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] + ly * (hx * a[j] + lx * bb[j]);     // the scalar kernel's association
    f16f8_store8(dst + (((long long)n * Ho + oy) * Wo + ox) * dp, FD, c0, v);
  }
}

int bilinear_launch(const cpn_op_t& op, const void* src, void* dst, cudaStream_t st) {
  CPN_REQUIRE(op.src.c == op.dst.c, "bilinear: channel mismatch");
  if (dtype_has_lo(op.src.dtype)) {
    CPN_REQUIRE(op.dst.dtype == op.src.dtype, "bilinear: split dtype mismatch");
    if (op.src.dtype == CPN_DT_F16F8 && op.src.c % 8 == 0 && op.src.pitch % 8 == 0 && op.dst.pitch % 8 == 0 &&
        (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
      const long long tot8 = (long long)op.dst.n * op.dst.h * op.dst.w * (op.dst.c / 8);
      bilinear_f8v8_kernel<<<grid_for(tot8, 256), 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h, op.src.w,
                                                                op.src.c, op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch,
                                                                lo_fmt(op.src), lo_fmt(op.dst));
      CPN_CHECK_LAUNCH();
      return 0;
    }
    const long long tot = (long long)op.dst.n * op.dst.h * op.dst.w * op.dst.c;
    bilinear_split_kernel<<<grid_for(tot, 256), 256, 0, st>>>((const __half*)src, (__half*)dst, op.src.n, op.src.h,
                                                             op.src.w, op.src.c, op.src.pitch, op.src.lo_delta, op.dst.h,
                                                             op.dst.w, op.dst.pitch, op.dst.lo_delta, lo_fmt(op.src),
                                                             lo_fmt(op.dst));
    CPN_CHECK_LAUNCH();
    return 0;
  }
  const long long total = (long long)op.dst.n * op.dst.h * op.dst.w * op.dst.c;
  const int grid = grid_for(total, 256);
  if (op.src.dtype == CPN_DT_F16 && op.dst.dtype == CPN_DT_F16 && op.src.c % 8 == 0 && op.src.pitch % 8 == 0 &&
      op.dst.pitch % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0) {
    bilinear_h8_kernel<<<grid_for(total / 8, 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, op.src.n, op.src.h,
                                                               op.src.w, op.src.c / 8, op.src.pitch / 8, op.dst.h,
                                                               op.dst.w, op.dst.pitch / 8);
    CPN_CHECK_LAUNCH();
    return 0;
  }
#define CPN_BIL(TI, TO)                                                                                             \
  bilinear_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, (TO*)dst, op.src.n, op.src.h, op.src.w, op.src.c, \
                                                op.src.pitch, op.dst.h, op.dst.w, op.dst.pitch)
  if (op.src.dtype == CPN_DT_F32 && op.dst.dtype == CPN_DT_F32) CPN_BIL(float, float);
  else if (op.src.dtype == CPN_DT_F16 && op.dst.dtype == CPN_DT_F16) CPN_BIL(__half, __half);
  else if (op.src.dtype == CPN_DT_F16 && op.dst.dtype == CPN_DT_F32) CPN_BIL(__half, float);
  else { set_error("bilinear: unsupported dtypes"); return 1; }
#undef CPN_BIL
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// PROJ: ReadOut's final 1x1 convolution (commons.py:499) + final activation (ScaledTanh for the refinement head,
// cpn.py:230) to fp32 pixel records.  One warp per pixel: lanes split the input channels (coalesced 8/16-byte loads),
// weights live in shared memory, warp-shuffle reduction per output channel.
//   dst[pixel * dp + j] = act( bias[j] + sum_c src[pixel * sp + cin_off + c] * w[j][c] ),  j < cout <= 32
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) proj_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                   const float* __restrict__ wgt, const float* __restrict__ bias,
                                                   long long pixels, int sp, int cin_off, int cin, int cout, int dp,
                                                   int act, float act_scale, int lo_delta, int f8, float lo_inv) {
  extern __shared__ float wsm[];  // [cout][cin]
  for (int i = threadIdx.x; i < cout * cin; i += blockDim.x) wsm[i] = wgt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (long long pix = (long long)blockIdx.x * wpb + warp; pix < pixels; pix += (long long)gridDim.x * wpb) {
    const T* sp_ = src + pix * sp + cin_off;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int cc = lane * 4; cc < (lo_delta > 0 ? 2 * cin : cin); cc += 128) {
      const int c = cc < cin ? cc : cc - cin;            // split sources: second sweep adds the lo halves
      float f[4];
      if (sizeof(T) == 2 && f8 && cc >= cin) {           // CPN_DT_F16F8: four lo8 bytes of channels cin_off + c ..
        const uint8_t* q = f8_block(reinterpret_cast<const __half*>(src) + pix * sp, lo_delta, cin_off + c);
        unpack_e4m3x4(*reinterpret_cast<const uint32_t*>(q), f[0], f[1], f[2], f[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] *= lo_inv;
      } else {
        Pack4<T> pk;
        pk.load(sp_ + c + (cc < cin ? 0 : lo_delta));
        pk.get(f);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < cout) {
          const float* w = wsm + j * cin + c;
          acc[j] = fmaf(f[0], w[0], acc[j]);
          acc[j] = fmaf(f[1], w[1], acc[j]);
          acc[j] = fmaf(f[2], w[2], acc[j]);
          acc[j] = fmaf(f[3], w[3], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < cout) {
        float v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[j] = v;
      }
    }
    // lane j writes output channel j (coalesced record write)
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j == lane) mine = acc[j];
    if (lane < cout) {
      float v = mine + (bias ? bias[lane] : 0.f);
      if (act == CPN_ACT_SCALED_TANH) v = tanhf(v) * act_scale;
      else if (act == CPN_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == CPN_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      dst[pix * dp + lane] = v;
    }
  }
}

int proj_launch(const cpn_op_t& op, const void* src, void* dst, const float* wgt, const float* bias, cudaStream_t st) {
  CPN_REQUIRE(op.dst.dtype == CPN_DT_F32, "proj: fp32 output required");
  CPN_REQUIRE(op.dst.c >= 1 && op.dst.c <= 32, "proj: cout %d must be in [1, 32]", op.dst.c);
  CPN_REQUIRE(op.proj_cin % 4 == 0 && op.proj_cin_off % 4 == 0 && op.src.pitch % 4 == 0,
              "proj: cin/cin_off/pitch must be multiples of 4");
  const long long pixels = (long long)op.src.n * op.src.h * op.src.w;
  const size_t smem = (size_t)op.dst.c * op.proj_cin * sizeof(float);
  CPN_REQUIRE(smem <= 48 * 1024, "proj: weights (%zu B) exceed 48 KB of shared memory", smem);
  long long blocks = (pixels + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (op.src.dtype == CPN_DT_F32)
    proj_kernel<float><<<(int)blocks, 256, smem, st>>>((const float*)src, (float*)dst, wgt, bias, pixels, op.src.pitch,
                                                       op.proj_cin_off, op.proj_cin, op.dst.c, op.dst.pitch, op.act,
                                                       op.act_scale, 0, 0, 1.f);
  else
    proj_kernel<__half><<<(int)blocks, 256, smem, st>>>((const __half*)src, (float*)dst, wgt, bias, pixels,
                                                        op.src.pitch, op.proj_cin_off, op.proj_cin, op.dst.c,
                                                        op.dst.pitch, op.act, op.act_scale,
                                                        dtype_has_lo(op.src.dtype) ? op.src.lo_delta : 0,
                                                        op.src.dtype == CPN_DT_F16F8, lo_fmt(op.src).lo_inv);
  CPN_CHECK_LAUNCH();
  return 0;
}

}  // namespace cpn

// ---------------------------------------------------------------------------------------------------------------------
// GATHER PATCHES: the k x k x C neighbourhood of P selected pixels of an NHWC feature map as rows of a [P, k*k*C] matrix
// (same element format as the source; zero outside the image = the convolution's zero padding), channel order
// [C/64 blocks][k*k taps][64 channels] -- the K order in which the dense k x k convolution contracts, so that a 1x1
// convolution over these rows reproduces it term by term.  Sparse evaluation of the location / fourier ReadOut heads:
// the reference computes them on every pixel (models/cpn.py:253-263) but reads them only where the score selects a
// proposal (:620-623).  One thread moves 16 bytes; 128 contiguous bytes (64 fp16 values, or the 64 channels' 8-bit chunks
// / lo halves) per (pixel, block, tap).
// ---------------------------------------------------------------------------------------------------------------------
namespace cpn {
__global__ void __launch_bounds__(256) gather_patches_kernel(const __half* __restrict__ src, const int32_t* __restrict__ idx,
                                                             long long P, int H, int W, int sp, int slo, int cblocks, int k,
                                                             int pad, __half* __restrict__ dst, int dp, int dlo, int has_lo) {
  const int taps = k * k;
  const long long per_row = (long long)cblocks * taps * (has_lo ? 2 : 1) * 8;      // 16-byte items per gathered row
  const long long total = P * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pr = i / per_row;
    int rem = (int)(i - pr * per_row);
    const int seg = rem & 7;
    rem >>= 3;
    const int blk = has_lo ? (rem & 1) : 0;          // 0: fp16 values, 1: second block (8-bit chunks / lo halves)
    if (has_lo) rem >>= 1;
    const int tap = rem % taps, cb = rem / taps;
    const long long pix = idx[pr];
    const long long hw = (long long)H * W;
    const int b = (int)(pix / hw);
    const int r2 = (int)(pix - (long long)b * hw);
    const int y = r2 / W + tap / k - pad, x = r2 % W + tap % k - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = __ldg(reinterpret_cast<const uint4*>(src + ((long long)b * hw + (long long)y * W + x) * sp + (blk ? slo : 0) + cb * 64) + seg);
    reinterpret_cast<uint4*>(dst + pr * dp + (blk ? dlo : 0) + (long long)(cb * taps + tap) * 64)[seg] = v;
  }
}
}  // namespace cpn

extern "C" int cpn_gather_patches(const void* src, const cpn_view_t* src_view_host, const int32_t* idx, int64_t P, int k,
                                  void* dst, const cpn_view_t* dst_view_host, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(src && src_view_host && dst && dst_view_host && k >= 1 && (k & 1), "gather_patches: bad arguments");
  const cpn_view_t& sv = *src_view_host;
  const cpn_view_t& dv = *dst_view_host;
  CPN_REQUIRE(sv.dtype == dv.dtype && (sv.dtype == CPN_DT_F16 || dtype_has_lo(sv.dtype)), "gather_patches: fp16-family views required");
  CPN_REQUIRE(sv.c % 64 == 0 && dv.c == sv.c * k * k && sv.pitch % 8 == 0 && dv.pitch % 8 == 0 && sv.lo_delta % 8 == 0 &&
                  dv.lo_delta % 8 == 0 && (long long)dv.n * dv.h * dv.w >= P,
              "gather_patches: channel / pitch mismatch (src c %d, dst c %d, k %d)", sv.c, dv.c, k);
  CPN_REQUIRE(sv.dtype != CPN_DT_F16F8 || sv.fp8_exp == dv.fp8_exp, "gather_patches: fp8 scales must match");
  if (P <= 0) return 0;
  const int has_lo = dtype_has_lo(sv.dtype) ? 1 : 0;
  const long long items = P * (long long)(sv.c / 64) * k * k * (has_lo ? 2 : 1) * 8;
  gather_patches_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half*>(reinterpret_cast<const char*>(src) + sv.offset), idx, P, sv.h, sv.w, sv.pitch,
      sv.lo_delta, sv.c / 64, k, k / 2, reinterpret_cast<__half*>(reinterpret_cast<char*>(dst) + dv.offset), dv.pitch,
      dv.lo_delta, has_lo);
  CPN_CHECK_LAUNCH();
  return 0;
}

// F.interpolate(x, size, mode='bilinear', align_corners=False) on a stand-alone fp32 NHWC tensor: the resize of the
// score bounds in _apply_score_bounds / _equal_size (models/cpn.py:109-123; c == 1 makes NHWC and NCHW coincide).
extern "C" int cpn_resize_bilinear(const float* src, int n, int h, int w, int c, float* dst, int ho, int wo,
                                   void* stream) {
  CPN_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && ho > 0 && wo > 0, "resize_bilinear: bad arguments");
  cpn_op_t op;
  memset(&op, 0, sizeof(op));
  op.kind = CPN_OP_BILINEAR;
  op.src.n = op.dst.n = n; op.src.c = op.dst.c = c; op.src.pitch = op.dst.pitch = c;
  op.src.h = h; op.src.w = w; op.dst.h = ho; op.dst.w = wo;
  op.src.dtype = op.dst.dtype = CPN_DT_F32;
  return cpn::bilinear_launch(op, src, dst, (cudaStream_t)stream);
}


// ---------------------------------------------------------------------------------------------------------------------
// PREPROCESS (celldetection_scripts/cpn_inference.py:196-222, cd.data.normalize_percentile data/misc.py:156-161): the
// input-side conditioning of a whole slide before tiling.  Integer work, HBM-bound: an exact histogram (the host derives
// np.percentile's order statistics, the image mean and every look-up table from it) and ONE table look-up per element that
// composes percentile normalisation -> uint8, gamma and brightness / contrast.
// ---------------------------------------------------------------------------------------------------------------------
namespace cpn {
// uint8: every warp owns a private 256-bin histogram in shared memory (no inter-warp contention), 16 values per 16-byte load;
// the block's eight copies are flushed with one global atomic per non-empty bin.
__global__ void __launch_bounds__(256) histogram_u8_kernel(const uint8_t* __restrict__ data, long long n,
                                                           unsigned int* __restrict__ hist) {
  __shared__ unsigned int h[8][256];
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  unsigned int* mine = h[threadIdx.x >> 5];
  const long long nvec = (reinterpret_cast<uintptr_t>(data) & 15) ? 0 : n / 16;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uint4* v4 = reinterpret_cast<const uint4*>(data);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 q = __ldg(v4 + i);
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(mine + (w[k] & 255u), 1u);
      atomicAdd(mine + ((w[k] >> 8) & 255u), 1u);
      atomicAdd(mine + ((w[k] >> 16) & 255u), 1u);
      atomicAdd(mine + (w[k] >> 24), 1u);
    }
  }
  for (long long i = nvec * 16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(mine + data[i], 1u);
  __syncthreads();
  unsigned int total = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) total += h[k][threadIdx.x];
  if (total) atomicAdd(hist + threadIdx.x, total);
}

// uint16: the block counts into 65 536 PACKED 16-bit counters in shared memory (128 KB, two bins per word) and flushes them to
// the global histogram after every round of at most 57 344 values, so no packed counter can overflow; the flush touches
// global memory only for the bins the round actually hit (a microscopy image concentrates on a few hundred).
constexpr int H16_THREADS = 1024;
constexpr int H16_ROUND_VECS = 7;          // 7 x 1024 threads x 8 values = 57 344 <= 65 535
__global__ void __launch_bounds__(H16_THREADS, 1) histogram_u16_kernel(const uint16_t* __restrict__ data, long long n,
                                                                      unsigned int* __restrict__ hist) {
  extern __shared__ unsigned int cnt[];    // [32768]: bin v -> half (v & 1) of word v >> 1
  for (int i = threadIdx.x; i < 32768; i += H16_THREADS) cnt[i] = 0;
  __syncthreads();
  const bool aligned = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
  const long long nvec = aligned ? n / 8 : 0;
  const uint4* v4 = reinterpret_cast<const uint4*>(data);
  const long long per_round = (long long)H16_ROUND_VECS * H16_THREADS;               // vectors per block and round
  const long long rounds = (nvec + per_round - 1) / per_round;
  auto flush = [&]() {
    __syncthreads();
    for (int i = threadIdx.x; i < 32768; i += H16_THREADS) {
      const unsigned int w = cnt[i];
      if (w) {
        if (w & 0xffffu) atomicAdd(hist + 2 * i, w & 0xffffu);
        if (w >> 16) atomicAdd(hist + 2 * i + 1, w >> 16);
        cnt[i] = 0;
      }
    }
    __syncthreads();
  };
  for (long long r = blockIdx.x; r < rounds; r += gridDim.x) {
    const long long base = r * per_round;
#pragma unroll
    for (int k = 0; k < H16_ROUND_VECS; ++k) {
      const long long i = base + (long long)k * H16_THREADS + threadIdx.x;
      if (i < nvec) {
        const uint4 q = __ldg(v4 + i);
        const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned int lo = w[j] & 0xffffu, hi = w[j] >> 16;
          atomicAdd(cnt + (lo >> 1), 1u << ((lo & 1) * 16));
          atomicAdd(cnt + (hi >> 1), 1u << ((hi & 1) * 16));
        }
      }
    }
    flush();
  }
  // scalar tail (and unaligned inputs): rounds of 57 344 values, one value per thread and step
  const long long tail0 = nvec * 8, n_tail = n - tail0, tail_round = per_round * 8;
  const long long tail_rounds = (n_tail + tail_round - 1) / tail_round;
  for (long long r = blockIdx.x; r < tail_rounds; r += gridDim.x) {
    const long long lo_i = tail0 + r * tail_round, hi_i = lo_i + tail_round < n ? lo_i + tail_round : n;
    for (long long i = lo_i + threadIdx.x; i < hi_i; i += H16_THREADS) {
      const unsigned int v = data[i];
      atomicAdd(cnt + (v >> 1), 1u << ((v & 1) * 16));
    }
    flush();
  }
}

// dst[i] = lut[src[i]] with the table in shared memory; 16 input bytes per load (16 uint8 / 8 uint16 values), packed stores.
template <typename T>
__global__ void __launch_bounds__(256) lut_kernel(const T* __restrict__ src, long long n, const uint8_t* __restrict__ lut,
                                                  int lut_size, uint8_t* __restrict__ dst) {
  extern __shared__ uint8_t lut_s[];
  for (int i = threadIdx.x * 4; i < lut_size; i += blockDim.x * 4)
    *reinterpret_cast<unsigned int*>(lut_s + i) = __ldg(reinterpret_cast<const unsigned int*>(lut + i));
  __syncthreads();
  constexpr int PER = 16 / (int)sizeof(T);
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const long long nvec = aligned ? n / PER : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uint4* v4 = reinterpret_cast<const uint4*>(src);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 q = __ldg(v4 + i);
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
    if (sizeof(T) == 1) {
      unsigned int o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = (unsigned int)lut_s[w[k] & 255u] | ((unsigned int)lut_s[(w[k] >> 8) & 255u] << 8) |
               ((unsigned int)lut_s[(w[k] >> 16) & 255u] << 16) | ((unsigned int)lut_s[w[k] >> 24] << 24);
      reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      unsigned int o[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        o[k] = (unsigned int)lut_s[w[2 * k] & 0xffffu] | ((unsigned int)lut_s[w[2 * k] >> 16] << 8) |
               ((unsigned int)lut_s[w[2 * k + 1] & 0xffffu] << 16) | ((unsigned int)lut_s[w[2 * k + 1] >> 16] << 24);
      reinterpret_cast<uint2*>(dst)[i] = make_uint2(o[0], o[1]);
    }
  }
  for (long long i = nvec * PER + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = lut_s[src[i]];
}
// cv2.COLOR_RGB2GRAY / COLOR_RGBA2GRAY for 8-bit images (OpenCV 4.x: 15-bit fixed point, pinned against the installed cv2 on
// all 2^24 colours), applied AFTER an optional per-element table (the percentile normalisation that precedes it in the
// reference's chain, cpn_inference.py:198-213): dst[p] = (9798 R + 19235 G + 3735 B + 2^14) >> 15.
template <typename T>
__global__ void __launch_bounds__(256) gray_kernel(const T* __restrict__ src, long long n_px, int channels,
                                                   const uint8_t* __restrict__ lut, int lut_size, uint8_t* __restrict__ dst) {
  extern __shared__ uint8_t lut_s[];
  if (lut) {
    for (int i = threadIdx.x * 4; i < lut_size; i += blockDim.x * 4)
      *reinterpret_cast<unsigned int*>(lut_s + i) = __ldg(reinterpret_cast<const unsigned int*>(lut + i));
    __syncthreads();
  }
  auto mix = [&](unsigned int r, unsigned int g, unsigned int b) -> unsigned int {
    if (lut) { r = lut_s[r]; g = lut_s[g]; b = lut_s[b]; }
    return (9798u * r + 19235u * g + 3735u * b + (1u << 14)) >> 15;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long done = 0;
  if (sizeof(T) == 1 && (reinterpret_cast<uintptr_t>(src) & (channels == 4 ? 15 : 3)) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
    // four pixels per thread: 3 (RGB) or 4 (RGBA) aligned 32-bit loads, one 32-bit store
    const long long quads = n_px / 4;
    const unsigned int* w32 = reinterpret_cast<const unsigned int*>(src);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += stride) {
      unsigned int o;
      if (channels == 3) {
        const unsigned int a = __ldg(w32 + 3 * q), b = __ldg(w32 + 3 * q + 1), c = __ldg(w32 + 3 * q + 2);
        o = mix(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u) | (mix(a >> 24, b & 255u, (b >> 8) & 255u) << 8) |
            (mix((b >> 16) & 255u, b >> 24, c & 255u) << 16) | (mix((c >> 8) & 255u, (c >> 16) & 255u, c >> 24) << 24);
      } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + q);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
        o = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) o |= mix(w[k] & 255u, (w[k] >> 8) & 255u, (w[k] >> 16) & 255u) << (8 * k);
      }
      reinterpret_cast<unsigned int*>(dst)[q] = o;
    }
    done = quads * 4;
  }
  for (long long i = done + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_px; i += stride) {
    const T* px = src + i * channels;
    dst[i] = (uint8_t)mix(px[0], px[1], px[2]);
  }
}
}  // namespace cpn

extern "C" int cpn_histogram(const void* data, int dtype, int64_t n, uint32_t* hist, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(hist && (data || n == 0) && n >= 0 && (dtype == CPN_DT_U8 || dtype == CPN_DT_U16), "histogram: uint8 / uint16 data required");
  cudaStream_t st = (cudaStream_t)stream;
  const int bins = dtype == CPN_DT_U8 ? 256 : 65536;
  CPN_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)bins * sizeof(uint32_t), st));
  if (n == 0) return 0;
  if (dtype == CPN_DT_U8) histogram_u8_kernel<<<grid_for(n, 256 * 64), 256, 0, st>>>((const uint8_t*)data, n, hist);
  else {
    static bool attr = false;
    if (!attr) {
      CPN_CHECK_CUDA(cudaFuncSetAttribute(histogram_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
      attr = true;
    }
    long long blocks = (n + 57343) / 57344;
    if (blocks > sm_count()) blocks = sm_count();
    histogram_u16_kernel<<<(unsigned)blocks, H16_THREADS, 131072, st>>>((const uint16_t*)data, n, hist);
  }
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_apply_lut(const void* src, int dtype, int64_t n, const uint8_t* lut, uint8_t* dst, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(n >= 0 && (dtype == CPN_DT_U8 || dtype == CPN_DT_U16), "apply_lut: uint8 / uint16 source required");
  if (n == 0) return 0;
  CPN_REQUIRE(src && lut && dst, "apply_lut: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CPN_DT_U8) {
    lut_kernel<uint8_t><<<grid_for(n, 256 * 16), 256, 256, st>>>((const uint8_t*)src, n, lut, 256, dst);
  } else {
    static bool attr = false;
    if (!attr) {
      CPN_CHECK_CUDA(cudaFuncSetAttribute(lut_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      attr = true;
    }
    lut_kernel<uint16_t><<<grid_for(n, 256 * 64), 256, 65536, st>>>((const uint16_t*)src, n, lut, 65536, dst);
  }
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_rgb2gray(const void* src, int dtype, int64_t n_px, int channels, const uint8_t* lut, uint8_t* dst,
                            void* stream) {
  using namespace cpn;
  CPN_REQUIRE((n_px == 0 || (src && dst)) && n_px >= 0 && (channels == 3 || channels == 4), "rgb2gray: 3 or 4 interleaved channels required");
  CPN_REQUIRE(dtype == CPN_DT_U8 || (dtype == CPN_DT_U16 && lut), "rgb2gray: uint8 data, or uint16 data with a table to uint8");
  if (n_px == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n_px, 256);
  if (dtype == CPN_DT_U8) {
    gray_kernel<uint8_t><<<grid, 256, lut ? 256 : 0, st>>>((const uint8_t*)src, n_px, channels, lut, 256, dst);
  } else {
    static bool attr = false;
    if (!attr) {
      CPN_CHECK_CUDA(cudaFuncSetAttribute(gray_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      attr = true;
    }
    gray_kernel<uint16_t><<<grid_for(n_px, 256 * 8), 256, 65536, st>>>((const uint16_t*)src, n_px, channels, lut, 65536, dst);
  }
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Helpers of the phase-decomposed refinement head (bilinear x2 followed by a k x k convolution evaluated as four phase
// convolutions on the low-resolution map, models/cpn.py:274-279): the phase-packed records [N, h, w, 4 * c] are shuffled
// into the full-resolution map [N, 2h, 2w, c], and the image-border strips -- where bilinear clamping and the
// convolution's zero padding make the phase identity inexact -- are cropped, recomputed by the plain path on small tensors
// and pasted back.  Byte movers, HBM-bound.
// ---------------------------------------------------------------------------------------------------------------------
namespace cpn {
__global__ void __launch_bounds__(256) copy_window_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N,
                                                          int Hs, int Ws, int Hd, int Wd, int vec_per_px, int ys, int xs,
                                                          int yd, int xd, int hh, int ww) {
  const long long total = (long long)N * hh * ww * vec_per_px;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec_per_px);
    long long t = i / vec_per_px;
    const int x = (int)(t % ww); t /= ww;
    const int y = (int)(t % hh);
    const int n = (int)(t / hh);
    dst[(((long long)n * Hd + yd + y) * Wd + xd + x) * vec_per_px + v] =
        __ldg(src + (((long long)n * Hs + ys + y) * Ws + xs + x) * vec_per_px + v);
  }
}

__global__ void __launch_bounds__(256) copy_window4_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int Hs,
                                                           int Ws, int Hd, int Wd, int words_per_px, int ys, int xs, int yd,
                                                           int xd, int hh, int ww) {
  const long long total = (long long)N * hh * ww * words_per_px;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % words_per_px);
    long long t = i / words_per_px;
    const int x = (int)(t % ww); t /= ww;
    const int y = (int)(t % hh);
    const int n = (int)(t / hh);
    dst[(((long long)n * Hd + yd + y) * Wd + xd + x) * words_per_px + v] =
        __ldg(src + (((long long)n * Hs + ys + y) * Ws + xs + x) * words_per_px + v);
  }
}

// rec [N, h, w, 4 * c] (phase (a, b) at channels (2a + b) * c ...) -> out [N, 2h, 2w, c]: out[n, 2y + a, 2x + b, :]
__global__ void __launch_bounds__(256) unshuffle2_kernel(const float* __restrict__ rec, float* __restrict__ out, int N, int h,
                                                         int w, int c) {
  const long long total = (long long)N * h * w * 4 * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int ph = (int)(t & 3); t >>= 2;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h);
    const int n = (int)(t / h);
    out[(((long long)n * 2 * h + 2 * y + (ph >> 1)) * 2 * w + 2 * x + (ph & 1)) * c + ch] = __ldg(rec + i);
  }
}
}  // namespace cpn

extern "C" int cpn_copy_window(const void* src, void* dst, int N, int Hs, int Ws, int Hd, int Wd, int bytes_per_px, int ys,
                               int xs, int yd, int xd, int hh, int ww, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(src && dst && N >= 0 && hh >= 0 && ww >= 0 && bytes_per_px > 0 && bytes_per_px % 4 == 0, "copy_window: bad arguments");
  CPN_REQUIRE(ys >= 0 && xs >= 0 && yd >= 0 && xd >= 0 && ys + hh <= Hs && xs + ww <= Ws && yd + hh <= Hd && xd + ww <= Wd,
              "copy_window: window outside the source or the destination");
  const long long px = (long long)N * hh * ww;
  if (px == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (bytes_per_px % 16 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    copy_window_kernel<<<grid_for(px * (bytes_per_px / 16), 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, N, Hs, Ws, Hd, Wd,
                                                                             bytes_per_px / 16, ys, xs, yd, xd, hh, ww);
  } else {
    copy_window4_kernel<<<grid_for(px * (bytes_per_px / 4), 256), 256, 0, st>>>((const float*)src, (float*)dst, N, Hs, Ws, Hd, Wd,
                                                                             bytes_per_px / 4, ys, xs, yd, xd, hh, ww);
  }
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_unshuffle2(const float* rec, int N, int h, int w, int c, float* out, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(rec && out && N >= 0 && h >= 0 && w >= 0 && c >= 1, "unshuffle2: bad arguments");
  const long long total = (long long)N * h * w * 4 * c;
  if (total == 0) return 0;
  unshuffle2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(rec, out, N, h, w, c);
  CPN_CHECK_LAUNCH();
  return 0;
}
