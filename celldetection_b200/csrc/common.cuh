// Shared helpers for the cpn_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/cpn_b200.h"

namespace cpn {

// Thread-local error string + global launch counter (cpn_last_error / cpn_launch_count).
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CPN_CHECK_CUDA(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      cpn::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

#define CPN_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      cpn::set_error(__VA_ARGS__);   \
      return 1;                      \
    }                                \
  } while (0)

#define CPN_CHECK_LAUNCH()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      cpn::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,                  \
                     cudaGetErrorString(_e));                                                \
      return 1;                                                                              \
    }                                                                                        \
    cpn::count_launch();                                                                     \
  } while (0)

inline int dtype_size(int dt) {
  return dt == CPN_DT_F32 ? 4 : ((dt == CPN_DT_F16 || dt == CPN_DT_F16X2 || dt == CPN_DT_F16F8) ? 2 : 1);
}
inline bool dtype_has_lo(int dt) { return dt == CPN_DT_F16X2 || dt == CPN_DT_F16F8; }   // second block per pixel
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();  // cached SM count of the current device

// ---- engine entry points (one per translation unit) ----------------------------------------------------------------
struct ConvTcPlan;  // opaque, conv_tc.cu
int conv_simt_launch(const cpn_op_t& op, const void* src, void* dst, const void* res, const void* wgt,
                     const float* bias, cudaStream_t st);
int conv_tc_plan_create(const cpn_op_t& op, const void* src, void* dst, const void* res, const void* wgt,
                        const float* bias, ConvTcPlan** out);
int conv_tc_launch(const ConvTcPlan* p, cudaStream_t st);
int conv_tc_fuse_proj(ConvTcPlan* p, int n, const cpn_op_t* projs, const char* weights);
void conv_tc_bind_proj_out(ConvTcPlan* p, int head, void* out);
int conv_tc_limit_rows(ConvTcPlan* p, long long rows);
void conv_tc_plan_destroy(ConvTcPlan* p);

int prep_launch(const cpn_op_t& op, const void* input, int input_format, void* dst, int32_t* flags, cudaStream_t st);
int maxpool_launch(const cpn_op_t& op, const void* src, void* dst, cudaStream_t st);
int upsample_launch(const cpn_op_t& op, const void* src, void* dst, cudaStream_t st);
int bilinear_launch(const cpn_op_t& op, const void* src, void* dst, cudaStream_t st);
int proj_launch(const cpn_op_t& op, const void* src, void* dst, const float* wgt, const float* bias, cudaStream_t st);

// ---- uniform spatial grid over box centres (nms.cu): shared by the stitch NMS, box voting and label rasterisation ----
struct GridInfo {
  float x0, y0, inv_cell;
  int ncx, ncy;
};
struct GridBins {
  const uint64_t* cell_keys;  // [n] (cell << 32 | row), ascending
  const int32_t* cell_rows;   // [n] row of each entry
  const GridInfo* gi;         // device
};
size_t grid_bin_workspace_bytes(int64_t n_boxes);
int grid_bin_boxes(const float4* boxes, int n, void* workspace, GridBins* out, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ int cell_of(const GridInfo& g, const float4 b, int* cx_out, int* cy_out) {
  int cx = (int)((0.5f * (b.x + b.z) - g.x0) * g.inv_cell), cy = (int)((0.5f * (b.y + b.w) - g.y0) * g.inv_cell);
  cx = min(max(cx, 0), g.ncx - 1); cy = min(max(cy, 0), g.ncy - 1);
  if (cx_out) { *cx_out = cx; *cy_out = cy; }
  return cy * g.ncx + cx;
}
#endif

// ---- device helpers -----------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

#ifdef __CUDACC__
// F16F8 helpers: four fp32 -> four e4m3 bytes (round to nearest even, saturating), four e4m3 bytes -> four fp32
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
__device__ __forceinline__ void unpack_e4m3x4(uint32_t v, float& a, float& b, float& c, float& d) {
  const __half2_raw l = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(v & 0xffffu), __NV_E4M3);
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(v >> 16), __NV_E4M3);
  const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&l));
  const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&h));
  a = fl.x; b = fl.y; c = fh.x; d = fh.y;
}
// CPN_DT_F16F8 element (see cpn_b200.h): `px` = the view's first hi element of a pixel, c = channel.  8-bit chunk j =
// c / 32 holds 32 lo8 bytes then 32 hi8 bytes.
__device__ __forceinline__ float e4m3_to_f32(uint8_t b) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
__device__ __forceinline__ uint8_t f32_to_e4m3(float v) {
  return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ const uint8_t* f8_block(const __half* px, int lo_delta, int c) {
  return reinterpret_cast<const uint8_t*>(px + lo_delta) + (c >> 5) * 64 + (c & 31);
}
__device__ __forceinline__ uint8_t* f8_block(__half* px, int lo_delta, int c) {
  return reinterpret_cast<uint8_t*>(px + lo_delta) + (c >> 5) * 64 + (c & 31);
}
__device__ __forceinline__ float f16f8_load(const __half* px, int lo_delta, int c, float lo_inv) {
  return __half2float(px[c]) + e4m3_to_f32(f8_block(px, lo_delta, c)[0]) * lo_inv;
}
__device__ __forceinline__ void f16f8_store(__half* px, int lo_delta, int c, float v, float lo_scale, float hi8_scale) {
  const __half h = __float2half_rn(v);
  const float hf = __half2float(h);
  px[c] = h;
  uint8_t* q = f8_block(px, lo_delta, c);
  q[0] = f32_to_e4m3((v - hf) * lo_scale);
  q[32] = f32_to_e4m3(hf * hi8_scale);
}
#endif

}  // namespace cpn
