// Post-head chain of the Contour Proposal Network (compiled with -fmad=false: the fused decode must round exactly like
// the reference's un-fused torch ops, because torch.round() in the refinement loop is discontinuous).
//
//   select   : sigmoid / score bounds / threshold / ordered compaction      models/cpn.py:575-587, 616-620
//   decode   : rel->abs location, inverse DFT, rescale, 4x local refinement,
//              clamp, boxes, offsets -- one kernel, one warp per proposal    ops/cpn.py:15-165, models/cpn.py:63-85,
//                                                                            :661-670, :695-702
//   f2c      : stand-alone ops.cpn.fouriers2contours (register-tiled)        ops/cpn.py:44-95
//   border   : remove_border_contours                                        ops/cpn.py:258-290
//   gather   : row gather (resolve_keep_indices)                             models/cpn.py:53-60
// All of it is HBM-/latency-bound gather-scatter work; no tensor cores.
#include "common.cuh"
#include <cstdlib>

namespace cpn {

// ---------------------------------------------------------------------------------------------------------------------
// select
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256, SEL_PER_THREAD = 4, SEL_CHUNK = SEL_THREADS * SEL_PER_THREAD;

// Device copy of cpn_select_params_t (pointers are device pointers already).
struct SelParams {
  const float* logits;
  const float* lower;
  const float* upper;
  const float* unc;
  int channels, use_certainty;
  float thresh, certainty_limit;
};

constexpr int SEL_MAX_CLASSES = 32;

// Score, class and foreground decision of pixel i for every scoring variant (models/cpn.py:575-587, 616-618).
__device__ __forceinline__ bool eval_pixel(const SelParams& p, long long i, float& score, int& cls) {
  bool fg;
  if (p.channels == 1) {
    float s = 1.f / (1.f + expf(-p.logits[i]));  // torch.sigmoid
    if (p.upper) s = fminf(s, p.upper[i]);       // cpn.py:118-123 (upper first, then lower)
    if (p.lower) s = fmaxf(s, p.lower[i]);
    score = s;
    fg = s > p.thresh;
    cls = fg ? 1 : 0;
  } else {
    // F.softmax(scores, dim=1): exp(x - max) / sum; bounds broadcast over the channels; argmax = first maximum
    const float* l = p.logits + i * p.channels;
    float m = l[0];
    for (int c = 1; c < p.channels; ++c) m = fmaxf(m, l[c]);
    float sum = 0.f;
    for (int c = 0; c < p.channels; ++c) sum = sum + expf(l[c] - m);
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < p.channels; ++c) {
      float s = expf(l[c] - m) / sum;
      if (p.upper) s = fminf(s, p.upper[i]);
      if (p.lower) s = fmaxf(s, p.lower[i]);
      if (s > best) { best = s; arg = c; }
    }
    score = best;
    cls = arg;
    fg = arg > 0;
  }
  if (fg && p.use_certainty) {  // fg_mask &= uncertainty.mean(1) < (1 - certainty_thresh)   (cpn.py:617-618)
    const float4 u = *reinterpret_cast<const float4*>(p.unc + i * 4);
    const float mean = (((u.x + u.y) + u.z) + u.w) / 4.f;
    fg = mean < p.certainty_limit;
  }
  return fg;
}

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const SelParams p, long long pixels,
                                                                   int* __restrict__ block_counts) {
  __shared__ int warp_sums[SEL_THREADS / 32];
  const long long base = (long long)blockIdx.x * SEL_CHUNK + threadIdx.x * SEL_PER_THREAD;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < SEL_PER_THREAD; ++j) {
    const long long i = base + j;
    float sc;
    int cl;
    if (i < pixels) cnt += eval_pixel(p, i, sc, cl) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) t += warp_sums[w];
    block_counts[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block_counts -> block_offsets[0..nblocks) (int64), block_offsets[nblocks] = total
__global__ void __launch_bounds__(1024) select_scan_kernel(const int* __restrict__ block_counts, int nblocks,
                                                           long long* __restrict__ block_offsets,
                                                           long long* __restrict__ total) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s, chunk_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const long long v = b < nblocks ? block_counts[b] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const long long w = warp_tot[lane];
      long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;  // exclusive offset of each warp inside this chunk
      if (lane == 31) chunk_s = winc;
    }
    __syncthreads();
    if (b < nblocks) block_offsets[b] = carry_s + warp_tot[warp] + (inc - v);
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    block_offsets[nblocks] = carry_s;
    *total = carry_s;
  }
}

__global__ void __launch_bounds__(SEL_THREADS) select_write_kernel(const SelParams p, long long pixels,
                                                                   const long long* __restrict__ block_offsets,
                                                                   int32_t* __restrict__ idx, float* __restrict__ score,
                                                                   long long* __restrict__ classes,
                                                                   long long capacity) {
  __shared__ int warp_sums[SEL_THREADS / 32];
  const long long base = (long long)blockIdx.x * SEL_CHUNK + threadIdx.x * SEL_PER_THREAD;
  float sc[SEL_PER_THREAD];
  int cl[SEL_PER_THREAD];
  bool fg[SEL_PER_THREAD];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < SEL_PER_THREAD; ++j) {
    const long long i = base + j;
    sc[j] = 0.f;
    cl[j] = 0;
    fg[j] = false;
    if (i < pixels) fg[j] = eval_pixel(p, i, sc[j], cl[j]);
    cnt += fg[j] ? 1 : 0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_sums[w];
  long long pos = block_offsets[blockIdx.x] + woff + (inc - cnt);
#pragma unroll
  for (int j = 0; j < SEL_PER_THREAD; ++j) {
    if (fg[j]) {
      if (pos < capacity) {
        idx[pos] = (int32_t)(base + j);
        score[pos] = sc[j];
        if (classes) classes[pos] = cl[j];
      }
      ++pos;
    }
  }
}

// seg_offsets[n] = lower_bound(idx, n * hw) for n in [0, N]
__global__ void select_segments_kernel(const int32_t* __restrict__ idx, const long long* __restrict__ total_p,
                                       long long capacity, int n_images, long long hw,
                                       int32_t* __restrict__ seg_offsets) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n > n_images) return;
  long long total = *total_p;
  if (total > capacity) total = capacity;
  const long long key = (long long)n * hw;
  long long lo = 0, hi = total;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)idx[mid] < key) lo = mid + 1; else hi = mid;
  }
  seg_offsets[n] = (int32_t)lo;
}

}  // namespace cpn

using namespace cpn;

extern "C" size_t cpn_select_workspace_bytes(int64_t pixels) {
  const int64_t nblocks = (pixels + SEL_CHUNK - 1) / SEL_CHUNK;
  // block_counts int32[nblocks] (padded to 8) + block_offsets int64[nblocks + 1] + total int64
  return (size_t)(((nblocks * 4 + 15) / 16) * 16 + (nblocks + 2) * 8 + 64);
}

static inline void select_ws(void* ws, int64_t pixels, int** counts, long long** offsets) {
  const int64_t nblocks = (pixels + SEL_CHUNK - 1) / SEL_CHUNK;
  *counts = reinterpret_cast<int*>(ws);
  *offsets = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + ((nblocks * 4 + 15) / 16) * 16);
}

static int make_sel(const cpn_select_params_t* q, SelParams* p) {
  CPN_REQUIRE(q != nullptr && q->logits != nullptr, "select: logits required");
  CPN_REQUIRE(q->channels == 1 || (q->channels > 2 && q->channels <= SEL_MAX_CLASSES),
              "select: score channels %d unsupported (1 or 3..%d)", q->channels, SEL_MAX_CLASSES);
  CPN_REQUIRE(!q->use_certainty || q->uncertainty != nullptr, "select: certainty filter without uncertainty map");
  CPN_REQUIRE(q->uncertainty == nullptr || ((uintptr_t)q->uncertainty % 16) == 0, "select: uncertainty must be 16-byte aligned");
  p->logits = q->logits; p->lower = q->lower; p->upper = q->upper; p->unc = q->uncertainty;
  p->channels = q->channels; p->use_certainty = q->use_certainty; p->thresh = q->thresh;
  p->certainty_limit = q->certainty_limit;
  return 0;
}

extern "C" int cpn_select_count_ex(const cpn_select_params_t* params, int64_t pixels, void* workspace,
                                   int64_t* total_dev, void* stream) {
  CPN_REQUIRE(pixels > 0 && pixels < (1ll << 31), "select: pixels %lld out of range", (long long)pixels);
  SelParams p;
  if (make_sel(params, &p)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  int* counts; long long* offsets;
  select_ws(workspace, pixels, &counts, &offsets);
  const int nblocks = (int)((pixels + SEL_CHUNK - 1) / SEL_CHUNK);
  select_count_kernel<<<nblocks, SEL_THREADS, 0, st>>>(p, pixels, counts);
  CPN_CHECK_LAUNCH();
  select_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, offsets, reinterpret_cast<long long*>(total_dev));
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_select_write_ex(const cpn_select_params_t* params, int n_images, int64_t hw, void* workspace,
                                   int32_t* idx, float* score, int64_t* classes, int64_t capacity,
                                   int32_t* seg_offsets, void* stream) {
  const int64_t pixels = (int64_t)n_images * hw;
  CPN_REQUIRE(pixels > 0 && pixels < (1ll << 31), "select: pixels %lld out of range", (long long)pixels);
  SelParams p;
  if (make_sel(params, &p)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  int* counts; long long* offsets;
  select_ws(workspace, pixels, &counts, &offsets);
  const int nblocks = (int)((pixels + SEL_CHUNK - 1) / SEL_CHUNK);
  select_write_kernel<<<nblocks, SEL_THREADS, 0, st>>>(p, pixels, offsets, idx, score,
                                                       reinterpret_cast<long long*>(classes), capacity);
  CPN_CHECK_LAUNCH();
  select_segments_kernel<<<(n_images + 1 + 127) / 128, 128, 0, st>>>(idx, offsets + nblocks, capacity, n_images, hw,
                                                                     seg_offsets);
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_select_count(const float* logits, const float* lower, const float* upper, int64_t pixels,
                                float thresh, void* workspace, int64_t* total_dev, void* stream) {
  cpn_select_params_t q = {logits, lower, upper, nullptr, 1, 0, thresh, 0.f};
  return cpn_select_count_ex(&q, pixels, workspace, total_dev, stream);
}

extern "C" int cpn_select_write(const float* logits, const float* lower, const float* upper, int n_images, int64_t hw,
                                float thresh, void* workspace, int32_t* idx, float* score, int64_t capacity,
                                int32_t* seg_offsets, void* stream) {
  cpn_select_params_t q = {logits, lower, upper, nullptr, 1, 0, thresh, 0.f};
  return cpn_select_write_ex(&q, n_images, hw, workspace, idx, score, nullptr, capacity, seg_offsets, stream);
}

namespace cpn {
__global__ void nms_weights_kernel(const float* __restrict__ scores, const float* __restrict__ unc,
                                   const int32_t* __restrict__ idx, long long P, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 u = *reinterpret_cast<const float4*>(unc + (long long)idx[i] * 4);
  const float mean = (((u.x + u.y) + u.z) + u.w) / 4.f;
  out[i] = scores[i] * (1.f - mean);   // s * (1. - u.mean(1))   (models/cpn.py:724)
}
}  // namespace cpn

extern "C" int cpn_nms_weights(const float* scores, const float* uncertainty, const int32_t* idx, int64_t P, float* out,
                               void* stream) {
  if (P <= 0) return 0;
  CPN_REQUIRE(scores && uncertainty && idx && out, "nms_weights: null pointer");
  cpn::nms_weights_kernel<<<(int)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(scores, uncertainty, idx, P, out);
  CPN_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// decode + refine (one warp per proposal)
// ---------------------------------------------------------------------------------------------------------------------
namespace cpn {

constexpr int DEC_WARPS = 8, DEC_MAX_ORDER = 32, DEC_MAX_SPL = 8;  // samples <= 32 * DEC_MAX_SPL

struct DecodeParams {
  const int32_t* idx;
  long long P;
  const float* locfou;
  int order_core, order, n_images, h, w, H, W;
  const float* trig;
  int samples;
  const float* refinement;
  int iters;
  const float* offsets;
  float *contours, *proposals, *boxes, *locations, *fourier_out;
  int trig_in_smem;
  int buckets;                  // refinement_buckets (1: plain [N,H,W,2] map)
  const int32_t* bucket_idx;    // [samples][3]
  const float* bucket_w;        // [samples][3]
  int records_by_row;           // locfou holds one record per PROPOSAL (sparse heads) instead of one per pixel
};

__global__ void __launch_bounds__(DEC_WARPS * 32) decode_refine_kernel(const DecodeParams p) {
  extern __shared__ float dsm[];
  float* trig_s = dsm;  // [2][order][samples] when trig_in_smem
  const int trig_n = 2 * p.order * p.samples;
  float* rec_s = dsm + (p.trig_in_smem ? trig_n : 0);  // [DEC_WARPS][2 + 4 * order]
  if (p.trig_in_smem)
    for (int i = threadIdx.x; i < trig_n; i += blockDim.x) trig_s[i] = p.trig[i];
  __syncthreads();
  const float* trig = p.trig_in_smem ? trig_s : p.trig;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rec_len = 2 + 4 * p.order;
  const int rec_core = 2 + 4 * p.order_core;
  float* rec = rec_s + warp * (2 + 4 * DEC_MAX_ORDER);
  const long long hw = (long long)p.h * p.w;
  // scale = original / actual, xy order (ops/cpn.py:98-104); computed in fp32 like torch.as_tensor(...)/...
  const float sx = (float)p.W / (float)p.w, sy = (float)p.H / (float)p.h;
  const float xmax = (float)(p.W - 1), ymax = (float)(p.H - 1);

  for (long long pr = (long long)blockIdx.x * DEC_WARPS + warp; pr < p.P; pr += (long long)gridDim.x * DEC_WARPS) {
    const long long pix = p.idx[pr];
    const int b = (int)(pix / hw);
    const int rem = (int)(pix - (long long)b * hw);
    const int py = rem / p.w, px = rem - py * p.w;
    __syncwarp();
    for (int i = lane; i < rec_len; i += 32) rec[i] = p.locfou[(p.records_by_row ? pr : pix) * rec_core + i];
    __syncwarp();
    // rel -> abs location (ops/cpn.py:15-41): x += column index, y += row index
    const float lx = rec[0] + (float)px, ly = rec[1] + (float)py;
    float ox = 0.f, oy = 0.f;
    if (p.offsets) { ox = p.offsets[b * 2]; oy = p.offsets[b * 2 + 1]; }

    float bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;
    for (int s = lane; s < p.samples; s += 32) {
      // con = (0 + loc) + sum_k f[k,(1,3)] * sin + sum_k f[k,(0,2)] * cos   (ops/cpn.py:92-94, sequential in k)
      float ssx = 0.f, ssy = 0.f, scx = 0.f, scy = 0.f;
      const float* tc = trig + s;
      const float* ts = trig + p.order * p.samples + s;
      for (int k = 0; k < p.order; ++k) {
        const float c = tc[k * p.samples], sn = ts[k * p.samples];
        const float* f = rec + 2 + 4 * k;
        ssx = ssx + f[1] * sn;
        ssy = ssy + f[3] * sn;
        scx = scx + f[0] * c;
        scy = scy + f[2] * c;
      }
      float cx = (lx + ssx) + scx, cy = (ly + ssy) + scy;
      cx = cx * sx;  // scale_contours
      cy = cy * sy;
      if (p.proposals) {
        float qx = cx, qy = cy;
        if (!(p.refinement && p.iters > 0)) {  // proposals alias contours when refinement is off (cpn.py:658-663)
          qx = fminf(fmaxf(qx, 0.f), xmax);
          qy = fminf(fmaxf(qy, 0.f), ymax);
          if (p.offsets) { qx = qx + ox; qy = qy + oy; }  // added twice in the reference (same tensor, cpn.py:698-699)
        }
        if (p.offsets) { qx = qx + ox; qy = qy + oy; }
        reinterpret_cast<float2*>(p.proposals)[pr * p.samples + s] = make_float2(qx, qy);
      }
      if (p.refinement && p.iters > 0 && p.buckets == 1) {  // models/cpn.py:63-85
        const float2* ref = reinterpret_cast<const float2*>(p.refinement) + (long long)b * p.H * p.W;
        for (int it = 0; it < p.iters; ++it) {
          float rx = rintf(cx), ry = rintf(cy);  // torch.round: half to even
          rx = fminf(fmaxf(rx, 0.f), xmax);
          ry = fminf(fmaxf(ry, 0.f), ymax);
          const float2 d = __ldg(ref + (long long)((int)ry) * p.W + (int)rx);
          cx = rx + d.x;
          cy = ry + d.y;
        }
      } else if (p.refinement && p.iters > 0) {
        // bucketed refinement (models/cpn.py:73-82): three neighbouring buckets of this sample, weights from the host
        // table (resolve_refinement_buckets); summed in the reference's order a, b, c with un-fused multiply/add
        const float2* ref = reinterpret_cast<const float2*>(p.refinement) + (long long)b * p.H * p.W * p.buckets;
        const int b0 = p.bucket_idx[s * 3], b1 = p.bucket_idx[s * 3 + 1], b2 = p.bucket_idx[s * 3 + 2];
        const float w0 = p.bucket_w[s * 3], w1 = p.bucket_w[s * 3 + 1], w2 = p.bucket_w[s * 3 + 2];
        for (int it = 0; it < p.iters; ++it) {
          float rx = rintf(cx), ry = rintf(cy);
          rx = fminf(fmaxf(rx, 0.f), xmax);
          ry = fminf(fmaxf(ry, 0.f), ymax);
          const float2* px_ref = ref + ((long long)((int)ry) * p.W + (int)rx) * p.buckets;
          const float2 d0 = __ldg(px_ref + b0), d1 = __ldg(px_ref + b1), d2 = __ldg(px_ref + b2);
          const float dx = (d0.x * w0 + d1.x * w1) + d2.x * w2;
          const float dy = (d0.y * w0 + d1.y * w1) + d2.y * w2;
          cx = rx + dx;
          cy = ry + dy;
        }
      }
      cx = fminf(fmaxf(cx, 0.f), xmax);  // cpn.py:661-663
      cy = fminf(fmaxf(cy, 0.f), ymax);
      bx0 = fminf(bx0, cx); by0 = fminf(by0, cy); bx1 = fmaxf(bx1, cx); by1 = fmaxf(by1, cy);
      if (p.contours) {
        float qx = cx, qy = cy;
        if (p.offsets) {
          qx = qx + ox; qy = qy + oy;
          if (!(p.refinement && p.iters > 0)) { qx = qx + ox; qy = qy + oy; }
        }
        reinterpret_cast<float2*>(p.contours)[pr * p.samples + s] = make_float2(qx, qy);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
      by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
      by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0) {
      if (p.boxes) reinterpret_cast<float4*>(p.boxes)[pr] = make_float4(bx0 + ox, by0 + oy, bx1 + ox, by1 + oy);
      if (p.locations) reinterpret_cast<float2*>(p.locations)[pr] = make_float2(lx * sx + ox, ly * sy + oy);
    }
    if (p.fourier_out) {  // scale_fourier: cols (0,1) * sx, (2,3) * sy (ops/cpn.py:133-137)
      for (int i = lane; i < 4 * p.order; i += 32)
        p.fourier_out[pr * 4 * p.order + i] = rec[2 + i] * (((i & 3) < 2) ? sx : sy);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// stand-alone fouriers2contours, default sampling: register-tiled, uses t_{S-1-s} = 1 - t_s (cos even, sin odd) so
// each (cos, sin) partial sum serves two output samples.  Thread tile: 4 proposals x up to 4 sample pairs.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int F2C_TX = 16, F2C_TY = 8, F2C_TP = 4, F2C_PB = F2C_TY * F2C_TP;  // 32 proposals per block iteration

__global__ void __launch_bounds__(F2C_TX * F2C_TY) f2c_tiled_kernel(const float* __restrict__ fourier,
                                                                    const float* __restrict__ locations, long long P,
                                                                    int order, int samples,
                                                                    const float* __restrict__ trig,
                                                                    float* __restrict__ out) {
  extern __shared__ float fsm[];
  const int np = (samples + 1) / 2;  // sample pairs (s, S-1-s); the middle sample of an odd S pairs with itself
  float* cos_s = fsm;                 // [order][np]
  float* sin_s = fsm + order * np;    // [order][np]
  float* coef = fsm + ((2 * order * np + 3) & ~3);  // [F2C_PB][order * 4 + 4] (+2 used: location), 16-byte aligned
  const int cstride = order * 4 + 4;
  for (int i = threadIdx.x; i < order * np; i += blockDim.x) {
    const int k = i / np, j = i - k * np;
    cos_s[i] = trig[k * samples + j];
    sin_s[i] = trig[(order + k) * samples + j];
  }
  const int tx = threadIdx.x % F2C_TX, ty = threadIdx.x / F2C_TX;
  const long long nblk = (P + F2C_PB - 1) / F2C_PB;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long p0 = blk * F2C_PB;
    __syncthreads();
    // stage coefficients (coalesced: the fourier rows of 32 consecutive proposals are contiguous)
    const long long nf = min((long long)F2C_PB, P - p0) * order * 4;
    for (long long i = threadIdx.x; i < nf; i += blockDim.x) {
      const int pl = (int)(i / (order * 4)), r = (int)(i - (long long)pl * order * 4);
      coef[pl * cstride + r] = __ldg(fourier + p0 * order * 4 + i);
    }
    for (int i = threadIdx.x; i < F2C_PB * 2; i += blockDim.x) {
      const int pl = i >> 1;
      if (p0 + pl < P) coef[pl * cstride + order * 4 + (i & 1)] = __ldg(locations + (p0 + pl) * 2 + (i & 1));
    }
    __syncthreads();
    for (int j0 = 0; j0 < np; j0 += F2C_TX * 4) {
      float cx[F2C_TP][4], sx_[F2C_TP][4], cy[F2C_TP][4], sy_[F2C_TP][4];
#pragma unroll
      for (int a = 0; a < F2C_TP; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) cx[a][q] = sx_[a][q] = cy[a][q] = sy_[a][q] = 0.f;
      for (int k = 0; k < order; ++k) {
        float c[4], s[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = j0 + tx + q * F2C_TX;
          c[q] = j < np ? cos_s[k * np + j] : 0.f;
          s[q] = j < np ? sin_s[k * np + j] : 0.f;
        }
#pragma unroll
        for (int a = 0; a < F2C_TP; ++a) {
          const float4 f = *reinterpret_cast<const float4*>(coef + (ty * F2C_TP + a) * cstride + k * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            cx[a][q] = __fmaf_rn(f.x, c[q], cx[a][q]);
            sx_[a][q] = __fmaf_rn(f.y, s[q], sx_[a][q]);
            cy[a][q] = __fmaf_rn(f.z, c[q], cy[a][q]);
            sy_[a][q] = __fmaf_rn(f.w, s[q], sy_[a][q]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < F2C_TP; ++a) {
        const int pl = ty * F2C_TP + a;
        const long long pr = p0 + pl;
        if (pr >= P) continue;
        const float lx = coef[pl * cstride + order * 4], ly = coef[pl * cstride + order * 4 + 1];
        float2* o = reinterpret_cast<float2*>(out) + pr * samples;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = j0 + tx + q * F2C_TX;
          if (j >= np) continue;
          o[j] = make_float2((lx + sx_[a][q]) + cx[a][q], (ly + sy_[a][q]) + cy[a][q]);
          const int jm = samples - 1 - j;
          if (jm != j) o[jm] = make_float2((lx - sx_[a][q]) + cx[a][q], (ly - sy_[a][q]) + cy[a][q]);
        }
      }
    }
  }
}

// Fast path for samples % 8 == 0 (the common 32 / 64 / 128): same register tile, but
//  * the coefficient tile of the NEXT 32 proposals is prefetched with cp.async (double buffer) while this one computes,
//  * each thread owns 4 ADJACENT sample pairs, so trig factors are one LDS.128 and every store is a 16-byte
//    st.global.cs (two (x, y) vertices), on both the forward and the mirrored half of the contour.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_cs_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MINB>
__global__ void __launch_bounds__(F2C_TX * F2C_TY, MINB) f2c_fast_kernel(const float* __restrict__ fourier,
                                                                       const float* __restrict__ locations,
                                                                       long long P, int order, int samples,
                                                                       const float* __restrict__ trig,
                                                                       float* __restrict__ out, int txe) {
  // Warp-granular software pipeline: every warp owns a double-buffered coefficient tile of PBW proposals and runs its
  // own cp.async prefetch / compute / store loop, synchronising with __syncwarp only (no block barriers in the loop).
  extern __shared__ __align__(16) float fsm[];
  const int np = samples / 2;             // multiple of 4
  const int row = order * 4;              // floats per proposal
  const int cstride = row + 4;            // padded row (16-byte multiple)
  constexpr int NWARP = F2C_TX * F2C_TY / 32;
  const int rows_w = 32 / txe;            // proposal rows (of F2C_TP proposals) per warp
  const int PBW = rows_w * F2C_TP;        // proposals per warp iteration
  float* cos_s = fsm;                     // [order][np]
  float* sin_s = fsm + order * np;        // [order][np]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* coef0 = fsm + 2 * order * np + warp * 2 * PBW * cstride;          // this warp: 2 x [PBW][cstride]
  float* loc0 = fsm + 2 * order * np + NWARP * 2 * PBW * cstride + warp * 2 * PBW * 2;  // this warp: 2 x [PBW][2]
  for (int i = threadIdx.x; i < order * np; i += blockDim.x) {
    const int k = i / np, j = i - k * np;
    cos_s[i] = trig[k * samples + j];
    sin_s[i] = trig[(order + k) * samples + j];
  }
  __syncthreads();
  const int tx = lane % txe, ty = lane / txe;
  const long long ntile = (P + PBW - 1) / PBW;
  const long long wid = (long long)blockIdx.x * NWARP + warp, nw = (long long)gridDim.x * NWARP;
  const int chunks_per_row = row / 4;     // 16-byte chunks per proposal

  auto prefetch = [&](long long t, int buf) {
    const long long p0 = t * PBW;
    const int npr = (int)min((long long)PBW, P - p0);
    float* cb = coef0 + buf * PBW * cstride;
    for (int i = lane; i < npr * chunks_per_row; i += 32) {
      const int pl = i / chunks_per_row, c = i - pl * chunks_per_row;
      cp_async16(cb + pl * cstride + c * 4, fourier + (p0 + pl) * row + c * 4);
    }
    float* lb = loc0 + buf * PBW * 2;
    for (int i = lane; i < npr / 2; i += 32) cp_async16(lb + i * 4, locations + p0 * 2 + i * 4);
    if ((npr & 1) && lane == 0) {          // odd tail proposal: plain loads (visible after __syncwarp)
      lb[(npr - 1) * 2] = locations[(p0 + npr - 1) * 2];
      lb[(npr - 1) * 2 + 1] = locations[(p0 + npr - 1) * 2 + 1];
    }
  };

  int buf = 0;
  if (wid < ntile) prefetch(wid, 0);
  cp_async_commit();
  for (long long t = wid; t < ntile; t += nw, buf ^= 1) {
    if (t + nw < ntile) prefetch(t + nw, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const long long p0 = t * PBW;
    const float* cb = coef0 + buf * PBW * cstride;
    const float* lb = loc0 + buf * PBW * 2;
    for (int j0 = 0; j0 < np; j0 += txe * 4) {
      const int j = j0 + tx * 4;          // this thread's 4 adjacent pairs j .. j+3
      if (j < np) {
        float cx[F2C_TP][4], sx_[F2C_TP][4], cy[F2C_TP][4], sy_[F2C_TP][4];
#pragma unroll
        for (int a = 0; a < F2C_TP; ++a)
#pragma unroll
          for (int q = 0; q < 4; ++q) cx[a][q] = sx_[a][q] = cy[a][q] = sy_[a][q] = 0.f;
#pragma unroll 2
        for (int k = 0; k < order; ++k) {
          const float4 c4 = *reinterpret_cast<const float4*>(cos_s + k * np + j);
          const float4 s4 = *reinterpret_cast<const float4*>(sin_s + k * np + j);
          const float c[4] = {c4.x, c4.y, c4.z, c4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int a = 0; a < F2C_TP; ++a) {
            const float4 f = *reinterpret_cast<const float4*>(cb + (ty * F2C_TP + a) * cstride + k * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              cx[a][q] = __fmaf_rn(f.x, c[q], cx[a][q]);
              sx_[a][q] = __fmaf_rn(f.y, s[q], sx_[a][q]);
              cy[a][q] = __fmaf_rn(f.z, c[q], cy[a][q]);
              sy_[a][q] = __fmaf_rn(f.w, s[q], sy_[a][q]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < F2C_TP; ++a) {
          const int pl = ty * F2C_TP + a;
          const long long pr = p0 + pl;
          if (pr >= P) continue;
          const float lx = lb[pl * 2], ly = lb[pl * 2 + 1];
          float* o = out + pr * samples * 2;
          float fx[4], fy[4], mx[4], my[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            fx[q] = (lx + sx_[a][q]) + cx[a][q];
            fy[q] = (ly + sy_[a][q]) + cy[a][q];
            mx[q] = (lx - sx_[a][q]) + cx[a][q];
            my[q] = (ly - sy_[a][q]) + cy[a][q];
          }
          st_cs_f4(o + (j + 0) * 2, fx[0], fy[0], fx[1], fy[1]);
          st_cs_f4(o + (j + 2) * 2, fx[2], fy[2], fx[3], fy[3]);
          const int jm = samples - 4 - j;   // mirrored samples S-1-j-3 .. S-1-j, ascending order
          st_cs_f4(o + (jm + 0) * 2, mx[3], my[3], mx[2], my[2]);
          st_cs_f4(o + (jm + 2) * 2, mx[1], my[1], mx[0], my[0]);
        }
      }
    }
    __syncwarp();   // all lanes are done with `buf` before the next iteration's prefetch overwrites it
  }
  cp_async_wait<0>();
}

// explicit per-proposal sampling [P, S] (ops/cpn.py:67-71): one warp per proposal, trig evaluated on the fly
__global__ void __launch_bounds__(256) f2c_sampling_kernel(const float* __restrict__ fourier,
                                                           const float* __restrict__ locations, long long P, int order,
                                                           int samples, const float* __restrict__ sampling,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float two_pi = 6.2831854820251465f;  // float(np.pi) * 2 rounded to fp32
  for (long long pr = (long long)blockIdx.x * 8 + warp; pr < P; pr += (long long)gridDim.x * 8) {
    const float lx = locations[pr * 2], ly = locations[pr * 2 + 1];
    const float* f = fourier + pr * order * 4;
    for (int s = lane; s < samples; s += 32) {
      const float t = sampling[pr * samples + s];
      float ssx = 0.f, ssy = 0.f, scx = 0.f, scy = 0.f;
      for (int k = 0; k < order; ++k) {
        const float arg = (two_pi * (float)(k + 1)) * t;
        const float c = cosf(arg), sn = sinf(arg);
        ssx = ssx + f[k * 4 + 1] * sn;
        ssy = ssy + f[k * 4 + 3] * sn;
        scx = scx + f[k * 4 + 0] * c;
        scy = scy + f[k * 4 + 2] * c;
      }
      reinterpret_cast<float2*>(out)[pr * samples + s] = make_float2((lx + ssx) + scx, (ly + ssy) + scy);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// border filter + gather
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) border_filter_kernel(const float* __restrict__ contours,
                                                            const int32_t* __restrict__ tile_of_row,
                                                            const float* __restrict__ tile_meta, long long K,
                                                            int samples, float padding, uint8_t* __restrict__ keep) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < K; r += (long long)gridDim.x * 8) {
    const float* m = tile_meta + (long long)tile_of_row[r] * CPN_TILE_META;
    const float offx = m[0], offy = m[1], h = m[2], w = m[3];
    const bool top = m[4] != 0.f, right = m[5] != 0.f, bottom = m[6] != 0.f, left = m[7] != 0.f;
    const bool ex_br = m[8] != 0.f;           // filter_contours_by_stitching_rule 'ex_br' (ops/cpn.py:316-319)
    const float stop_x = m[9], stop_y = m[10];
    bool ok = true, all_rb = true;
    for (int s = lane; s < samples; s += 32) {
      const float2 v = reinterpret_cast<const float2*>(contours)[r * samples + s];
      const float x = v.x + (-offx), y = v.y + (-offy);  // contours + offsets with offsets = -tile offset
      if (top) ok = ok && (y > padding);
      if (right) ok = ok && (x < (w - padding));
      if (bottom) ok = ok && (y < (h - padding));
      if (left) ok = ok && (x > padding);
      all_rb = all_rb && ((x >= stop_x) || (y >= stop_y));   // (contours >= stop).any(-1)
    }
    ok = __all_sync(0xffffffffu, ok);
    all_rb = __all_sync(0xffffffffu, all_rb);                // ....all(-1): every vertex lies in the right/bottom overlap
    if (lane == 0) keep[r] = (ok && !(ex_br && all_rb)) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ src, long long row_words,
                                                          const int32_t* __restrict__ index, long long n_rows,
                                                          uint32_t* __restrict__ dst) {
  const long long total = n_rows * row_words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_words, c = i - r * row_words;
    dst[i] = src[(long long)index[r] * row_words + c];
  }
}

static inline int capped_grid(long long blocks, int per_sm) {
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace cpn

extern "C" int cpn_decode_refine(const int32_t* idx, int64_t P, const float* locfou, int order_core, int order,
                                 int n_images, int h, int w, int H, int W, const float* trig, int samples,
                                 const float* refinement, int iters, const float* offsets, float* contours,
                                 float* proposals, float* boxes, float* locations, float* fourier_out, void* stream) {
  return cpn_decode_refine_buckets(idx, P, locfou, order_core, order, n_images, h, w, H, W, trig, samples, refinement,
                                   iters, 1, nullptr, nullptr, offsets, contours, proposals, boxes, locations,
                                   fourier_out, stream);
}

extern "C" int cpn_decode_refine_buckets(const int32_t* idx, int64_t P, const float* locfou, int order_core, int order,
                                         int n_images, int h, int w, int H, int W, const float* trig, int samples,
                                         const float* refinement, int iters, int buckets, const int32_t* bucket_idx,
                                         const float* bucket_w, const float* offsets, float* contours,
                                         float* proposals, float* boxes, float* locations, float* fourier_out,
                                         void* stream) {
  return cpn_decode_refine_rows(idx, P, locfou, 0, order_core, order, n_images, h, w, H, W, trig, samples, refinement,
                                iters, buckets, bucket_idx, bucket_w, offsets, contours, proposals, boxes, locations,
                                fourier_out, stream);
}

extern "C" int cpn_decode_refine_rows(const int32_t* idx, int64_t P, const float* locfou, int records_by_row,
                                      int order_core, int order, int n_images, int h, int w, int H, int W,
                                      const float* trig, int samples, const float* refinement, int iters, int buckets,
                                      const int32_t* bucket_idx, const float* bucket_w, const float* offsets,
                                      float* contours, float* proposals, float* boxes, float* locations,
                                      float* fourier_out, void* stream) {
  CPN_REQUIRE(buckets >= 1, "decode: refinement buckets must be >= 1");
  CPN_REQUIRE(buckets == 1 || refinement == nullptr || iters <= 0 || (bucket_idx != nullptr && bucket_w != nullptr),
              "decode: bucketed refinement needs the bucket tables");
  CPN_REQUIRE(order >= 1 && order <= DEC_MAX_ORDER && order <= order_core, "decode: order %d out of range (core %d)",
              order, order_core);
  CPN_REQUIRE(samples >= 1, "decode: samples must be >= 1");
  if (P <= 0) return 0;
  DecodeParams p;
  p.idx = idx; p.P = P; p.locfou = locfou; p.order_core = order_core; p.order = order; p.n_images = n_images;
  p.h = h; p.w = w; p.H = H; p.W = W; p.trig = trig; p.samples = samples; p.refinement = refinement; p.iters = iters;
  p.offsets = offsets; p.contours = contours; p.proposals = proposals; p.boxes = boxes; p.locations = locations;
  p.fourier_out = fourier_out;
  p.buckets = buckets; p.bucket_idx = bucket_idx; p.bucket_w = bucket_w;
  p.records_by_row = records_by_row ? 1 : 0;
  const size_t trig_bytes = (size_t)2 * order * samples * sizeof(float);
  const size_t rec_bytes = (size_t)DEC_WARPS * (2 + 4 * DEC_MAX_ORDER) * sizeof(float);
  p.trig_in_smem = trig_bytes + rec_bytes <= 48 * 1024;
  const size_t smem = rec_bytes + (p.trig_in_smem ? trig_bytes : 0);
  const int grid = capped_grid((P + DEC_WARPS - 1) / DEC_WARPS, 8);
  decode_refine_kernel<<<grid, DEC_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_fouriers2contours(const float* fourier, const float* locations, int64_t P, int order, int samples,
                                     const float* trig, const float* sampling, float* out, void* stream) {
  CPN_REQUIRE(order >= 1 && samples >= 1, "fouriers2contours: bad order/samples");
  CPN_REQUIRE((trig != nullptr) != (sampling != nullptr), "fouriers2contours: pass exactly one of trig / sampling");
  if (P <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (sampling) {
    f2c_sampling_kernel<<<capped_grid((P + 7) / 8, 8), 256, 0, st>>>(fourier, locations, P, order, samples, sampling,
                                                                     out);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  if (samples % 8 == 0 && ((uintptr_t)fourier % 16 == 0) && ((uintptr_t)locations % 16 == 0) &&
      ((uintptr_t)out % 16 == 0)) {
    const int nph = samples / 2;
    int txe = 16;
    while (txe > 1 && txe * 4 > nph) txe >>= 1;          // threads per proposal row: 4 adjacent pairs each
    const int pb = (F2C_TX * F2C_TY / txe) * F2C_TP;     // proposals per block iteration (all warps together)
    const size_t fsmem = ((size_t)2 * order * nph + 2 * (size_t)pb * (order * 4 + 4) + 2 * (size_t)pb * 2) * sizeof(float);
    if (fsmem <= 200 * 1024) {
      static size_t fconfigured = 0;
      if (fsmem > 48 * 1024 && fsmem > fconfigured) {
        CPN_CHECK_CUDA(cudaFuncSetAttribute(f2c_fast_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        CPN_CHECK_CUDA(cudaFuncSetAttribute(f2c_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        fconfigured = fsmem;
      }
      const long long nb = (P + pb - 1) / pb;
      static int minb = 0;
      if (minb == 0) { const char* e = getenv("CPN_F2C_MINB"); minb = (e && atoi(e) == 4) ? 4 : 3; }
      if (minb == 4)
        f2c_fast_kernel<4><<<capped_grid(nb, 4), F2C_TX * F2C_TY, fsmem, st>>>(fourier, locations, P, order, samples,
                                                                               trig, out, txe);
      else
        f2c_fast_kernel<3><<<capped_grid(nb, 3), F2C_TX * F2C_TY, fsmem, st>>>(fourier, locations, P, order, samples,
                                                                               trig, out, txe);
      CPN_CHECK_LAUNCH();
      return 0;
    }
  }
  const int np = (samples + 1) / 2;
  const size_t smem = ((size_t)((2 * order * np + 3) & ~3) + (size_t)F2C_PB * (order * 4 + 4)) * sizeof(float);
  CPN_REQUIRE(smem <= 200 * 1024, "fouriers2contours: order*samples too large for shared memory (%zu B)", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(f2c_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long nblk = (P + F2C_PB - 1) / F2C_PB;
  f2c_tiled_kernel<<<capped_grid(nblk, 8), F2C_TX * F2C_TY, smem, st>>>(fourier, locations, P, order, samples, trig,
                                                                        out);
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_border_filter(const float* contours, const int32_t* tile_of_row, const float* tile_meta, int64_t K,
                                 int samples, float padding, uint8_t* keep, void* stream) {
  if (K <= 0) return 0;
  border_filter_kernel<<<capped_grid((K + 7) / 8, 8), 256, 0, (cudaStream_t)stream>>>(contours, tile_of_row, tile_meta,
                                                                                      K, samples, padding, keep);
  CPN_CHECK_LAUNCH();
  return 0;
}

extern "C" int cpn_gather_rows(const void* src, int64_t row_bytes, const int32_t* index, int64_t n_rows, void* dst,
                               void* stream) {
  CPN_REQUIRE(row_bytes > 0 && row_bytes % 4 == 0, "gather_rows: row_bytes must be a positive multiple of 4");
  if (n_rows <= 0) return 0;
  const long long total = n_rows * (row_bytes / 4);
  gather_rows_kernel<<<capped_grid((total + 255) / 256, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(src), row_bytes / 4, index, n_rows, reinterpret_cast<uint32_t*>(dst));
  CPN_CHECK_LAUNCH();
  return 0;
}
