// contours2labels: rasterisation of (stitched) contours into an overlap-aware multi-channel label image -- the step
// that follows the CPN hot path.  Replaces /root/reference/celldetection/data/cpn.py:292-358
// (contours2labels) and :246-256 (render_contour -> cv2.drawContours(thickness=-1), OpenCV imgproc/drawing.cpp:
// CollectPolyEdges + Line + FillEdgeCollection), as called by celldetection_scripts/cpn_inference.py:809-813.
//
// Reference semantics (sequential over contours k = 0, 1, ...):  round (half to even) and clip the vertices, render the
// polygon (boundary lines by 8-connected left-to-right Bresenham + integer scan-line interior on 16.16 fixed-point edge
// x coordinates) with label k + 1, and add it to the FIRST channel whose gap-dilated bounding-box region holds no label
// yet (a new channel is appended when all are occupied).
//
// B200 design: the greedy channel choice only depends on earlier contours whose bounding box meets this contour's
// dilated box, so the sequence is a sparse dependency DAG.  One persistent warp per contour, contours taken in index
// order (warp w handles k = w, w + G, ...): the warp (1) waits until every earlier neighbour (found through the same
// uniform grid the stitch NMS uses) has published its `done` flag, (2) ORs the occupied-channel bits over its dilated
// box (L2 loads, lanes across pixels, all channels of a pixel in one vector), (3) paints boundary + interior with plain
// stores, (4) fences and publishes its own flag.  The smallest unfinished index never waits, so the schedule cannot
// deadlock as long as all warps of the grid are resident (the grid is sized by the occupancy API).  Integer / byte work,
// L2-resident; no tensor cores.  Results are bit-identical to the reference (tests/golden/contours2labels.npz).
#include "common.cuh"
#include <cstring>

namespace cpn {

constexpr int C2L_WARPS = 8;
constexpr int C2L_MAX_CH = 64;          // occupied-channel bitmask width
constexpr long long XY_ONE = 1ll << 16;

struct C2LParams {
  const float* contours;     // [K, S, 2] (x, y)
  int K, S, H, W;
  int rounded, clip, gap, C;
  int2* pts;                 // [K, S] integer vertices
  int4* bbox;                // [K] (xmin, ymin, xmax, ymax), from the float contour like render_contour
  float4* ebox;              // [K] dilated boxes as floats (grid binning)
  int32_t* labels;           // [H, W, C]
  uint32_t* done;            // [K]
  int32_t* info;             // [0] channels used, [1] channels needed, [2] error flags
};

// ---- pass 1: vertices -> integers, boxes -------------------------------------------------------------------------------
__global__ void c2l_prep_kernel(const C2LParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.K) return;
  const float2* c = reinterpret_cast<const float2*>(p.contours) + (long long)warp * p.S;
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int s = lane; s < p.S; s += 32) {
    float2 v = c[s];
    if (p.rounded) { v.x = rintf(v.x); v.y = rintf(v.y); }                   // np.round: half to even
    if (p.clip) {                                                            // clip_contour_(contour, size - 1)
      v.x = fminf(fmaxf(v.x, 0.f), (float)(p.W - 1));
      v.y = fminf(fmaxf(v.y, 0.f), (float)(p.H - 1));
    }
    mnx = fminf(mnx, v.x); mny = fminf(mny, v.y); mxx = fmaxf(mxx, v.x); mxy = fmaxf(mxy, v.y);
    p.pts[(long long)warp * p.S + s] = make_int2((int)v.x, (int)v.y);        // np.array(contour, dtype=np.int32)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) {
    const int x0 = (int)floorf(mnx), y0 = (int)floorf(mny), x1 = (int)ceilf(mxx), y1 = (int)ceilf(mxy);
    p.bbox[warp] = make_int4(x0, y0, x1, y1);
    p.ebox[warp] = make_float4((float)(x0 - p.gap), (float)(y0 - p.gap), (float)(x1 + p.gap + 1), (float)(y1 + p.gap + 1));
    p.done[warp] = 0u;
  }
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- pass 2: ordered painting ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(C2L_WARPS * 32) c2l_paint_kernel(const C2LParams p, const GridBins bins) {
  extern __shared__ long long c2l_sm[];                 // per warp: [S] edge slopes (16.16), then [S] vertices
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  long long* slope = c2l_sm + (size_t)wib * 2 * p.S;    // slope[e]: dx per scan line of edge (v[e-1] -> v[e])
  int2* v = reinterpret_cast<int2*>(slope + p.S);
  const int gw = blockIdx.x * C2L_WARPS + wib, nw = gridDim.x * C2L_WARPS;
  const GridInfo g = *bins.gi;
  const int C = p.C;
  for (int k = gw; k < p.K; k += nw) {
    __syncwarp();
    for (int s = lane; s < p.S; s += 32) v[s] = p.pts[(long long)k * p.S + s];
    __syncwarp();
    for (int e = lane; e < p.S; e += 32) {              // OpenCV: edge.dx = (x1 - x0) / (y1 - y0), truncating, 16.16
      const int2 a = v[e == 0 ? p.S - 1 : e - 1], b = v[e];
      slope[e] = a.y == b.y ? 0 : (((long long)b.x - a.x) * XY_ONE) / ((long long)b.y - a.y);
    }
    const int4 bb = p.bbox[k];
    const int ex0 = bb.x - p.gap, ey0 = bb.y - p.gap, ex1 = bb.z + p.gap, ey1 = bb.w + p.gap;   // inclusive
    __syncwarp();
    // (1) wait for every earlier contour whose bounding box meets the dilated box of this one
    {
      int cx, cy;
      cell_of(g, p.ebox[k], &cx, &cy);
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = cy + dy;
        if (yy < 0 || yy >= g.ncy) continue;
        const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, g.ncx - 1);
        const uint64_t k_lo = (uint64_t)(uint32_t)(yy * g.ncx + x_lo) << 32;
        const uint64_t k_hi = (uint64_t)(uint32_t)(yy * g.ncx + x_hi + 1) << 32;
        int lo = 0, hi = p.K;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (bins.cell_keys[mid] < k_lo) lo = mid + 1; else hi = mid; }
        int end = lo, hi2 = p.K;
        while (end < hi2) { const int mid = (end + hi2) >> 1; if (bins.cell_keys[mid] < k_hi) end = mid + 1; else hi2 = mid; }
        for (int q = lo + lane; q < end; q += 32) {
          const int j = bins.cell_rows[q];
          if (j >= k) continue;
          const int4 bj = p.bbox[j];
          if (bj.x > ex1 || bj.z < ex0 || bj.y > ey1 || bj.w < ey0) continue;
          long long spins = 0;
          while (ld_acquire(p.done + j) == 0u) {
            __nanosleep(64);
            if (++spins > (1ll << 22)) { atomicOr(p.info + 2, 1); break; }     // never expected: report, do not hang
          }
        }
      }
      __syncwarp();
    }
    // (2) first channel whose dilated-box region is empty (labels[region] > 0).sum((0, 1)) == 0
    const int rx0 = max(ex0, 0), ry0 = max(ey0, 0), rx1 = min(ex1, p.W - 1), ry1 = min(ey1, p.H - 1);
    unsigned long long occ = 0;
    if (rx1 >= rx0 && ry1 >= ry0) {
      const int rw = rx1 - rx0 + 1;
      const long long npx = (long long)rw * (ry1 - ry0 + 1);
      for (long long i = lane; i < npx; i += 32) {
        const int y = ry0 + (int)(i / rw), x = rx0 + (int)(i % rw);
        const int32_t* px = p.labels + ((long long)y * p.W + x) * C;
        if ((C & 3) == 0) {
          for (int c = 0; c < C; c += 4) {
            const int4 l = __ldcg(reinterpret_cast<const int4*>(px + c));
            occ |= (unsigned long long)((l.x > 0 ? 1u : 0u) | (l.y > 0 ? 2u : 0u) | (l.z > 0 ? 4u : 0u) | (l.w > 0 ? 8u : 0u)) << c;
          }
        } else {
          for (int c = 0; c < C; ++c) occ |= (unsigned long long)(__ldcg(px + c) > 0 ? 1u : 0u) << c;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) occ |= __shfl_xor_sync(0xffffffffu, occ, o);
    const int ch = __ffsll((long long)~occ) - 1;         // first free channel (-1: all 64 mask bits taken)
    if (ch < 0 || ch >= C) {
      if (lane == 0) { atomicMax(p.info + 1, (ch < 0 ? C2L_MAX_CH : ch) + 1); }
    } else {
      if (lane == 0) { atomicMax(p.info + 0, ch + 1); atomicMax(p.info + 1, ch + 1); }
      const int32_t lbl = k + 1;
      int32_t* L = p.labels + ch;
      // (3a) boundary: cv::Line, 8-connected Bresenham from the left end point (LineIterator leftToRight)
      for (int e = lane; e < p.S; e += 32) {
        int2 a = v[e == 0 ? p.S - 1 : e - 1], b = v[e];
        int dx = b.x - a.x, dy = b.y - a.y;
        if (dx < 0) { const int2 t = a; a = b; b = t; dx = -dx; dy = -dy; }
        const int sy = dy >= 0 ? 1 : -1;
        dy = dy >= 0 ? dy : -dy;
        const bool steep = dy > dx;
        if (steep) { const int t = dx; dx = dy; dy = t; }
        int err = dx - (dy + dy);
        const int plus = dx + dx, minus = -(dy + dy);
        int x = a.x, y = a.y;
        for (int i = 0; i <= dx; ++i) {
          if ((unsigned)x < (unsigned)p.W && (unsigned)y < (unsigned)p.H) L[((long long)y * p.W + x) * C] = lbl;
          const bool neg = err < 0;
          err += minus + (neg ? plus : 0);
          if (steep) { y += sy; if (neg) x += 1; } else { x += 1; if (neg) y += sy; }
        }
      }
      // (3b) interior: FillEdgeCollection.  Row y takes the crossings x_e(y) = x_e(y0) + dx_e * (y - y0) of the edges
      // with y0 <= y < y1 in ascending order and fills [ceil(x_1), floor(x_2)], [ceil(x_3), floor(x_4)], ...
      int ymin = 0x7fffffff, ymax = -0x7fffffff;
      for (int s = lane; s < p.S; s += 32) { ymin = min(ymin, v[s].y); ymax = max(ymax, v[s].y); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
      }
      for (int y = ymin + lane; y < ymax; y += 32) {
        if ((unsigned)y >= (unsigned)p.H) continue;
        long long last_x = -(1ll << 62);
        int last_e = -1;
        bool open = false;
        long long xa = 0;
        for (;;) {
          // next crossing in (x, edge index) order after (last_x, last_e)
          long long best_x = 1ll << 62;
          int best_e = -1;
          for (int e = 0; e < p.S; ++e) {
            const int2 a = v[e == 0 ? p.S - 1 : e - 1], b = v[e];
            if (a.y == b.y) continue;
            const int2 lo = a.y < b.y ? a : b, hi = a.y < b.y ? b : a;
            if (y < lo.y || y >= hi.y) continue;
            const long long xe = (long long)lo.x * XY_ONE + slope[e] * (y - lo.y);
            if ((xe > last_x || (xe == last_x && e > last_e)) && (xe < best_x || (xe == best_x && e < best_e))) {
              best_x = xe; best_e = e;
            }
          }
          if (best_e < 0) break;
          last_x = best_x; last_e = best_e;
          if (!open) { xa = best_x; open = true; }
          else {
            open = false;
            int x1 = (int)((xa + XY_ONE - 1) >> 16), x2 = (int)(best_x >> 16);
            if (x1 < p.W && x2 >= 0) {
              x1 = max(x1, 0); x2 = min(x2, p.W - 1);
              for (int x = x1; x <= x2; ++x) L[((long long)y * p.W + x) * C] = lbl;
            }
          }
        }
      }
    }
    // (4) publish
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(p.done + k, 1u);
  }
}

static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace cpn

using namespace cpn;

extern "C" size_t cpn_contours2labels_workspace_bytes(int64_t K, int samples) {
  const size_t k = (size_t)(K > 0 ? K : 1), s = (size_t)(samples > 0 ? samples : 1);
  return al(k * s * sizeof(int2)) + al(k * sizeof(int4)) + al(k * sizeof(float4)) + al(k * sizeof(uint32_t)) +
         al(grid_bin_workspace_bytes(K)) + 1024;
}

extern "C" int cpn_contours2labels(const float* contours, int64_t K, int samples, int H, int W, int rounded, int clip,
                                   int gap, int32_t* labels, int channels, void* workspace, int32_t* info_dev,
                                   void* stream) {
  CPN_REQUIRE(H > 0 && W > 0 && samples >= 1 && samples <= 4096, "contours2labels: bad size / samples");
  CPN_REQUIRE(channels >= 1 && channels <= C2L_MAX_CH, "contours2labels: channels %d must be in [1, %d]", channels,
              C2L_MAX_CH);
  CPN_REQUIRE(K >= 0 && K < (1ll << 31) - 1 && gap >= 0, "contours2labels: bad K / gap");
  CPN_REQUIRE(labels && info_dev && ((uintptr_t)labels % 16) == 0, "contours2labels: labels must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CPN_CHECK_CUDA(cudaMemsetAsync(labels, 0, (size_t)H * W * channels * sizeof(int32_t), st));
  CPN_CHECK_CUDA(cudaMemsetAsync(info_dev, 0, 4 * sizeof(int32_t), st));
  if (K == 0) return 0;
  char* b = reinterpret_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = b + off; off += al(bytes); return q; };
  C2LParams p;
  memset(&p, 0, sizeof(p));
  p.contours = contours; p.K = (int)K; p.S = samples; p.H = H; p.W = W; p.rounded = rounded; p.clip = clip; p.gap = gap;
  p.C = channels;
  p.pts = (int2*)take((size_t)K * samples * sizeof(int2));
  p.bbox = (int4*)take((size_t)K * sizeof(int4));
  p.ebox = (float4*)take((size_t)K * sizeof(float4));
  p.done = (uint32_t*)take((size_t)K * sizeof(uint32_t));
  void* grid_ws = take(grid_bin_workspace_bytes(K));
  p.labels = labels; p.info = info_dev;
  c2l_prep_kernel<<<(int)((K * 32 + 255) / 256), 256, 0, st>>>(p);
  CPN_CHECK_LAUNCH();
  GridBins bins;
  if (grid_bin_boxes(p.ebox, (int)K, grid_ws, &bins, st)) return 1;
  // persistent grid: every warp must be resident (ordered spin-wait), so size it by occupancy
  const size_t smem = (size_t)C2L_WARPS * samples * (sizeof(int2) + sizeof(long long));
  static bool attr_set = false;
  if (!attr_set) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(c2l_paint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  CPN_REQUIRE(smem <= 200 * 1024, "contours2labels: %d samples per contour exceed the shared-memory budget", samples);
  int per_sm = 0;
  CPN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c2l_paint_kernel, C2L_WARPS * 32, smem));
  CPN_REQUIRE(per_sm >= 1, "contours2labels: kernel does not fit");
  long long grid = (long long)per_sm * sm_count();
  const long long need = (K + C2L_WARPS - 1) / C2L_WARPS;
  if (grid > need) grid = need;
  c2l_paint_kernel<<<(int)grid, C2L_WARPS * 32, smem, st>>>(p, bins);
  CPN_CHECK_LAUNCH();
  return 0;
}

// =====================================================================================================================
// resolve_label_channels (data/cpn.py:361-398, method='dilation', 3x3 cross kernel): channel image [H,W,C] -> flat label
// image [H,W].  Pixels covered by exactly one object keep its label; pixels covered by several start at 0 and are filled
// by repeated grey dilation (max over the pixel and its 4 neighbours of the previous iterate -- Jacobi sweeps, exactly
// like `lbl[m] = cv2.dilate(lbl)[m]`) until nothing changes or max_iter sweeps ran.  HBM-bound integer work.
// =====================================================================================================================
namespace cpn {

__global__ void rlc_init_kernel(const int32_t* __restrict__ labels, long long pixels, int C, int32_t* __restrict__ flat,
                                uint8_t* __restrict__ overlap, int* __restrict__ flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const int32_t* px = labels + i * C;
  int cnt = 0;
  int32_t mx = px[0];
  for (int c = 0; c < C; ++c) { const int32_t v = px[c]; cnt += v > 0 ? 1 : 0; mx = max(mx, v); }
  // no overlap anywhere: the result is labels.max(-1); otherwise cores keep max, everything else starts at 0
  overlap[i] = cnt > 1 ? 1 : 0;
  flat[i] = cnt > 1 ? 0 : mx;             // provisional: plain max for non-overlap pixels (fixed up below if needed)
  if (cnt > 1) atomicOr(flags, 1);
  if (cnt == 0 && mx != 0) atomicOr(flags, 2);   // negative labels present (max over a pixel without foreground)
}

// with overlaps the reference zero-fills every pixel that is not a core (mask_sm == 1), including negative labels
__global__ void rlc_fix_kernel(const int32_t* __restrict__ labels, long long pixels, int C, int32_t* __restrict__ flat) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const int32_t* px = labels + i * C;
  int cnt = 0;
  for (int c = 0; c < C; ++c) cnt += px[c] > 0 ? 1 : 0;
  if (cnt == 0) flat[i] = 0;
}

__global__ void rlc_sweep_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                 const uint8_t* __restrict__ overlap, int H, int W, int* __restrict__ changed) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  int32_t v = in[i];
  if (overlap[i] && v <= 0) {
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    int32_t m = v;
    if (y > 0) m = max(m, in[i - W]);
    if (y + 1 < H) m = max(m, in[i + W]);
    if (x > 0) m = max(m, in[i - 1]);
    if (x + 1 < W) m = max(m, in[i + 1]);
    if (m != v) { v = m; *changed = 1; }
  }
  out[i] = v;
}

}  // namespace cpn

extern "C" size_t cpn_resolve_label_channels_workspace_bytes(int H, int W) {
  return al((size_t)H * W * sizeof(int32_t)) + al((size_t)H * W) + 1024;
}

extern "C" int cpn_resolve_label_channels(const int32_t* labels, int H, int W, int channels, int max_iter, int32_t* flat,
                                          void* workspace, int* sweeps_host, void* stream) {
  CPN_REQUIRE(labels && flat && workspace && H > 0 && W > 0 && channels >= 1, "resolve_label_channels: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long pixels = (long long)H * W;
  char* b = reinterpret_cast<char*>(workspace);
  int32_t* tmp = reinterpret_cast<int32_t*>(b);
  uint8_t* overlap = reinterpret_cast<uint8_t*>(b + al((size_t)pixels * sizeof(int32_t)));
  int* flags = reinterpret_cast<int*>(b + al((size_t)pixels * sizeof(int32_t)) + al((size_t)pixels));
  CPN_CHECK_CUDA(cudaMemsetAsync(flags, 0, 64, st));
  const int tb = 256;
  const int gb = (int)((pixels + tb - 1) / tb);
  rlc_init_kernel<<<gb, tb, 0, st>>>(labels, pixels, channels, flat, overlap, flags);
  CPN_CHECK_LAUNCH();
  int h_flags = 0;
  CPN_CHECK_CUDA(cudaMemcpyAsync(&h_flags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CPN_CHECK_CUDA(cudaStreamSynchronize(st));
  int sweeps = 0;
  if (h_flags & 1) {
    if (h_flags & 2) { rlc_fix_kernel<<<gb, tb, 0, st>>>(labels, pixels, channels, flat); CPN_CHECK_LAUNCH(); }
    int32_t *cur = flat, *nxt = tmp;
    int* changed = flags + 4;
    while (sweeps < max_iter) {
      CPN_CHECK_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), st));
      rlc_sweep_kernel<<<gb, tb, 0, st>>>(cur, nxt, overlap, H, W, changed);
      CPN_CHECK_LAUNCH();
      ++sweeps;
      int32_t* t = cur; cur = nxt; nxt = t;
      int h_changed = 0;
      CPN_CHECK_CUDA(cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
      CPN_CHECK_CUDA(cudaStreamSynchronize(st));
      if (!h_changed) break;
    }
    if (cur != flat) CPN_CHECK_CUDA(cudaMemcpyAsync(flat, cur, (size_t)pixels * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  if (sweeps_host) *sweeps_host = sweeps;
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// LABEL PROPERTIES: the region statistics behind cd.data.labels2property_table (data/misc.py:320-345 -> third-party
// skimage.measure.regionprops_table), as used for the csv output of cpn_inference.py:824-837.  One pass over the label
// image (integer work, HBM-bound): per (channel, label) the pixel count, the bounding box and the coordinate sums, from
// which the host derives area, bbox, centroid, area_bbox, extent and equivalent_diameter_area exactly (integer sums,
// one float64 division).  Lanes that hold the same (channel, label) are merged with __match_any_sync and the hardware
// warp reductions before they reach memory: a run of equal labels along a row costs one set of atomics per warp.
// ---------------------------------------------------------------------------------------------------------------------
namespace cpn {
__global__ void __launch_bounds__(256) label_props_kernel(const int32_t* __restrict__ labels, long long n_px, int W, int C,
                                                          int max_label, uint32_t* __restrict__ area,
                                                          int32_t* __restrict__ bbox,            // [.., 4] min_r, min_c, max_r, max_c
                                                          unsigned long long* __restrict__ sums,  // [.., 2] sum_r, sum_c
                                                          int32_t* __restrict__ flags) {
  const long long n = n_px * C;
  const long long n_pad = (n + 31) / 32 * 32;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    int lab = 0;
    if (i < n) lab = __ldg(labels + i);
    if (lab > max_label) { atomicOr(flags, 1); lab = 0; }
    if (!__any_sync(0xffffffffu, lab > 0)) continue;          // background (most of a label image) costs one load
    const long long px = i / C;
    const int ch = (int)(i - px * C);
    const unsigned r = (unsigned)(px / W), c = (unsigned)(px - (long long)r * W);
    const unsigned key = lab > 0 ? (unsigned)ch * (unsigned)(max_label + 1) + (unsigned)lab : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (lab > 0) {
      const unsigned cnt = __popc(peers);
      const unsigned sr = __reduce_add_sync(peers, r), sc = __reduce_add_sync(peers, c);
      const unsigned r0 = __reduce_min_sync(peers, r), r1 = __reduce_max_sync(peers, r);
      const unsigned c0 = __reduce_min_sync(peers, c), c1 = __reduce_max_sync(peers, c);
      if ((unsigned)(__ffs(peers) - 1) == lane) {
        atomicAdd(area + key, cnt);
        atomicAdd(sums + 2ull * key, (unsigned long long)sr);
        atomicAdd(sums + 2ull * key + 1, (unsigned long long)sc);
        atomicMin(bbox + 4ull * key, (int)r0);
        atomicMin(bbox + 4ull * key + 1, (int)c0);
        atomicMax(bbox + 4ull * key + 2, (int)r1 + 1);
        atomicMax(bbox + 4ull * key + 3, (int)c1 + 1);
      }
    }
  }
}

__global__ void label_props_init_kernel(long long n, uint32_t* area, int32_t* bbox, unsigned long long* sums, int32_t* flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) flags[0] = 0;
  if (i >= n) return;
  area[i] = 0;
  sums[2 * i] = 0; sums[2 * i + 1] = 0;
  bbox[4 * i] = 0x7fffffff; bbox[4 * i + 1] = 0x7fffffff; bbox[4 * i + 2] = 0; bbox[4 * i + 3] = 0;
}
}  // namespace cpn

extern "C" int cpn_label_props(const int32_t* labels, int H, int W, int channels, int max_label, uint32_t* area,
                               int32_t* bbox, uint64_t* sums, int32_t* flags, void* stream) {
  using namespace cpn;
  CPN_REQUIRE(labels && area && bbox && sums && flags, "label_props: null pointer");
  CPN_REQUIRE(H >= 0 && W >= 0 && channels >= 1 && max_label >= 0, "label_props: bad shape");
  CPN_REQUIRE((long long)channels * ((long long)max_label + 1) < (1ll << 31), "label_props: channels * (max_label + 1) too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long long slots = (long long)channels * ((long long)max_label + 1);
  label_props_init_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(slots, area, bbox, (unsigned long long*)sums, flags);
  CPN_CHECK_LAUNCH();
  const long long n_px = (long long)H * W;
  if (n_px == 0) return 0;
  long long blocks = (n_px * channels + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  label_props_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels, n_px, W, channels, max_label, area, bbox,
                                                       (unsigned long long*)sums, flags);
  CPN_CHECK_LAUNCH();
  return 0;
}
