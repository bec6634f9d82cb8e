// Segmented greedy NMS with torchvision semantics (replaces torch.ops.torchvision.nms as called by
// /root/reference/celldetection/ops/cpn.py:189-227 and celldetection_scripts/cpn_inference.py:405-408).
//
// Semantics mirrored (torchvision 0.26): candidates in stable descending score order (ties keep
// the lower original index first); box j is suppressed by an earlier kept box i when
//   inter / (area_i + area_j - inter) > thr   (fp32, strict, NaN never suppresses; areas (x2-x1)*(y2-y1), no +1);
// the result lists kept indices in descending score order.  The 50 000-chunk rule of batched_box_nmsi is reproduced
// with two passes (chunks of the ORIGINAL row order, then NMS over the concatenated survivors).
//
// Algorithm: one stable radix sort of (segment, ~score) keys (radix_sort_pairs below), then one CTA per segment runs a
// blocked greedy scan with O(P) memory: 256 candidates at a time are (a) tested against all previously kept boxes
// (staged through shared memory), (b) cross-tested inside the tile into a 256x256 bit matrix, (c) resolved by one warp
// that jumps from kept box to kept box (__ffs over the alive bitmap), so the sequential chain is K long, not P long.
// Compiled with -fmad=false so the IoU arithmetic rounds exactly like the reference's.
#include "common.cuh"

namespace cpn {

constexpr int NMS_T = 256;  // candidates per tile == threads per CTA

// ---------------------------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (uint64 key, int32 value) pairs, 8 bits per pass over bits [0, end_bit): the ordering step of
// every NMS variant here (score order, (segment, score) order, (grid cell, rank) order).  Three kernels per pass: per-block
// digit histograms (shared-memory atomics), one exclusive scan over (digit-major, block-minor) counts, and a scatter in
// which every block walks its 4096-pair chunk in order -- ranks among equal digits come from __match_any_sync inside a warp
// and per-warp digit counts across warps, so equal keys keep their input order (the NMS tie-break depends on it).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256, RS_TILES = 16, RS_CHUNK = RS_THREADS * RS_TILES;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, int n, int shift, uint32_t mask,
                                                             uint32_t* __restrict__ counts, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_CHUNK;
#pragma unroll 4
  for (int t = 0; t < RS_TILES; ++t) {
    const int i = base + t * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ counts, int total) {
  __shared__ uint32_t part[1024];
  const int per = (total + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, total);
  uint32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += counts[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {               // Hillis-Steele inclusive scan of the 1024 partial sums
    const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;                  // exclusive prefix of this thread's segment
  for (int i = lo; i < hi; ++i) { const uint32_t c = counts[i]; counts[i] = run; run += c; }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int n,
                                                                int shift, uint32_t mask, const uint32_t* __restrict__ offsets,
                                                                int nblocks) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t wcnt[RS_THREADS / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  base[threadIdx.x] = offsets[(size_t)threadIdx.x * nblocks + blockIdx.x];
  for (int w = 0; w < RS_THREADS / 32; ++w) wcnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int start = blockIdx.x * RS_CHUNK;
  for (int t = 0; t < RS_TILES; ++t) {
    const int i = start + t * RS_THREADS + threadIdx.x;
    if (start + t * RS_THREADS >= n) break;                 // block-uniform
    const bool ok = i < n;
    uint64_t k = 0; int32_t v = 0;
    if (ok) { k = keys_in[i]; v = vals_in[i]; }
    const uint32_t d = ok ? ((uint32_t)(k >> shift) & mask) : (0x100u + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (ok) {
      uint32_t pos = base[d] + rank;
      for (int w = 0; w < warp; ++w) pos += wcnt[w][d];
      keys_out[pos] = k;
      vals_out[pos] = v;
    }
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) { add += wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = 0; }
    base[threadIdx.x] += add;
    __syncthreads();
  }
}

static inline size_t rs_align(size_t v) { return (v + 255) / 256 * 256; }

size_t radix_sort_tmp_bytes(int64_t n) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const size_t nblocks = (nn + RS_CHUNK - 1) / RS_CHUNK;
  return rs_align(256 * nblocks * 4) + rs_align(nn * 8) + rs_align(nn * 4) + 256;
}

// keys_in / vals_in are not modified; the sorted pairs land in keys_out / vals_out.
int radix_sort_pairs(void* tmp, const uint64_t* keys_in, uint64_t* keys_out, const int32_t* vals_in, int32_t* vals_out, int n,
                     int end_bit, cudaStream_t st) {
  if (n <= 0) return 0;
  const int nblocks = (n + RS_CHUNK - 1) / RS_CHUNK;
  char* b = reinterpret_cast<char*>(tmp);
  uint32_t* counts = reinterpret_cast<uint32_t*>(b);
  uint64_t* keys_t = reinterpret_cast<uint64_t*>(b + rs_align(256 * (size_t)nblocks * 4));
  int32_t* vals_t = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(keys_t) + rs_align((size_t)n * 8));
  const int passes = end_bit <= 0 ? 1 : (end_bit + 7) / 8;
  const uint64_t* ksrc = keys_in;
  const int32_t* vsrc = vals_in;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    const int bits = end_bit - shift < 8 ? (end_bit - shift > 0 ? end_bit - shift : 8) : 8;
    const uint32_t mask = (1u << bits) - 1u;
    const bool to_out = ((passes - 1 - p) & 1) == 0;
    uint64_t* kdst = to_out ? keys_out : keys_t;
    int32_t* vdst = to_out ? vals_out : vals_t;
    rs_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(ksrc, n, shift, mask, counts, nblocks);
    CPN_CHECK_LAUNCH();
    rs_scan_kernel<<<1, 1024, 0, st>>>(counts, 256 * nblocks);
    CPN_CHECK_LAUNCH();
    rs_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(ksrc, vsrc, kdst, vdst, n, shift, mask, counts, nblocks);
    CPN_CHECK_LAUNCH();
    count_launch(3);
    ksrc = kdst;
    vsrc = vdst;
  }
  return 0;
}

__device__ __forceinline__ uint32_t float_desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-order key
  return ~u;                                       // descending
}

__device__ __forceinline__ int find_segment(const int32_t* __restrict__ seg_offsets, int n_segments, long long row) {
  int lo = 0, hi = n_segments;  // largest s with seg_offsets[s] <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)seg_offsets[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// keys for pass A.  sub-segment id = sub_base[seg] + (row - seg_start) / chunk   (chunk <= 0: one sub per segment)
__global__ void nms_keys_kernel(const float* __restrict__ scores, const int32_t* __restrict__ seg_offsets,
                                int n_segments, long long n, int chunk, uint64_t* __restrict__ keys,
                                int32_t* __restrict__ vals) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int seg = find_segment(seg_offsets, n_segments, i);
  uint32_t sub = (uint32_t)seg;
  if (chunk > 0) {
    uint32_t base = 0;
    for (int s = 0; s < seg; ++s) {
      const int sz = seg_offsets[s + 1] - seg_offsets[s];
      base += sz > 0 ? (uint32_t)((sz + chunk - 1) / chunk) : 1u;
    }
    sub = base + (uint32_t)((i - seg_offsets[seg]) / chunk);
  }
  keys[i] = ((uint64_t)sub << 32) | float_desc_key(scores[i]);
  vals[i] = (int32_t)i;
}

// keys for pass B: rows are the survivors of pass A (in sub-segment order); group by the ORIGINAL segment
__global__ void nms_keys_rows_kernel(const float* __restrict__ scores, const int32_t* __restrict__ seg_offsets,
                                     int n_segments, const int32_t* __restrict__ rows,
                                     const long long* __restrict__ n_ptr, uint64_t* __restrict__ keys,
                                     int32_t* __restrict__ vals, long long capacity) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= capacity) return;
  if (i >= *n_ptr) {  // padding sorts to the very end
    keys[i] = ~0ull;
    vals[i] = -1;
    return;
  }
  const int r = rows[i];
  const int seg = find_segment(seg_offsets, n_segments, r);
  keys[i] = ((uint64_t)(uint32_t)seg << 32) | float_desc_key(scores[r]);
  vals[i] = r;
}

// group_start[g] = lower_bound over the high 32 bits of the sorted keys, g in [0, n_groups]
__global__ void nms_bounds_kernel(const uint64_t* __restrict__ keys, long long n, int n_groups,
                                  int32_t* __restrict__ group_start) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_groups) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((uint32_t)(keys[mid] >> 32) < (uint32_t)g) lo = mid + 1; else hi = mid;
  }
  group_start[g] = (int32_t)lo;
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const float area_b,
                                       const float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
  const float inter = w * h;
  const float ovr = inter / (area_a + area_b - inter);
  return ovr > thr;
}

// One CTA per group (sorted rows perm[group_start[g] .. group_start[g+1])).
// kept rows are written to keep_out at out_offsets[g] (== group_start[g] when out_offsets is NULL); counts to counts[g].
__global__ void __launch_bounds__(NMS_T) nms_group_kernel(const float* __restrict__ boxes,
                                                          const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ group_start,
                                                          const int32_t* __restrict__ out_offsets, float thr,
                                                          float4* __restrict__ kept_boxes,
                                                          int32_t* __restrict__ keep_out,
                                                          int32_t* __restrict__ counts) {
  __shared__ float4 kb[NMS_T];
  __shared__ float4 cand[NMS_T];
  __shared__ int32_t crow[NMS_T];
  __shared__ uint32_t mask[NMS_T][NMS_T / 32];
  __shared__ uint32_t alive_bm[NMS_T / 32];
  __shared__ int nkept_s;

  const int g = blockIdx.x;
  const int start = group_start[g];
  const int n = group_start[g + 1] - start;
  const int out0 = out_offsets ? out_offsets[g] : start;
  const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
  if (j == 0) nkept_s = 0;
  __syncthreads();

  for (int t0 = 0; t0 < n; t0 += NMS_T) {
    const int ntile = min(NMS_T, n - t0);
    const bool valid = j < ntile;
    int row = -1;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      row = perm[start + t0 + j];
      box = reinterpret_cast<const float4*>(boxes)[row];
    }
    const float area = (box.z - box.x) * (box.w - box.y);
    bool alive = valid;
    const int nkept = nkept_s;
    // (a) against everything kept so far
    for (int k0 = 0; k0 < nkept; k0 += NMS_T) {
      const int m = min(NMS_T, nkept - k0);
      if (j < m) kb[j] = kept_boxes[out0 + k0 + j];
      __syncthreads();
      if (alive) {
        for (int k = 0; k < m; ++k) {
          const float4 b = kb[k];
          if (iou_gt(b, (b.z - b.x) * (b.w - b.y), box, area, thr)) { alive = false; break; }
        }
      }
      __syncthreads();
    }
    // (b) tile-internal bit matrix (rows of dead candidates are never read)
    cand[j] = box;
    crow[j] = row;
    const uint32_t bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) alive_bm[warp] = bal;
    __syncthreads();
    if (alive) {
#pragma unroll 1
      for (int wd = 0; wd < NMS_T / 32; ++wd) {
        uint32_t bits = 0;
        if (wd * 32 + 31 > j) {
          for (int b = 0; b < 32; ++b) {
            const int k = wd * 32 + b;
            if (k > j && k < ntile) {
              const float4 c = cand[k];
              if (iou_gt(box, area, c, (c.z - c.x) * (c.w - c.y), thr)) bits |= 1u << b;
            }
          }
        }
        mask[j][wd] = bits;
      }
    }
    __syncthreads();
    // (c) resolve: warp 0, lanes 0..7 own one 32-bit word of the removed bitmap each
    if (warp == 0) {
      uint32_t rem = (lane < NMS_T / 32) ? ~alive_bm[lane] : 0xffffffffu;
      int nk = nkept;
      while (true) {
        const uint32_t z = ~rem;
        int first = z ? (lane * 32 + __ffs(z) - 1) : 1 << 20;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= ntile) break;
        if (lane == 0) {
          keep_out[out0 + nk] = crow[first];
          kept_boxes[out0 + nk] = cand[first];
        }
        ++nk;
        if (lane < NMS_T / 32) {
          rem |= mask[first][lane];
          if ((first >> 5) == lane) rem |= 1u << (first & 31);
        }
      }
      if (lane == 0) nkept_s = nk;
    }
    __syncthreads();
  }
  if (j == 0) counts[g] = nkept_s;
}

// pass-B glue: compact the per-sub-segment survivor lists (sub-segment order) into rows2, total -> n2
__global__ void nms_compact_kernel(const int32_t* __restrict__ keepA, const int32_t* __restrict__ sub_start,
                                   const int32_t* __restrict__ countsA, int n_sub, int32_t* __restrict__ rows2,
                                   long long* __restrict__ n2) {
  const int s = blockIdx.x;
  long long off = 0;
  for (int q = 0; q < s; ++q) off += countsA[q];
  const int c = countsA[s];
  for (int i = threadIdx.x; i < c; i += blockDim.x) rows2[off + i] = keepA[sub_start[s] + i];
  if (s == n_sub - 1 && threadIdx.x == 0) *n2 = off + c;
}

// number of sub-segments on the device (pass B needs the true count for its grid guard)
__global__ void nms_count_subs_kernel(const int32_t* __restrict__ seg_offsets, int n_segments, int chunk,
                                      int32_t* __restrict__ n_sub_out) {
  if (threadIdx.x || blockIdx.x) return;
  int t = 0;
  for (int s = 0; s < n_segments; ++s) {
    const int sz = seg_offsets[s + 1] - seg_offsets[s];
    t += sz > 0 ? (sz + chunk - 1) / chunk : 1;
  }
  *n_sub_out = t;
}

__global__ void nms_zero_counts_kernel(int32_t* __restrict__ counts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = 0;
}

struct NmsWs {
  uint64_t *keys_a, *keys_b;
  int32_t *vals_a, *vals_b;
  float4* kept_boxes;
  int32_t *keepA, *rows2, *group_start, *countsA;
  long long* n2;
  void* sort_tmp;
  size_t sort_bytes;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int capped_blocks(int blocks) { const int cap = sm_count() * 8; return blocks > cap ? cap : (blocks < 1 ? 1 : blocks); }

static size_t carve(void* base, int64_t n, int n_segments, NmsWs* ws) {
  const int64_t n_sub_max = (int64_t)n_segments + n / 1024 + 2;  // chunk >= 1024 enforced
  size_t off = 0;
  char* b = reinterpret_cast<char*>(base);
  auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off = align_up(off + bytes, 256); return p; };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  uint64_t* ka = (uint64_t*)take(nn * 8);
  uint64_t* kb2 = (uint64_t*)take(nn * 8);
  int32_t* va = (int32_t*)take(nn * 4);
  int32_t* vb = (int32_t*)take(nn * 4);
  float4* kept = (float4*)take(nn * 16);
  int32_t* keepA = (int32_t*)take(nn * 4);
  int32_t* rows2 = (int32_t*)take(nn * 4);
  int32_t* gs = (int32_t*)take((size_t)(n_sub_max + 2) * 4);
  int32_t* ca = (int32_t*)take((size_t)(n_sub_max + 2) * 4);
  long long* n2 = (long long*)take(64);
  const size_t sort_bytes = radix_sort_tmp_bytes((int64_t)nn);
  void* tmp = take(sort_bytes + 256);
  if (ws) {
    ws->keys_a = ka; ws->keys_b = kb2; ws->vals_a = va; ws->vals_b = vb; ws->kept_boxes = kept; ws->keepA = keepA;
    ws->rows2 = rows2; ws->group_start = gs; ws->countsA = ca; ws->n2 = n2; ws->sort_tmp = tmp;
    ws->sort_bytes = sort_bytes + 256;
  }
  return off;
}

}  // namespace cpn

using namespace cpn;

extern "C" size_t cpn_nms_workspace_bytes(int64_t n_boxes, int n_segments) {
  return carve(nullptr, n_boxes, n_segments, nullptr);
}

extern "C" int cpn_nms_segments(const float* boxes, const float* scores, const int32_t* seg_offsets, int n_segments,
                                int64_t n_boxes, float iou_threshold, int chunk, void* workspace, int32_t* keep,
                                int32_t* keep_counts, void* stream) {
  CPN_REQUIRE(n_segments >= 1, "nms: need at least one segment");
  CPN_REQUIRE(n_boxes < (1ll << 31), "nms: too many boxes");
  CPN_REQUIRE(chunk <= 0 || chunk >= 1024, "nms: chunk must be <= 0 (off) or >= 1024");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_boxes <= 0) {
    nms_zero_counts_kernel<<<(n_segments + 127) / 128, 128, 0, st>>>(keep_counts, n_segments);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  NmsWs ws;
  carve(workspace, n_boxes, n_segments, &ws);
  const int n = (int)n_boxes;
  const bool two_pass = chunk > 0 && n_boxes > chunk;
  const int tb = 256, gb = (n + tb - 1) / tb;

  // ---- pass A: sort by (sub-segment, score desc), NMS per sub-segment ----
  nms_keys_kernel<<<gb, tb, 0, st>>>(scores, seg_offsets, n_segments, n, two_pass ? chunk : 0, ws.keys_a, ws.vals_a);
  CPN_CHECK_LAUNCH();
  const int n_groups_a = two_pass ? (int)(n_segments + n_boxes / chunk + 1) : n_segments;
  int end_bit = 32;
  while ((1ll << (end_bit - 32)) < n_groups_a + 1 && end_bit < 64) ++end_bit;
  if (radix_sort_pairs(ws.sort_tmp, ws.keys_a, ws.keys_b, ws.vals_a, ws.vals_b, n, end_bit, st)) return 1;
  nms_bounds_kernel<<<(n_groups_a + 1 + 127) / 128, 128, 0, st>>>(ws.keys_b, n, n_groups_a, ws.group_start);
  CPN_CHECK_LAUNCH();
  if (!two_pass) {
    nms_group_kernel<<<n_groups_a, NMS_T, 0, st>>>(boxes, ws.vals_b, ws.group_start, nullptr, iou_threshold,
                                                   ws.kept_boxes, keep, keep_counts);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  nms_group_kernel<<<n_groups_a, NMS_T, 0, st>>>(boxes, ws.vals_b, ws.group_start, nullptr, iou_threshold,
                                                 ws.kept_boxes, ws.keepA, ws.countsA);
  CPN_CHECK_LAUNCH();
  // ---- pass B: survivors (sub-segment order) -> sort by (segment, score desc) -> NMS per segment ----
  nms_compact_kernel<<<n_groups_a, 256, 0, st>>>(ws.keepA, ws.group_start, ws.countsA, n_groups_a, ws.rows2, ws.n2);
  CPN_CHECK_LAUNCH();
  nms_keys_rows_kernel<<<gb, tb, 0, st>>>(scores, seg_offsets, n_segments, ws.rows2, ws.n2, ws.keys_a, ws.vals_a, n);
  CPN_CHECK_LAUNCH();
  if (radix_sort_pairs(ws.sort_tmp, ws.keys_a, ws.keys_b, ws.vals_a, ws.vals_b, n, 64, st)) return 1;
  nms_bounds_kernel<<<(n_segments + 1 + 127) / 128, 128, 0, st>>>(ws.keys_b, n, n_segments, ws.group_start);
  CPN_CHECK_LAUNCH();
  // results are packed at each segment's ORIGINAL offset (seg_offsets), as documented
  nms_group_kernel<<<n_segments, NMS_T, 0, st>>>(boxes, ws.vals_b, ws.group_start, seg_offsets, iou_threshold,
                                                 ws.kept_boxes, keep, keep_counts);
  CPN_CHECK_LAUNCH();
  return 0;
}

// =====================================================================================================================
// Grid NMS: exact greedy NMS for ONE large segment (the global stitch of a tiled whole-slide image,
// celldetection_scripts/cpn_inference.py:405-408, 1e5-1e6 boxes) in parallel.
//
// Greedy NMS keeps box i iff no KEPT box of higher priority overlaps it (IoU > thr).  That fixed point is computed by
// parallel rounds over a uniform spatial grid: a box becomes SUPPRESSED as soon as a higher-priority overlapping box is
// KEPT, and KEPT once every higher-priority overlapping box is SUPPRESSED; the highest-priority undecided box always
// resolves, so the rounds terminate, and the fixed point is unique = the sequential greedy result (same IoU arithmetic,
// same stable score order).  Cell size = the largest box extent, so overlapping boxes lie in 3x3 neighbouring cells.
// =====================================================================================================================
namespace cpn {

// pass 1: extents (min corner, max corner, max box size) via atomics on ordered-int floats
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void grid_extent_kernel(const float4* __restrict__ boxes, int n, int* __restrict__ ext) {
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY, ms = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 b = boxes[i];
    const float cx = 0.5f * (b.x + b.z), cy = 0.5f * (b.y + b.w);
    mnx = fminf(mnx, cx); mny = fminf(mny, cy); mxx = fmaxf(mxx, cx); mxy = fmaxf(mxy, cy);
    ms = fmaxf(ms, fmaxf(b.z - b.x, b.w - b.y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(ext + 0, f2ord(mnx)); atomicMin(ext + 1, f2ord(mny));
    atomicMax(ext + 2, f2ord(mxx)); atomicMax(ext + 3, f2ord(mxy)); atomicMax(ext + 4, f2ord(ms));
  }
}

__global__ void grid_setup_kernel(const int* __restrict__ ext, GridInfo* __restrict__ gi) {
  if (threadIdx.x || blockIdx.x) return;
  const float mnx = ord2f(ext[0]), mny = ord2f(ext[1]), mxx = ord2f(ext[2]), mxy = ord2f(ext[3]);
  float cell = fmaxf(ord2f(ext[4]), 1e-3f) * 1.0001f;
  const float spanx = fmaxf(mxx - mnx, 0.f), spany = fmaxf(mxy - mny, 0.f);
  cell = fmaxf(cell, fmaxf(spanx, spany) / 4000.f);   // at most ~4000 x 4000 cells
  gi->x0 = mnx; gi->y0 = mny; gi->inv_cell = 1.f / cell;
  gi->ncx = (int)(spanx / cell) + 1; gi->ncy = (int)(spany / cell) + 1;
}

// keys: (cell << 32) | rank, where rank = position in the stable descending score order; value = rank
__global__ void grid_cellkeys_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ perm, int n,
                                     const GridInfo* __restrict__ gi, uint64_t* __restrict__ keys,
                                     int32_t* __restrict__ vals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const GridInfo g = *gi;
  const int c = cell_of(g, boxes[perm[r]], nullptr, nullptr);
  keys[r] = ((uint64_t)(uint32_t)c << 32) | (uint32_t)r;
  vals[r] = r;
}

// state per rank: 0 undecided, 1 kept, 2 suppressed.  cell_rows: ranks sorted by (cell, rank).
__global__ void grid_round_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ perm,
                                  const uint64_t* __restrict__ cell_keys, const int32_t* __restrict__ cell_rows, int n,
                                  const GridInfo* __restrict__ gi, float thr, const uint8_t* __restrict__ state_in,
                                  uint8_t* __restrict__ state_out, int* __restrict__ undecided) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint8_t st = state_in[r];
  if (st != 0) { state_out[r] = st; return; }
  const GridInfo g = *gi;
  const float4 b = boxes[perm[r]];
  const float area = (b.z - b.x) * (b.w - b.y);
  int cx, cy;
  cell_of(g, b, &cx, &cy);
  bool any_kept = false, any_open = false;
  for (int dy = -1; dy <= 1 && !any_kept; ++dy) {
    const int yy = cy + dy;
    if (yy < 0 || yy >= g.ncy) continue;
    const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, g.ncx - 1);
    // the three cells of this row are contiguous in the (cell, rank) order
    const uint64_t k_lo = (uint64_t)(uint32_t)(yy * g.ncx + x_lo) << 32;
    const uint64_t k_hi = (uint64_t)(uint32_t)(yy * g.ncx + x_hi + 1) << 32;
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cell_keys[mid] < k_lo) lo = mid + 1; else hi = mid; }
    for (int q = lo; q < n && cell_keys[q] < k_hi; ++q) {
      const int rj = cell_rows[q];
      if (rj >= r) continue;                       // only higher-priority boxes can suppress
      const uint8_t sj = state_in[rj];
      if (sj == 2) continue;
      const float4 c = boxes[perm[rj]];
      if (!iou_gt(c, (c.z - c.x) * (c.w - c.y), b, area, thr)) continue;
      if (sj == 1) { any_kept = true; break; }
      any_open = true;
    }
  }
  uint8_t ns = 0;
  if (any_kept) ns = 2;
  else if (!any_open) ns = 1;
  state_out[r] = ns;
  if (ns == 0) atomicAdd(undecided, 1);
}

__global__ void grid_flags_kernel(const uint8_t* __restrict__ state, int n, int32_t* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flags[r] = state[r] == 1 ? 1 : 0;
}

// single block: exclusive scan of flags -> write kept rows (perm[r]) in rank order
__global__ void __launch_bounds__(1024) grid_compact_kernel(const int32_t* __restrict__ flags,
                                                            const int32_t* __restrict__ perm, int n,
                                                            int32_t* __restrict__ keep, int32_t* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s, chunk_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int r = b0 + threadIdx.x;
    const int v = r < n ? flags[r] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
      warp_tot[lane] = winc - w;
      if (lane == 31) chunk_s = winc;
    }
    __syncthreads();
    if (v) keep[carry_s + warp_tot[warp] + inc - v] = perm[r];
    __syncthreads();
    if (threadIdx.x == 0) carry_s += chunk_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry_s;
}

}  // namespace cpn

extern "C" size_t cpn_nms_grid_workspace_bytes(int64_t n_boxes) {
  const size_t nn = (size_t)(n_boxes > 0 ? n_boxes : 1);
  const size_t sort_bytes = radix_sort_tmp_bytes((int64_t)nn);
  // keys a/b (8), vals a/b (4), perm (4), cell_keys (8), cell_rows (4), state a/b (1), flags (4) + small
  return align_up(nn * 8, 256) * 3 + align_up(nn * 4, 256) * 5 + align_up(nn, 256) * 2 + align_up(sort_bytes + 256, 256) +
         4096;
}

extern "C" int cpn_nms_grid(const float* boxes, const float* scores, int64_t n_boxes, float iou_threshold,
                            void* workspace, int32_t* keep, int32_t* keep_count, int* rounds_host, void* stream) {
  CPN_REQUIRE(n_boxes < (1ll << 31), "nms_grid: too many boxes");
  cudaStream_t st = (cudaStream_t)stream;
  if (rounds_host) *rounds_host = 0;
  if (n_boxes <= 0) {
    nms_zero_counts_kernel<<<1, 32, 0, st>>>(keep_count, 1);
    CPN_CHECK_LAUNCH();
    return 0;
  }
  const int n = (int)n_boxes;
  const size_t nn = (size_t)n;
  char* b = reinterpret_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = b + off; off += align_up(bytes, 256); return p; };
  uint64_t* keys_a = (uint64_t*)take(nn * 8);
  uint64_t* keys_b = (uint64_t*)take(nn * 8);
  uint64_t* cell_keys = (uint64_t*)take(nn * 8);
  int32_t* vals_a = (int32_t*)take(nn * 4);
  int32_t* vals_b = (int32_t*)take(nn * 4);
  int32_t* perm = (int32_t*)take(nn * 4);
  int32_t* cell_rows = (int32_t*)take(nn * 4);
  int32_t* flags = (int32_t*)take(nn * 4);
  uint8_t* state_a = (uint8_t*)take(nn);
  uint8_t* state_b = (uint8_t*)take(nn);
  int* small = (int*)take(2048);   // [0..4] extents, [8] undecided, GridInfo at +64 bytes
  GridInfo* gi = reinterpret_cast<GridInfo*>(reinterpret_cast<char*>(small) + 64);
  const size_t sort_bytes = radix_sort_tmp_bytes(n);
  void* sort_tmp = take(sort_bytes + 256);
  const int tb = 256, gb = (n + tb - 1) / tb;
  // 1. stable descending score order -> perm[rank] = row
  const int32_t seg_host[2] = {0, n};
  int32_t* seg_dev = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(small) + 256);
  CPN_CHECK_CUDA(cudaMemcpyAsync(seg_dev, seg_host, sizeof(seg_host), cudaMemcpyHostToDevice, st));
  nms_keys_kernel<<<gb, tb, 0, st>>>(scores, seg_dev, 1, n, 0, keys_a, vals_a);
  CPN_CHECK_LAUNCH();
  if (radix_sort_pairs(sort_tmp, keys_a, keys_b, vals_a, perm, n, 33, st)) return 1;
  // 2. grid: extents -> cell size -> (cell, rank) order
  const int ext_init[5] = {0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  CPN_CHECK_CUDA(cudaMemcpyAsync(small, ext_init, sizeof(ext_init), cudaMemcpyHostToDevice, st));
  grid_extent_kernel<<<capped_blocks(gb), tb, 0, st>>>(reinterpret_cast<const float4*>(boxes), n, small);
  CPN_CHECK_LAUNCH();
  grid_setup_kernel<<<1, 32, 0, st>>>(small, gi);
  CPN_CHECK_LAUNCH();
  grid_cellkeys_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const float4*>(boxes), perm, n, gi, keys_a, vals_a);
  CPN_CHECK_LAUNCH();
  if (radix_sort_pairs(sort_tmp, keys_a, cell_keys, vals_a, cell_rows, n, 64, st)) return 1;
  // 3. rounds until every box is decided (the undecided count is read back every 4 rounds)
  CPN_CHECK_CUDA(cudaMemsetAsync(state_a, 0, nn, st));
  uint8_t *s_in = state_a, *s_out = state_b;
  int rounds = 0;
  int undecided = n;
  while (undecided > 0) {
    for (int k = 0; k < 4; ++k) {
      CPN_CHECK_CUDA(cudaMemsetAsync(small + 8, 0, sizeof(int), st));
      grid_round_kernel<<<gb, tb, 0, st>>>(reinterpret_cast<const float4*>(boxes), perm, cell_keys, cell_rows, n, gi,
                                           iou_threshold, s_in, s_out, small + 8);
      CPN_CHECK_LAUNCH();
      uint8_t* t = s_in; s_in = s_out; s_out = t;
      ++rounds;
    }
    CPN_CHECK_CUDA(cudaMemcpyAsync(&undecided, small + 8, sizeof(int), cudaMemcpyDeviceToHost, st));
    CPN_CHECK_CUDA(cudaStreamSynchronize(st));
    CPN_REQUIRE(rounds < 100000, "nms_grid: did not converge");
  }
  // 4. kept rows in descending score order
  grid_flags_kernel<<<gb, tb, 0, st>>>(s_in, n, flags);
  CPN_CHECK_LAUNCH();
  grid_compact_kernel<<<1, 1024, 0, st>>>(flags, perm, n, keep, keep_count);
  CPN_CHECK_LAUNCH();
  if (rounds_host) *rounds_host = rounds;
  return 0;
}

// =====================================================================================================================
// Box voting (ensembles): cd.ops.filter_by_box_voting / get_iou_voting, ops/boxes.py:53-83, as called by
// cpn_inference.py:419-424.   votes[i] = sum_j iou(i, j) * (iou(i, j) > thr)   over ALL boxes j including i itself
// (torchvision box_iou arithmetic: inter / (area_i + area_j - inter)).  The reference builds the dense K x K matrix; here
// only the 3x3 neighbouring cells of the same uniform grid the stitch NMS uses are visited -- every other pair has
// inter = 0 and contributes an exact zero.  A zero-area box has iou(i, i) = 0/0 = NaN and therefore a NaN vote, like the
// reference (NaN * False = NaN).
// =====================================================================================================================
namespace cpn {

__global__ void votes_cellkeys_kernel(const float4* __restrict__ boxes, int n, const GridInfo* __restrict__ gi,
                                      uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const GridInfo g = *gi;
  const int c = cell_of(g, boxes[r], nullptr, nullptr);
  keys[r] = ((uint64_t)(uint32_t)c << 32) | (uint32_t)r;
  vals[r] = r;
}

__global__ void votes_kernel(const float4* __restrict__ boxes, const uint64_t* __restrict__ cell_keys,
                             const int32_t* __restrict__ cell_rows, int n, const GridInfo* __restrict__ gi, float thr,
                             float* __restrict__ votes) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const GridInfo g = *gi;
  const float4 b = boxes[r];
  const float area = (b.z - b.x) * (b.w - b.y);
  int cx, cy;
  cell_of(g, b, &cx, &cy);
  float sum = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = cy + dy;
    if (yy < 0 || yy >= g.ncy) continue;
    const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, g.ncx - 1);
    const uint64_t k_lo = (uint64_t)(uint32_t)(yy * g.ncx + x_lo) << 32;
    const uint64_t k_hi = (uint64_t)(uint32_t)(yy * g.ncx + x_hi + 1) << 32;
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cell_keys[mid] < k_lo) lo = mid + 1; else hi = mid; }
    for (int q = lo; q < n && cell_keys[q] < k_hi; ++q) {
      const float4 c = boxes[cell_rows[q]];
      const float xx1 = fmaxf(b.x, c.x), yy1 = fmaxf(b.y, c.y), xx2 = fminf(b.z, c.z), yy2 = fminf(b.w, c.w);
      const float inter = fmaxf(0.f, xx2 - xx1) * fmaxf(0.f, yy2 - yy1);
      const float iou = inter / (area + (c.z - c.x) * (c.w - c.y) - inter);
      sum = sum + iou * (iou > thr ? 1.f : 0.f);       // iou *= iou > thresh; NaN stays NaN
    }
  }
  votes[r] = sum;
}

}  // namespace cpn

namespace cpn {

size_t grid_bin_workspace_bytes(int64_t n_boxes) {
  const size_t nn = (size_t)(n_boxes > 0 ? n_boxes : 1);
  const size_t sort_bytes = radix_sort_tmp_bytes((int64_t)nn);
  return align_up(nn * 8, 256) * 2 + align_up(nn * 4, 256) * 2 + 2048 + align_up(sort_bytes + 256, 256) + 1024;
}

// Bins `n` boxes (by centre) into the uniform grid whose cell is the largest box extent: boxes that overlap lie in
// 3x3 neighbouring cells.  Returns device arrays inside `workspace`: cell_keys[n] = (cell << 32 | row) ascending,
// cell_rows[n] = the row of each entry, and the GridInfo.
int grid_bin_boxes(const float4* boxes, int n, void* workspace, GridBins* out, cudaStream_t st) {
  const size_t nn = (size_t)n;
  char* b = reinterpret_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = b + off; off += align_up(bytes, 256); return p; };
  uint64_t* keys_a = (uint64_t*)take(nn * 8);
  uint64_t* cell_keys = (uint64_t*)take(nn * 8);
  int32_t* vals_a = (int32_t*)take(nn * 4);
  int32_t* cell_rows = (int32_t*)take(nn * 4);
  int* small = (int*)take(2048);
  GridInfo* gi = reinterpret_cast<GridInfo*>(reinterpret_cast<char*>(small) + 64);
  const size_t sort_bytes = radix_sort_tmp_bytes(n);
  void* sort_tmp = take(sort_bytes + 256);
  const int tb = 256, gb = (n + tb - 1) / tb;
  const int ext_init[5] = {0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  CPN_CHECK_CUDA(cudaMemcpyAsync(small, ext_init, sizeof(ext_init), cudaMemcpyHostToDevice, st));
  grid_extent_kernel<<<capped_blocks(gb), tb, 0, st>>>(boxes, n, small);
  CPN_CHECK_LAUNCH();
  grid_setup_kernel<<<1, 32, 0, st>>>(small, gi);
  CPN_CHECK_LAUNCH();
  votes_cellkeys_kernel<<<gb, tb, 0, st>>>(boxes, n, gi, keys_a, vals_a);
  CPN_CHECK_LAUNCH();
  if (radix_sort_pairs(sort_tmp, keys_a, cell_keys, vals_a, cell_rows, n, 64, st)) return 1;
  out->cell_keys = cell_keys; out->cell_rows = cell_rows; out->gi = gi;
  return 0;
}

}  // namespace cpn

extern "C" size_t cpn_box_votes_workspace_bytes(int64_t n_boxes) { return cpn::grid_bin_workspace_bytes(n_boxes); }

extern "C" int cpn_box_votes(const float* boxes, int64_t n_boxes, float iou_threshold, void* workspace, float* votes,
                             void* stream) {
  CPN_REQUIRE(n_boxes < (1ll << 31), "box_votes: too many boxes");
  if (n_boxes <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)n_boxes;
  GridBins bins;
  if (grid_bin_boxes(reinterpret_cast<const float4*>(boxes), n, workspace, &bins, st)) return 1;
  votes_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(boxes), bins.cell_keys, bins.cell_rows, n,
                                                bins.gi, iou_threshold, votes);
  CPN_CHECK_LAUNCH();
  return 0;
}
