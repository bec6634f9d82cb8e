// Tensor-core implicit-GEMM convolutions for sm_100a: tcgen05.mma (fp16 x fp16 -> fp32 in TMEM), operands staged by
// TMA, mbarrier pipelines, persistent warp-specialised CTAs, double-buffered TMEM accumulators.
//
// Replaces nn.Conv2d (+ folded BatchNorm2d) (+ residual) (+ ReLU) (+ ReadOut's final 1x1 projection) on the path:
//   /root/reference/celldetection/models/commons.py:494-511 (ReadOut 7x7 + 1x1), :120-149 (TwoConvNormRelu 3x3),
//   models/resnet.py:88-116 (Bottleneck 1x1 / grouped 3x3 / 1x1, forward = torchvision), models/unet.py:121-128
//   (1x1 "inner" convs), models/fpn.py:106-121 (lateral 1x1, output 3x3), models/resnet.py:273 (7x7 s2 stem, as a 1x1
//   over the im2col operand written by prep_im2col_kernel).
//
// GEMM view: D[m, n] = sum_{tap, c} A[pixel m shifted by tap, c] * Wt[tap][n][c]
//   M = 128 output pixels per accumulator (UMMA M = 128, cta_group::1), N tile = BN in {64, 128, 256} output channels,
//   K block = 64 input channels of one filter tap (one 128-byte swizzle row = 4 x UMMA K = 16).
// Two kernels share the descriptors, the epilogue (epilogue_rows) and the host-side plan:
//   conv_tc_kernel<BN>         any stride-1/2 convolution: per (tap, channel block) one TMA box {64ch, 16, 8, 1} of the
//                              NHWC activations (out-of-bounds zero fill = padding; stride 2 via four parity views) and
//                              one box {64, BN, 1} of the K-major weights, both SWIZZLE_128B, 4-8 stage ring.
//   conv_halo_kernel<BN,MSUB>  stride-1 kxk: the input patch of a tile is loaded ONCE per channel block and every filter
//                              tap is a UMMA descriptor on a shifted window of it (see the kernel's header).
// Warp roles: TMA producer(s), MMA issuer(s), 8 epilogue warps (two per TMEM lane quadrant).  The MMA warp runs fully
// converged and elects the issuing lane inside the asm (umma_f16_elect): with a single divergent thread ptxas moved every
// descriptor through R2UR one MMA at a time and the tensor pipe idled 20 % (BN = 256) to 75 % (BN = 64) of the time.
// fp32 accumulators never leave TMEM until the epilogue: bias (+ residual, optionally through nearest up-sampling)
// (+ ReLU) -> fp16 NHWC store, or -> fused ReadOut projection (+ 3*tanh) -> fp32 head records.
// Split-precision mode (CPN_DT_F16X2): three passes per K block, A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, outputs re-split.
// Grouped convolutions use BN = 64 and contract only over their 64-channel block-diagonal slab (cpn_op_t::kslab).
// Layers without a fused projection use the COAL kernel variants: a line-coalesced, shared-memory-staged epilogue
// (epilogue_coalesced) for residual loads and output stores.
// Round 2: the 2-pass engine (CPN_DT_F16F8: one kind::f8f6f4 pass over e4m3 copies of the operands' rounding residuals, then
// the kind::f16 pass, same TMEM accumulator, epilogue scale 1/S), conv_tap2_kernel (two adjacent taps per N = 128 instruction
// for the 64-wide 7x7 layers), CPN_CONV_UP2 (conv3x3 o nearest x2 as phase kernels on the low-res map, pixel-shuffle store,
// 2 x 2 of 3 x 3 taps per phase N tile), conv_pair_kernel (CTA pairs, cta_group::2; opt-in).
// Environment switches (experiments, see profiles/r01_summary.md / r02_summary.md): CPN_HALO=0, CPN_HALO_ALL=0,
// CPN_HALO_SW128=0, CPN_HALO_BASEOFF=1, CPN_HALO_SWAP=1, CPN_ROTATE=1, CPN_COALESCE=0, CPN_COALESCE_HALO=0,
// CPN_COALESCE_SPLIT=0, CPN_SPLIT_LOFIRST=0, CPN_PAIR=1, CPN_TAP2=0|2, CPN_UP2_SKIP=0, CPN_BN_1X1=64|128, CPN_TC_STAGES=n,
// CPN_DBG_EPI=1 (timing experiments only: the coalesced epilogue skips its global traffic, results are garbage).
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <memory>
#include <cuda.h>
#include <cudaTypedefs.h>

namespace cpn {

constexpr int TC_BW = 16, TC_BH = 8, TC_BM = 128, TC_BK = 64;
constexpr int TC_EPI_WARPS = 8;                        // two warps per TMEM lane quadrant, each takes every other chunk
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SMEM_BUDGET = 196608;  // bytes of operand stages
constexpr int TC_STAGING_BYTES = TC_EPI_WARPS * 2048;   // epilogue_coalesced: 32 rows x 64 B per epilogue warp
constexpr int TC_PROJ_SMEM_MAX = 32768; // fused projection weights (fp32): up to 4 heads x 8 outputs x 256 mid channels

constexpr int TC_MAX_HEADS = 4;    // fused ReadOut projections: one per N tile
constexpr int TC_PROJ_MAX = 24;    // max output channels of one fused projection

// ReadOut's final 1x1 convolution (+ final activation) fused into the epilogue of the kxk convolution that feeds it
// (models/commons.py:494-511): the BN mid channels of one N tile never leave registers.
struct ProjHead {
  const float* w;   // [cout][BN] fp32
  const float* b;   // [cout] or nullptr
  float* out;       // fp32 records: out[pixel * pitch + j]
  int cout, pitch, act;
  float act_scale;
};

struct ConvTcParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  ProjHead proj[TC_MAX_HEADS];
  int nproj;
  __half* out;
  const __half* res;
  const float* bias;
  int out_pitch, res_pitch, res_h, res_w;
  int N, Ho, Wo, cout;
  int R, S, stride, pad;
  int cblocks, kslab, slab_mode;
  int relu;
  int tiles_x, tiles_y, tiles_n;
  long long total_tiles;
  // halo variant (stride-1 kxk): activations of one 64-channel block are staged ONCE per tile as a (16+R-1) x
  // (8*MSUB+S-1) pixel patch and every filter tap addresses a shifted window of it
  CUtensorMap tmH;          // (C, W, H, N), box {8, PW, PH, 1}, no swizzle
  int halo, pw, ph, plane_stride, nb_stages, swap_lbo_sbo;
  int halo_sw128, halo_baseoff;
  int split_lofirst;                 // split mode: run the two correction passes before the main pass
  int split, a_lo, out_lo, res_lo;   // split = 1, CPN_DT_F16X2: 3 passes per 64-channel block (A_hi W_hi, A_lo W_hi,
                                     // A_hi W_lo); element distance hi -> lo half in the A / output / residual buffers.
                                     // split = 2, CPN_DT_F16F8: 2 passes -- one kind::f8f6f4 K block over the pixel's
                                     // 128-byte (lo8 | hi8) row, then the kind::f16 K block; *_lo = distance to the 8-bit block
  float acc_scale;                   // accumulator scale applied first in every epilogue (1/S of the F16F8 weight packing)
  float out_lo_scale, out_hi8_scale; // F16F8 output: lo8 = e4m3((v - hi) * out_lo_scale), hi8 = e4m3(hi * out_hi8_scale)
  float res_lo_inv;                  // F16F8 residual: value = hi + lo8 * res_lo_inv
  int up2;                  // CPN_CONV_UP2: the accumulator's 4 x 64-channel column groups are the four phases of a 2x
                            // up-sampled output: pixel (y, x), column (2a+b)*cup + c -> out[(2y+a, 2x+b), c]
  int cup, Ho2, Wo2;        // up2: channels per phase, output extent
  int up2_skip;             // up2 with cup % BN == 0: an N tile is ONE phase (a, b) and needs only its 2 x 2 of the 3 x 3 taps
  int tap2;                 // conv_tap2_kernel: 64-wide layers, two horizontally adjacent filter taps per N = 128 instruction
  int dbg_epi;              // experiment switch CPN_DBG_EPI=1: the coalesced epilogue skips its global loads / stores
  int coalesce;             // conv_tc_kernel: smem-staged, line-coalesced residual loads / output stores (epilogue_coalesced)
  int rotate;               // start each CTA's K loop at a different (tap, block): de-correlates the L2 reads of the shared weights   // 1: the patch is ONE box {64, PW, PH, 1} with SWIZZLE_128B (128-byte pixel rows)
};

struct ConvTcPlan {
  ConvTcParams p;
  int pair;               // 1: conv_pair_kernel (CTA pairs, tcgen05 cta_group::2, M = 256 per instruction)
  int proj_smem_bytes;
  int msub;
  int bn;
  int stages;
  int smem_bytes;
  int grid;
  long long full_tiles;   // total_tiles / grid of the whole tensor (conv_tc_limit_rows trims the launch to leading rows)
  int full_grid;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; fp16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent variants: the WHOLE warp executes the surrounding code (so descriptors stay in uniform registers and
// ptxas does not serialise R2UR moves on a single divergent thread); elect.sync picks the one lane that issues.
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// e4m3 x e4m3 -> fp32 (kind::f8f6f4, K = 32 per instruction): same descriptors, same instruction-descriptor bits (format
// code 0 is E4M3 for this kind and F16 for kind::f16), same accumulator -- the correction pass of the F16F8 engine.
__device__ __forceinline__ void umma_f8_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar)
      : "memory");
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// conv_tap2_kernel: an accumulator of sub-tile j holds 128 columns per pixel row -- [0, 64) the sums of the taps issued in
// the "lo" half, [64, 128) the sums of their right-hand neighbour taps, which belong to the output pixel one to the LEFT
// (see the kernel's header).  Output (y, x) = lo(y, x) + hi(y, x + 1): lane + 1 inside an 8-pixel row group, lane - 7 of
// the NEXT sub-tile's accumulator for x == 7.  Returns the 32 combined fp32 values of chunk `ch` (wait included).
__device__ __forceinline__ void tap2_combine32(const uint32_t taddr, const int ch, const int j, const int lane,
                                               uint32_t (&v)[32]) {
  const bool inner = (lane & 7) < 7;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t lo[16], hi[16], hn[16];
    tmem_ld16(taddr + ch * 32 + h * 16, lo);
    tmem_ld16(taddr + 64 + ch * 32 + h * 16, hi);
    if (j == 0) {
      tmem_ld16(taddr + 128 + 64 + ch * 32 + h * 16, hn);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) hn[c] = 0u;
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float r1 = __shfl_down_sync(0xffffffffu, __uint_as_float(hi[c]), 1);
      const float r2 = __shfl_sync(0xffffffffu, __uint_as_float(hn[c]), lane & ~7);
      v[h * 16 + c] = __float_as_uint(__uint_as_float(lo[c]) + (inner ? r1 : r2));
    }
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14),
// LBO >> 4 in [16,30) (= 1, unused for swizzled K-major), SBO >> 4 in [32,46) (= 1024 B between 8-row groups),
// version = 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// SW128 K-major descriptor whose 8-row groups are `sbo` bytes apart and whose start may sit on any 128-byte row of
// a 1024-byte-aligned swizzled region (base_offset = row phase, optional).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc_ex(uint32_t smem_addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// K-major, no-swizzle ("interleaved") descriptor: core matrix = 8 rows x 16 bytes with rows 16 bytes apart;
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6), a_format F16 = 0 @ [7,10),
// b_format F16 = 0 @ [10,13), a_major = b_major = K (0), n_dim = N >> 3 @ [17,23), m_dim = M >> 4 @ [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Split-precision (CPN_DT_F16X2) K order.  Pass ids: 0 = A_hi*W_hi, 1 = A_lo*W_hi, 2 = A_hi*W_lo.  The tensor core's fp32
// accumulator truncates the aligned sum of every MMA step, an error of a fraction of ulp(|D|) per step that does not
// average out.  Running the two correction passes FIRST, while |D| is still ~2^-11 of its final magnitude, leaves only
// the K/16 steps of the main pass exposed (a third of the steps of an interleaved order).
// (CPN_SPLIT_LOFIRST=0 restores the interleaved-era order hi, lo, lo for A/B measurements.)
__device__ __forceinline__ int split_pass_order(int i, int lofirst) { return lofirst ? (i == 2 ? 0 : i + 1) : i; }

// ---------------------------------------------------------------------------------------------------------------------
// Epilogue of one accumulator (128 rows x BN columns in TMEM at `taddr`): this thread owns output pixel (img, y, x).
// Either the fused ReadOut projection (p.nproj > 0) or bias / residual / ReLU / fp16 store of every other 32-column
// chunk (`half` selects which of the two warps of a lane quadrant this is).
// ---------------------------------------------------------------------------------------------------------------------
// Residual row of output pixel (img, y, x) at channel n0 (nullptr: no residual / pixel outside the image).  A residual
// of a different spatial size is read through the nearest-neighbour index map (torchvision FPN top-down path).
__device__ __forceinline__ const __half* residual_row(const ConvTcParams& p, const int img, const int y, const int x,
                                                      const int n0) {
  if (!p.res || y >= p.Ho || x >= p.Wo || img >= p.N) return nullptr;
  const int ry = (p.res_h == p.Ho) ? y : (int)(((long long)y * p.res_h) / p.Ho);
  const int rx = (p.res_w == p.Wo) ? x : (int)(((long long)x * p.res_w) / p.Wo);
  return p.res + (((long long)img * p.res_h + ry) * p.res_w + rx) * p.res_pitch + n0;
}

// The 32 residual halves (hi part) of one 32-column chunk: four 16-byte loads.
struct ResChunk { uint4 v[4]; };
__device__ __forceinline__ void load_res_chunk(const __half* rp, const int ch, ResChunk& r) {
  const uint4* q = reinterpret_cast<const uint4*>(rp + ch * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __ldg(q + i);
}

// PF: software-pipelined residual loads -- `rfirst` holds the residual of this warp's first chunk (loaded by the caller
// BEFORE it waited for the accumulator: the residual does not depend on the MMAs), and the next chunk's residual is
// requested into the same registers as soon as the current chunk's values have been consumed, i.e. ahead of the stores,
// the next TMEM load and its wait.  Without it the epilogue of a 1x1 + residual layer, whose mainloop is only 4-16 K
// blocks long, exposed one global-memory latency per chunk and ran longer than the mainloop (ncu: 72.8 us with the
// residual vs 49.6 us without, same 1024 -> 1024 shape).
template <int BN, bool PF, bool TAP2 = false>
__device__ __forceinline__ void epilogue_rows(const ConvTcParams& p, const uint32_t taddr, const int img, const int y,
                                              const int x, const int n_tile, const int n0, const int half,
                                              const float* proj_w, const ResChunk* rfirst = nullptr, const int tapj = 0,
                                              const int lane = 0, const bool keep = true) {
  const bool valid = (y < p.Ho) && (x < p.Wo) && (img < p.N) && keep;
  __half* op = p.out + (((long long)img * p.Ho + y) * p.Wo + x) * p.out_pitch + n0;
  const __half* rp = residual_row(p, img, y, x, n0);
  if (p.nproj > 0) {
    // ---- fused ReadOut: bias + ReLU on the fp32 accumulators, then the 1x1 projection of this N tile's head ----
    const ProjHead& H = p.proj[n_tile];
    int woff = 0;
    for (int h = 0; h < n_tile; ++h) woff += p.proj[h].cout * BN;
    const float* wh = proj_w + woff;
    float pacc[TC_PROJ_MAX];
#pragma unroll
    for (int j = 0; j < TC_PROJ_MAX; ++j) pacc[j] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      uint32_t v[32];
      if (TAP2) {
        tap2_combine32(taddr, ch, tapj, lane, v);
      } else {
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
      }
      float f[32];
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + ch * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b = p.bias ? __ldg(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[q * 4 + 0] = __uint_as_float(v[q * 4 + 0]) * p.acc_scale + b.x;
        f[q * 4 + 1] = __uint_as_float(v[q * 4 + 1]) * p.acc_scale + b.y;
        f[q * 4 + 2] = __uint_as_float(v[q * 4 + 2]) * p.acc_scale + b.z;
        f[q * 4 + 3] = __uint_as_float(v[q * 4 + 3]) * p.acc_scale + b.w;
      }
      if (p.relu) {
#pragma unroll
        for (int c = 0; c < 32; ++c) f[c] = fmaxf(f[c], 0.f);
      }
#pragma unroll
      for (int j = 0; j < TC_PROJ_MAX; ++j) {
        if (j < H.cout) {
          const float4* w4 = reinterpret_cast<const float4*>(wh + j * BN + ch * 32);
          float a = pacc[j];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w = w4[q];
            a = fmaf(f[q * 4 + 0], w.x, a);
            a = fmaf(f[q * 4 + 1], w.y, a);
            a = fmaf(f[q * 4 + 2], w.z, a);
            a = fmaf(f[q * 4 + 3], w.w, a);
          }
          pacc[j] = a;
        }
      }
    }
    if (valid) {
      float* o = H.out + (((long long)img * p.Ho + y) * p.Wo + x) * H.pitch;
#pragma unroll
      for (int j = 0; j < TC_PROJ_MAX; ++j) {
        if (j < H.cout) {
          float r = pacc[j] + (H.b ? __ldg(H.b + j) : 0.f);
          if (H.act == CPN_ACT_SCALED_TANH) r = tanhf(r) * H.act_scale;
          else if (H.act == CPN_ACT_RELU) r = fmaxf(r, 0.f);
          else if (H.act == CPN_ACT_SIGMOID) r = 1.f / (1.f + expf(-r));
          o[j] = r;
        }
      }
    }
  } else {
  ResChunk rc;
  if (PF && rp) rc = *rfirst;
#pragma unroll 1
  for (int ch = half; ch < BN / 32; ch += 2) {
    uint32_t v[32];
    if (TAP2) {
      if (!PF && rp) load_res_chunk(rp, ch, rc);
      tap2_combine32(taddr, ch, tapj, lane, v);
    } else {
      tmem_ld32(taddr + ch * 32, v);
      if (!PF && rp) load_res_chunk(rp, ch, rc);
      tmem_ld_wait();
    }
    if (valid) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + ch * 32);
      uint4 packed[4], packed_lo[4];
      uint32_t* pk = reinterpret_cast<uint32_t*>(packed);
      uint32_t* pl = reinterpret_cast<uint32_t*>(packed_lo);
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(rc.v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 b = p.bias ? __ldg(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float f0 = __uint_as_float(v[q * 4 + 0]) * p.acc_scale + b.x, f1 = __uint_as_float(v[q * 4 + 1]) * p.acc_scale + b.y;
        float f2 = __uint_as_float(v[q * 4 + 2]) * p.acc_scale + b.z, f3 = __uint_as_float(v[q * 4 + 3]) * p.acc_scale + b.w;
        if (rp) {
          const __half2 r0 = *reinterpret_cast<const __half2*>(&rw[q * 2]);
          const __half2 r1 = *reinterpret_cast<const __half2*>(&rw[q * 2 + 1]);
          f0 += __low2float(r0); f1 += __high2float(r0); f2 += __low2float(r1); f3 += __high2float(r1);
          if (p.split == 1) {   // residual = hi + lo
            const uint2 rl = __ldg(reinterpret_cast<const uint2*>(rp + p.res_lo + ch * 32 + q * 4));
            const __half2 l0 = *reinterpret_cast<const __half2*>(&rl.x), l1 = *reinterpret_cast<const __half2*>(&rl.y);
            f0 += __low2float(l0); f1 += __high2float(l0); f2 += __low2float(l1); f3 += __high2float(l1);
          } else if (p.split == 2) {   // residual = hi + lo8 * 2^-(8+e): the chunk's first 32 bytes are the lo8 values
            float l0, l1, l2, l3;
            unpack_e4m3x4(__ldg(reinterpret_cast<const uint32_t*>(rp + p.res_lo + ch * 32) + q), l0, l1, l2, l3);
            f0 += l0 * p.res_lo_inv; f1 += l1 * p.res_lo_inv; f2 += l2 * p.res_lo_inv; f3 += l3 * p.res_lo_inv;
          }
        }
        if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
        __half2 h0 = __floats2half2_rn(f0, f1), h1 = __floats2half2_rn(f2, f3);
        pk[q * 2 + 0] = *reinterpret_cast<uint32_t*>(&h0);
        pk[q * 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
        if (p.split == 1) {     // lo = fp16(v - hi)
          __half2 g0 = __floats2half2_rn(f0 - __low2float(h0), f1 - __high2float(h0));
          __half2 g1 = __floats2half2_rn(f2 - __low2float(h1), f3 - __high2float(h1));
          pl[q * 2 + 0] = *reinterpret_cast<uint32_t*>(&g0);
          pl[q * 2 + 1] = *reinterpret_cast<uint32_t*>(&g1);
        } else if (p.split == 2) {   // 8-bit block of the chunk: 32 x lo8 then 32 x hi8
          const float a0 = __low2float(h0), a1 = __high2float(h0), a2 = __low2float(h1), a3 = __high2float(h1);
          pl[q] = pack_e4m3x4((f0 - a0) * p.out_lo_scale, (f1 - a1) * p.out_lo_scale, (f2 - a2) * p.out_lo_scale,
                              (f3 - a3) * p.out_lo_scale);
          pl[8 + q] = pack_e4m3x4(a0 * p.out_hi8_scale, a1 * p.out_hi8_scale, a2 * p.out_hi8_scale, a3 * p.out_hi8_scale);
        }
      }
      if (PF && rp && ch + 2 < BN / 32) load_res_chunk(rp, ch + 2, rc);   // rc is dead: request the next chunk now
      uint4* o4 = reinterpret_cast<uint4*>(op + ch * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) o4[q] = packed[q];
      if (p.split) {
        uint4* l4 = reinterpret_cast<uint4*>(op + p.out_lo + ch * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) l4[q] = packed_lo[q];
      }
    }
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Line-coalesced epilogue of conv_tc_kernel (fp16, no fused projection).  TMEM hands every thread one output pixel (row)
// and 32 consecutive channels (64 bytes), so direct global accesses touch 32 different lines per warp instruction; ncu
// showed the 1x1 layers of the encoder bound by exactly that (time proportional to the bytes moved through the LSU at
// ~2 TB/s, tensor pipe 8-45 % active, DRAM 12-35 %).  Here each warp owns a 2 KB staging tile (32 rows x 64 B, 16-byte
// slots XOR-swizzled with (row >> 1) & 3: conflict-free for both access patterns): residual and output cross it so that
// the GLOBAL accesses are made in the transposed role -- lane l moves 16-byte segment l & 3 of rows i * 8 + (l >> 2) --
// i.e. 8 rows x 64 contiguous bytes per instruction, a quarter of the wavefronts.  The residual of the next chunk is in
// flight (registers) while the current one is computed; the first chunk's is requested before the accumulator barrier.
// ---------------------------------------------------------------------------------------------------------------------
struct CoalRows {
  uint32_t ooff[4], roff[4];   // element offsets (out / residual) of this lane's four rows at its 16-byte segment
  uint32_t okmask;             // bit i: row i is inside the image
};

// WLOG2: log2 of the accumulator's pixel-tile width (row r of the accumulator is pixel (r >> WLOG2, r & (2^WLOG2 - 1)))
template <int WLOG2>
__device__ __forceinline__ void coal_rows(const ConvTcParams& p, const int img, const int ty0, const int tx0,
                                          const int n0, const int quad, const int lane, CoalRows& c,
                                          const int x_limit = 0x7fffffff) {
  c.okmask = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = quad * 32 + i * 8 + (lane >> 2);
    const int y = ty0 + (r >> WLOG2), x = tx0 + (r & ((1 << WLOG2) - 1));
    const bool ok = y < p.Ho && x < p.Wo && x < x_limit && img < p.N;   // (img == N: phantom tile of an odd CTA pair)
    c.okmask |= ok ? (1u << i) : 0u;
    c.ooff[i] = ok ? (uint32_t)((((long long)img * p.Ho + y) * p.Wo + x) * p.out_pitch + n0 + (lane & 3) * 8) : 0u;
    if (p.up2 && ok)    // phase (0, 0) pixel of the up-sampled output; the chunk's phase / channel offset is added at the store
      c.ooff[i] = (uint32_t)((((long long)img * p.Ho2 + 2 * y) * p.Wo2 + 2 * x) * p.out_pitch + (lane & 3) * 8);
    c.roff[i] = 0u;
    if (p.res && ok) {
      const int ry = (p.res_h == p.Ho) ? y : (int)(((long long)y * p.res_h) / p.Ho);
      const int rx = (p.res_w == p.Wo) ? x : (int)(((long long)x * p.res_w) / p.Wo);
      c.roff[i] = (uint32_t)((((long long)img * p.res_h + ry) * p.res_w + rx) * p.res_pitch + n0 + (lane & 3) * 8);
    }
  }
}

// lo_delta: 0 for the hi (or only) half of the residual, p.res_lo for the lo half of a split (CPN_DT_F16X2) tensor
__device__ __forceinline__ void coal_load_res(const ConvTcParams& p, const CoalRows& c, const int ch, const int lo_delta,
                                              uint4 (&rg)[4]) {
  if (p.dbg_epi) return;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (c.okmask & (1u << i)) rg[i] = __ldg(reinterpret_cast<const uint4*>(p.res + c.roff[i] + lo_delta + ch * 32));
}

// staging-tile transposition helpers (see the header comment above): "t" = transposed / coalesced role, "o" = row owner
__device__ __forceinline__ void stg_put_t(uint4* stg, const int lane, const uint4 (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { const int r = i * 8 + (lane >> 2); stg[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))] = x[i]; }
}
__device__ __forceinline__ void stg_get_o(const uint4* stg, const int lane, uint4 (&x)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = stg[lane * 4 + (k ^ ((lane >> 1) & 3))];
}
__device__ __forceinline__ void stg_put_o(uint4* stg, const int lane, const uint4 (&x)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) stg[lane * 4 + (k ^ ((lane >> 1) & 3))] = x[k];
}
__device__ __forceinline__ void stg_store_t(const uint4* stg, const int lane, const CoalRows& c, __half* base) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2);
    if (c.okmask & (1u << i)) *reinterpret_cast<uint4*>(base + c.ooff[i]) = stg[r * 4 + ((lane & 3) ^ ((r >> 1) & 3))];
  }
}

// rg / rgl: residual (hi / lo half) of this warp's first chunk in the transposed role, requested by the caller.
// Split tensors (p.split): residual = hi + lo, the result is re-split into hi = fp16(v), lo = fp16(v - hi), and both
// halves cross the staging tile one after the other.
template <int BN, bool TAP2 = false>
__device__ __forceinline__ void epilogue_coalesced(const ConvTcParams& p, const uint32_t taddr, const CoalRows& c,
                                                   const int n0, const int half, const int lane, uint4* stg,
                                                   uint4 (&rg)[4], uint4 (&rgl)[4], const int tapj = 0) {
  const bool has_res = p.res != nullptr;
  const bool split = p.split != 0;
  const bool f8 = p.split == 2;
  // bias of the chunk: one value per lane, requested a whole chunk ahead (ncu, round 2: the float4 bias loads issued right
  // before their first use were 10-12 % of all stall samples of the 1x1 layers); lane c's value reaches column c by shuffle
  float bl = p.bias ? __ldg(p.bias + n0 + half * 32 + lane) : 0.f;
#pragma unroll 1
  for (int ch = half; ch < BN / 32; ch += 2) {
    uint32_t v[32];
    if (TAP2) tap2_combine32(taddr, ch, tapj, lane, v);
    else tmem_ld32(taddr + ch * 32, v);
    const float bcur = bl;
    if (p.bias && ch + 2 < BN / 32) bl = __ldg(p.bias + n0 + (ch + 2) * 32 + lane);
    uint4 rr[4], rl[4];
    if (has_res) {
      stg_put_t(stg, lane, rg);
      __syncwarp();
      stg_get_o(stg, lane, rr);
      if (split) {
        __syncwarp();
        stg_put_t(stg, lane, rgl);
        __syncwarp();
        stg_get_o(stg, lane, rl);
      }
      if (ch + 2 < BN / 32) {                                        // next chunk's residual while this one computes
        coal_load_res(p, c, ch + 2, 0, rg);
        if (split) coal_load_res(p, c, ch + 2, p.res_lo, rgl);
      }
    }
    if (!TAP2) tmem_ld_wait();
    uint32_t* pk = reinterpret_cast<uint32_t*>(rr);                  // results overwrite the residual registers in place
    uint32_t* pl = reinterpret_cast<uint32_t*>(rl);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 b = make_float4(__shfl_sync(0xffffffffu, bcur, k * 4 + 0), __shfl_sync(0xffffffffu, bcur, k * 4 + 1),
                                   __shfl_sync(0xffffffffu, bcur, k * 4 + 2), __shfl_sync(0xffffffffu, bcur, k * 4 + 3));
      float f0 = __uint_as_float(v[k * 4 + 0]) * p.acc_scale + b.x, f1 = __uint_as_float(v[k * 4 + 1]) * p.acc_scale + b.y;
      float f2 = __uint_as_float(v[k * 4 + 2]) * p.acc_scale + b.z, f3 = __uint_as_float(v[k * 4 + 3]) * p.acc_scale + b.w;
      if (has_res) {
        const __half2 r0 = *reinterpret_cast<const __half2*>(&pk[k * 2]);
        const __half2 r1 = *reinterpret_cast<const __half2*>(&pk[k * 2 + 1]);
        f0 += __low2float(r0); f1 += __high2float(r0); f2 += __low2float(r1); f3 += __high2float(r1);
        if (f8) {                                                      // words 0..7 of the 8-bit chunk are the lo8 values
          float l0, l1, l2, l3;
          unpack_e4m3x4(pl[k], l0, l1, l2, l3);
          f0 += l0 * p.res_lo_inv; f1 += l1 * p.res_lo_inv; f2 += l2 * p.res_lo_inv; f3 += l3 * p.res_lo_inv;
        } else if (split) {
          const __half2 l0 = *reinterpret_cast<const __half2*>(&pl[k * 2]);
          const __half2 l1 = *reinterpret_cast<const __half2*>(&pl[k * 2 + 1]);
          f0 += __low2float(l0); f1 += __high2float(l0); f2 += __low2float(l1); f3 += __high2float(l1);
        }
      }
      if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
      const __half2 h0 = __floats2half2_rn(f0, f1), h1 = __floats2half2_rn(f2, f3);
      pk[k * 2 + 0] = *reinterpret_cast<const uint32_t*>(&h0);
      pk[k * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      if (f8) {                                                      // 8-bit chunk: words 0..7 lo8, words 8..15 hi8
        const float a0 = __low2float(h0), a1 = __high2float(h0), a2 = __low2float(h1), a3 = __high2float(h1);
        pl[k] = pack_e4m3x4((f0 - a0) * p.out_lo_scale, (f1 - a1) * p.out_lo_scale, (f2 - a2) * p.out_lo_scale,
                            (f3 - a3) * p.out_lo_scale);   // (the residual's lo8 word k was consumed above)
        pl[8 + k] = pack_e4m3x4(a0 * p.out_hi8_scale, a1 * p.out_hi8_scale, a2 * p.out_hi8_scale, a3 * p.out_hi8_scale);
      } else if (split) {                                            // lo = fp16(v - hi)
        const __half2 g0 = __floats2half2_rn(f0 - __low2float(h0), f1 - __high2float(h0));
        const __half2 g1 = __floats2half2_rn(f2 - __low2float(h1), f3 - __high2float(h1));
        pl[k * 2 + 0] = *reinterpret_cast<const uint32_t*>(&g0);
        pl[k * 2 + 1] = *reinterpret_cast<const uint32_t*>(&g1);
      }
    }

    __syncwarp();                                                    // every lane has read its residual row(s)
    if (p.dbg_epi) continue;
    long long obase = ch * 32;
    if (p.up2) {                                                     // global column n0 + ch * 32 -> (phase, channel)
      const int col = n0 + ch * 32, ph = col / p.cup, cc = col - ph * p.cup;
      obase = ((long long)(ph >> 1) * p.Wo2 + (ph & 1)) * p.out_pitch + cc;
    }
    stg_put_o(stg, lane, rr);
    __syncwarp();
    stg_store_t(stg, lane, c, p.out + obase);
    if (split) {
      __syncwarp();
      stg_put_o(stg, lane, rl);
      __syncwarp();
      stg_store_t(stg, lane, c, p.out + p.out_lo + obase);
    }
    __syncwarp();                                                    // staging tile free for the next chunk
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------------------------------
template <int BN, bool COAL>
// Register budget: 10 warps spread 3/3/2/2 over the four SM sub-partitions of 16 K registers each, so a thread may use at
// most 168 registers (3 x 32 x 168 <= 16384) -- ptxas derives exactly that cap from __launch_bounds__(320, 1); forcing
// more with __maxnreg__ compiles but fails at launch ("too many resources requested").
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p, int stages) {
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // double-buffered accumulator (power of two >= 32)
  constexpr int MAX_STAGES = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // operand stages must be 1024-byte aligned for SWIZZLE_128B
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = p.R * p.S * p.cblocks;  // K blocks per tile
  // fused projection weights live behind the operand stages: [head][cout][BN] fp32
  // behind the operand stages: the epilogue staging tiles (p.coalesce) or the fused projection weights (p.nproj > 0)
  uint8_t* tail = smem_raw + ((smem_base - smem_u32(smem_raw)) + stages * STAGE_BYTES);
  float* proj_w = reinterpret_cast<float*>(tail);
  if (!COAL && p.nproj > 0) {
    int off = 0;
    for (int h = 0; h < p.nproj; ++h) {
      const int nw = p.proj[h].cout * BN;
      for (int i = threadIdx.x; i < nw; i += blockDim.x) proj_w[off + i] = p.proj[h].w[i];
      off += nw;
    }
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[0]);
    prefetch_tmap(&p.tmB);
    for (int i = 0; i < stages; ++i) {
      mbar_init(smem_u32(&bar_full[i]), 1);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_tfull[i]), 1);
      mbar_init(smem_u32(&bar_tempty[i]), p.nproj > 0 ? 4 : TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = (int)(tile % p.tiles_n);
        const long long m_tile = tile / p.tiles_n;
        const int img = (int)(m_tile / tiles_per_img);
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int y0 = (t_in / p.tiles_x) * TC_BH, x0 = (t_in % p.tiles_x) * TC_BW;
        const int n0 = n_tile * BN;
        const int cbase = p.slab_mode ? (n0 / p.kslab) * p.kslab : 0;
        int kk = p.rotate ? (int)(blockIdx.x % (unsigned)nk) : 0;   // rotated start of the K loop (sum order is free)
        for (int it = 0; it < nk; ++it) {
          int tap = kk / p.cblocks, cb = kk - tap * p.cblocks;   // cb runs over 3 * logical blocks when split
          if (++kk == nk) kk = 0;
          int a_ch = cb * TC_BK;
          if (p.split == 1) {
            // pass-major K order, the two correction passes first (see split_pass_order)
            const int cbl = p.cblocks / 3, per_pass = p.R * p.S * cbl;
            const int kq = tap * p.cblocks + cb, pi = kq / per_pass, rem = kq - pi * per_pass;
            const int pass = split_pass_order(pi, p.split_lofirst);
            tap = rem / cbl;
            const int c = rem - tap * cbl;
            cb = pass * cbl + c;
            a_ch = c * TC_BK + (pass == 1 ? p.a_lo : 0);
          } else if (p.split == 2) {
            // pass-major: the 8-bit correction blocks of every (tap, channel block) first, then the fp16 blocks.  Weight
            // K blocks [0, cbl) are the e4m3 rows, [cbl, 2 cbl) the fp16 rows; the activations' 8-bit rows sit a_lo
            // elements behind the fp16 ones
            const int cbl = p.cblocks / 2, per_pass = p.R * p.S * cbl;
            const int kq = tap * p.cblocks + cb, pi = kq / per_pass, rem = kq - pi * per_pass;
            tap = rem / cbl;
            const int c = rem - tap * cbl;
            cb = pi * cbl + c;
            a_ch = c * TC_BK + (pi == 0 ? p.a_lo : 0);
          }
          const int r = tap / p.S, s = tap - r * p.S;
          int qy = r - p.pad, qx = s - p.pad, map = 0;
          if (p.stride == 2) {
            const int py = qy & 1, px = qx & 1;
            map = py * 2 + px;
            qy = (qy - py) >> 1;
            qx = (qx - px) >> 1;
          }
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full, STAGE_BYTES);
          tma_load_4d(sa, &p.tmA[map], full, cbase + a_ch, x0 + qx, y0 + qy, img);
          tma_load_3d(sb, &p.tmB, full, cb * TC_BK, n0, tap);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = make_idesc_f16(TC_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        const uint64_t da = make_sw128_kmajor_desc(sa), db = make_sw128_kmajor_desc(sb);
        // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the >>4 address field
        if (p.split == 2 && kb < (nk >> 1)) {   // 8-bit correction blocks (K = 32 bytes per instruction: same +2 steps)
          umma_f8_elect(d_tmem, da, db, idesc, (uint32_t)(kb != 0));
#pragma unroll
          for (int k = 1; k < TC_BK / 16; ++k) umma_f8_elect(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        } else {
          umma_f16_elect(d_tmem, da, db, idesc, (uint32_t)(kb != 0));
#pragma unroll
          for (int k = 1; k < TC_BK / 16; ++k) umma_f16_elect(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
        umma_commit_elect(smem_u32(&bar_empty[stage]));  // frees the smem stage when these MMAs retire
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(smem_u32(&bar_tfull[acc]));  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================ epilogue (8 warps, 128 rows x 2 column halves) ================================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which of the two warps of this quadrant (takes chunks ch % 2 == half)
    if (!COAL && p.nproj > 0 && half == 1) {     // the fused-projection epilogue keeps a whole row per thread: 4 warps only
      // fall through to the teardown barrier
    } else {
    const int row = quad * 32 + lane;
    const int py = row / TC_BW, px = row % TC_BW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = (int)(tile % p.tiles_n);
      const long long m_tile = tile / p.tiles_n;
      const int img = (int)(m_tile / tiles_per_img);
      const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
      const int ty0 = (t_in / p.tiles_x) * TC_BH, tx0 = (t_in % p.tiles_x) * TC_BW;
      const int y = ty0 + py, x = tx0 + px;
      const int n0 = n_tile * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * BN;
      if (COAL) {
        CoalRows cr;
        coal_rows<4>(p, img, ty0, tx0, n0, quad, lane, cr);   // TC_BW = 16
        uint4 rg[4], rgl[4];
        if (p.res) {                                 // first chunk's residual: in flight while the MMAs still run
          coal_load_res(p, cr, half, 0, rg);
          if (p.split) coal_load_res(p, cr, half, p.res_lo, rgl);
        }
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tc_fence_after();
        epilogue_coalesced<BN>(p, taddr, cr, n0, half, lane, reinterpret_cast<uint4*>(tail) + (warp - 2) * 128, rg, rgl);
      } else {
      ResChunk rfirst;
      {   // residual of the first chunk: requested while the MMAs of this tile are still running
        const __half* rp0 = residual_row(p, img, y, x, n0);
        if (rp0) load_res_chunk(rp0, half, rfirst);
      }
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      epilogue_rows<BN, true>(p, taddr, img, y, x, n_tile, n0, half, proj_w, &rfirst);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair kernel (tcgen05 cta_group::2): the 1x1 layers of the encoder / decoder with C_out % 256 == 0.
// Hypothesis tested in round 2: the 1x1 layers (tensor pipe 41-44 % active, DRAM 23 %, L2 28 % in ncu) are bound by operand
// delivery -- every CTA of conv_tc_kernel<256> pulls 16 KB of activations + 32 KB of weights per K block.  Result: NOT
// confirmed.  This kernel moves a third fewer bytes per CTA and is 3 % (ncu, cold) to 20 % (back-to-back launches) slower,
// so it is opt-in (CPN_PAIR=1) and kept as a verified cta_group::2 building block.  Two CTAs of a cluster (the two SMs of a
// TPC) share ONE 256 x 256 accumulator tile: each loads the activations of its own 128 pixels and only HALF of the weight tile
// (128 of the 256 output channels), the leader issues M = 256 instructions that read both halves, and each CTA finds its
// 128 rows x 256 columns of the result in its own tensor memory: 32 KB per CTA and K block (-33 %), six stages instead of four.
//   barriers  full[stage]   leader only; expect_tx covers the bytes of BOTH CTAs (their TMA loads signal the leader's barrier)
//             empty[stage]  one per CTA, released by the leader's multicast commit
//             tfull[acc]    one per CTA, multicast commit after the tile's last K block
//             tempty[acc]   leader only; the epilogue warps of both CTAs arrive (the peer through its cluster address)
// Tiles: pair tile = (two consecutive M tiles, one N tile), N fastest; CTA rank r owns M tile 2 * mp + r (an odd tile count
// leaves a phantom tile: its TMA boxes are out of bounds = zero filled, its epilogue stores nothing).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in the executing CTA's shared memory, the bytes are counted on `bar` (a
// shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <bool F8>
__device__ __forceinline__ void umma_pair_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  if (F8) {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the pair's previously issued MMAs have retired
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(bar)
      : "memory");
}

constexpr int TCP_BN = 256;                         // N of the pair's accumulator tile
constexpr uint32_t TCP_A_BYTES = TC_BM * TC_BK * 2;           // 16 KB: this CTA's 128 pixels x 64 channels
constexpr uint32_t TCP_B_BYTES = (TCP_BN / 2) * TC_BK * 2;    // 16 KB: this CTA's half of the weight tile
constexpr uint32_t TCP_STAGE_BYTES = TCP_A_BYTES + TCP_B_BYTES;

template <bool COAL>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_pair_kernel(const __grid_constant__ ConvTcParams p, int stages) {
  constexpr int BN = TCP_BN;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int MAX_STAGES = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int nk = p.R * p.S * p.cblocks;
  uint8_t* tail = smem_raw + ((smem_base - smem_u32(smem_raw)) + stages * TCP_STAGE_BYTES);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[0]);
    prefetch_tmap(&p.tmB);
    for (int i = 0; i < stages; ++i) {
      mbar_init(smem_u32(&bar_full[i]), 1);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_tfull[i]), 1);
      mbar_init(smem_u32(&bar_tempty[i]), 2 * TC_EPI_WARPS);      // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {      // the same warp of both CTAs allocates the pair's tensor memory
    tmem_alloc_pair(smem_u32(&tmem_base_slot), TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const long long m_tiles = (long long)p.N * tiles_per_img;
  const long long pair_tiles = ((m_tiles + 1) >> 1) * p.tiles_n;
  const long long first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = first; tile < pair_tiles; tile += step) {
        const int n_tile = (int)(tile % p.tiles_n);
        const long long m_tile = (tile / p.tiles_n) * 2 + rank;
        const int img = (int)(m_tile / tiles_per_img);                 // == N for the phantom tile: out of bounds, zeros
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int y0 = (t_in / p.tiles_x) * TC_BH, x0 = (t_in % p.tiles_x) * TC_BW;
        const int n0 = n_tile * BN;
        for (int it = 0; it < nk; ++it) {
          int tap = it / p.cblocks, cb = it - tap * p.cblocks;
          int a_ch = cb * TC_BK;
          if (p.split == 1) {
            const int cbl = p.cblocks / 3, per_pass = p.R * p.S * cbl;
            const int pi = it / per_pass, rem = it - pi * per_pass;
            const int pass = split_pass_order(pi, p.split_lofirst);
            tap = rem / cbl;
            const int c = rem - tap * cbl;
            cb = pass * cbl + c;
            a_ch = c * TC_BK + (pass == 1 ? p.a_lo : 0);
          } else if (p.split == 2) {
            const int cbl = p.cblocks / 2, per_pass = p.R * p.S * cbl;
            const int pi = it / per_pass, rem = it - pi * per_pass;
            tap = rem / cbl;
            const int c = rem - tap * cbl;
            cb = pi * cbl + c;
            a_ch = c * TC_BK + (pi == 0 ? p.a_lo : 0);
          }
          const int r = tap / p.S, s = tap - r * p.S;
          int qy = r - p.pad, qx = s - p.pad, map = 0;
          if (p.stride == 2) {
            const int py = qy & 1, px = qx & 1;
            map = py * 2 + px;
            qy = (qy - py) >> 1;
            qx = (qx - px) >> 1;
          }
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
          const uint32_t full = mapa_u32(smem_u32(&bar_full[stage]), 0);       // the leader's barrier
          const uint32_t sa = smem_base + stage * TCP_STAGE_BYTES, sb = sa + TCP_A_BYTES;
          if (leader) mbar_expect_tx(smem_u32(&bar_full[stage]), 2 * TCP_STAGE_BYTES);
          tma_load_4d_pair(sa, &p.tmA[map], full, a_ch, x0 + qx, y0 + qy, img);
          tma_load_3d_pair(sb, &p.tmB, full, cb * TC_BK, n0 + (int)rank * (BN / 2), tap);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(2 * TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = first; tile < pair_tiles; tile += step) {
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * TCP_STAGE_BYTES, sb = sa + TCP_A_BYTES;
          const uint64_t da = make_sw128_kmajor_desc(sa), db = make_sw128_kmajor_desc(sb);
          if (p.split == 2 && kb < (nk >> 1)) {
            umma_pair_elect<true>(d_tmem, da, db, idesc, (uint32_t)(kb != 0));
#pragma unroll
            for (int k = 1; k < TC_BK / 16; ++k) umma_pair_elect<true>(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
          } else {
            umma_pair_elect<false>(d_tmem, da, db, idesc, (uint32_t)(kb != 0));
#pragma unroll
            for (int k = 1; k < TC_BK / 16; ++k) umma_pair_elect<false>(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
          }
          umma_commit_pair_elect(smem_u32(&bar_empty[stage]));   // frees this stage in both CTAs
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair_elect(smem_u32(&bar_tfull[acc]));       // accumulator complete -> both epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (8 warps per CTA) ================================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int py = row / TC_BW, px = row % TC_BW;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&bar_tempty[0]), 0), tempty_leader1 = mapa_u32(smem_u32(&bar_tempty[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = first; tile < pair_tiles; tile += step) {
      const int n_tile = (int)(tile % p.tiles_n);
      const long long m_tile = (tile / p.tiles_n) * 2 + rank;
      const int img = (int)(m_tile / tiles_per_img);
      const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
      const int ty0 = (t_in / p.tiles_x) * TC_BH, tx0 = (t_in % p.tiles_x) * TC_BW;
      const int n0 = n_tile * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * BN;
      if (COAL) {
        CoalRows cr;
        coal_rows<4>(p, img, ty0, tx0, n0, quad, lane, cr);
        uint4 rg[4], rgl[4];
        if (p.res) {
          coal_load_res(p, cr, half, 0, rg);
          if (p.split) coal_load_res(p, cr, half, p.res_lo, rgl);
        }
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tc_fence_after();
        epilogue_coalesced<BN>(p, taddr, cr, n0, half, lane, reinterpret_cast<uint4*>(tail) + (warp - 2) * 128, rg, rgl);
      } else {
        ResChunk rfirst;
        {
          const __half* rp0 = residual_row(p, img, ty0 + py, tx0 + px, n0);
          if (rp0) load_res_chunk(rp0, half, rfirst);
        }
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tc_fence_after();
        epilogue_rows<BN, true>(p, taddr, img, ty0 + py, tx0 + px, n_tile, n0, half, nullptr, &rfirst);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Halo kernel: stride-1 kxk convolutions with operand reuse across filter taps.
//   tile      = 16 rows x (8 * MSUB) columns of output pixels: MSUB accumulators of M = 128 (8 x 16 pixels each)
//   A operand = per 64-channel block ONE patch of (16+R-1) x (8*MSUB+S-1) input pixels: one TMA box
//               {64ch, PW, PH, 1} with SWIZZLE_128B, i.e. 128-byte pixel rows in (y, x) order.  Tap (r, s) of sub-tile j
//               is a K-major SW128 descriptor that STARTS at pixel row (r, s + 8j) with the 8-row groups one patch row
//               (PW * 128 B) apart.  The tensor core applies the 128-byte swizzle on absolute shared-memory address
//               bits, so a start that is not 1024-byte aligned needs no base_offset (verified on hardware; a phase-
//               corrected base_offset gives wrong results).  No re-load per tap: L2->SM activation traffic drops from
//               R*S to ~(1 + halo) reads.  (CPN_HALO_SW128=0 selects the first working variant: 8 no-swizzle planes
//               [16-byte channel chunk][y][x] with rows 16 B apart, LBO = one plane, SBO = PW*16.)
//   B operand = weights of one (tap, channel block): BN x 64 halves, SWIZZLE_128B, shared by the MSUB sub-tiles.
// Warps: 0 = B producer, 1 = MMA issuer (sub-tiles [0, MSUB/2) or all), 2 = A producer, 3 = second MMA issuer (MSUB = 2:
// a 64-wide MMA occupies the tensor pipe for only 32 cycles, less than one warp needs to issue it), 4..11 = epilogue.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TCH_THREADS = 128 + 32 * TC_EPI_WARPS;   // B producer, MMA issuer 0, A producer, MMA issuer 1, epilogue

template <int BN, int MSUB, bool COAL>
__global__ void __launch_bounds__(TCH_THREADS, 1) conv_halo_kernel(const __grid_constant__ ConvTcParams p) {
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * MSUB * BN;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  constexpr int MAX_NB = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_bfull[MAX_NB];
  __shared__ __align__(8) uint64_t bar_bempty[MAX_NB];
  __shared__ __align__(8) uint64_t bar_afull[2];
  __shared__ __align__(8) uint64_t bar_aempty[2];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nb_stages;
  const uint32_t patch_bytes = 8u * (uint32_t)p.plane_stride;         // (sw128 mode: plane_stride = patch / 8)
  const uint32_t a_base = smem_base + nb * B_BYTES;                    // 2 patches behind the B ring
  const uint32_t a_tx = 8u * (uint32_t)(p.ph * p.pw * 16);            // bytes the box load(s) deliver
  // behind the patches: the epilogue staging tiles (COAL) or the fused projection weights (p.nproj > 0)
  uint8_t* tail = smem_raw + ((a_base - smem_u32(smem_raw)) + 2 * patch_bytes);
  float* proj_w = reinterpret_cast<float*>(tail);
  if (!COAL && p.nproj > 0) {
    int off = 0;
    for (int h = 0; h < p.nproj; ++h) {
      const int nw = p.proj[h].cout * BN;
      for (int i = threadIdx.x; i < nw; i += blockDim.x) proj_w[off + i] = p.proj[h].w[i];
      off += nw;
    }
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmH);
    prefetch_tmap(&p.tmB);
    constexpr uint32_t NISSUE = MSUB >= 2 ? 2 : 1;   // MMA issuer warps; each commits once per barrier phase
    for (int i = 0; i < nb; ++i) { mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), NISSUE); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1);
      mbar_init(smem_u32(&bar_aempty[i]), NISSUE);
      mbar_init(smem_u32(&bar_tfull[i]), NISSUE);
      mbar_init(smem_u32(&bar_tempty[i]), p.nproj > 0 ? 4 : TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int taps = p.R * p.S;
  constexpr int TW = 8 * MSUB, TH = 16;

  if (warp == 0) {
    // ================================ B (weights) producer ================================
    if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % p.tiles_n) * BN;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          int cbw = cb;                                   // K block of the (W_hi | W_hi | W_lo) weight tensor
          if (p.split == 1) {
            const int cbl = p.cblocks / 3, pi = cb / cbl;
            cbw = split_pass_order(pi, p.split_lofirst) * cbl + (cb - pi * cbl);
          }                                               // split == 2: (W8 | W16) is already in issue order
          int tap = p.rotate ? (int)(blockIdx.x % (unsigned)taps) : 0;
          const int ph_ = p.up2_skip ? n0 / p.cup : 0;                 // phase (a, b) of this N tile
          const int ntap = p.up2_skip ? 4 : taps;
          for (int it = 0; it < ntap; ++it, tap = (tap + 1 == taps ? 0 : tap + 1)) {
            if (p.up2_skip) tap = ((ph_ >> 1) + (it >> 1)) * 3 + (ph_ & 1) + (it & 1);   // rows a, a+1 x columns b, b+1
            mbar_wait(smem_u32(&bar_bempty[sb]), phb ^ 1);
            const uint32_t full = smem_u32(&bar_bfull[sb]);
            mbar_expect_tx(full, B_BYTES);
            tma_load_3d(smem_base + sb * B_BYTES, &p.tmB, full, cbw * TC_BK, n0, tap);
            if (++sb == nb) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ A (activation patch) producer ================================
    if (lane == 0) {
      int ab = 0;
      uint32_t pha = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const long long m_tile = tile / p.tiles_n;
        const int img = (int)(m_tile / tiles_per_img);
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int y0 = (t_in / p.tiles_x) * TH, x0 = (t_in % p.tiles_x) * TW;
        const int n0 = (int)(tile % p.tiles_n) * BN;
        const int cbase = p.slab_mode ? (n0 / p.kslab) * p.kslab : 0;   // grouped: this N tile's 64-channel slab
        for (int cb = 0; cb < p.cblocks; ++cb) {
          int a_ch = cb * TC_BK;
          if (p.split == 1) {
            const int cbl = p.cblocks / 3, pi = cb / cbl, pass = split_pass_order(pi, p.split_lofirst);
            a_ch = (cb - pi * cbl) * TC_BK + (pass == 1 ? p.a_lo : 0);
          } else if (p.split == 2) {
            const int cbl = p.cblocks / 2;
            a_ch = cb < cbl ? cb * TC_BK + p.a_lo : (cb - cbl) * TC_BK;
          }
          mbar_wait(smem_u32(&bar_aempty[ab]), pha ^ 1);
          const uint32_t full = smem_u32(&bar_afull[ab]);
          mbar_expect_tx(full, a_tx);
          const uint32_t dst = a_base + ab * patch_bytes;
          if (p.halo_sw128) {
            tma_load_4d(dst, &p.tmH, full, cbase + a_ch, x0 - p.pad, y0 - p.pad, img);
          } else {
#pragma unroll
            for (int kc = 0; kc < 8; ++kc)
              tma_load_4d(dst + kc * p.plane_stride, &p.tmH, full, cbase + a_ch + kc * 8, x0 - p.pad, y0 - p.pad, img);
          }
          ab ^= 1;
          if (ab == 0) pha ^= 1;
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && MSUB >= 2)) {
    // ================================ MMA issuer(s) ================================
    constexpr uint32_t idesc = make_idesc_f16(TC_BM, BN);
    constexpr int JN = MSUB >= 2 ? MSUB / 2 : 1;           // sub-tiles per issuer warp
    const int jb = (warp == 3) ? JN : 0;                   // first sub-tile of this issuer
    int sb = 0, ab = 0, acc = 0;
    uint32_t phb = 0, pha = 0, acc_phase = 0;
    const uint32_t lbo = p.swap_lbo_sbo ? (uint32_t)(p.pw * 16) : (uint32_t)p.plane_stride;
    const uint32_t sbo = p.swap_lbo_sbo ? (uint32_t)p.plane_stride : (uint32_t)(p.pw * 16);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1);
      tc_fence_after();
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(smem_u32(&bar_afull[ab]), pha);
        tc_fence_after();
        const uint32_t patch = a_base + ab * patch_bytes;
        int tap = p.rotate ? (int)(blockIdx.x % (unsigned)taps) : 0;
        const int ph_ = p.up2_skip ? (int)(tile % p.tiles_n) * BN / p.cup : 0;
        const int ntap = p.up2_skip ? 4 : taps;
        for (int it = 0; it < ntap; ++it, tap = (tap + 1 == taps ? 0 : tap + 1)) {
          if (p.up2_skip) tap = ((ph_ >> 1) + (it >> 1)) * 3 + (ph_ & 1) + (it & 1);
          const int r = tap / p.S, s_ = tap - r * p.S;
          mbar_wait(smem_u32(&bar_bfull[sb]), phb);
          tc_fence_after();
          const uint64_t db = make_sw128_kmajor_desc(smem_base + sb * B_BYTES);
          const uint32_t first = (uint32_t)((cb | it) != 0);
          if (p.halo_sw128) {
            // window start = 128-byte pixel row (r, s) of the swizzled patch; 8-row groups one patch row apart;
            // sub-tile j starts 8 pixel rows (1024 B -> +64 in the >>4 field) further, K steps +32 B (+2)
            const uint64_t da0 = make_sw128_kmajor_desc_ex(patch + (uint32_t)((r * p.pw + s_) * 128),
                                                           (uint32_t)(p.pw * 128), 0);
            if (p.split == 2 && cb < (p.cblocks >> 1)) {   // 8-bit correction block: the patch rows are (lo8 | hi8) bytes
#pragma unroll
              for (int jj = 0; jj < JN; ++jj) {
                const int j = jb + jj;
                const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MSUB + j) * BN);
                umma_f8_elect(d_tmem, da0 + (uint64_t)(j * 64), db, idesc, first);
#pragma unroll
                for (int k = 1; k < TC_BK / 16; ++k)
                  umma_f8_elect(d_tmem, da0 + (uint64_t)(j * 64 + k * 2), db + (uint64_t)(k * 2), idesc, 1u);
              }
            } else {
#pragma unroll
            for (int jj = 0; jj < JN; ++jj) {
              const int j = jb + jj;
              const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MSUB + j) * BN);
              umma_f16_elect(d_tmem, da0 + (uint64_t)(j * 64), db, idesc, first);
#pragma unroll
              for (int k = 1; k < TC_BK / 16; ++k)
                umma_f16_elect(d_tmem, da0 + (uint64_t)(j * 64 + k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            }
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < JN; ++jj) {
              const int j = jb + jj;
              const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MSUB + j) * BN);
              const uint32_t win = patch + (uint32_t)((r * p.pw + s_ + 8 * j) * 16);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint64_t da = make_nosw_kmajor_desc(win + (uint32_t)(2 * k) * p.plane_stride, lbo, sbo);
                umma_f16_elect(d_tmem, da, db + (uint64_t)(k * 2), idesc, k == 0 ? first : 1u);
              }
            }
          }
          umma_commit_elect(smem_u32(&bar_bempty[sb]));
          if (++sb == nb) { sb = 0; phb ^= 1; }
        }
        umma_commit_elect(smem_u32(&bar_aempty[ab]));   // patch free once every tap's MMAs have retired
        ab ^= 1;
        if (ab == 0) pha ^= 1;
      }
      umma_commit_elect(smem_u32(&bar_tfull[acc]));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================ epilogue ================================
    if (warp == 3) {
      // MSUB == 1: the second issuer slot is idle
    } else {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    if (COAL || !(p.nproj > 0 && half == 1)) {
      const int row = quad * 32 + lane;
      const int py = row >> 3, px = row & 7;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = (int)(tile % p.tiles_n);
        const long long m_tile = tile / p.tiles_n;
        const int img = (int)(m_tile / tiles_per_img);
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int ty0 = (t_in / p.tiles_x) * TH, tx0 = (t_in % p.tiles_x) * TW;
        const int y = ty0 + py, xb = tx0 + px;
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < MSUB; ++j) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * MSUB + j) * BN);
          if (COAL) {     // line-coalesced stores / residual loads through the warp's staging tile (8-pixel-wide sub-tile)
            CoalRows cr;
            coal_rows<3>(p, img, ty0, tx0 + 8 * j, n_tile * BN, quad, lane, cr);
            uint4 rg[4], rgl[4];
            if (p.res) {
              coal_load_res(p, cr, half, 0, rg);
              if (p.split) coal_load_res(p, cr, half, p.res_lo, rgl);
            }
            epilogue_coalesced<BN>(p, taddr, cr, n_tile * BN, half, lane, reinterpret_cast<uint4*>(tail) + (warp - 4) * 128,
                                   rg, rgl);
          } else {
            epilogue_rows<BN, false>(p, taddr, img, y, xb + 8 * j, n_tile, n_tile * BN, half, proj_w);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tap-pair kernel: the 64-wide stride-1 kxk layers (7x7 refinement head 64 -> 64 @512^2, the 3x3 64 -> 64 decoder convs,
// the grouped 3x3 convs whose N tile is one 64-channel slab).
// An M128 x N64 x K16 instruction occupies the tensor pipe for 32 cycles but reads 4 KB (A) + 2 KB (B) of shared memory,
// 48 cycles at 128 B/cycle: conv_halo_kernel<64, 2> tops out at ~2/3 of the pipe (ncu: 38-45 % active).  Here ONE
// instruction handles TWO horizontally adjacent taps (r, s) and (r, s + 1) over the same activation window: B stacks their
// weight tiles (rows 0-63 / 64-127, two consecutive taps of the K-major weight tensor = one 16 KB ring slot), N = 128,
// 4 KB + 4 KB per 64 cycles.  Columns 64-127 of the accumulator then hold A(p + (r, s)) W(r, s + 1): the contribution of
// tap (r, s + 1) to the output pixel one to the LEFT of p, so the epilogue adds lo(y, x) + hi(y, x + 1) (tap2_combine32:
// a lane shuffle, across the two sub-tiles at x = 7).  The last column of a 16-pixel-wide tile has no right neighbour in
// the tile: tiles advance by 15 columns and column 15 is recomputed by the next tile (6 % extra work).  An odd S leaves
// one single tap per filter row, issued as an N = 64 instruction into the lo columns.  k = 7: 4 instead of 7 issue slots
// per row, k = 3: 2 instead of 3.  Same warp roles and barrier protocol as conv_halo_kernel<64, 2> (two MMA issuers, one
// per sub-tile); TMEM: 2 accumulator sets x 2 sub-tiles x 128 columns = 512.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TAP2_TW = 15;      // output columns per tile (of 16 accumulator columns)

template <bool COAL>
__global__ void __launch_bounds__(TCH_THREADS, 1) conv_tap2_kernel(const __grid_constant__ ConvTcParams p) {
  constexpr int BN = 64, ACC = 128, MSUB = 2;
  constexpr uint32_t SLOT_BYTES = 2 * BN * TC_BK * 2;    // two taps
  constexpr uint32_t TAP_BYTES = BN * TC_BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * MSUB * ACC;
  constexpr int MAX_NB = 8;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_bfull[MAX_NB];
  __shared__ __align__(8) uint64_t bar_bempty[MAX_NB];
  __shared__ __align__(8) uint64_t bar_afull[2];
  __shared__ __align__(8) uint64_t bar_aempty[2];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nb_stages;
  const uint32_t patch_bytes = 8u * (uint32_t)p.plane_stride;
  const uint32_t a_base = smem_base + nb * SLOT_BYTES;
  const uint32_t a_tx = 8u * (uint32_t)(p.ph * p.pw * 16);
  uint8_t* tail = smem_raw + ((a_base - smem_u32(smem_raw)) + 2 * patch_bytes);
  float* proj_w = reinterpret_cast<float*>(tail);
  if (!COAL && p.nproj > 0) {
    int off = 0;
    for (int h = 0; h < p.nproj; ++h) {
      const int nw = p.proj[h].cout * BN;
      for (int i = threadIdx.x; i < nw; i += blockDim.x) proj_w[off + i] = p.proj[h].w[i];
      off += nw;
    }
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmH);
    prefetch_tmap(&p.tmB);
    for (int i = 0; i < nb; ++i) { mbar_init(smem_u32(&bar_bfull[i]), 1); mbar_init(smem_u32(&bar_bempty[i]), 2); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1);
      mbar_init(smem_u32(&bar_aempty[i]), 2);
      mbar_init(smem_u32(&bar_tfull[i]), 2);
      mbar_init(smem_u32(&bar_tempty[i]), p.nproj > 0 ? 4 : TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int npair = (p.S + 1) >> 1;                  // issue slots per filter row
  constexpr int TH = 16;

  if (warp == 0) {
    // ================================ B (weights) producer: one slot = taps (r, 2q) and (r, 2q + 1) ================================
    if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % p.tiles_n) * BN;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          int cbw = cb;
          if (p.split == 1) {
            const int cbl = p.cblocks / 3, pi = cb / cbl;
            cbw = split_pass_order(pi, p.split_lofirst) * cbl + (cb - pi * cbl);
          }
          for (int r = 0; r < p.R; ++r) {
            for (int q = 0; q < npair; ++q) {
              const int s0 = 2 * q;
              const bool two = s0 + 1 < p.S;
              mbar_wait(smem_u32(&bar_bempty[sb]), phb ^ 1);
              const uint32_t full = smem_u32(&bar_bfull[sb]);
              mbar_expect_tx(full, two ? SLOT_BYTES : TAP_BYTES);
              tma_load_3d(smem_base + sb * SLOT_BYTES, &p.tmB, full, cbw * TC_BK, n0, r * p.S + s0);
              if (two) tma_load_3d(smem_base + sb * SLOT_BYTES + TAP_BYTES, &p.tmB, full, cbw * TC_BK, n0, r * p.S + s0 + 1);
              if (++sb == nb) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ A (activation patch) producer ================================
    if (lane == 0) {
      int ab = 0;
      uint32_t pha = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const long long m_tile = tile / p.tiles_n;
        const int img = (int)(m_tile / tiles_per_img);
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int y0 = (t_in / p.tiles_x) * TH, x0 = (t_in % p.tiles_x) * TAP2_TW;
        const int n0 = (int)(tile % p.tiles_n) * BN;
        const int cbase = p.slab_mode ? (n0 / p.kslab) * p.kslab : 0;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          int a_ch = cb * TC_BK;
          if (p.split == 1) {
            const int cbl = p.cblocks / 3, pi = cb / cbl, pass = split_pass_order(pi, p.split_lofirst);
            a_ch = (cb - pi * cbl) * TC_BK + (pass == 1 ? p.a_lo : 0);
          } else if (p.split == 2) {
            const int cbl = p.cblocks / 2;
            a_ch = cb < cbl ? cb * TC_BK + p.a_lo : (cb - cbl) * TC_BK;
          }
          mbar_wait(smem_u32(&bar_aempty[ab]), pha ^ 1);
          const uint32_t full = smem_u32(&bar_afull[ab]);
          mbar_expect_tx(full, a_tx);
          tma_load_4d(a_base + ab * patch_bytes, &p.tmH, full, cbase + a_ch, x0 - p.pad, y0 - p.pad, img);
          ab ^= 1;
          if (ab == 0) pha ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ================================ MMA issuers: warp 1 -> sub-tile 0, warp 3 -> sub-tile 1 ================================
    constexpr uint32_t idesc2 = make_idesc_f16(TC_BM, ACC), idesc1 = make_idesc_f16(TC_BM, BN);
    const int j = (warp == 3) ? 1 : 0;
    int sb = 0, ab = 0, acc = 0;
    uint32_t phb = 0, pha = 0, acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MSUB + j) * ACC);
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(smem_u32(&bar_afull[ab]), pha);
        tc_fence_after();
        const uint32_t patch = a_base + ab * patch_bytes;
        const bool f8 = p.split == 2 && cb < (p.cblocks >> 1);
        for (int r = 0; r < p.R; ++r) {
          for (int q = 0; q < npair; ++q) {
            const int s0 = 2 * q;
            const bool two = s0 + 1 < p.S;
            mbar_wait(smem_u32(&bar_bfull[sb]), phb);
            tc_fence_after();
            const uint64_t db = make_sw128_kmajor_desc(smem_base + sb * SLOT_BYTES);
            const uint64_t da = make_sw128_kmajor_desc_ex(patch + (uint32_t)((r * p.pw + s0) * 128), (uint32_t)(p.pw * 128), 0)
                                + (uint64_t)(j * 64);
            const uint32_t first = (uint32_t)((cb | r | q) != 0);
            const uint32_t idesc = two ? idesc2 : idesc1;
            if (f8) {
              umma_f8_elect(d_tmem, da, db, idesc, first);
#pragma unroll
              for (int k = 1; k < TC_BK / 16; ++k) umma_f8_elect(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            } else {
              umma_f16_elect(d_tmem, da, db, idesc, first);
#pragma unroll
              for (int k = 1; k < TC_BK / 16; ++k) umma_f16_elect(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            }
            umma_commit_elect(smem_u32(&bar_bempty[sb]));
            if (++sb == nb) { sb = 0; phb ^= 1; }
          }
        }
        umma_commit_elect(smem_u32(&bar_aempty[ab]));
        ab ^= 1;
        if (ab == 0) pha ^= 1;
      }
      umma_commit_elect(smem_u32(&bar_tfull[acc]));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    if (COAL || !(p.nproj > 0 && half == 1)) {
      const int row = quad * 32 + lane;
      const int py = row >> 3, px = row & 7;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = (int)(tile % p.tiles_n);
        const long long m_tile = tile / p.tiles_n;
        const int img = (int)(m_tile / tiles_per_img);
        const int t_in = (int)(m_tile - (long long)img * tiles_per_img);
        const int ty0 = (t_in / p.tiles_x) * TH, tx0 = (t_in % p.tiles_x) * TAP2_TW;
        mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < MSUB; ++j) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * MSUB + j) * ACC);
          if (COAL) {
            CoalRows cr;
            coal_rows<3>(p, img, ty0, tx0 + 8 * j, n_tile * BN, quad, lane, cr, tx0 + TAP2_TW);
            uint4 rg[4], rgl[4];
            if (p.res) {
              coal_load_res(p, cr, half, 0, rg);
              if (p.split) coal_load_res(p, cr, half, p.res_lo, rgl);
            }
            epilogue_coalesced<BN, true>(p, taddr, cr, n_tile * BN, half, lane,
                                         reinterpret_cast<uint4*>(tail) + (warp - 4) * 128, rg, rgl, j);
          } else {
            epilogue_rows<BN, false, true>(p, taddr, img, ty0 + py, tx0 + px + 8 * j, n_tile, n_tile * BN, half, proj_w,
                                           nullptr, j, lane, px + 8 * j < TAP2_TW);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch configuration
// ---------------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

static int encode_map(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                      const cuuint32_t* box, bool swizzle128 = true) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
  CPN_REQUIRE(fn != nullptr, "conv_tc: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CPN_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2]);
  return 0;
}

static int pick_bn(int cout, int slab_mode) {
  if (slab_mode) return 64;
  if (cout % 256 == 0) return 256;
  if (cout % 128 == 0) return 128;
  return 64;
}

int conv_tc_plan_create(const cpn_op_t& op_in, const void* src, void* dst, const void* res, const void* wgt,
                        const float* bias, ConvTcPlan** out) {
  cpn_op_t op = op_in;      // (CPN_CONV_UP2 rewrites the logical output dims below)
  const bool f16f8 = op.src.dtype == CPN_DT_F16F8;
  const bool split = op.src.dtype == CPN_DT_F16X2 || f16f8;    // tensors with a second (lo / 8-bit) block per pixel
  const int npass = f16f8 ? 2 : (split ? 3 : 1);
  CPN_REQUIRE((op.src.dtype == CPN_DT_F16 || split) && op.dst.dtype == op.src.dtype,
              "conv_tc: fp16 (or split fp16 / fp16+e4m3) activations required");
  CPN_REQUIRE(!f16f8 || (op.src.lo_delta % 32 == 0 && op.dst.lo_delta % 32 == 0 && (op.res.n == 0 || op.res.lo_delta % 32 == 0)),
              "conv_tc: fp16+e4m3 tensors need lo_delta multiples of 32 elements");
  CPN_REQUIRE(op.res.n == 0 || op.res.dtype == op.src.dtype, "conv_tc: residual dtype must match");
  CPN_REQUIRE(!split || (op.src.lo_delta % 8 == 0 && op.dst.lo_delta % 8 == 0 && (op.res.n == 0 || op.res.lo_delta % 8 == 0)),
              "conv_tc: lo_delta must be a multiple of 8 elements");
  CPN_REQUIRE(op.kslab % TC_BK == 0, "conv_tc: kslab %d must be a multiple of 64", op.kslab);
  CPN_REQUIRE(op.dst.c % 64 == 0, "conv_tc: cout %d must be a multiple of 64", op.dst.c);
  CPN_REQUIRE(op.stride == 1 || op.stride == 2, "conv_tc: stride %d unsupported", op.stride);
  CPN_REQUIRE(op.src.pitch % 8 == 0 && op.dst.pitch % 8 == 0 && (op.res.n == 0 || op.res.pitch % 8 == 0),
              "conv_tc: pitches must be multiples of 8 elements (16 bytes)");
  CPN_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) && ((uintptr_t)wgt % 16 == 0) &&
                  (res == nullptr || (uintptr_t)res % 16 == 0) && (bias == nullptr || (uintptr_t)bias % 16 == 0),
              "conv_tc: base pointers must be 16-byte aligned");
  const bool up2 = (op_in.flags & CPN_CONV_UP2) != 0;
  if (up2) {
    // the GEMM runs at the source resolution with 4 * cout columns; only the epilogue knows about the 2x output
    CPN_REQUIRE(op_in.r == 3 && op_in.s == 3 && op_in.stride == 1 && op_in.pad == 1 && op_in.res.n == 0 && op_in.slab_mode == 0 &&
                    op_in.fuse_next == 0 && op_in.dst.h == 2 * op_in.src.h && op_in.dst.w == 2 * op_in.src.w &&
                    (4 * op_in.dst.c) % 256 == 0 && op_in.dst.c % 32 == 0,
                "conv_tc: CPN_CONV_UP2 needs a 3x3/s1/p1 dense conv without residual, dst = 2x src, 4*cout %% 256 == 0");
    op.dst.h = op_in.src.h; op.dst.w = op_in.src.w; op.dst.c = 4 * op_in.dst.c;
  }
  const int eh = (op.src.h + 2 * op.pad - op.r) / op.stride + 1, ew = (op.src.w + 2 * op.pad - op.s) / op.stride + 1;
  CPN_REQUIRE(eh == op.dst.h && ew == op.dst.w && op.src.n == op.dst.n,
              "conv_tc: output shape mismatch (%dx%d expected %dx%d)", op.dst.h, op.dst.w, eh, ew);
  CPN_REQUIRE(op.slab_mode == 0 ? op.kslab <= op.src.c : true, "conv_tc: kslab exceeds input channels");

  std::unique_ptr<ConvTcPlan> owner(new ConvTcPlan());     // released to the caller on success only
  ConvTcPlan* pl = owner.get();
  ConvTcParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  int bn = pick_bn(op.dst.c, op.slab_mode);
  {
    // experiment switch: N tile of the 1x1 layers (CPN_BN_1X1=128|64); smaller tiles = more pipeline stages in flight
    static int bn1x1_env = -1;
    if (bn1x1_env < 0) { const char* e = getenv("CPN_BN_1X1"); bn1x1_env = e ? atoi(e) : 0; }
    if (bn1x1_env > 0 && op.r * op.s == 1 && !op.slab_mode && op.fuse_next == 0 && op.dst.c % bn1x1_env == 0 &&
        (bn1x1_env == 64 || bn1x1_env == 128 || bn1x1_env == 256))
      bn = bn1x1_env;
  }
  pl->bn = bn;
  // --- A maps: (C, W, H, N) with box {64, 16, 8, 1}; stride 2 -> four parity views
  const int nmaps = op.stride == 2 ? 4 : 1;
  for (int m = 0; m < nmaps; ++m) {
    const int py = m >> 1, px = m & 1;
    const int st = op.stride;
    const long long wv = (op.src.w - px + st - 1) / st, hv = (op.src.h - py + st - 1) / st;
    if (wv <= 0 || hv <= 0) {  // degenerate parity view (1-pixel maps): alias view 0, never selected with data
      p.tmA[m] = p.tmA[0];
      continue;
    }
    const char* base = reinterpret_cast<const char*>(src) + ((long long)py * op.src.w + px) * op.src.pitch * 2;
    cuuint64_t dims[4] = {(cuuint64_t)(op.src.c + (split ? op.src.lo_delta : 0)), (cuuint64_t)wv, (cuuint64_t)hv,
                          (cuuint64_t)op.src.n};
    cuuint64_t strides[3] = {(cuuint64_t)st * op.src.pitch * 2, (cuuint64_t)st * op.src.w * op.src.pitch * 2,
                             (cuuint64_t)op.src.h * op.src.w * op.src.pitch * 2};
    cuuint32_t box[4] = {TC_BK, TC_BW, TC_BH, 1};
    if (encode_map(&p.tmA[m], const_cast<char*>(base), 4, dims, strides, box)) return 1;
  }
  for (int m = nmaps; m < 4; ++m) p.tmA[m] = p.tmA[0];
  {
    const cuuint64_t kw = (cuuint64_t)op.kslab * npass;   // F16X2: (W_hi | W_hi | W_lo) along K; F16F8: (W8 | S * W_hi)
    cuuint64_t dims[3] = {kw, (cuuint64_t)op.dst.c, (cuuint64_t)(op.r * op.s)};
    cuuint64_t strides[2] = {kw * 2, kw * op.dst.c * 2};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)bn, 1};
    if (encode_map(&p.tmB, const_cast<void*>(wgt), 3, dims, strides, box)) return 1;
  }
  p.out = reinterpret_cast<__half*>(dst);
  p.res = op.res.n ? reinterpret_cast<const __half*>(res) : nullptr;
  p.bias = bias;
  p.out_pitch = op.dst.pitch; p.res_pitch = op.res.pitch;
  p.res_h = op.res.n ? op.res.h : 0; p.res_w = op.res.n ? op.res.w : 0;
  p.N = op.dst.n; p.Ho = op.dst.h; p.Wo = op.dst.w; p.cout = op.dst.c;
  p.R = op.r; p.S = op.s; p.stride = op.stride; p.pad = op.pad;
  p.cblocks = op.kslab / TC_BK * npass; p.kslab = op.kslab; p.slab_mode = op.slab_mode;
  p.split = f16f8 ? 2 : (split ? 1 : 0); p.a_lo = op.src.lo_delta; p.out_lo = op.dst.lo_delta;
  p.res_lo = op.res.n ? op.res.lo_delta : 0;
  p.acc_scale = op.acc_scale != 0.f ? op.acc_scale : 1.f;
  p.out_lo_scale = ldexpf(1.f, 8 + op.dst.fp8_exp); p.out_hi8_scale = ldexpf(1.f, -2 + op.dst.fp8_exp);
  p.res_lo_inv = ldexpf(1.f, -(8 + (op.res.n ? op.res.fp8_exp : 0)));
  p.relu = op.act == CPN_ACT_RELU;
  {
    static int rot_env = -1;
    // Off by default: measured no gain on B200 (profiles/r01_summary.md), and a CTA-dependent summation order would
    // make a tile's result depend on the batch it is computed in (the tests assert batch invariance bit for bit).
    if (rot_env < 0) { const char* e = getenv("CPN_ROTATE"); rot_env = (e && atoi(e) == 1) ? 1 : 0; }
    p.rotate = rot_env;
    static int lofirst_env = -1;
    if (lofirst_env < 0) { const char* e = getenv("CPN_SPLIT_LOFIRST"); lofirst_env = (e && atoi(e) == 0) ? 0 : 1; }
    p.split_lofirst = lofirst_env;
  }
  p.tiles_x = (op.dst.w + TC_BW - 1) / TC_BW; p.tiles_y = (op.dst.h + TC_BH - 1) / TC_BH;
  p.tiles_n = op.dst.c / bn;
  p.total_tiles = (long long)op.dst.n * p.tiles_x * p.tiles_y * p.tiles_n;
  // ---- halo variant for stride-1 kxk dense convolutions (CPN_HALO=0 disables) ----
  pl->msub = 1;
  {
    static int halo_env = -1, swap_env = -1, sw128_env = -1, baseoff_env = -1, halo33_env = -1;
    if (halo_env < 0) { const char* e = getenv("CPN_HALO"); halo_env = (e && atoi(e) == 0) ? 0 : 1; }
    if (sw128_env < 0) { const char* e = getenv("CPN_HALO_SW128"); sw128_env = (e && atoi(e) == 0) ? 0 : 1; }
    if (baseoff_env < 0) { const char* e = getenv("CPN_HALO_BASEOFF"); baseoff_env = (e && atoi(e) == 1) ? 1 : 0; }
    if (halo33_env < 0) { const char* e = getenv("CPN_HALO_ALL"); halo33_env = (e && atoi(e) == 0) ? 0 : 1; }
    if (swap_env < 0) { const char* e = getenv("CPN_HALO_SWAP"); swap_env = (e && atoi(e) == 1) ? 1 : 0; }
    // measured (profiles/r01_summary.md): with the warp-convergent MMA issue the halo variant wins for every stride-1
    // kxk layer (3x3 @256-wide: 1500-1650 vs 1360-1590 TF/s); CPN_HALO_ALL=0 restricts it to 7x7 and <= 128-wide tiles.
    const bool worth = halo33_env || op.r * op.s >= 25 || bn <= 128;
    if (halo_env && worth && op.stride == 1 && op.r * op.s > 1 && op.r <= 16 && op.s <= 16) {
      const int msub = bn == 256 ? 1 : 2;
      const int pw = 8 * msub + op.s - 1, ph = 16 + op.r - 1;
      const int plane_stride = sw128_env ? ((ph * pw * 128 + 1023) / 1024 * 1024) / 8 : (ph * pw * 16 + 127) / 128 * 128;
      const int patch = 8 * plane_stride;
      const int b_bytes = bn * TC_BK * 2;
      int nbs = (TC_SMEM_BUDGET - 2 * patch) / b_bytes;
      if (nbs > 8) nbs = 8;
      if (nbs >= 2 && pw <= 256 && ph <= 256) {
        cuuint64_t dims[4] = {(cuuint64_t)(op.src.c + (split ? op.src.lo_delta : 0)), (cuuint64_t)op.src.w,
                              (cuuint64_t)op.src.h, (cuuint64_t)op.src.n};
        cuuint64_t strides[3] = {(cuuint64_t)op.src.pitch * 2, (cuuint64_t)op.src.w * op.src.pitch * 2,
                                 (cuuint64_t)op.src.h * op.src.w * op.src.pitch * 2};
        cuuint32_t box[4] = {(cuuint32_t)(sw128_env ? 64 : 8), (cuuint32_t)pw, (cuuint32_t)ph, 1};
        if (encode_map(&p.tmH, const_cast<void*>(src), 4, dims, strides, box, sw128_env != 0)) return 1;
        p.halo_sw128 = sw128_env; p.halo_baseoff = baseoff_env;
        CPN_REQUIRE(!f16f8 || sw128_env, "conv_tc: the fp16+e4m3 engine needs the SWIZZLE_128B halo patch (CPN_HALO_SW128)");
        p.halo = 1; p.pw = pw; p.ph = ph; p.plane_stride = plane_stride; p.nb_stages = nbs; p.swap_lbo_sbo = swap_env;
        pl->msub = msub;
        p.tiles_x = (op.dst.w + 8 * msub - 1) / (8 * msub);
        p.tiles_y = (op.dst.h + 15) / 16;
        p.total_tiles = (long long)op.dst.n * p.tiles_x * p.tiles_y * p.tiles_n;
        {
          // tap-pair kernel for the 64-wide tiles: two taps per N = 128 instruction.  Measured (profiles/r02h_ab_tap2.log):
          // 7x7 refinement head 3.93 -> 2.52 ms (2-pass engine), 2.08 -> 1.88 ms (fp16); 3x3 layers get SLOWER (1.02 ->
          // 1.15 ms: 2 instead of 3 issue slots per filter row do not pay for the three-fold TMEM read + shuffles of the
          // epilogue), so it is used for >= 25 taps only.  CPN_TAP2=0 disables it, CPN_TAP2=2 forces it for every k >= 3.
          static int tap2_env = -1;
          if (tap2_env < 0) { const char* e = getenv("CPN_TAP2"); tap2_env = e ? atoi(e) : 1; }
          const int nbs2 = (TC_SMEM_BUDGET - 2 * patch) / (2 * b_bytes);
          if (tap2_env && (op.r * op.s >= 25 || tap2_env == 2) && bn == 64 && msub == 2 && sw128_env && op.s >= 3 && nbs2 >= 2) {
            p.tap2 = 1;
            p.nb_stages = nbs2 > 8 ? 8 : nbs2;
            p.tiles_x = (op.dst.w + TAP2_TW - 1) / TAP2_TW;
            p.total_tiles = (long long)op.dst.n * p.tiles_x * p.tiles_y * p.tiles_n;
          }
        }
      }
    }
  }
  const int stage_bytes = TC_BM * TC_BK * 2 + bn * TC_BK * 2;
  pl->stages = TC_SMEM_BUDGET / stage_bytes;
  if (pl->stages > 8) pl->stages = 8;
  {
    static int stages_env = -1;   // experiment switch: cap the operand ring (CPN_TC_STAGES)
    if (stages_env < 0) { const char* e = getenv("CPN_TC_STAGES"); stages_env = e ? atoi(e) : 0; }
    if (stages_env >= 2 && stages_env < pl->stages) pl->stages = stages_env;
  }
  pl->proj_smem_bytes = 0;
  {
    // line-coalesced epilogue (CPN_COALESCE=0 disables): single-precision-storage layers whose tensors can be
    // addressed with 32-bit element offsets; conv_tc_fuse_proj switches it off again for fused ReadOut heads
    static int coal_env = -1;
    if (coal_env < 0) { const char* e = getenv("CPN_COALESCE"); coal_env = (e && atoi(e) == 0) ? 0 : 1; }
    const long long out_elems = (long long)op_in.dst.n * op_in.dst.h * op_in.dst.w * op_in.dst.pitch;
    const long long res_elems = op.res.n ? (long long)op.res.n * op.res.h * op.res.w * op.res.pitch : 0;
    static int coal_halo_env = -1;   // CPN_COALESCE_HALO=0: keep the direct epilogue in conv_halo_kernel only (A/B switch)
    if (coal_halo_env < 0) { const char* e = getenv("CPN_COALESCE_HALO"); coal_halo_env = (e && atoi(e) == 0) ? 0 : 1; }
    static int dbg_epi_env = -1;
    if (dbg_epi_env < 0) { const char* e = getenv("CPN_DBG_EPI"); dbg_epi_env = (e && atoi(e) == 1) ? 1 : 0; }
    p.dbg_epi = dbg_epi_env;
    static int coal_split_env = -1;  // CPN_COALESCE_SPLIT=0: direct epilogue for the split-precision engine (A/B switch)
    if (coal_split_env < 0) { const char* e = getenv("CPN_COALESCE_SPLIT"); coal_split_env = (e && atoi(e) == 0) ? 0 : 1; }
    p.coalesce = (coal_env && (coal_split_env || !split) && (coal_halo_env || !p.halo) && out_elems < (1ll << 31) &&
                  res_elems < (1ll << 31)) ? 1 : 0;
  }
  if (up2) {
    CPN_REQUIRE(p.halo && p.coalesce && !p.tap2, "conv_tc: CPN_CONV_UP2 needs the halo kernel with the coalesced epilogue");
    p.up2 = 1; p.cup = op_in.dst.c; p.Ho2 = op_in.dst.h; p.Wo2 = op_in.dst.w;
    {
      static int skip_env = -1;   // CPN_UP2_SKIP=0: issue all nine taps for every phase (A/B switch)
      if (skip_env < 0) { const char* e = getenv("CPN_UP2_SKIP"); skip_env = (e && atoi(e) == 0) ? 0 : 1; }
      p.up2_skip = (skip_env && p.cup % bn == 0 && !p.rotate) ? 1 : 0;
    }
  }
  pl->smem_bytes = pl->stages * stage_bytes + 1024 + (p.coalesce ? TC_STAGING_BYTES : 0);
  if (p.halo) pl->smem_bytes = p.nb_stages * bn * TC_BK * 2 * (p.tap2 ? 2 : 1) + 2 * 8 * p.plane_stride + 1024 +
                               (p.coalesce ? TC_STAGING_BYTES : 0);
  const long long sms = sm_count();
  pl->grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  {
    // CTA-pair kernel for the 1x1 layers with 256-wide N tiles: opt-in (CPN_PAIR=1).  Measured on B200 (profiles/
    // r02_summary.md): bit-identical results, but 3-20 % SLOWER than conv_tc_kernel<256> per layer although it moves a
    // third fewer operand bytes per CTA -- these layers are not bound by operand bytes (see the kernel's header).
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("CPN_PAIR"); pair_env = (e && atoi(e) == 1) ? 1 : 0; }
    const long long m_tiles = (long long)op.dst.n * p.tiles_x * p.tiles_y;
    pl->pair = 0;
    if (pair_env && !p.halo && bn == TCP_BN && op.r * op.s == 1 && op.fuse_next == 0 && m_tiles >= 2 && sms >= 2) {
      {   // each CTA of a pair loads HALF of the weight tile: box of 128 output channels
        const cuuint64_t kw = (cuuint64_t)op.kslab * npass;
        cuuint64_t dims[3] = {kw, (cuuint64_t)op.dst.c, (cuuint64_t)(op.r * op.s)};
        cuuint64_t strides[2] = {kw * 2, kw * op.dst.c * 2};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)(TCP_BN / 2), 1};
        if (encode_map(&p.tmB, const_cast<void*>(wgt), 3, dims, strides, box)) return 1;
      }
      pl->pair = 1;
      pl->stages = TC_SMEM_BUDGET / (int)TCP_STAGE_BYTES;
      if (pl->stages > 8) pl->stages = 8;
      pl->smem_bytes = pl->stages * (int)TCP_STAGE_BYTES + 1024 + (p.coalesce ? TC_STAGING_BYTES : 0);
      const long long pair_tiles = ((m_tiles + 1) / 2) * p.tiles_n;
      const long long pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
      pl->grid = (int)(2 * pairs);
    }
  }
  pl->full_tiles = p.total_tiles;
  pl->full_grid = pl->grid;
  *out = owner.release();
  return 0;
}

// Run only the M tiles that hold the first `rows` pixels (row-major over the whole batch) of the next launches: the sparse
// heads' row matrix is sized for a power-of-two row count but only the leading P rows are proposals.  Tiles are ordered N
// fastest, so trimming the tile count trims whole 128-pixel row blocks.  rows < 0 restores the full launch.
int conv_tc_limit_rows(ConvTcPlan* pl, long long rows) {
  ConvTcParams& p = pl->p;
  if (rows < 0) { p.total_tiles = pl->full_tiles; pl->grid = pl->full_grid; return 0; }
  CPN_REQUIRE(!pl->pair && !p.halo && p.tiles_x == 1, "conv_tc: row limit needs a plain 1x1 layer over a [1, rows/16, 16] tensor");
  long long m_tiles = (rows + TC_BM - 1) / TC_BM;
  if (m_tiles < 1) m_tiles = 1;
  long long tiles = m_tiles * p.tiles_n;
  if (tiles > pl->full_tiles) tiles = pl->full_tiles;
  p.total_tiles = tiles;
  pl->grid = (int)(tiles < pl->full_grid ? tiles : pl->full_grid);
  return 0;
}

template <bool COAL>
static int launch_pair2(const ConvTcPlan* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(conv_pair_kernel<COAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TC_SMEM_BUDGET + 1024 + TC_PROJ_SMEM_MAX));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)pl->grid, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = (size_t)pl->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<COAL>, pl->p, pl->stages));
  CPN_CHECK_LAUNCH();
  return 0;
}

template <int BN, bool COAL>
static int launch_bn2(const ConvTcPlan* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, COAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TC_SMEM_BUDGET + 1024 + TC_PROJ_SMEM_MAX));
    attr_set = true;
  }
  conv_tc_kernel<BN, COAL><<<pl->grid, TC_THREADS, pl->smem_bytes, st>>>(pl->p, pl->stages);
  CPN_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int launch_bn(const ConvTcPlan* pl, cudaStream_t st) {
  return pl->p.coalesce ? launch_bn2<BN, true>(pl, st) : launch_bn2<BN, false>(pl, st);
}

template <int BN, int MSUB, bool COAL>
static int launch_halo2(const ConvTcPlan* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, MSUB, COAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TC_SMEM_BUDGET + 1024 + TC_PROJ_SMEM_MAX));
    attr_set = true;
  }
  conv_halo_kernel<BN, MSUB, COAL><<<pl->grid, TCH_THREADS, pl->smem_bytes, st>>>(pl->p);
  CPN_CHECK_LAUNCH();
  return 0;
}

template <int BN, int MSUB>
static int launch_halo(const ConvTcPlan* pl, cudaStream_t st) {
  return pl->p.coalesce ? launch_halo2<BN, MSUB, true>(pl, st) : launch_halo2<BN, MSUB, false>(pl, st);
}

template <bool COAL>
static int launch_tap2(const ConvTcPlan* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CPN_CHECK_CUDA(cudaFuncSetAttribute(conv_tap2_kernel<COAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TC_SMEM_BUDGET + 1024 + TC_PROJ_SMEM_MAX));
    attr_set = true;
  }
  conv_tap2_kernel<COAL><<<pl->grid, TCH_THREADS, pl->smem_bytes, st>>>(pl->p);
  CPN_CHECK_LAUNCH();
  return 0;
}

int conv_tc_launch(const ConvTcPlan* pl, cudaStream_t st) {
  if (pl->pair) return pl->p.coalesce ? launch_pair2<true>(pl, st) : launch_pair2<false>(pl, st);
  if (pl->p.halo && pl->p.tap2) return pl->p.coalesce ? launch_tap2<true>(pl, st) : launch_tap2<false>(pl, st);
  if (pl->p.halo) {
    if (pl->bn == 64 && pl->msub == 2) return launch_halo<64, 2>(pl, st);
    if (pl->bn == 128 && pl->msub == 2) return launch_halo<128, 2>(pl, st);
    if (pl->bn == 256 && pl->msub == 1) return launch_halo<256, 1>(pl, st);
    set_error("conv_tc: bad halo configuration BN %d MSUB %d", pl->bn, pl->msub);
    return 1;
  }
  switch (pl->bn) {
    case 64: return launch_bn<64>(pl, st);
    case 128: return launch_bn<128>(pl, st);
    case 256: return launch_bn<256>(pl, st);
  }
  set_error("conv_tc: bad BN %d", pl->bn);
  return 1;
}

// Fuse `n` ReadOut projections (one per N tile, in tile order) into the epilogue; the convolution then writes the fp32
// head records instead of its fp16 output.  Output pointers are (re)bound with conv_tc_bind_proj_out before launches.
int conv_tc_fuse_proj(ConvTcPlan* pl, int n, const cpn_op_t* projs, const char* weights) {
  CPN_REQUIRE(n >= 1 && n <= TC_MAX_HEADS && n == pl->p.tiles_n, "conv_tc: %d fused projections for %d N tiles", n,
              pl->p.tiles_n);
  CPN_REQUIRE(pl->p.res == nullptr, "conv_tc: fused projection with residual is not supported");
  int total = 0;
  for (int h = 0; h < n; ++h) {
    const cpn_op_t& q = projs[h];
    CPN_REQUIRE(q.kind == CPN_OP_PROJ && q.proj_cin == pl->bn && q.proj_cin_off == h * pl->bn && q.dst.c <= TC_PROJ_MAX &&
                    q.dst.dtype == CPN_DT_F32,
                "conv_tc: projection %d cannot be fused (cin %d off %d cout %d)", h, q.proj_cin, q.proj_cin_off, q.dst.c);
    ProjHead& H = pl->p.proj[h];
    H.w = reinterpret_cast<const float*>(weights + q.w_offset);
    H.b = q.b_offset >= 0 ? reinterpret_cast<const float*>(weights + q.b_offset) : nullptr;
    H.out = nullptr;
    H.cout = q.dst.c; H.pitch = q.dst.pitch; H.act = q.act; H.act_scale = q.act_scale;
    total += q.dst.c * pl->bn * 4;
  }
  CPN_REQUIRE(total <= TC_PROJ_SMEM_MAX, "conv_tc: fused projection weights (%d B) exceed %d B", total, TC_PROJ_SMEM_MAX);
  pl->p.nproj = n;
  pl->proj_smem_bytes = total;
  if (pl->p.coalesce) { pl->p.coalesce = 0; pl->smem_bytes -= TC_STAGING_BYTES; }   // the tail holds the projection weights
  pl->smem_bytes += total;
  return 0;
}

void conv_tc_bind_proj_out(ConvTcPlan* pl, int head, void* out) { pl->p.proj[head].out = reinterpret_cast<float*>(out); }

void conv_tc_plan_destroy(ConvTcPlan* p) { delete p; }

}  // namespace cpn
