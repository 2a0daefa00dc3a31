// tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (TF32 operands, FP32 accumulate — the
// numerics class of the reference's own cuDNN path, which allows TF32 by default).
//
//   fprop / dgrad  (tc_conv_kernel):   D[pixels(128) x Cout(BN)] += A_tap[pixels x 32ch] * B_tap[32ch x BN]
//       A: TMA 4-D box (32 ch, BW, BH, 1) of the NHWC activation at the tap-shifted coordinate; zero padding
//          is the TMA out-of-bounds fill, stride 2 is the tensor map's element stride, a channel concat is a
//          second tensor map.  K-major, SWIZZLE_128B.
//       B: TMA 2-D box (32 k, BN) of the [Cout][kh*kw*Cin] weight matrix.  K-major, SWIZZLE_128B.
//       dgrad is the same kernel run on dy with the transposed + tap-flipped weight matrix.
//   wgrad (tc_wgrad_kernel):  dW^T[(tap,ci)(128) x Cout(BN)] += X_tap[32 px x 128 (tap,ci)]^T * dY[32 px x BN]
//       both operands MN-major straight out of NHWC (pixels are the GEMM-K), split over pixel ranges,
//       partial sums reduced with red.global.add.f32.
//
// Warp roles (1 CTA / SM, persistent): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
// allocator, warps 4.. = epilogue (TMEM -> registers -> bias/activation -> smem transpose -> global): eight of them
// in the fprop/dgrad kernels (384 threads; a single warp per SM sub-partition ran the epilogue latency-bound at
// ~2.3 TB/s of output, measured on the 64->256 1x1 layers), four in the wgrad kernels (256 threads).
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace segsde {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Split form for software pipelining: issue the load of the next accumulator chunk, work on the current one, then
// wait.  The wait names all 32 registers as read-write operands so the compiler cannot move a use above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor: version 1, layout type 2)
// layout_type: 2 = SWIZZLE_128B (16-byte swizzle atoms, 8-row period; K-major operands),
//              1 = SWIZZLE_128B_BASE32B (32-byte atoms, 4-row period) - the only layout tcgen05 accepts for
//                  MN-major tf32 operands (cutlass sm100_common.inl:92)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2, uint32_t base_offset = 0) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7u) << 49) |
         ((uint64_t)layout_type << 61);
}
// instruction descriptor for kind::tf32, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// epilogue activation: ELU through ex2.approx (abs error ~1e-7, far below the TF32 operand rounding of this path)
__device__ __forceinline__ float act_tc(float v, int act) {
  if (act == SEGSDE_ACT_ELU) return v > 0.f ? v : __expf(v) - 1.f;
  if (act == SEGSDE_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SEGSDE_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

// Epilogue of one 32-row x 32-column accumulator chunk owned by one warp: bias + activation in registers, then a
// transpose through a padded shared-memory tile so that every st.global.v4 instruction of the warp writes four
// complete 128-byte lines (the direct form - each lane storing 16 B of its own row - cost 0.5 ms of a 1.5 ms layer).
// dst8[i] = output pointer (first channel of the CTA's N tile) of accumulator row (lane>>3) + 4*i of this warp,
// nullptr when that pixel is outside the image; valid = ballot of "row `lane` is a real pixel".  Both are computed
// once per tile by the caller (the div/mod of the pixel decomposition stays out of the per-chunk path).
constexpr int EPI_LD = 36;     // floats per staged row: 144 B keeps both the 128-bit writes and reads conflict free
// csum (nullable): CTA-level shared accumulators [2][Cout] of sum / sum of squares per output channel (BatchNorm
// batch statistics fused into the producing convolution); csum_c = first channel of this chunk.
__device__ __forceinline__ void epilogue_chunk(float (&v)[32], const float* __restrict__ bias32, int act,
                                               float* __restrict__ stage, int lane, float* const (&dst8)[8], int col0,
                                               unsigned valid, float* __restrict__ csum = nullptr, int csum_c = 0,
                                               int cout = 0) {
  if (bias32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias32) + j);
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  switch (act) {
    case SEGSDE_ACT_ELU:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : __expf(v[i]) - 1.f;
      break;
    case SEGSDE_ACT_RELU:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      break;
    case SEGSDE_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
      break;
    default: break;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * EPI_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int c16 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i;
    if (dst8[i]) *reinterpret_cast<float4*>(dst8[i] + col0 + c16 * 4) = *reinterpret_cast<const float4*>(stage + r * EPI_LD + c16 * 4);
  }
  if (csum) {      // lane = channel: sum the staged column over the rows that are real pixels
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains
    if (valid == 0xffffffffu) {
#pragma unroll
      for (int r = 0; r < 32; ++r) { const float t = stage[r * EPI_LD + lane]; s1[r & 3] += t; s2[r & 3] = fmaf(t, t, s2[r & 3]); }
    } else {
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (valid & (1u << r)) { const float t = stage[r * EPI_LD + lane]; s1[r & 3] += t; s2[r & 3] = fmaf(t, t, s2[r & 3]); }
      }
    }
    atomicAdd(csum + csum_c + lane, (s1[0] + s1[1]) + (s1[2] + s1[3]));
    atomicAdd(csum + cout + csum_c + lane, (s2[0] + s2[1]) + (s2[2] + s2[3]));
  }
  __syncwarp();
}

// One warp's share of a tile's epilogue: chunks first, first + step, ... < nch (32 accumulator columns each) with the
// TMEM load of the next chunk in flight while the current one is activated / staged / stored.
template <int MAXI>
__device__ __forceinline__ void epilogue_chunks(uint32_t taddr0, int first, int step, int nch, const float* bias, int act,
                                                float* __restrict__ stage, int lane, float* const (&dst8)[8],
                                                unsigned valid, float* __restrict__ csum, int csum_c0, int cout) {
  uint32_t ra[32], rb[32];
  if (first < nch) tmem_ld32_issue(taddr0 + first * 32, ra);
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int cc = first + i * step;
    if (cc < nch) {
      float v[32];
      if (i & 1) {
        tmem_ld32_wait(rb);
        if (cc + step < nch) tmem_ld32_issue(taddr0 + (cc + step) * 32, ra);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(rb[k]);
      } else {
        tmem_ld32_wait(ra);
        if (cc + step < nch) tmem_ld32_issue(taddr0 + (cc + step) * 32, rb);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(ra[k]);
      }
      epilogue_chunk(v, bias ? bias + cc * 32 : nullptr, act, stage, lane, dst8, cc * 32, valid, csum, csum_c0 + cc * 32, cout);
    }
  }
}

constexpr int MAX_STAT_C = 2048;
constexpr int A_BYTES = 128 * 128;        // 128 rows x 32 fp32
constexpr int WSTAGES = 3;                // generic wgrad: 3 stages so that two CTAs share an SM (measured on the
                                          // <=64-channel 3x3 wgrad: a second resident CTA hides the TMA round trip
                                          // far better than a deeper pipeline in one CTA, 1.8 -> 1.15 ms)
constexpr int NT = 256;                   // wgrad kernels: 4 epilogue warps
constexpr int NT_CONV = 384;              // fprop / dgrad kernels: 8 epilogue warps (two per TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;
constexpr int MAX_DYN_SMEM = 226 * 1024;         // 227 KB per CTA minus the 1 KB of static barriers / slots

struct TcConvP {
  View y;                     // output view
  const float* bias;
  int act;
  int C[2];                   // channels of the two sources (multiples of 32; C[1] may be 0)
  int Ctot, Cout;
  int kh, kw, stride, pad, dil;
  int stride_w;               // horizontal stride (= stride except for the stem's row-band view)
  int Ho, Wo, N;
  int BW, BH, tiles_w, tiles_h, tiles_n;   // pixel tile = BW x BH (=128), tiles_n = Cout / BN
  long long total_tiles;
  double* stats;              // nullable: [0,C) += sum y, [C,2C) += sum y^2 (BatchNorm statistics of the output)
};

// NACC = 1: one M = 128 accumulator per tile (128 pixels).  NACC = 2: a tile is 256 pixels = two M = 128 accumulators that
// share every weight tile (one (32 ch, BW, 2*BH) activation box per stage): operand bytes per MMA clock drop from
// (16 + BN/8) KB / 256 clk to (32 + BN/8) KB / 512 clk, i.e. by 1.33x at BN = 128 — the generic kernel is bound by its
// L2 -> shared-memory operand stream (profiles/r1_generic_kernel.md), so that is its speed-up on layers with enough tiles.
template <int BN, int NACC>
__host__ __device__ constexpr int conv_stages() {
  constexpr int stage = NACC * A_BYTES + BN * 128;
  constexpr int fit = (MAX_DYN_SMEM - 1024 - EPI_BYTES) / stage;    // the optional statistics buffer is checked at launch
  return fit > 5 ? 5 : fit;
}

template <int BN, int NACC>
__global__ void __launch_bounds__(NT_CONV, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const TcConvP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = NACC * A_BYTES + B_BYTES;
  constexpr int STAGES = conv_stages<BN, NACC>();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_slot;
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);      // [EPI_WARPS][32 * EPI_LD]
  float* csum = epi_stage + EPI_WARPS * 32 * EPI_LD;                             // [2][Cout] when p.stats
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) csum[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB);
    if (p.C[1]) tma_prefetch_desc(&tmA1);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(2 * NACC * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int kchunks = p.Ctot / 32;
  const int kiters = p.kh * p.kw * kchunks;
  const int TH = NACC * p.BH;               // output rows per tile

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (splitting the activation / weight copies over two issuing warps was measured and does not help here: this
    // kernel is bound by the latency of its operand stream - 160 KB in flight, a stage completing every ~800 clocks,
    // profiles/r1_generic_kernel.md - not by the issue rate of the bulk copies, unlike wgrad)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tn = (int)(t % p.tiles_n); long long q = t / p.tiles_n;
        const int tw = (int)(q % p.tiles_w); q /= p.tiles_w;
        const int th = (int)(q % p.tiles_h); const int n = (int)(q / p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * TH;
        for (int tap = 0; tap < p.kh * p.kw; ++tap) {
          const int r = tap / p.kw, s = tap % p.kw;
          const int wi = w0 * p.stride_w - p.pad + s * p.dil, hi = h0 * p.stride - p.pad + r * p.dil;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + NACC * A_BYTES;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, STAGE_BYTES);
            const int c = kc * 32;
            if (c < p.C[0]) tma_load_4d(sa, &tmA0, fb, c, wi, hi, n);
            else tma_load_4d(sa, &tmA1, fb, c - p.C[0], wi, hi, n);
            tma_load_2d(sb, &tmB, fb, tap * p.Ctot + c, tn * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty0 + 8 * as, aphase ^ 1);
      tc_fence_after();
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + NACC * A_BYTES;
#pragma unroll
          for (int j = 0; j < NACC; ++j) {
            const uint32_t d_tmem = tmem_base + (as * NACC + j) * BN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
              const uint64_t ad = make_desc(sa + j * A_BYTES + 32 * k, 16, 1024), bd = make_desc(sb + 32 * k, 16, 1024);
              tc_mma_tf32(d_tmem, ad, bd, idesc, (it | k) != 0);
            }
          }
          tc_commit(empty0 + 8 * stage);            // frees the smem slot when these MMAs retire
          if (it == kiters - 1) tc_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;       // the two warps of a quarter take alternate 32-column chunks
    float* stage = epi_stage + (warp - 4) * 32 * EPI_LD;
    int as = 0; uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int tn = (int)(t % p.tiles_n); long long qq = t / p.tiles_n;
      const int tw = (int)(qq % p.tiles_w); qq /= p.tiles_w;
      const int th = (int)(qq % p.tiles_h); const int n = (int)(qq / p.tiles_h);
      // NACC == 2: the two warps of a lane quarter take one accumulator (= 128-pixel half of the tile) each;
      // NACC == 1: they take alternate 32-column chunks of the single accumulator
      const int jacc = NACC == 2 ? half : 0;
      auto row_ptr = [&](int r) -> float* {
        const int m = jacc * 128 + q * 32 + r;
        const int h = th * TH + m / p.BW, w = tw * p.BW + m % p.BW;
        return (h < p.Ho && w < p.Wo) ? p.y.p + p.y.off(n, h, w) + tn * BN : nullptr;
      };
      float* dst8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dst8[i] = row_ptr((lane >> 3) + 4 * i);
      const unsigned valid = __ballot_sync(0xffffffffu, row_ptr(lane) != nullptr);
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      if (NACC == 2)
        epilogue_chunks<BN / 32>(tmem_base + ((uint32_t)(q * 32) << 16) + (as * 2 + jacc) * BN, 0, 1, BN / 32,
                                 p.bias ? p.bias + tn * BN : nullptr, p.act, stage, lane, dst8, valid,
                                 p.stats ? csum : nullptr, tn * BN, p.Cout);
      else
        epilogue_chunks<(BN / 32 + 1) / 2>(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN, half, 2, BN / 32,
                                           p.bias ? p.bias + tn * BN : nullptr, p.act, stage, lane, dst8, valid,
                                           p.stats ? csum : nullptr, tn * BN, p.Cout);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) atomicAdd(p.stats + i, (double)csum[i]);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * NACC * BN));
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 / stride 1 fprop (and dgrad) with row-halo reuse, for layers at least 128 pixels wide.
// One CTA tile = 2 output rows x 128 pixels x BN channels (two M=128 accumulators sharing every weight tile).
// A stage is one (32-channel chunk, tap row r): two input rows are fetched as (128 + 2*dil)-pixel TMA boxes and
// the three horizontal taps are the SAME shared-memory rows addressed through descriptors whose start is shifted
// by s*dil*128 B (the 128B swizzle is a function of absolute smem address bits, base offset 0), so an input row is read from L2 3x per tile instead of 9x,
// and a weight tile 4.5x instead of 9x.  (ncu, round 1: the one-box-per-tap kernel moves 21.7 GB through L2 for a
// 3.2 GB layer and is L2-bandwidth bound.)
// ------------------------------------------------------------------------------------------------
struct TcRowP {
  View y;
  const float* bias;
  int act;
  int C[2], Ctot, Cout;
  int pad, dil;
  int Ho, Wo, N;
  int tiles_w, tiles_h, tiles_n;     // tiles_h = ceil(Ho / 2)
  long long total_tiles;
  int row_bytes;                     // (128 + 2*dil) * 128 rounded up to 1024
  int base_off_mode;                 // 1: descriptor base_offset = (addr >> 7) & 7 ; 0: always 0
  double* stats;
};

template <int BN, int NSTAGE>
__global__ void __launch_bounds__(NT_CONV, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const TcRowP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_BYTES = BN * 128;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 2 * p.row_bytes + 3 * B_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 4];
  __shared__ uint32_t tmem_base_slot;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)NSTAGE * stage_bytes);   // [EPI_WARPS][32 * EPI_LD]
  float* csum = epi_stage + EPI_WARPS * 32 * EPI_LD;                                   // [2][Cout] when p.stats
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) csum[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NSTAGE]);
  const uint32_t tfull0 = smem_u32(&bars[2 * NSTAGE]), tempty0 = smem_u32(&bars[2 * NSTAGE + 2]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB);
    if (p.C[1]) tma_prefetch_desc(&tmA1);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int kchunks = p.Ctot / 32;
  const int nstg = kchunks * 3;            // stages per tile

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_bytes = 2u * (uint32_t)((128 + 2 * p.dil) * 128) + 3u * B_BYTES;
      for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tn = (int)(t % p.tiles_n); long long q = t / p.tiles_n;
        const int tw = (int)(q % p.tiles_w); q /= p.tiles_w;
        const int th = (int)(q % p.tiles_h); const int n = (int)(q / p.tiles_h);
        const int wi = tw * 128 - p.pad, h0 = th * 2;
        for (int kc = 0; kc < kchunks; ++kc) {
          const int c = kc * 32;
          const bool s0 = c < p.C[0];
          for (int r = 0; r < 3; ++r) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + 2 * p.row_bytes;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, tx_bytes);
            const int hi = h0 - p.pad + r * p.dil;
            tma_load_4d(sa, s0 ? &tmA0 : &tmA1, fb, s0 ? c : c - p.C[0], wi, hi, n);
            tma_load_4d(sa + p.row_bytes, s0 ? &tmA0 : &tmA1, fb, s0 ? c : c - p.C[0], wi, hi + 1, n);
#pragma unroll
            for (int s = 0; s < 3; ++s) tma_load_2d(sb + s * B_BYTES, &tmB, fb, (r * 3 + s) * p.Ctot + c, tn * BN);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(BN, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty0 + 8 * as, aphase ^ 1);
      tc_fence_after();
      for (int it = 0; it < nstg; ++it) {
        mbar_wait(full0 + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + 2 * p.row_bytes;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t d_tmem = tmem_base + (as * 2 + j) * BN;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t arow = sa + j * p.row_bytes + s * p.dil * 128;
              const uint32_t bo = p.base_off_mode ? ((arow >> 7) & 7u) : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_desc(arow + 32 * k, 16, 1024, 2, bo);
                const uint64_t bd = make_desc(sb + s * B_BYTES + 32 * k, 16, 1024);
                tc_mma_tf32(d_tmem, ad, bd, idesc, (it | s | k) != 0);
              }
            }
          }
          tc_commit(empty0 + 8 * stage);
          if (it == nstg - 1) tc_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                 // TMEM lane quarter
    const int j = (warp - 4) >> 2;          // output row of the tile this warp writes
    float* stage = epi_stage + (warp - 4) * 32 * EPI_LD;
    int as = 0; uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int tn = (int)(t % p.tiles_n); long long qq = t / p.tiles_n;
      const int tw = (int)(qq % p.tiles_w); qq /= p.tiles_w;
      const int th = (int)(qq % p.tiles_h); const int n = (int)(qq / p.tiles_h);
      const int h = th * 2 + j;
      auto row_ptr = [&](int r) -> float* {
        const int w = tw * 128 + q * 32 + r;
        return (h < p.Ho && w < p.Wo) ? p.y.p + p.y.off(n, h, w) + tn * BN : nullptr;
      };
      float* dst8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dst8[i] = row_ptr((lane >> 3) + 4 * i);
      const unsigned valid = __ballot_sync(0xffffffffu, row_ptr(lane) != nullptr);
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      epilogue_chunks<BN / 32>(tmem_base + ((uint32_t)(q * 32) << 16) + (as * 2 + j) * BN, 0, 1, BN / 32,
                               p.bias ? p.bias + tn * BN : nullptr, p.act, stage, lane, dst8, valid,
                               p.stats ? csum : nullptr, tn * BN, p.Cout);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) for (int i = threadIdx.x; i < 2 * p.Cout; i += blockDim.x) atomicAdd(p.stats + i, (double)csum[i]);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * BN));
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
struct TcWgradP {
  float* dw;                  // [Cout][Ktot]
  int C[2], Ctot, Cout, Ktot;
  int kh, kw, pad, dil;
  int stride_h, stride_w;     // strides of the convolution (the x boxes use TMA element strides horizontally)
  int xmerge, chb, rhb;       // xmerge: the x operand comes through the 5-D map in boxes of chb chunks x rhb rows
                              // (= chb*rhb consecutive 32-row M groups per bulk copy); 0: one 4-D copy per group
  int Ho, Wo, N;
  int wchunks;                // Wo / 32
  long long chunks;           // N * Ho * wchunks
  int mtiles, ntiles, splits;
  long long chunks_per_split;
};

template <int BN>
__global__ void __launch_bounds__(NT, 2)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                const __grid_constant__ CUtensorMap tmX5, const __grid_constant__ CUtensorMap tmDy, const TcWgradP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2 * WSTAGES + 1];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[WSTAGES]), tfull = smem_u32(&bars[2 * WSTAGES]);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WSTAGES; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(p.xmerge ? &tmX5 : &tmX0); tma_prefetch_desc(&tmDy);
    if (p.C[1]) tma_prefetch_desc(&tmX1);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(BN < 32 ? 32 : BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // work item: (mt, nt, split)
  // M-tile fastest, pixel range slowest: CTAs sharing a pixel range are co-scheduled (x / dY served from L2)
  int b = blockIdx.x;
  const int mt = b % p.mtiles; b /= p.mtiles;
  const int nt = b % p.ntiles; const int split = b / p.ntiles;
  const long long c_beg = (long long)split * p.chunks_per_split;
  const long long c_end = min(p.chunks, c_beg + p.chunks_per_split);
  const int niter = (int)max(0LL, c_end - c_beg);
  // the 4 (tap, channel-chunk) groups that make up this CTA's 128 GEMM-M rows
  const int ngroups = min(4, (p.Ktot - mt * 128 + 31) / 32);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int g_src[4], g_c[4], g_dw[4], g_dh[4];
      for (int g = 0; g < 4; ++g) {
        const int kf = (mt * 4 + g) * 32;
        const int kk = kf < p.Ktot ? kf : 0;
        const int tap = kk / p.Ctot, coff = kk % p.Ctot;
        g_src[g] = coff < p.C[0] ? 0 : 1;
        g_c[g] = g_src[g] ? coff - p.C[0] : coff;
        g_dh[g] = -p.pad + (tap / p.kw) * p.dil;
        g_dw[g] = -p.pad + (tap % p.kw) * p.dil;
      }
      for (long long ch = c_beg; ch < c_end; ++ch) {
        const int wc = (int)(ch % p.wchunks); long long q = ch / p.wchunks;
        const int h = (int)(q % p.Ho); const int n = (int)(q / p.Ho);
        const int w0 = wc * 32;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
        const uint32_t fb = full0 + 8 * stage;
        if (p.xmerge) {
          const int per = p.chb * p.rhb;                      // M groups per bulk copy (a full box always lands)
          const int nload = (ngroups + per - 1) / per;
          mbar_expect_tx(fb, (nload * per + BN / 32) * 4096);
          for (int l = 0; l < nload; ++l) {
            const int g = l * per;
            tma_load_5d(sa + g * 4096, &tmX5, fb, 0, w0 * p.stride_w + g_dw[g], g_c[g] >> 5, h * p.stride_h + g_dh[g], n);
          }
        } else {
          mbar_expect_tx(fb, (ngroups + BN / 32) * 4096);
          for (int g = 0; g < ngroups; ++g)
            tma_load_4d(sa + g * 4096, g_src[g] ? &tmX1 : &tmX0, fb, g_c[g], w0 * p.stride_w + g_dw[g], h * p.stride_h + g_dh[g], n);
        }
        tma_load_5d(sb, &tmDy, fb, 0, w0, nt * (BN / 32), h, n);     // all BN/32 output-channel chunks in one copy
        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(BN, 1, 1);
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < niter; ++it) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // 8 pixels per MMA = two 4-row (512 B) BASE32B swizzle atoms along K
          const uint64_t ad = make_desc(sa + 1024 * k, 4096, 512, 1), bd = make_desc(sb + 1024 * k, 4096, 512, 1);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (it | k) != 0);
        }
        tc_commit(empty0 + 8 * stage);
        if (it == niter - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4 && niter > 0) {
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int kf = mt * 128 + m;
    mbar_wait(tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, v);
      if (kf < p.Ktot) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = nt * BN + cc * 32 + j;
          atomicAdd(p.dw + (long long)co * p.Ktot + kf, v[j]);   // lanes = consecutive kf -> coalesced RED
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN < 32 ? 32 : BN));
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad of 3x3 / stride 1 convolutions with halo reuse: one CTA owns a tap row r, a group of <= 128 input channels
// and an output-channel tile; per 32-pixel step it fetches ONE (32 + 2*dil)-pixel x box per 32-channel chunk and
// derives the three horizontal taps from it by shifting the descriptor start by s*dil pixel rows (MN-major,
// SWIZZLE_128B_BASE32B), into three accumulators that share the dY tile.  L2 bytes per MAC drop ~2.8x vs the
// one-box-per-tap kernel above (ncu round 1: 22 GB of L2 traffic, 5-9 GB of DRAM re-reads per launch).
// ------------------------------------------------------------------------------------------------
struct TcWg3P {
  float* dw;
  int C[2], Ctot, Cout, Ktot;
  int pad, dil;
  int Ho, Wo, N;
  int wchunks; long long chunks;
  int groups0, groups;          // channel groups (<=128 ch) of source 0 / both sources
  int units;                    // 3: one CTA per tap row.  2 (every group <= 64 channels): tap rows {0,1} share a
                                // CTA — its four M groups are (row, 32-channel chunk) pairs — and row 2 has its own
  int ntiles, splits; long long chunks_per_split;
  int xbox;                     // bytes reserved per 32-channel x box in a stage (1 KB multiple)
  int merge;                    // 1: 5-D maps - one bulk copy brings all (tap row, chunk) groups of a stage, packed
                                // at (32 + 2*dil) * 128 bytes per group; 0: one 4-D copy per group at xbox strides
};

template <int BN, int NSTAGE, int MINB>
__global__ void __launch_bounds__(NT, MINB)
tc_wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                   const __grid_constant__ CUtensorMap tmX0p, const __grid_constant__ CUtensorMap tmX1p,
                   const __grid_constant__ CUtensorMap tmDy, const TcWg3P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int B_BYTES = BN * 128;
  constexpr int TCOLS = (3 * BN <= 256) ? 256 : 512;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = 4 * p.xbox + B_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 1];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NSTAGE]), tfull = smem_u32(&bars[2 * NSTAGE]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmX0); tma_prefetch_desc(&tmDy);
    if (p.C[1]) tma_prefetch_desc(&tmX1);
    if (p.units == 2) tma_prefetch_desc(&tmX0p);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // tap row fastest, then channel group / output tile, pixel range slowest: the CTAs that stream the same pixels
  // (x and dY) are co-scheduled, so those tensors come from L2 instead of being re-read from DRAM per tap row
  int b = blockIdx.x;
  const int unit = b % p.units; b /= p.units;
  const bool paired = p.units == 2 && unit == 0;                   // this CTA accumulates tap rows 0 and 1
  const int r = p.units == 2 ? unit * 2 : unit;                    // first tap row of this CTA
  const int grp = b % p.groups; b /= p.groups;
  const int nt = b % p.ntiles; const int split = b / p.ntiles;
  const int src = grp < p.groups0 ? 0 : 1;
  const int cbase = (src ? grp - p.groups0 : grp) * 128;           // within the source
  const int nch = min(4, (p.C[src] - cbase) / 32);                 // valid 32-channel chunks in this group
  const int ng = paired ? 2 * nch : nch;                           // M groups of 32 rows: (tap row, chunk) pairs
  const int cabs = (src ? p.C[0] : 0) + cbase;                     // channel offset in the concatenated K index
  const long long c_beg = (long long)split * p.chunks_per_split;
  const long long c_end = min(p.chunks, c_beg + p.chunks_per_split);
  const int niter = (int)max(0LL, c_end - c_beg);
  // merged copies pack the M groups as [tap row][chunk] at the exact box pitch; the per-group form keeps xbox strides
  const uint32_t bx = (uint32_t)((32 + 2 * p.dil) * 128);
  const uint32_t gstride = p.merge ? bx : (uint32_t)p.xbox;
  const int nbox_ch = min(4, p.C[src] / 32);                       // chunks per merged box of this source's map

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (p.merge ? (uint32_t)(nbox_ch * (paired ? 2 : 1)) : (uint32_t)ng) * bx + (BN / 32) * 4096u;
      for (long long ch = c_beg; ch < c_end; ++ch) {
        const int wc = (int)(ch % p.wchunks); long long q = ch / p.wchunks;
        const int h = (int)(q % p.Ho); const int n = (int)(q / p.Ho);
        const int w0 = wc * 32;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sb = sa + 4 * p.xbox;
        const uint32_t fb = full0 + 8 * stage;
        mbar_expect_tx(fb, tx);
        if (p.merge) {
          const CUtensorMap* xm = paired ? (src ? &tmX1p : &tmX0p) : (src ? &tmX1 : &tmX0);
          tma_load_5d(sa, xm, fb, 0, w0 - p.pad, cbase >> 5, h - p.pad + r * p.dil, n);
        } else {
          for (int g = 0; g < ng; ++g) {
            const int chunk = paired ? g % nch : g, rr = paired ? r + g / nch : r;
            tma_load_4d(sa + g * p.xbox, src ? &tmX1 : &tmX0, fb, cbase + chunk * 32, w0 - p.pad, h - p.pad + rr * p.dil, n);
          }
        }
        tma_load_5d(sb, &tmDy, fb, 0, w0, nt * (BN / 32), h, n);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(BN, 1, 1);
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < niter; ++it) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes), sb = sa + 4 * p.xbox;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc(sa + s * p.dil * 128 + 1024 * k, gstride, 512, 1);
            const uint64_t bd = make_desc(sb + 1024 * k, 4096, 512, 1);
            tc_mma_tf32(tmem_base + s * BN, ad, bd, idesc, (it | k) != 0);
          }
        }
        tc_commit(empty0 + 8 * stage);
        if (it == niter - 1) tc_commit(tfull);
      }
      __syncwarp();
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4 && niter > 0) {
    const int q = warp - 4;                // = M group of this warp's 32 accumulator rows
    const int m = q * 32 + lane;
    const int rr = paired ? r + q / nch : r;
    const int ch = paired ? (q % nch) * 32 + lane : m;
    mbar_wait(tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
      const int kf = (rr * 3 + s) * p.Ctot + cabs + ch;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + s * BN + cc * 32, v);
        if (m < ng * 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(p.dw + (long long)(nt * BN + cc * 32 + j) * p.Ktot + kf, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_encode_state = 0;   // 0 unknown, 1 ok, -1 unavailable

static bool tc_init() {
  if (g_encode_state == 0) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) { g_encode = (EncodeTiledFn)fn; g_encode_state = 1; }
    else { cudaGetLastError(); g_encode_state = -1; }
  }
  return g_encode_state == 1;
}

// NHWC activation view -> 4-D map (C, W, H, N), box (32, bw, bh, 1), optional element stride on W/H
static bool make_act_map(CUtensorMap* m, const View& v, int bw, int bh, int estride,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int estride_w = 0) {
  if (!estride_w) estride_w = estride;
  if ((reinterpret_cast<uintptr_t>(v.p) & 15) || (v.sw % 4) || (v.sh % 4) || (v.sn % 4) || (v.c % 4)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 4, (cuuint64_t)v.sh * 4, (cuuint64_t)v.sn * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(bw * estride_w), (cuuint32_t)(bh * estride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)estride_w, (cuuint32_t)estride, 1};
  if (box[1] > 256 || box[2] > 256) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
// NHWC activation view -> 5-D map (32 ch, W, C/32, H, N), box (32, bw, nchi, bh, 1): ONE bulk copy lands nchi
// 32-channel chunks x bh image rows as consecutive [row][chunk][pixel][32 ch] groups of bw*128 bytes — the MN-major
// operand layout of the wgrad kernels, which used to take one copy per chunk (the single producer thread was the
// bottleneck: 6-8 bulk copies per 12 MMAs).
static bool make_act_map5(CUtensorMap* m, const View& v, int bw, int bh, int nchi, CUtensorMapSwizzle swz, int estride_w = 1) {
  if ((reinterpret_cast<uintptr_t>(v.p) & 15) || (v.sw % 4) || (v.sh % 4) || (v.sn % 4) || (v.c % 32)) return false;
  cuuint64_t dims[5] = {32, (cuuint64_t)v.w, (cuuint64_t)(v.c / 32), (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[4] = {(cuuint64_t)v.sw * 4, 128, (cuuint64_t)v.sh * 4, (cuuint64_t)v.sn * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)(bw * estride_w), (cuuint32_t)nchi, (cuuint32_t)bh, 1};
  cuuint32_t es[5] = {1, (cuuint32_t)estride_w, 1, 1, 1};
  if (box[1] > 256 || box[2] > 256 || box[3] > 256) return false;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
static bool make_w_map(CUtensorMap* m, const float* w, int ktot, int cout, int bn) {
  if ((reinterpret_cast<uintptr_t>(w) & 15) || (ktot % 4)) return false;
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// which tensor-core kernel the calling thread launched last (SEGSDE_TC_KERNEL_*): lets tests assert the route
static thread_local int g_last_tc_kernel = 0;

static int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

template <int BN, int NACC>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcConvP& p, cudaStream_t st) {
  const int smem = conv_stages<BN, NACC>() * (NACC * A_BYTES + BN * 128) + 1024 + EPI_BYTES + (p.stats ? 2 * p.Cout * 4 : 0);
  static int attr = 0;
  if (smem > MAX_DYN_SMEM) return SEGSDE_E_UNSUPPORTED;
  if (smem > attr) {
    if (cudaFuncSetAttribute(tc_conv_kernel<BN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return SEGSDE_E_UNSUPPORTED;
    }
    attr = smem;
  }
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  tc_conv_kernel<BN, NACC><<<(int)grid, NT_CONV, smem, st>>>(a0, a1, b, p);
  g_last_tc_kernel = NACC == 2 ? SEGSDE_TC_KERNEL_CONV256 : SEGSDE_TC_KERNEL_CONV;
  return launched();
}
template <int BN>
static int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& x5, const CUtensorMap& dy,
                        const TcWgradP& p, cudaStream_t st) {
  constexpr int smem = WSTAGES * (A_BYTES + BN * 128) + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  tc_wgrad_kernel<BN><<<p.mtiles * p.ntiles * p.splits, NT, smem, st>>>(x0, x1, x5, dy, p);
  g_last_tc_kernel = SEGSDE_TC_KERNEL_WGRAD;
  return launched();
}

template <int BN, int NSTAGE>
static int launch_conv3x3(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcRowP& p, cudaStream_t st) {
  const int smem = NSTAGE * (2 * p.row_bytes + 3 * BN * 128) + 1024 + EPI_BYTES + (p.stats ? 2 * p.Cout * 4 : 0);
  static int attr = 0;
  if (smem > MAX_DYN_SMEM) return SEGSDE_E_UNSUPPORTED;
  if (smem > attr) {
    if (cudaFuncSetAttribute(tc_conv3x3_kernel<BN, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return SEGSDE_E_UNSUPPORTED;
    }
    attr = smem;
  }
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  tc_conv3x3_kernel<BN, NSTAGE><<<(int)grid, NT_CONV, smem, st>>>(a0, a1, b, p);
  g_last_tc_kernel = SEGSDE_TC_KERNEL_ROWHALO;
  return launched();
}

// SEGSDE_TC_ROWHALO: 0 = off, 2 (default) = on, descriptor base_offset 0, 1 = on with base_offset = (addr>>7)&7.
// Measured on B200: the swizzle XOR is taken from the absolute shared-memory address bits, so a start address
// shifted by whole 128-byte rows needs base_offset 0 (mode 1 produces wrong results; kept for the record).
static int rowhalo_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("SEGSDE_TC_ROWHALO"); m = e ? atoi(e) : 2; }
  return m;
}

template <int BN, int NSTAGE, int MINB>
static int launch_wgrad3x3(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& x0p, const CUtensorMap& x1p,
                           const CUtensorMap& dy, const TcWg3P& p, cudaStream_t st) {
  const int smem = NSTAGE * (4 * p.xbox + BN * 128) + 1024;
  static int attr = 0;
  if (smem > 200 * 1024) return SEGSDE_E_UNSUPPORTED;
  if (smem > attr) {
    if (cudaFuncSetAttribute(tc_wgrad3x3_kernel<BN, NSTAGE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return SEGSDE_E_UNSUPPORTED;
    }
    attr = smem;
  }
  tc_wgrad3x3_kernel<BN, NSTAGE, MINB><<<p.groups * p.units * p.ntiles * p.splits, NT, smem, st>>>(x0, x1, x0p, x1p, dy, p);
  g_last_tc_kernel = SEGSDE_TC_KERNEL_WGRAD3X3;
  return launched();
}
// SEGSDE_TC_M256: 0 = never use 256-pixel tiles in the generic kernel, 1 (default) = when every SM keeps >= 2 tiles,
// 2 = always (tests)
static int m256_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("SEGSDE_TC_M256"); m = e ? atoi(e) : 1; }
  return m;
}
static int wg3_mode() {     // SEGSDE_TC_WGRAD3: 0 = off, 1 (default) = halo-reuse wgrad for 3x3 / stride 1
  static int m = -1;
  if (m < 0) { const char* e = getenv("SEGSDE_TC_WGRAD3"); m = e ? atoi(e) : 1; }
  return m;
}
// SEGSDE_TC_WG64 (experiments on the <= 64-channel wgrad): bit 0 = pair tap rows 0/1 in one CTA (M = 2 x 64),
// bit 1 = 3 stages with two CTAs per SM instead of 4 stages with one.  Default 3 (64->64 @512x1024, B=12:
// 2.52 ms with neither, 1.82 pairing only, 1.55 two CTAs only, 1.15 ms with both).
static int wgmerge_mode() {   // SEGSDE_TC_WGMERGE: 0 = one bulk copy per 32-channel chunk (round-1 form), 1 (default) = merged
  static int m = -1;
  if (m < 0) { const char* e = getenv("SEGSDE_TC_WGMERGE"); m = e ? atoi(e) : 1; }
  return m;
}
static int wg64_mode() {
  static int m = -1;
  if (m < 0) { const char* e = getenv("SEGSDE_TC_WG64"); m = e ? atoi(e) : 3; }
  return m;
}

static int pow2_floor(int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; }

}  // namespace segsde
using namespace segsde;

extern "C" int segsde_tc_available(void) { return tc_init() ? 1 : 0; }
extern "C" int segsde_tc_last_kernel(void) { return g_last_tc_kernel; }

// y = act(conv(cat(x1, x2), w) + bias), zero padding, stride 1 or 2, any dilation; channel counts multiples
// of 32 (inputs) / 64 (outputs).  Anything else -> SEGSDE_E_UNSUPPORTED (caller uses the generic path).
extern "C" int segsde_conv2d_fwd_tc_stats(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                                          const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                                          double* stats, void* stream);
extern "C" int segsde_conv2d_fwd_tc(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                                    const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                                    void* stream) {
  return segsde_conv2d_fwd_tc_stats(x1, x2, w, bias, y, d, nullptr, stream);
}
// Same, and additionally accumulates the per-channel sum / sum of squares of the (pre-activation) output into
// stats[0..C) / stats[C..2C) (fp64, zero-filled by the caller): BatchNorm batch statistics without a second pass.
extern "C" int segsde_conv2d_fwd_tc_stats(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                                          const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                                          double* stats, void* stream) {
  if (!x1 || !x1->ptr || !w || !y || !y->ptr || !d) return SEGSDE_E_ARG;
  if (stats && (y->c > MAX_STAT_C || d->act != SEGSDE_ACT_NONE)) return SEGSDE_E_UNSUPPORTED;
  if (!tc_init()) return SEGSDE_E_UNSUPPORTED;
  if (d->pad_mode != SEGSDE_PAD_ZERO || d->up1 || d->nchw_norm_in) return SEGSDE_E_UNSUPPORTED;
  if (d->stride != 1 && d->stride != 2) return SEGSDE_E_UNSUPPORTED;   // stride 2 = TMA element strides
  View v1 = mk(x1), v2 = mk(x2), vy = mk(y);
  const int C1 = v1.c, C2 = v2.p ? v2.c : 0, Cout = vy.c;
  if (C1 % 32 || C2 % 32 || Cout % 32) return SEGSDE_E_UNSUPPORTED;
  if (v2.p && (v2.h != v1.h || v2.w != v1.w || v2.n != v1.n)) return SEGSDE_E_ARG;
  const int stride_w = d->stride_w ? d->stride_w : d->stride;
  if (stride_w != 1 && stride_w != 2) return SEGSDE_E_UNSUPPORTED;
  int Ho = (v1.h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  int Wo = (v1.w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / stride_w + 1;
  // An output that is up to dil*(k-1) rows / columns LARGER than the formula asks for implicit extra zero padding at
  // the bottom / right (stride 1 only): the taps that fall outside the input are the TMA out-of-bounds fill.  Used by
  // the phase decomposition of the stride-2 dgrad (2-tap phases with offsets {0, +1}).
  if (vy.n != v1.n || vy.h < Ho || vy.w < Wo) return SEGSDE_E_ARG;
  if (vy.h != Ho || vy.w != Wo) {
    if (d->stride != 1 || stride_w != 1 || vy.h > Ho + d->dil * (d->kh - 1) || vy.w > Wo + d->dil * (d->kw - 1)) return SEGSDE_E_ARG;
    Ho = vy.h; Wo = vy.w;
  }
  if (!vec4_ok(vy)) return SEGSDE_E_UNSUPPORTED;
  const int BN = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
  TcConvP p;
  p.y = vy; p.bias = bias; p.act = d->act; p.stats = stats;
  p.C[0] = C1; p.C[1] = C2; p.Ctot = C1 + C2; p.Cout = Cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil; p.stride_w = stride_w;
  p.Ho = Ho; p.Wo = Wo; p.N = v1.n;
  p.BW = Wo >= 128 ? 128 : pow2_floor(Wo < 1 ? 1 : Wo);
  if (p.BW < 8) return SEGSDE_E_UNSUPPORTED;
  p.BH = 128 / p.BW;
  p.tiles_w = cdiv(Wo, p.BW); p.tiles_h = cdiv(Ho, p.BH); p.tiles_n = Cout / BN;
  p.total_tiles = (long long)p.tiles_w * p.tiles_h * p.N * p.tiles_n;
  CUtensorMap a0, a1, b;
  if (rowhalo_mode() && d->kh == 3 && d->kw == 3 && d->stride == 1 && Wo >= 128 && BN >= 64 && 128 + 2 * d->dil <= 256) {
    TcRowP r;
    r.y = vy; r.bias = bias; r.act = d->act; r.C[0] = C1; r.C[1] = C2; r.Ctot = C1 + C2; r.Cout = Cout;
    r.pad = d->pad; r.dil = d->dil; r.Ho = Ho; r.Wo = Wo; r.N = v1.n;
    r.tiles_w = cdiv(Wo, 128); r.tiles_h = cdiv(Ho, 2); r.tiles_n = Cout / BN;
    r.total_tiles = (long long)r.tiles_w * r.tiles_h * r.N * r.tiles_n;
    r.row_bytes = (((128 + 2 * d->dil) * 128) + 1023) / 1024 * 1024;
    r.base_off_mode = rowhalo_mode() == 1;
    r.stats = stats;
    if (make_act_map(&a0, v1, 128 + 2 * d->dil, 1, 1) && (!C2 || make_act_map(&a1, v2, 128 + 2 * d->dil, 1, 1)) &&
        make_w_map(&b, w, 9 * r.Ctot, Cout, BN)) {
      if (!C2) a1 = a0;
      return BN == 128 ? launch_conv3x3<128, 2>(a0, a1, b, r, as_stream(stream))
                       : launch_conv3x3<64, 3>(a0, a1, b, r, as_stream(stream));
    }
  }
  // 256-pixel tiles (two accumulators sharing the weight tile) when that still leaves every SM at least two tiles
  {
    const long long tiles2 = (long long)p.tiles_w * cdiv(Ho, 2 * p.BH) * p.N * p.tiles_n;
    const int mode = m256_mode();
    if (mode && (mode == 2 || tiles2 >= 2LL * num_sms()) && 2 * p.BH * d->stride <= 256) {
      TcConvP p2 = p;
      p2.tiles_h = cdiv(Ho, 2 * p.BH);
      p2.total_tiles = tiles2;
      if (make_act_map(&a0, v1, p.BW, 2 * p.BH, d->stride, CU_TENSOR_MAP_SWIZZLE_128B, stride_w) &&
          (!C2 || make_act_map(&a1, v2, p.BW, 2 * p.BH, d->stride, CU_TENSOR_MAP_SWIZZLE_128B, stride_w)) &&
          make_w_map(&b, w, d->kh * d->kw * p.Ctot, Cout, BN)) {
        if (!C2) a1 = a0;
        int rc = BN == 128 ? launch_conv<128, 2>(a0, a1, b, p2, as_stream(stream))
                           : (BN == 64 ? launch_conv<64, 2>(a0, a1, b, p2, as_stream(stream))
                                       : launch_conv<32, 2>(a0, a1, b, p2, as_stream(stream)));
        if (rc != SEGSDE_E_UNSUPPORTED) return rc;
      }
    }
  }
  if (!make_act_map(&a0, v1, p.BW, p.BH, d->stride, CU_TENSOR_MAP_SWIZZLE_128B, stride_w)) return SEGSDE_E_UNSUPPORTED;
  if (C2) { if (!make_act_map(&a1, v2, p.BW, p.BH, d->stride, CU_TENSOR_MAP_SWIZZLE_128B, stride_w)) return SEGSDE_E_UNSUPPORTED; } else a1 = a0;
  if (!make_w_map(&b, w, d->kh * d->kw * p.Ctot, Cout, BN)) return SEGSDE_E_UNSUPPORTED;
  if (BN == 128) return launch_conv<128, 1>(a0, a1, b, p, as_stream(stream));
  if (BN == 64) return launch_conv<64, 1>(a0, a1, b, p, as_stream(stream));
  return launch_conv<32, 1>(a0, a1, b, p, as_stream(stream));
}

// dgrad on the tensor cores = fprop of dy with the transposed/tap-flipped weights; the Python layer prepares
// those weights (segsde_weight_transpose_flip) and calls segsde_conv2d_fwd_tc, so this entry only reports
// that the direct form is not offered.
extern "C" int segsde_conv2d_dgrad_tc(const segsde_nhwc_t*, const float*, const segsde_nhwc_t*, const segsde_nhwc_t*,
                                      const segsde_conv_desc_t*, void*) {
  return SEGSDE_E_UNSUPPORTED;
}

// Split-K factor of the weight-gradient kernels.  Every split CTA pays a prologue and a full-tile epilogue of fp32
// atomics (a 3 x 128 x 128 tile = 196 KB of REDs, about what ten 32-pixel k-steps load), and CTAs run in waves of
// (SMs x resident CTAs per SM): minimise  waves(s) * (chunks / s + EPI)  over s.  (Round-1 rule: 2 x SMs / ctas - two
// waves of half-length CTAs for the one-CTA-per-SM kernel, i.e. twice the atomics for the same SM time.)
static int pick_splits(long long ctas, long long chunks, long long maxs, int ctas_per_sm) {
  const long long slots = (long long)num_sms() * ctas_per_sm;
  const double EPI = 10.0;
  long long best = 1; double best_cost = 1e30;
  const long long smax = maxs < 4 * slots ? maxs : 4 * slots;
  for (long long s = 1; s <= smax; ++s) {
    const long long waves = (ctas * s + slots - 1) / slots;
    const double per = (double)((chunks + s - 1) / s) + EPI;
    const double cost = (double)waves * per;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return (int)best;
}

extern "C" int segsde_conv2d_wgrad_tc(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const segsde_nhwc_t* dy,
                                      float* dw, float* dbias, const segsde_conv_desc_t* d, void* stream) {
  if (!x1 || !x1->ptr || !dy || !dy->ptr || !dw || !d) return SEGSDE_E_ARG;
  if (!tc_init()) return SEGSDE_E_UNSUPPORTED;
  if (dbias) return SEGSDE_E_UNSUPPORTED;      // bias gradient is produced by segsde_act_bwd_bias
  if (d->pad_mode != SEGSDE_PAD_ZERO || d->up1 || d->nchw_norm_in) return SEGSDE_E_UNSUPPORTED;
  // stride 2 (either direction): the generic kernel strides its x boxes (TMA element strides horizontally, the
  // row coordinate vertically); dy stays dense
  if (d->stride != 1 && d->stride != 2) return SEGSDE_E_UNSUPPORTED;
  const int stride_w = d->stride_w ? d->stride_w : d->stride;
  if (stride_w != 1 && stride_w != 2) return SEGSDE_E_UNSUPPORTED;
  View v1 = mk(x1), v2 = mk(x2), vd = mk(dy);
  const int C1 = v1.c, C2 = v2.p ? v2.c : 0, Cout = vd.c;
  if (C1 % 32 || C2 % 32 || Cout % 32) return SEGSDE_E_UNSUPPORTED;
  const int Ho = (v1.h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  const int Wo = (v1.w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / stride_w + 1;
  if (vd.h != Ho || vd.w != Wo || vd.n != v1.n) return SEGSDE_E_ARG;
  if (Wo % 32) return SEGSDE_E_UNSUPPORTED;
  const int BN = (Cout % 128 == 0) ? 128 : (Cout % 64 == 0 ? 64 : 32);
  if (wg3_mode() && d->kh == 3 && d->kw == 3 && d->stride == 1 && stride_w == 1 && BN >= 64 && 32 + 2 * d->dil <= 256) {
    TcWg3P q;
    q.dw = dw; q.C[0] = C1; q.C[1] = C2; q.Ctot = C1 + C2; q.Cout = Cout; q.Ktot = 9 * q.Ctot;
    q.pad = d->pad; q.dil = d->dil; q.Ho = Ho; q.Wo = Wo; q.N = v1.n; q.wchunks = Wo / 32;
    q.chunks = (long long)q.N * Ho * q.wchunks;
    q.groups0 = cdiv(C1, 128); q.groups = q.groups0 + (C2 ? cdiv(C2, 128) : 0);
    q.ntiles = Cout / BN;
    q.xbox = (((32 + 2 * d->dil) * 128) + 1023) / 1024 * 1024;
    q.units = ((wg64_mode() & 1) && C1 <= 64 && C2 <= 64 && d->dil == 1) ? 2 : 3;
    long long ctas = (long long)q.groups * q.units * q.ntiles;
    long long maxs = q.chunks / 16; if (maxs < 1) maxs = 1;
    q.splits = pick_splits(ctas, q.chunks, maxs, (BN == 128 || !(wg64_mode() & 2)) ? 1 : 2);
    q.chunks_per_split = (q.chunks + q.splits - 1) / q.splits;
    q.splits = (int)((q.chunks + q.chunks_per_split - 1) / q.chunks_per_split);
    CUtensorMap x0m, x1m, x0p, x1p, dym;
    const CUtensorMapSwizzle swz3 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const int bw = 32 + 2 * d->dil;
    const int ch1 = C1 / 32 < 4 ? C1 / 32 : 4, ch2 = C2 ? (C2 / 32 < 4 ? C2 / 32 : 4) : 0;
    bool ok;
    q.merge = wgmerge_mode();
    if (q.merge) {
      ok = make_act_map5(&x0m, v1, bw, 1, ch1, swz3) && (!C2 || make_act_map5(&x1m, v2, bw, 1, ch2, swz3));
      if (ok && q.units == 2) ok = make_act_map5(&x0p, v1, bw, 2, ch1, swz3) && (!C2 || make_act_map5(&x1p, v2, bw, 2, ch2, swz3));
    } else {
      ok = make_act_map(&x0m, v1, bw, 1, 1, swz3) && (!C2 || make_act_map(&x1m, v2, bw, 1, 1, swz3));
    }
    if (ok && make_act_map5(&dym, vd, 32, 1, BN / 32, swz3)) {
      if (!C2) x1m = x0m;
      if (!(q.merge && q.units == 2)) { x0p = x0m; x1p = x1m; } else if (!C2) x1p = x0p;
      cudaStream_t st3 = as_stream(stream);
      int rc;
      if (BN == 128) {     // TMEM (3 x 128 columns) allows one CTA per SM: as many stages as shared memory holds
        rc = launch_wgrad3x3<128, 5, 1>(x0m, x1m, x0p, x1p, dym, q, st3);
        if (rc == SEGSDE_E_UNSUPPORTED) rc = launch_wgrad3x3<128, 3, 1>(x0m, x1m, x0p, x1p, dym, q, st3);
      } else {
        rc = (wg64_mode() & 2) ? launch_wgrad3x3<64, 3, 2>(x0m, x1m, x0p, x1p, dym, q, st3)
                               : launch_wgrad3x3<64, 4, 1>(x0m, x1m, x0p, x1p, dym, q, st3);
      }
      if (rc != SEGSDE_E_UNSUPPORTED) return rc;
    }
  }
  TcWgradP p;
  p.dw = dw; p.C[0] = C1; p.C[1] = C2; p.Ctot = C1 + C2; p.Cout = Cout; p.Ktot = d->kh * d->kw * p.Ctot;
  p.kh = d->kh; p.kw = d->kw; p.pad = d->pad; p.dil = d->dil; p.stride_h = d->stride; p.stride_w = stride_w;
  p.Ho = Ho; p.Wo = Wo; p.N = v1.n; p.wchunks = Wo / 32;
  p.chunks = (long long)p.N * Ho * p.wchunks;
  p.mtiles = cdiv(p.Ktot, 128); p.ntiles = Cout / BN;
  long long maxs = p.chunks / 16; if (maxs < 1) maxs = 1;
  p.splits = pick_splits((long long)p.mtiles * p.ntiles, p.chunks, maxs, 2);
  p.chunks_per_split = (p.chunks + p.splits - 1) / p.splits;
  p.splits = (int)((p.chunks + p.chunks_per_split - 1) / p.chunks_per_split);
  CUtensorMap x0m, x1m, x5m, dym;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (!make_act_map(&x0m, v1, 32, 1, 1, swz, stride_w)) return SEGSDE_E_UNSUPPORTED;
  if (C2) { if (!make_act_map(&x1m, v2, 32, 1, 1, swz, stride_w)) return SEGSDE_E_UNSUPPORTED; } else x1m = x0m;
  // merged x copies (single source): chb consecutive chunks of one tap, and - when the taps of an M tile are
  // vertically consecutive rows (kw == 1, dil == 1: the stem band) - rhb tap rows per bulk copy
  p.xmerge = 0; p.chb = p.rhb = 1;
  x5m = x0m;
  if (!C2 && wgmerge_mode()) {
    const int nch = C1 / 32;
    p.chb = nch % 4 == 0 ? 4 : (nch % 2 == 0 ? 2 : 1);
    p.rhb = (nch == p.chb && d->kw == 1 && d->dil == 1) ? 4 / p.chb : 1;
    if (p.chb * p.rhb > 1 && make_act_map5(&x5m, v1, 32, p.rhb, p.chb, swz, stride_w)) p.xmerge = 1;
    else { p.chb = p.rhb = 1; x5m = x0m; }
  }
  if (!make_act_map5(&dym, vd, 32, 1, BN / 32, swz)) return SEGSDE_E_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (BN == 128) return launch_wgrad<128>(x0m, x1m, x5m, dym, p, st);
  if (BN == 64) return launch_wgrad<64>(x0m, x1m, x5m, dym, p, st);
  return launch_wgrad<32>(x0m, x1m, x5m, dym, p, st);
}
