// The GEMM engine of the OCR forward path: D = A * B^T for K-major bf16 operands on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA into 128B-swizzled shared memory,
// persistent warp-specialised CTAs (one per SM):
//
//     warp 0          : TMA producer       (one elected lane)
//     warp 1          : TMEM allocator + tcgen05.mma issuer (one elected lane)
//     warps 2..9      : epilogue           (TMEM -> registers -> smem transpose -> fused bias/act/residual/
//                                           scatter -> fully coalesced global stores)
//
// Precision: the reference computes in fp32.  bf16 tensor-core operands alone would miss the
// "logits within 1e-3" bar, so every fp32 matrix is carried as a (hi, lo) bf16 pair with hi + lo == x to
// ~2^-17 and each K-step issues three MMAs (hi*hi + lo*hi + hi*lo) into the same fp32 TMEM accumulator
// (NSPLIT = 3).  NSPLIT = 1 is the plain single-pass bf16 mode.
//
// Every linear layer of Swin-B / the decoder / ViT, the FPN 1x1 convolutions, the decoder cross-attention
// (Q K^T and P V as batched GEMMs) and the vocabulary heads go through this one kernel.
#include <algorithm>

#include "alm_internal.h"
#include "ptx.cuh"

namespace alm {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 352;  // warp 0 TMA(A), warp 1 MMA, warps 2..9 epilogue, warp 10 TMA(B)
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 16;

struct GemmParams {
  int M, N, K;
  int nb0, nb1;          // batch extents of the problem
  int a_b0, a_b1;        // 1 if operand A varies along that batch dim (else coordinate 0 = broadcast)
  int b_b0, b_b1;
  int m_blocks, n_blocks, k_blocks;
  long num_tiles;
  Epilogue e;
  int vec_ok;  // rows are 16-byte addressable -> vector stores
  int stage_tx;               // bytes one pipeline stage receives (the A box shrinks to the live rows when M < 128)
  int a_bytes;                // shared-memory footprint of one A operand tile (live rows only)
  int stage_bytes;            // one stage: kNumA A tiles + kNumB B tiles
  int stages;                 // pipeline depth: as many stages as fit (<= 16) -- deep for the small-M decode GEMMs
  unsigned long long* trace;  // optional [cap][6] records: start ns, end ns (CTA 0), M, N, K, tiles
  int* trace_idx;
  int trace_cap;
  unsigned long long* detail;  // optional [64 tiles][6] per-role timestamps of CTA 0 (debug)
};

template <int BLOCK_N, int NSPLIT>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kNumA = (NSPLIT == 3) ? 2 : 1;
  static constexpr int kNumB = (NSPLIT == 3) ? 2 : 1;
  static constexpr int kStageBytes = kNumA * kABytes + kNumB * kBBytes;
  static constexpr int kBarBytes = 1024;
  static constexpr int kStagingBytes = kEpiWarps * 4096;  // one XOR-swizzled 32x32 fp32 block per epilogue warp
  static constexpr int kBudget = 227 * 1024 - 1024 - kBarBytes - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kTotal = kBudget + kBarBytes + kStagingBytes + 1024;  // +1024 alignment slack
};

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), exact-erf form (nn.GELU default).  The fc1 epilogue is instruction-issue
// bound (537 M activations per 16-page batch on 8 epilogue warps per SM), so erf is evaluated with the
// Abramowitz-Stegun 7.1.26 rational form (|error| <= 1.5e-7 absolute, i.e. at the level of fp32 erff itself and two
// orders below the path's 1e-5 accuracy) on the MUFU rcp / ex2 units: ~14 instructions instead of ~35.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // MUFU.RCP, <= 1 ulp
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));  // MUFU.EX2, <= 2 ulp
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// (x, y) -> packed bf16x2 hi and lo words with one cvt.rn.bf16x2.f32 each (hi + lo == x to ~2^-17)
__device__ __forceinline__ void split_pack_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));  // upper half <- first source
  const float rx = x - __uint_as_float(hi << 16);
  const float ry = y - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ry), "f"(rx));
}

// The same GELU on a PAIR of values with the packed fp32x2 pipe (FFMA2 / FMUL2): the polynomial, the products and the
// final blend are one instruction per pair; only the two MUFU ops and the sign transfer stay per lane.
//   erf(|x|/sqrt2) = 1 - p(t) t e,  t = 1 / (1 + 0.3275911 |x|/sqrt2),  e = 2^(-x^2 log2(e) / 2)
__device__ __forceinline__ f2 gelu_erf2(f2 x) {
  float x0, x1;
  f2_get(x, x0, x1);
  const f2 ax = f2_make(fabsf(x0), fabsf(x1));
  const f2 den = f2_fma(f2_splat(0.3275911f * 0.70710678118654752440f), ax, f2_splat(1.0f));
  const f2 arg = f2_mul(f2_mul(x, x), f2_splat(-0.5f * 1.4426950408889634f));
  float d0, d1, a0, a1, t0, t1, e0, e1;
  f2_get(den, d0, d1);
  f2_get(arg, a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f2 t = f2_make(t0, t1), e = f2_make(e0, e1);
  // q = -p(t): the Abramowitz-Stegun 7.1.26 coefficients with flipped signs, so that erf = 1 + q t e is one fma
  f2 q = f2_fma(t, f2_splat(-1.061405429f), f2_splat(1.453152027f));
  q = f2_fma(t, q, f2_splat(-1.421413741f));
  q = f2_fma(t, q, f2_splat(0.284496736f));
  q = f2_fma(t, q, f2_splat(-0.254829592f));
  const f2 erf_abs = f2_fma(f2_mul(q, t), e, f2_splat(1.0f));
  float r0, r1;
  f2_get(erf_abs, r0, r1);
  const f2 serf = f2_make(copysignf(r0, x0), copysignf(r1, x1));
  const f2 h = f2_mul(x, f2_splat(0.5f));
  return f2_fma(h, serf, h);   // 0.5 x (1 + erf)
}

// Epilogue inner loop for one thread: 8 rows x 4 consecutive columns, everything already in registers.
// Compile-time activation / output kinds keep this a short branch-free instruction stream (the epilogue warps
// have one or two warps per scheduler, so instruction count is what bounds small-K GEMMs); the arithmetic runs on
// fp32 pairs (alpha / bias fma, GELU, residual add, the hi/lo split) -- half the FMA-pipe instructions.
template <int ACT, int F32, int SPLIT, int ROWBIAS>
__device__ __forceinline__ void epi_store8(const float4 (&acc)[8], const int (&orow)[8], const float (&rbias)[8],
                                           const float4 (&res)[8], const float4& bb, float alpha, long ocol, long ldo,
                                           float* __restrict__ out_f32, bf16* __restrict__ out_hi,
                                           bf16* __restrict__ out_lo) {
  const f2 al = f2_splat(alpha);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (orow[i] < 0) continue;
    const float rb = ROWBIAS ? rbias[i] : 0.f;
    f2 a = f2_fma(f2_make(acc[i].x, acc[i].y), al, ROWBIAS ? f2_make(rb + bb.x, rb + bb.y) : f2_make(bb.x, bb.y));
    f2 b = f2_fma(f2_make(acc[i].z, acc[i].w), al, ROWBIAS ? f2_make(rb + bb.z, rb + bb.w) : f2_make(bb.z, bb.w));
    if (ACT == ACT_GELU) { a = gelu_erf2(a); b = gelu_erf2(b); }
    if (ACT == ACT_RELU) {
      float u0, u1, u2, u3;
      f2_get(a, u0, u1); f2_get(b, u2, u3);
      a = f2_make(fmaxf(u0, 0.f), fmaxf(u1, 0.f)); b = f2_make(fmaxf(u2, 0.f), fmaxf(u3, 0.f));
    }
    a = f2_add(a, f2_make(res[i].x, res[i].y));
    b = f2_add(b, f2_make(res[i].z, res[i].w));
    float f0, f1, f2_, f3;
    f2_get(a, f0, f1); f2_get(b, f2_, f3);
    const long o = ocol + static_cast<long>(orow[i]) * ldo;
    if (F32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(f0, f1, f2_, f3);
    if (SPLIT) {
      uint32_t h01, l01, h23, l23;
      split_pack_bf16x2(f0, f1, h01, l01);
      split_pack_bf16x2(f2_, f3, h23, l23);
      *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(h01, h23);
      if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(l01, l23);
    }
  }
}

// PLAIN = 1: the epilogue of a map-free launch (no output / residual row maps, no per-row bias, vector-aligned
// outputs): the row bookkeeping collapses to "row < M" and none of the predicated map / row-bias loads of the generic
// epilogue are even issued (ncu: those option paths were ~1/3 of the executed instructions of a GELU epilogue).
template <int BLOCK_N, int NSPLIT, int PLAIN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                    const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
                    const GemmParams p) {
  using L = SmemLayout<BLOCK_N, NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBudget);  // operand stages live in [0, kBudget)
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int trace_slot = -1;
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    trace_slot = atomicAdd(p.trace_idx, 1);
    if (trace_slot < p.trace_cap) {
      unsigned long long* r = p.trace + 6 * static_cast<long>(trace_slot);
      r[0] = ptx::globaltimer_ns(); r[2] = p.M; r[3] = p.N; r[4] = p.K; r[5] = p.num_tiles * 1000 + BLOCK_N;
    } else trace_slot = -1;
  }
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // double-buffered accumulator (power of two >= 32)
  static_assert(BLOCK_N == 32 || BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_b_hi);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);  // one arrive.expect_tx per producer warp (A tiles, B tiles)
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], kEpiWarps);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long tiles_per_batch = static_cast<long>(p.m_blocks) * p.n_blocks;

  if (warp == 0 || warp == 10) {
    // ===================================================================== TMA producers
    // Two single-thread producers in different warps (A tiles / B tiles): a lone thread issues one bulk-tensor
    // copy every ~0.25 us, which bounds both the small-tile decode GEMMs and the big ones; two issue in parallel.
    {
      // the whole warp runs the (warp-uniform) loop and one elected lane issues: operands of UTMALDG / UTCHMMA live
      // in uniform registers, and a `lane == 0` branch would make ptxas wrap every issue in an ELECT/R2UR loop
      const bool is_a = (warp == 0);
      const int my_bytes = is_a ? L::kNumA * p.a_bytes : L::kNumB * L::kBBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int batch = static_cast<int>(tile / tiles_per_batch);
        const int rem = static_cast<int>(tile % tiles_per_batch);
        const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
        const int b0 = batch % p.nb0, b1 = batch / p.nb0;
        const int c2 = is_a ? (p.a_b0 ? b0 : 0) : (p.b_b0 ? b0 : 0);
        const int c3 = is_a ? (p.a_b1 ? b1 : 0) : (p.b_b1 ? b1 : 0);
        const int rowc = is_a ? m_blk * BLOCK_M : n_blk * BLOCK_N;
        const void* tm = is_a ? static_cast<const void*>(&tm_a_hi) : static_cast<const void*>(&tm_b_hi);
        const long tl = (tile - blockIdx.x) / gridDim.x;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.detail && is_a && blockIdx.x == 0 && tl < 64 && kb == 0 && lane == 0) p.detail[tl * 6 + 0] = ptx::globaltimer_ns();
          uint8_t* dst = stage_base + stage * p.stage_bytes + (is_a ? 0 : L::kNumA * p.a_bytes);
          if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[stage], my_bytes);
          if (NSPLIT == 3) {
            // hi and lo planes of an operand sit at a fixed distance in HBM: ONE 5-D box {k, rows, plane = 2}
            // fetches both tiles back to back
            ptx::tma_load_5d(dst, tm, &full_bar[stage], kb * BLOCK_K, rowc, 0, c2, c3);
          } else {
            ptx::tma_load_4d(dst, tm, &full_bar[stage], kb * BLOCK_K, rowc, c2, c3);
          }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) |
                                 (uint32_t(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const long tl = (tile - blockIdx.x) / gridDim.x;
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
        ptx::tc_fence_after();
        if (p.detail && blockIdx.x == 0 && tl < 64 && lane == 0) p.detail[tl * 6 + 1] = ptx::globaltimer_ns();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (p.detail && blockIdx.x == 0 && tl < 64 && kb == 0 && lane == 0) p.detail[tl * 6 + 2] = ptx::globaltimer_ns();
          const uint32_t a_hi = ptx::smem_u32(stage_base + stage * p.stage_bytes);
          const uint32_t a_lo = a_hi + p.a_bytes;
          const uint32_t b_hi = a_hi + L::kNumA * p.a_bytes;
          const uint32_t b_lo = b_hi + L::kBBytes;
          if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;  // bytes inside the 128 B swizzle row
            const uint64_t da = ptx::make_kmajor_sw128_desc(a_hi + koff);
            const uint64_t db = ptx::make_kmajor_sw128_desc(b_hi + koff);
            ptx::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            if (NSPLIT == 3) {
              const uint64_t dal = ptx::make_kmajor_sw128_desc(a_lo + koff);
              const uint64_t dbl = ptx::make_kmajor_sw128_desc(b_lo + koff);
              ptx::umma_bf16(d_tmem, dal, db, idesc, 1);
              ptx::umma_bf16(d_tmem, da, dbl, idesc, 1);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        __syncwarp();
        if (p.detail && blockIdx.x == 0 && tl < 64 && lane == 0) p.detail[tl * 6 + 3] = ptx::globaltimer_ns();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (PLAIN == 2) {
    // ===================================================================== epilogue warps (8), TMA-store form
    // Launches whose only output is an unbatched split-bf16 matrix with column bias / activation (the qkv and fc1 GEMMs:
    // 7/8 of a transformer block's output elements).  The accumulator block stays ROW-PER-THREAD as tcgen05.ld delivers it
    // (thread = row, 32 consecutive columns): bias / GELU / the hi-lo split run on fp32 pairs in registers, the bf16 rows
    // go to a 64-byte-swizzled staging block with conflict-free 16-byte stores, and ONE lane hands the 32 x 32 block to
    // the TMA unit (cp.async.bulk.tensor store, one per plane), which clips the M tail.  No fp32 transpose through shared
    // memory, no per-thread global addressing or store instructions: the epilogue was the bound of every small-K launch
    // (ncu: ~30 thread-instructions per output element of a GELU + split epilogue).
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) (hardware: warp_id % 4)
    const int half = (warp - 2) >> 2;      // 32-column chunks c = half, half + 2, ... of the tile
    const uint32_t stg = ptx::smem_u32(smem + L::kBudget + L::kBarBytes) + (warp - 2) * 4096;  // hi block, lo block at +2048
    const uint32_t row_off = lane * 64, sw = (lane >> 1) & 3;   // 64B swizzle: 16-byte chunk k of row r sits at k ^ ((r >> 1) & 3)
    const Epilogue& e = p.e;
    const f2 al = f2_splat(e.alpha);
    int as = 0;
    uint32_t aphase = 0;
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int rem = static_cast<int>(tile % tiles_per_batch);
      const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const int row_base = m_blk * BLOCK_M + quarter * 32;
      const bool rows_live = row_base < p.M;
      // The warp's BLOCK_N / 64 chunks are software-pipelined: the tcgen05.ld of chunk i + 1 is in flight while chunk i is
      // computed (with two epilogue warps per scheduler the TMEM read latency was exposed once per chunk).
      constexpr int NCH = BLOCK_N / 64;
      const uint32_t t_acc = tmem_base + (uint32_t(quarter * 32) << 16) + as * BLOCK_N + half * 32;
      uint32_t v[2][32];
      ptx::tmem_ld_32x32(t_acc, v[0]);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int col0 = n_blk * BLOCK_N + (half + 2 * ci) * 32;
        const bool live = rows_live && col0 < p.N;   // warp-uniform; N % 32 == 0 for these launches
        float4 bb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          bb[q] = (e.bias && live) ? __ldg(reinterpret_cast<const float4*>(e.bias + col0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        ptx::tmem_ld_wait();
        if (ci + 1 < NCH) ptx::tmem_ld_32x32(t_acc + (ci + 1) * 64, v[(ci + 1) & 1]);
        if (live) {
          const uint32_t (&vc)[32] = v[ci & 1];
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            f2 a = f2_fma(f2_make(__uint_as_float(vc[4 * q]), __uint_as_float(vc[4 * q + 1])), al, f2_make(bb[q].x, bb[q].y));
            f2 b = f2_fma(f2_make(__uint_as_float(vc[4 * q + 2]), __uint_as_float(vc[4 * q + 3])), al, f2_make(bb[q].z, bb[q].w));
            if (e.act == ACT_GELU) { a = gelu_erf2(a); b = gelu_erf2(b); }
            float f0, f1, f2_, f3;
            f2_get(a, f0, f1); f2_get(b, f2_, f3);
            if (e.act == ACT_RELU) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2_ = fmaxf(f2_, 0.f); f3 = fmaxf(f3, 0.f); }
            split_pack_bf16x2(f0, f1, ph[2 * q], pl[2 * q]);
            split_pack_bf16x2(f2_, f3, ph[2 * q + 1], pl[2 * q + 1]);
          }
          // the previous block of this warp must have left the staging area before it is rewritten
          if (lane == 0) ptx::tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t dst = stg + row_off + ((k ^ sw) << 4);
            ptx::sts_v4(dst, ph[4 * k], ph[4 * k + 1], ph[4 * k + 2], ph[4 * k + 3]);
            if (e.out_lo) ptx::sts_v4(dst + 2048, pl[4 * k], pl[4 * k + 1], pl[4 * k + 2], pl[4 * k + 3]);
          }
          ptx::fence_proxy_async();   // generic-proxy writes -> visible to the async proxy (TMA)
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tm_o_hi, stg, col0, row_base);
            if (e.out_lo) ptx::tma_store_2d(&tm_o_lo, stg + 2048, col0, row_base);
            ptx::tma_store_commit();
          }
        }
      }
      // release the accumulator stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) ptx::tma_store_wait_all();   // the staging blocks must outlive the last stores
  } else {
    // ===================================================================== epilogue warps (8)
    // warp -> (TMEM lane quarter, column half).  Each 32x32 accumulator block is read row-per-thread from TMEM,
    // transposed through an XOR-swizzled 4 KB staging block in shared memory, and then handled four columns per
    // thread / eight threads per row, so that every global access (bias, residual, outputs) is a full 128 B line.
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) (hardware: warp_id % 4)
    const int half = (warp - 2) >> 2;      // 32-column chunks c = half, half + 2, ... of the tile
    float* stg = reinterpret_cast<float*>(smem + L::kBudget + L::kBarBytes) + (warp - 2) * 1024;
    const Epilogue& e = p.e;
    const int rl_base = lane >> 3, j4 = lane & 7;
    int as = 0;
    uint32_t aphase = 0;
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int batch = static_cast<int>(tile / tiles_per_batch);
      const int rem = static_cast<int>(tile % tiles_per_batch);
      const int m_blk = rem / p.n_blocks, n_blk = rem % p.n_blocks;
      const int b0 = batch % p.nb0, b1 = batch / p.nb0;
      const long tl = (tile - blockIdx.x) / gridDim.x;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      if (p.detail && blockIdx.x == 0 && tl < 64 && warp == 4 && lane == 0) p.detail[tl * 6 + 4] = ptx::globaltimer_ns();
      const long obatch = static_cast<long>(b0) * e.obs0 + static_cast<long>(b1) * e.obs1;
      const float* rbatch = e.resid ? e.resid + static_cast<long>(b0) * e.rbs0 + static_cast<long>(b1) * e.rbs1 : nullptr;
      const float* bias = e.bias ? e.bias + static_cast<long>(b0) * e.bias_bs0 : nullptr;
      const int row_base = m_blk * BLOCK_M + quarter * 32;
      const bool rows_live = row_base < p.M;

#pragma unroll 1
      for (int c = half; c < BLOCK_N / 32; c += 2) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + as * BLOCK_N + c * 32, v);
        ptx::tmem_ld_wait();
        const int col0 = n_blk * BLOCK_N + c * 32;
        if (!rows_live || col0 >= p.N) continue;  // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int col = col0 + j4 * 4;
        const bool col_ok = col < p.N;
        const bool vec = p.vec_ok && col + 3 < p.N;
        // phase 1: everything this thread needs for its 8 rows is requested up front (independent loads in
        // flight: staging reads, row maps, row biases, residual lines) -- the epilogue is latency-bound otherwise.
        float4 acc[8];
        int orow[8], rrow[8];
        float rbias[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + rl_base;
          acc[i] = *reinterpret_cast<const float4*>(stg + rl * 32 + ((j4 ^ (rl & 7)) << 2));
          const int row = row_base + rl;
          int o = (row < p.M && col_ok) ? row : -1;
          if (!PLAIN && o >= 0 && e.out_map) o = e.out_map[row];
          orow[i] = o;
          rrow[i] = (!PLAIN && o >= 0 && e.resid_map) ? e.resid_map[row] : o;
          rbias[i] = (!PLAIN && o >= 0 && bias && e.bias_mode == BIAS_ROW) ? bias[row] : 0.0f;
        }
        if (PLAIN || vec) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && e.bias_mode == BIAS_COL) bb = *reinterpret_cast<const float4*>(bias + col);
          float4 res[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            res[i] = (rbatch && rrow[i] >= 0)
                         ? *reinterpret_cast<const float4*>(rbatch + static_cast<long>(rrow[i]) * e.ldr + col)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          // phase 2: straight-line math + stores; the option checks are hoisted into one switch per chunk
          const int kind = e.act * 4 + (e.out_f32 ? 1 : 0) + (e.out_hi ? 2 : 0);
          const long ocol = obatch + (e.col_group ? static_cast<long>(col / e.col_group) * e.col_group_stride + col % e.col_group : col);
          switch (kind) {
#define ALM_EPI_CASE(ACTV, F32V, SPLV)                                                                         \
  case (ACTV) * 4 + (F32V) + 2 * (SPLV):                                                                        \
    epi_store8<ACTV, F32V, SPLV, !PLAIN>(acc, orow, rbias, res, bb, e.alpha, ocol, e.ldo, e.out_f32, e.out_hi, e.out_lo); \
    break;
            ALM_EPI_CASE(0, 1, 0) ALM_EPI_CASE(0, 0, 1) ALM_EPI_CASE(0, 1, 1)
            ALM_EPI_CASE(1, 1, 0) ALM_EPI_CASE(1, 0, 1) ALM_EPI_CASE(1, 1, 1)
            ALM_EPI_CASE(2, 1, 0) ALM_EPI_CASE(2, 0, 1) ALM_EPI_CASE(2, 1, 1)
#undef ALM_EPI_CASE
            default: break;
          }
        } else if (!PLAIN) {
          // ragged / unaligned tail: scalar, guarded (fully unrolled so the row arrays stay in registers)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (orow[i] < 0) continue;
            const long o = obatch + static_cast<long>(orow[i]) * e.ldo +
                           (e.col_group ? static_cast<long>(col / e.col_group) * e.col_group_stride + col % e.col_group : col);
            const float f[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (col + q >= p.N) break;
              float x = f[q] * e.alpha + rbias[i];
              if (bias && e.bias_mode == BIAS_COL) x += bias[col + q];
              if (e.act == ACT_GELU) x = gelu_erf(x);
              else if (e.act == ACT_RELU) x = fmaxf(x, 0.0f);
              if (rbatch) x += rbatch[static_cast<long>(rrow[i]) * e.ldr + col + q];
              if (e.out_f32) e.out_f32[o + q] = x;
              if (e.out_hi) {
                bf16 h, l;
                split_bf16(x, h, l);
                e.out_hi[o + q] = h;
                if (e.out_lo) e.out_lo[o + q] = l;
              }
            }
          }
        }
        __syncwarp();  // staging block is reused by the next chunk
      }
      // release the accumulator stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
      if (p.detail && blockIdx.x == 0 && tl < 64 && warp == 4 && lane == 0) p.detail[tl * 6 + 5] = ptx::globaltimer_ns();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  if (trace_slot >= 0) p.trace[6 * static_cast<long>(trace_slot) + 1] = ptx::globaltimer_ns();
}

// ------------------------------------------------------------------------------------------------
// SIMT debug kernel: same operands (hi + lo re-joined to fp32), same epilogue, plain FMA.  Used to bisect
// the tensor-core path on the GPU box (alm_set_option "gemm_impl" 1); never the default.
// ------------------------------------------------------------------------------------------------
struct SimtOperand {
  const bf16* hi; const bf16* lo; long ld, bs0, bs1;
};

__global__ void gemm_simt_kernel(SimtOperand A, SimtOperand B, GemmParams p, int nsplit) {
  __shared__ float sa[16][17];
  __shared__ float sb[16][17];
  const int batch = blockIdx.z;
  const int b0 = batch % p.nb0, b1 = batch / p.nb0;
  const long aoff = (p.a_b0 ? b0 : 0) * A.bs0 + (p.a_b1 ? b1 : 0) * A.bs1;
  const long boff = (p.b_b0 ? b0 : 0) * B.bs0 + (p.b_b1 ? b1 : 0) * B.bs1;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    const int ka = k0 + tx;
    float va = 0.f, vb = 0.f;
    if (row < p.M && ka < p.K) {
      va = __bfloat162float(A.hi[aoff + row * A.ld + ka]);
      if (nsplit == 3) va += __bfloat162float(A.lo[aoff + row * A.ld + ka]);
    }
    const int brow = blockIdx.x * 16 + ty;
    if (brow < p.N && ka < p.K) {
      vb = __bfloat162float(B.hi[boff + brow * B.ld + ka]);
      if (nsplit == 3) vb += __bfloat162float(B.lo[boff + brow * B.ld + ka]);
    }
    sa[ty][tx] = va;
    sb[ty][tx] = vb;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (row >= p.M || col >= p.N) return;
  const Epilogue& e = p.e;
  long orow = row;
  if (e.out_map) orow = e.out_map[row];
  if (orow < 0) return;
  long rrow = orow;
  if (e.resid_map) rrow = e.resid_map[row];
  float x = acc * e.alpha;
  if (e.bias) {
    const float* bias = e.bias + static_cast<long>(b0) * e.bias_bs0;
    x += (e.bias_mode == BIAS_ROW) ? bias[row] : bias[col];
  }
  if (e.act == ACT_GELU) x = gelu_erf(x);
  else if (e.act == ACT_RELU) x = fmaxf(x, 0.f);
  if (e.resid) x += e.resid[b0 * e.rbs0 + b1 * e.rbs1 + rrow * e.ldr + col];
  const long o = b0 * e.obs0 + b1 * e.obs1 + orow * e.ldo +
                 (e.col_group ? static_cast<long>(col / e.col_group) * e.col_group_stride + col % e.col_group : col);
  if (e.out_f32) e.out_f32[o] = x;
  if (e.out_hi) {
    bf16 h, l;
    split_bf16(x, h, l);
    e.out_hi[o] = h;
    if (e.out_lo) e.out_lo[o] = l;
  }
}

// Encoding a descriptor is a driver call (a few microseconds); the decode loop re-issues the same few hundred
// (pointer, geometry) pairs every token, so descriptors are memoised per context.
const CUtensorMap& make_tmap(Ctx* c, const bf16* base, const Operand& op, int box_rows, long plane_stride) {
  Ctx::TmapKey key{base, op.K, op.rows, op.ld, op.nb0, op.nb1, op.bs0, op.bs1, box_rows, plane_stride};
  auto it = c->tmap_cache.find(key);
  if (it != c->tmap_cache.end()) return it->second;
  if (c->tmap_cache.size() > 20000) c->tmap_cache.clear();
  CUtensorMap tm;
  ALM_REQUIRE(base != nullptr, ALM_ERR_INVALID, "gemm operand pointer is null");
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, ALM_ERR_INVALID, "gemm operand not 16-byte aligned");
  ALM_REQUIRE(op.ld % 8 == 0 && op.bs0 % 8 == 0 && op.bs1 % 8 == 0, ALM_ERR_INVALID,
              "gemm operand strides must be multiples of 8 elements");
  const cuuint64_t row_bytes = cuuint64_t(op.ld) * 2;
  cuuint64_t s1 = op.nb0 > 1 ? cuuint64_t(op.bs0) * 2 : row_bytes * cuuint64_t(op.rows);
  cuuint64_t s2 = op.nb1 > 1 ? cuuint64_t(op.bs1) * 2 : s1 * cuuint64_t(op.nb0);
  if (s1 == 0) s1 = row_bytes;  // degenerate broadcast dims still need a legal stride
  if (s2 == 0) s2 = s1;
  CUresult r;
  if (plane_stride > 0) {
    // 5-D: {k, row, plane(hi, lo), batch0, batch1}
    ALM_REQUIRE(plane_stride % 8 == 0, ALM_ERR_INVALID, "hi/lo planes must be a multiple of 16 bytes apart");
    cuuint64_t dims[5] = {cuuint64_t(op.K), cuuint64_t(op.rows), 2, cuuint64_t(op.nb0), cuuint64_t(op.nb1)};
    cuuint64_t strides[4] = {row_bytes, cuuint64_t(plane_stride) * 2, s1, s2};
    cuuint32_t box[5] = {cuuint32_t(BLOCK_K), cuuint32_t(box_rows), 2, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {cuuint64_t(op.K), cuuint64_t(op.rows), cuuint64_t(op.nb0), cuuint64_t(op.nb1)};
    cuuint64_t strides[3] = {row_bytes, s1, s2};
    cuuint32_t box[4] = {cuuint32_t(BLOCK_K), cuuint32_t(box_rows), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)) +
                                     " (K=" + std::to_string(op.K) + " rows=" + std::to_string(op.rows) +
                                     " ld=" + std::to_string(op.ld) + ")"};
  return c->tmap_cache.emplace(key, tm).first->second;
}

// 2-D map of one output plane for the TMA-store epilogue: [M rows, N columns] bf16 with leading dimension ldo, boxes of
// 32 x 32 elements (64-byte rows, 64-byte swizzle).  Memoised like the operand maps (the decode loop repeats its launches).
const CUtensorMap& make_out_tmap(Ctx* c, const bf16* base, int M, int N, long ldo) {
  Ctx::TmapKey key{base, N, M, ldo, -2, 0, 0, 0, 32, 0};
  auto it = c->tmap_cache.find(key);
  if (it != c->tmap_cache.end()) return it->second;
  if (c->tmap_cache.size() > 20000) c->tmap_cache.clear();
  CUtensorMap tm;
  cuuint64_t dims[2] = {cuuint64_t(N), cuuint64_t(M)};
  cuuint64_t strides[1] = {cuuint64_t(ldo) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled (gemm output plane) failed with CUresult " + std::to_string(int(r))};
  return c->tmap_cache.emplace(key, tm).first->second;
}

template <int BLOCK_N, int NSPLIT, int PLAIN>
void launch_tc(Ctx* c, const Operand& A, const Operand& B, GemmParams& p) {
  using L = SmemLayout<BLOCK_N, NSPLIT>;
  static DeviceOnce attr_set;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, NSPLIT, PLAIN>;
  if (attr_set.need()) {
    ALM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    pin_carveout(kern);
    attr_set.mark();
  }
  p.n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
  p.num_tiles = static_cast<long>(p.m_blocks) * p.n_blocks * p.nb0 * p.nb1;
  // rows of A beyond M are never stored, so the TMA box only covers the live rows (multiple of 8 = one swizzle
  // group); the MMA still reads 128 rows of shared memory, the stale ones only feed discarded accumulator rows.
  const int a_rows = std::min(BLOCK_M, (p.M + 7) & ~7);
  p.a_bytes = a_rows * BLOCK_K * 2;
  p.stage_bytes = L::kNumA * p.a_bytes + L::kNumB * L::kBBytes;
  p.stage_tx = p.stage_bytes;
  p.stages = std::min(kMaxStages, L::kBudget / p.stage_bytes);
  // by value: the descriptor cache may be cleared by a later call
  long pa = 0, pb = 0;
  if (NSPLIT == 3) {
    pa = A.lo - A.hi;
    pb = B.lo - B.hi;
    ALM_REQUIRE(pa > 0 && pb > 0, ALM_ERR_INVALID, "split operands must be allocated hi plane first, lo plane after");
  }
  const CUtensorMap ta_hi = make_tmap(c, A.hi, A, a_rows, pa);
  const CUtensorMap tb_hi = make_tmap(c, B.hi, B, BLOCK_N, pb);
  const CUtensorMap ta_lo = ta_hi, tb_lo = tb_hi;  // kept in the signature; the 5-D maps cover both planes
  // TMA-store epilogue: 2-D maps of the output planes (32 x 32 bf16 boxes, 64-byte swizzle); placeholders otherwise
  const CUtensorMap to_hi = PLAIN == 2 ? make_out_tmap(c, p.e.out_hi, p.M, p.N, p.e.ldo) : ta_hi;
  const CUtensorMap to_lo = (PLAIN == 2 && p.e.out_lo) ? make_out_tmap(c, p.e.out_lo, p.M, p.N, p.e.ldo) : to_hi;
  // Small launches (the per-token decode GEMMs) leave half of the SMs to the kernels of the other in-flight
  // streams / batches: every CTA of this kernel needs a whole SM (231 KB of shared memory).
  long cap = c->num_sms;
  if (c->gemm_grid_cap > 0) cap = std::min<long>(cap, c->gemm_grid_cap);
  if (c->small_grid_cap > 0 && p.num_tiles <= 2L * c->num_sms) cap = c->small_grid_cap;
  const int grid = static_cast<int>(std::min<long>(p.num_tiles, cap));
  kern<<<grid, kThreads, L::kTotal, c->stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, to_hi, to_lo, p);
}

}  // namespace

void gemm(Ctx* c, const Operand& A, const Operand& B, const Epilogue& E) {
  if (c->skipped(2)) return;
  ALM_REQUIRE(A.K == B.K, ALM_ERR_INVALID, "gemm: K mismatch");
  ALM_REQUIRE(A.rows > 0 && B.rows > 0 && A.K > 0, ALM_ERR_INVALID, "gemm: empty problem");
  ALM_REQUIRE(E.out_f32 || E.out_hi, ALM_ERR_INVALID, "gemm: no output");
  GemmParams p;
  p.M = A.rows; p.N = B.rows; p.K = A.K;
  p.nb0 = std::max(A.nb0, B.nb0);
  p.nb1 = std::max(A.nb1, B.nb1);
  ALM_REQUIRE((A.nb0 == 1 || A.nb0 == p.nb0) && (B.nb0 == 1 || B.nb0 == p.nb0) && (A.nb1 == 1 || A.nb1 == p.nb1) &&
                  (B.nb1 == 1 || B.nb1 == p.nb1), ALM_ERR_INVALID, "gemm: batch extents do not broadcast");
  ALM_REQUIRE(!(E.out_map || E.resid_map) || (p.nb0 == 1 && p.nb1 == 1), ALM_ERR_INVALID,
              "gemm: row maps require an unbatched problem");
  p.a_b0 = A.nb0 > 1; p.a_b1 = A.nb1 > 1; p.b_b0 = B.nb0 > 1; p.b_b1 = B.nb1 > 1;
  p.m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  p.e = E;
  if (c->lo_unused) p.e.out_lo = nullptr;  // single-pass bf16 ViT path: the consumers read only the hi planes
  bool vec = (E.ldo % 8 == 0) && (E.obs0 % 8 == 0) && (E.obs1 % 8 == 0);
  ALM_REQUIRE(E.col_group == 0 || (E.col_group % 4 == 0 && E.col_group_stride % 8 == 0 && !E.resid), ALM_ERR_INVALID,
              "gemm: column-group scatter needs groups of 4k columns and no residual");
  if (E.out_f32) vec = vec && (reinterpret_cast<uintptr_t>(E.out_f32) % 16 == 0);
  if (E.out_hi) vec = vec && (reinterpret_cast<uintptr_t>(E.out_hi) % 16 == 0);
  if (E.out_lo) vec = vec && (reinterpret_cast<uintptr_t>(E.out_lo) % 16 == 0);
  if (E.resid) vec = vec && (E.ldr % 4 == 0) && (E.rbs0 % 4 == 0) && (E.rbs1 % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(E.resid) % 16 == 0);
  if (E.bias && E.bias_mode == BIAS_COL) vec = vec && (reinterpret_cast<uintptr_t>(E.bias) % 16 == 0) && (E.bias_bs0 % 4 == 0);
  p.vec_ok = vec ? 1 : 0;
  p.trace = c->trace_buf; p.trace_idx = c->trace_idx; p.trace_cap = c->trace_cap;
  p.detail = c->detail_buf;
  const int nsplit = c->nsplit;
  ALM_REQUIRE(nsplit == 1 || (A.lo && B.lo), ALM_ERR_INVALID, "gemm: split mode needs lo operands");

  Ctx::GemmRec rec{nullptr, nullptr, 2.0 * p.M * p.N * p.K * p.nb0 * p.nb1};
  if (c->profile_gemm) {
    ALM_CHECK_CUDA(cudaEventCreate(&rec.a));
    ALM_CHECK_CUDA(cudaEventCreate(&rec.b));
    ALM_CHECK_CUDA(cudaEventRecord(rec.a, c->stream));
  }
  if (c->gemm_impl == 1) {
    p.n_blocks = 0; p.num_tiles = 0;
    SimtOperand sa{A.hi, A.lo, A.ld, A.bs0, A.bs1}, sb{B.hi, B.lo, B.ld, B.bs0, B.bs1};
    dim3 grid((p.N + 15) / 16, (p.M + 15) / 16, p.nb0 * p.nb1);
    gemm_simt_kernel<<<grid, dim3(16, 16), 0, c->stream>>>(sa, sb, p, nsplit);
  } else {
    // Few-tile problems (the per-token decoder GEMMs: M <= a few hundred rows) are bound by what ONE SM can pull
    // through its TMA port, so they run with 32-wide column tiles to spread the weight stream over 4x more SMs.
    const long tiles128 = static_cast<long>(p.m_blocks) * ((p.N + 127) / 128) * p.nb0 * p.nb1;
    const bool narrow = tiles128 < c->num_sms / 2 && p.N > 32;
    // Large problems are bound by what one SM can ingest through TMA (~58 B/clk): 256-wide tiles read A once per
    // 256 output columns (96 KB per K-step for twice the MMA work of a 128-wide tile's 64 KB).
    const bool wide = !narrow && c->wide_tiles && p.N >= 256 && (p.N % 256 == 0 || p.N >= 1024) &&
                      static_cast<long>(p.m_blocks) * ((p.N + 255) / 256) * p.nb0 * p.nb1 >= 2L * c->num_sms;
    // map-free, vector-aligned launches take the slim epilogue
    const bool plain = c->gemm_plain_epilogue && !E.out_map && !E.resid_map && E.bias_mode != BIAS_ROW && vec && p.N % 4 == 0;
    // ... and, when the only output is an unbatched split-bf16 matrix (qkv / fc1), the TMA-store form of it
    const bool tma_out = plain && c->gemm_plain_epilogue >= 2 && !narrow && !E.resid && !E.out_f32 && E.out_hi &&
                         p.nb0 * p.nb1 == 1 && p.N % 32 == 0 && E.col_group == 0;
#define ALM_LAUNCH(BNV, NSV)                                  \
  do {                                                        \
    if (plain) launch_tc<BNV, NSV, 1>(c, A, B, p);            \
    else launch_tc<BNV, NSV, 0>(c, A, B, p);                  \
  } while (0)
#define ALM_LAUNCH_T(BNV, NSV)                                \
  do {                                                        \
    if (tma_out) launch_tc<BNV, NSV, 2>(c, A, B, p);          \
    else ALM_LAUNCH(BNV, NSV);                                \
  } while (0)
    if (nsplit == 3) {
      if (narrow) ALM_LAUNCH(32, 3);
      else if (wide) ALM_LAUNCH_T(256, 3);
      else ALM_LAUNCH_T(128, 3);
    } else {
      if (narrow) ALM_LAUNCH(32, 1);
      else if (wide) ALM_LAUNCH_T(256, 1);
      else ALM_LAUNCH_T(128, 1);
    }
#undef ALM_LAUNCH_T
#undef ALM_LAUNCH
  }
  if (c->profile_gemm) {
    ALM_CHECK_CUDA(cudaEventRecord(rec.b, c->stream));
    c->gemm_recs.push_back(rec);
  }
  count_launch(c);
  check_launch("gemm");
}

}  // namespace alm
