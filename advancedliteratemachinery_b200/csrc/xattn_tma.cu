// Fused decoder cross-attention, TMA + mbarrier pipeline variant (alm_set_option "xattn_impl" 2) -- EXPERIMENTAL: written
// after round 1's GPU budget was spent, compiled for sm_100a, not yet run.  Same arithmetic as xattn.cu (scores, mask,
// online softmax and P.V on mma.sync with the 3-term split; partial merge by the last CTA); what changes is how K_c /
// V_c reach shared memory.  ncu on xattn.cu's 16-row (pt loop) variant: 4.0 TB/s, warps stalled on `cp.async` data --
// one 32 KB block in flight per CTA, i.e. latency-bound (Little: ~110 KB in flight per SM needed for 6.5 TB/s).  Here:
//   * one CTA per SM, persistent (same balanced run schedule);
//   * warp 0 is a producer: per 64-key block four bulk-tensor copies (K hi, K lo, V hi, V lo; 8 KB each, contiguous
//     in HBM thanks to the head-major cache) into a 4-deep ring, completion on `full` mbarriers -- up to 128 KB in
//     flight per SM and not a single load instruction in the math warps;
//   * tiles are unpadded 128-byte rows with the hardware 128-byte swizzle; ldmatrix addresses XOR the 16-byte chunk
//     index with (row & 7), conflict-free like the padded layout;
//   * the math warps release a slot by arriving on its `empty` mbarrier; there is no block-wide barrier in the loop.
//   MODE 1: 4 math warps share one 16-row query tile, 16 keys of each block each (pt loop);
//   MODE 2: 8 math warps = four 16-row tiles x two 32-key groups (poly / rec loops).
#include <algorithm>

#include "mma.cuh"
#include "omni.h"

namespace alm {
namespace {

// The softmax runs in base 2: scores are scaled by head_dim^-0.5 * log2(e) in the one multiply that applied the 1/8
// before, running maxima / partial maxima are kept in those units, and every exponential is one MUFU.EX2 (2 ulp) --
// expf cost ~8 instructions per score on a kernel whose issue slots compete with the HMMA stream.
constexpr float XA_SCALE_LOG2E = 0.125f * 1.4426950408889634f;
__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


constexpr int TX_KB = 64;            // keys per block
constexpr int TX_PLANE = 64 * 128;   // bytes of one plane tile: 64 keys x 64 bf16
constexpr int TX_STAGE = 4 * TX_PLANE;
constexpr int TX_STAGES = 4;
constexpr int TX_SMEM = TX_STAGES * TX_STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int TX_PART = 66;          // floats per partial row: m, l, o[64]

// byte offset of (row, 16-byte chunk) inside a 128-byte-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int NS, int MODE>
__global__ void __launch_bounds__(MODE == 2 ? 288 : 160, 1)
cross_attn_tma_kernel(const __grid_constant__ CUtensorMap tm_kh, const __grid_constant__ CUtensorMap tm_kl,
                      const __grid_constant__ CUtensorMap tm_vh, const __grid_constant__ CUtensorMap tm_vl,
                      const bf16* __restrict__ q_hi, const bf16* __restrict__ q_lo, const float* __restrict__ q_f32, int Ncap,
                      const uint8_t* __restrict__ kpm, int M, int nqb, int npairs, int max_parts, int z0,
                      float* __restrict__ partial, int* __restrict__ counters, bf16* __restrict__ out_hi,
                      bf16* __restrict__ out_lo, float* __restrict__ out_f32) {
  constexpr bool Q16 = MODE == 1;
  constexpr int NCW = MODE == 2 ? 8 : 4;              // math warps
  constexpr int NT = 32 * NCW;                        // math threads
  constexpr int NKG = MODE == 1 ? 4 : 2;              // key groups a 64-key block is split into
  constexpr int NJ = 8 / NKG;                         // 8-key n-tiles of a block one warp scores
  constexpr int NKK = 4 / NKG;                        // 16-key k-steps of a block one warp feeds into P.V
  extern __shared__ unsigned char tx_raw[];
  // 1024-byte aligned ring (the 128-byte swizzle pattern repeats every 8 rows = 1024 bytes)
  unsigned char* ring = tx_raw + ((1024u - (ptx::smem_u32(tx_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + TX_STAGES * TX_STAGE);
  uint64_t* empty = full + TX_STAGES;
  __shared__ int last_flag;

  const int warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (M + TX_KB - 1) / TX_KB;
  const long NB = static_cast<long>(npairs) * nkb, G = gridDim.x;
  const long b_begin = static_cast<long>(blockIdx.x) * NB / G, b_end = (static_cast<long>(blockIdx.x) + 1) * NB / G;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_kh); ptx::prefetch_tmap(&tm_vh);
    if (NS == 3) { ptx::prefetch_tmap(&tm_kl); ptx::prefetch_tmap(&tm_vl); }
    for (int s = 0; s < TX_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);      // the producer's arrive.expect_tx
      ptx::mbar_init(&empty[s], NCW);   // one arrival per math warp
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  if (warp_all == 0) {
    // ===================================================================== producer: one bulk-tensor copy per plane
    int stage = 0;
    uint32_t phase = 0;
    for (long b = b_begin; b < b_end; ++b) {
      ptx::mbar_wait(&empty[stage], phase ^ 1);
      if (ptx::elect_one()) {
        const int pair = static_cast<int>(b / nkb), kb = static_cast<int>(b - static_cast<long>(pair) * nkb);
        const int ih = pair / nqb, h = ih & 7, img = ih >> 3;
        const int z = img * 96 + z0 + h;  // (image, decoder-layer, head) slice of the [B*96][M][64] cache
        unsigned char* dst = ring + stage * TX_STAGE;
        ptx::mbar_expect_tx(&full[stage], (NS == 3 ? 4 : 2) * TX_PLANE);
        ptx::tma_load_3d(dst, &tm_kh, &full[stage], 0, kb * TX_KB, z);
        if (NS == 3) ptx::tma_load_3d(dst + TX_PLANE, &tm_kl, &full[stage], 0, kb * TX_KB, z);
        ptx::tma_load_3d(dst + 2 * TX_PLANE, &tm_vh, &full[stage], 0, kb * TX_KB, z);
        if (NS == 3) ptx::tma_load_3d(dst + 3 * TX_PLANE, &tm_vl, &full[stage], 0, kb * TX_KB, z);
      }
      __syncwarp();
      if (++stage == TX_STAGES) { stage = 0; phase ^= 1; }
    }
    return;
  }

  // ======================================================================= math warps
  const int tid = threadIdx.x - 32, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  auto consumer_sync = [&] { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); };
  auto owner = [&](long b) { return static_cast<int>(((b + 1) * G - 1) / NB); };  // CTA whose run contains block b
  int stage = 0;
  uint32_t phase = 0;
  for (long b = b_begin; b < b_end;) {
  const int pair = static_cast<int>(b / nkb), qb = pair % nqb, ih = pair / nqb, h = ih & 7, img = ih >> 3;
  const int kb0 = static_cast<int>(b - static_cast<long>(pair) * nkb);
  const int kb1 = static_cast<int>(min(static_cast<long>(nkb), kb0 + (b_end - b)));
  b += kb1 - kb0;
  const int first = owner(static_cast<long>(pair) * nkb);
  const int gs = owner(static_cast<long>(pair) * nkb + nkb - 1) - first + 1;  // CTAs sharing this pair
  const int split = static_cast<int>(blockIdx.x) - first;
  const int q0 = qb * 64;
  const int kg = Q16 ? warp : (warp >> 2);  // key group of this warp
  const int j0 = NJ * kg, kk0 = NKK * kg;
  const int r_lo = (Q16 ? 0 : (warp & 3) * 16) + g, r_hi = r_lo + 8;  // the two query rows (within the block) this lane holds
  const bool live_lo = q0 + r_lo < Ncap, live_hi = q0 + r_hi < Ncap;

  // ---- A fragments of q (hi, lo) for the four 16-dim k-steps, straight from global memory (read once per segment)
  constexpr int NP = (NS == 3) ? 2 : 1;
  uint32_t aq[NP][4][4];
  {
    const long row_lo = (static_cast<long>(img) * Ncap + q0 + r_lo) * 512 + h * 64;
    const long row_hi = (static_cast<long>(img) * Ncap + q0 + r_hi) * 512 + h * 64;
    if (q_f32) {  // fp32 queries (pt loop): split on the fly
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c0 = ks * 16 + 2 * t;
        const float2 z = make_float2(0.f, 0.f);
        const float2 v0 = live_lo ? *reinterpret_cast<const float2*>(q_f32 + row_lo + c0) : z;
        const float2 v1 = live_hi ? *reinterpret_cast<const float2*>(q_f32 + row_hi + c0) : z;
        const float2 v2 = live_lo ? *reinterpret_cast<const float2*>(q_f32 + row_lo + c0 + 8) : z;
        const float2 v3 = live_hi ? *reinterpret_cast<const float2*>(q_f32 + row_hi + c0 + 8) : z;
        uint32_t l0, l1, l2, l3;
        split_pack2(v0.x, v0.y, aq[0][ks][0], l0);
        split_pack2(v1.x, v1.y, aq[0][ks][1], l1);
        split_pack2(v2.x, v2.y, aq[0][ks][2], l2);
        split_pack2(v3.x, v3.y, aq[0][ks][3], l3);
        if (NS == 3) { aq[NP - 1][ks][0] = l0; aq[NP - 1][ks][1] = l1; aq[NP - 1][ks][2] = l2; aq[NP - 1][ks][3] = l3; }
      }
    } else {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const bf16* qp = p == 0 ? q_hi : q_lo;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int c0 = ks * 16 + 2 * t;
          aq[p][ks][0] = live_lo ? *reinterpret_cast<const uint32_t*>(qp + row_lo + c0) : 0u;
          aq[p][ks][1] = live_hi ? *reinterpret_cast<const uint32_t*>(qp + row_hi + c0) : 0u;
          aq[p][ks][2] = live_lo ? *reinterpret_cast<const uint32_t*>(qp + row_lo + c0 + 8) : 0u;
          aq[p][ks][3] = live_hi ? *reinterpret_cast<const uint32_t*>(qp + row_hi + c0 + 8) : 0u;
        }
      }
    }
  }

  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  const int lrow = lane & 7;
  // key-padding mask bytes of this lane's score columns, fetched one block ahead (keys >= M count as padding)
  auto load_mask = [&](int kb) {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = kb * TX_KB + 8 * (j0 + j) + 2 * t + e;
        const bool dead = key >= M || (kpm && kpm[static_cast<long>(img) * M + key]);
        bits |= (dead ? 1u : 0u) << (2 * j + e);
      }
    return bits;
  };
  uint32_t mask_next = load_mask(kb0);

  for (int kb = kb0; kb < kb1; ++kb) {
    const uint32_t mbits = mask_next;
    if (kb + 1 < kb1) mask_next = load_mask(kb + 1);
    ptx::mbar_wait(&full[stage], phase);
    const uint32_t sK = ptx::smem_u32(ring + stage * TX_STAGE);
    const uint32_t sKl = sK + TX_PLANE, sV = sK + 2 * TX_PLANE, sVl = sK + 3 * TX_PLANE;
    // ---- S = q K^T over this warp's keys of the block (n-tiles of 8 keys, 4 k-steps of 16 dims)
    float s[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t bh[2][4], bl[2][4];
      const int row = 8 * (j0 + j) + lrow;   // key; x4 matrices = 16-byte chunks (lane >> 3) [+ 4] of that row
      ldsm_x4(bh[0], sK + swz(row, lane >> 3));
      ldsm_x4(bh[1], sK + swz(row, 4 + (lane >> 3)));
      if (NS == 3) {
        ldsm_x4(bl[0], sKl + swz(row, lane >> 3));
        ldsm_x4(bl[1], sKl + swz(row, 4 + (lane >> 3)));
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = bh[ks >> 1][(ks & 1) * 2], b1 = bh[ks >> 1][(ks & 1) * 2 + 1];
        mma_bf16_16816(s[j], aq[0][ks], b0, b1);
        if (NS == 3) {
          mma_bf16_16816(s[j], aq[NP - 1][ks], b0, b1);
          mma_bf16_16816(s[j], aq[0][ks], bl[ks >> 1][(ks & 1) * 2], bl[ks >> 1][(ks & 1) * 2 + 1]);
        }
      }
    }
    // ---- scale (q / 8 == scores / 8 exactly), key-padding mask, online softmax update
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool dead = (mbits >> (2 * j + e)) & 1u;
        s[j][e] = dead ? -INFINITY : s[j][e] * XA_SCALE_LOG2E;
        s[j][2 + e] = dead ? -INFINITY : s[j][2 + e] * XA_SCALE_LOG2E;
        mx_lo = fmaxf(mx_lo, s[j][e]);
        mx_hi = fmaxf(mx_hi, s[j][2 + e]);
      }
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float mu_lo = (mn_lo == -INFINITY) ? 0.f : mn_lo, mu_hi = (mn_hi == -INFINITY) ? 0.f : mn_hi;
    const float sc_lo = xa_ex2(m_lo - mu_lo), sc_hi = xa_ex2(m_hi - mu_hi);  // 2^-inf == 0 on the first live block
    m_lo = mn_lo; m_hi = mn_hi;
    l_lo *= sc_lo; l_hi *= sc_hi;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= sc_lo; o[n][1] *= sc_lo;
      o[n][2] *= sc_hi; o[n][3] *= sc_hi;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[j][e] = xa_ex2(s[j][e] - mu_lo);
        s[j][2 + e] = xa_ex2(s[j][2 + e] - mu_hi);
        l_lo += s[j][e];
        l_hi += s[j][2 + e];
      }
    // ---- O += P V : P (unnormalised, split) is the A operand straight from the accumulator fragments
    uint32_t ph[NKK][4], pl[NKK][4];
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) {
      split_pack2(s[2 * kk][0], s[2 * kk][1], ph[kk][0], pl[kk][0]);
      split_pack2(s[2 * kk][2], s[2 * kk][3], ph[kk][1], pl[kk][1]);
      split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[kk][2], pl[kk][2]);
      split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[kk][3], pl[kk][3]);
    }
    // v tiles are [key][dim]: the B fragments come through ldmatrix.trans; matrix i of an x4 load = (key half i & 1,
    // dim tile 2 * np + (i >> 1)), i.e. (b0, b1) of two consecutive 8-dim output tiles for one 16-key step
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vh[4], vl[4];
        const int row = 16 * (kk0 + kk) + ((lane >> 3) & 1) * 8 + lrow;  // key
        const int chunk = 2 * np + (lane >> 4);                          // 8-dim tile
        ldsm_x4_trans(vh, sV + swz(row, chunk));
        if (NS == 3) ldsm_x4_trans(vl, sVl + swz(row, chunk));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          mma_bf16_16816(o[2 * np + q], ph[kk], vh[2 * q], vh[2 * q + 1]);
          if (NS == 3) {
            mma_bf16_16816(o[2 * np + q], pl[kk], vh[2 * q], vh[2 * q + 1]);
            mma_bf16_16816(o[2 * np + q], ph[kk], vl[2 * q], vl[2 * q + 1]);
          }
        }
      }
    }
    __syncwarp();                                  // every lane of this warp is done reading the slot ...
    if (lane == 0) ptx::mbar_arrive(&empty[stage]);  // ... hand it back to the producer
    if (++stage == TX_STAGES) { stage = 0; phase ^= 1; }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);

  // ---- publish (m, l, o[64]) per query row and partial; the last CTA of the pair merges.  A partial is one key
  //      split (64 rows) or, for Q16, one warp of one key split (16 rows).
  constexpr int PR = Q16 ? 16 : 64;           // rows per partial
  const int nparts = gs * NKG;
  const long pair_rows = static_cast<long>(max_parts) * NKG * PR;  // partial rows reserved per pair
  {
    const int part = split * NKG + kg;
    float* p_lo = partial + (static_cast<long>(pair) * pair_rows + part * PR + r_lo) * TX_PART;
    float* p_hi = partial + (static_cast<long>(pair) * pair_rows + part * PR + r_hi) * TX_PART;
    if (t == 0) {
      p_lo[0] = m_lo; p_lo[1] = l_lo;
      p_hi[0] = m_hi; p_hi[1] = l_hi;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<float2*>(p_lo + 2 + 8 * n + 2 * t) = make_float2(o[n][0], o[n][1]);
      *reinterpret_cast<float2*>(p_hi + 2 + 8 * n + 2 * t) = make_float2(o[n][2], o[n][3]);
    }
  }
  __threadfence();
  consumer_sync();
  if (tid == 0) last_flag = (atomicAdd(&counters[pair], 1) == gs - 1);
  consumer_sync();
  if (!last_flag) continue;
  __threadfence();
  {
    constexpr int DPT = 64 * PR / NT;         // dims per merging thread: 32 / 8 / 16
    constexpr int TPR = 64 / DPT;             // threads per row
    const int row = tid / TPR, half = (tid % TPR) * DPT;
    if (q0 + row < Ncap) {
      const float* base = partial + (static_cast<long>(pair) * pair_rows + row) * TX_PART;
      const long sstride = static_cast<long>(PR) * TX_PART;
      float mm = -INFINITY;
      for (int sidx = 0; sidx < nparts; ++sidx) mm = fmaxf(mm, base[sidx * sstride]);
      float ltot = 0.f, acc[DPT];
#pragma unroll
      for (int i = 0; i < DPT; ++i) acc[i] = 0.f;
      for (int sidx = 0; sidx < nparts; ++sidx) {
        const float* ps = base + sidx * sstride;
        const float ms = ps[0];
        const float w = (ms == -INFINITY) ? 0.f : xa_ex2(ms - mm);   // partial maxima are in log2 units
        ltot += w * ps[1];
#pragma unroll
        for (int i = 0; i < DPT; i += 2) {
          const float2 v = *reinterpret_cast<const float2*>(ps + 2 + half + i);
          acc[i] = fmaf(w, v.x, acc[i]);
          acc[i + 1] = fmaf(w, v.y, acc[i + 1]);
        }
      }
      const float inv = 1.0f / ltot;
      const long ob = (static_cast<long>(img) * Ncap + q0 + row) * 512 + h * 64 + half;
#pragma unroll
      for (int i = 0; i < DPT; i += 8) {
        if (out_f32) {
          *reinterpret_cast<float4*>(out_f32 + ob + i) = make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv);
          *reinterpret_cast<float4*>(out_f32 + ob + i + 4) = make_float4(acc[i + 4] * inv, acc[i + 5] * inv, acc[i + 6] * inv, acc[i + 7] * inv);
        }
        if (out_hi) {
          uint4 hh, ll;
          split_pack2(acc[i] * inv, acc[i + 1] * inv, hh.x, ll.x);
          split_pack2(acc[i + 2] * inv, acc[i + 3] * inv, hh.y, ll.y);
          split_pack2(acc[i + 4] * inv, acc[i + 5] * inv, hh.z, ll.z);
          split_pack2(acc[i + 6] * inv, acc[i + 7] * inv, hh.w, ll.w);
          *reinterpret_cast<uint4*>(out_hi + ob + i) = hh;
          if (out_lo) *reinterpret_cast<uint4*>(out_lo + ob + i) = ll;
        }
      }
    }
  }
  if (tid == 0) counters[pair] = 0;  // ready for the next launch (graph replay)
  }  // segments of this CTA's run
}

}  // namespace

void cross_attn_tma_plan(Ctx* c, int nimg, int Ncap, int M, int* grid, int* max_parts, int* pairs) {
  const int nqb = (Ncap + 63) / 64;
  const int np = nimg * 8 * nqb;
  const int nkb = (M + TX_KB - 1) / TX_KB;
  const long NB = static_cast<long>(np) * nkb;
  const int G = static_cast<int>(std::min<long>(NB, c->num_sms));
  const int bpc = static_cast<int>(NB / G);
  *grid = G;
  *max_parts = (nkb + bpc - 1) / bpc + 1;
  *pairs = np;
}

namespace {

// [B*96][M][64] bf16 plane as a 3-D tensor map, box = one 64-key x 64-dim tile, 128-byte swizzle, OOB keys read as zero
CUtensorMap plane_map(Ctx* c, const bf16* base, int M, long slices) {
  auto key = std::make_tuple(static_cast<const void*>(base), M, slices);
  auto it = c->xattn_tmaps.find(key);
  if (it != c->xattn_tmaps.end()) return it->second;
  if (c->xattn_tmaps.size() > 256) c->xattn_tmaps.clear();
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, ALM_ERR_INVALID, "K/V cache plane not 16-byte aligned");
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, cuuint64_t(M), cuuint64_t(slices)};
  cuuint64_t strides[2] = {128, cuuint64_t(M) * 128};
  cuuint32_t box[3] = {64, cuuint32_t(TX_KB), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled (K/V cache plane) failed with CUresult " + std::to_string(int(r))};
  return c->xattn_tmaps.emplace(key, tm).first->second;
}

}  // namespace

// kc_* / vc_* are the BASES of the whole cache ([nimg_total*96][M][64]); z0 = (first image) * 96 + (decoder-layer) * 8
void cross_attn_tma(Ctx* c, const bf16* q_hi, const bf16* q_lo, const float* q_f32, int nimg, int Ncap, const bf16* kc_hi,
                    const bf16* kc_lo, const bf16* vc_hi, const bf16* vc_lo, long slices, int z0, const uint8_t* kpm, int M,
                    int grid, int max_parts, float* partial, int* counters, bf16* out_hi, bf16* out_lo, float* out_f32) {
  if (c->skipped(1)) return;
  const int nqb = (Ncap + 63) / 64;
  static DeviceOnce attr;
  if (attr.need()) {
    auto prep = [](auto* k) {
      ALM_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM));
      pin_carveout(k);
    };
    prep(cross_attn_tma_kernel<3, 1>); prep(cross_attn_tma_kernel<1, 1>);
    prep(cross_attn_tma_kernel<3, 2>); prep(cross_attn_tma_kernel<1, 2>);
    attr.mark();
  }
  ALM_REQUIRE(q_f32 || q_hi, ALM_ERR_INVALID, "cross_attn_tma: no query operand");
  const bool q16 = Ncap <= 16;
  ALM_REQUIRE(q16 || (out_hi && !out_f32), ALM_ERR_INVALID, "cross_attn_tma: fp32 output only on the <= 16-query path");
  const bool three = c->nsplit == 3 && (q_f32 || q_lo) && kc_lo && vc_lo;
  const CUtensorMap tkh = plane_map(c, kc_hi, M, slices), tvh = plane_map(c, vc_hi, M, slices);
  const CUtensorMap tkl = three ? plane_map(c, kc_lo, M, slices) : tkh, tvl = three ? plane_map(c, vc_lo, M, slices) : tvh;
  const int npairs = nimg * 8 * nqb;
#define ALM_TX_LAUNCH(NS, MODE)                                                                                       \
  cross_attn_tma_kernel<NS, MODE><<<grid, MODE == 2 ? 288 : 160, TX_SMEM, c->stream>>>(                               \
      tkh, tkl, tvh, tvl, q_hi, q_lo, q_f32, Ncap, kpm, M, nqb, npairs, max_parts, z0, partial, counters, out_hi,     \
      out_lo, out_f32)
  if (three) { if (q16) ALM_TX_LAUNCH(3, 1); else ALM_TX_LAUNCH(3, 2); }
  else       { if (q16) ALM_TX_LAUNCH(1, 1); else ALM_TX_LAUNCH(1, 2); }
#undef ALM_TX_LAUNCH
  count_launch(c); check_launch("cross_attn_tma");
}

}  // namespace alm
