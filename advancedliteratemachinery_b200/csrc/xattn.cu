// Fused multi-query cross-attention for the batched poly / rec decode loops (transformer.py:444-447 with
// Ncap live sequences per image): per (image, head, 64-query block)
//     S = (q / 8) K_c^T  ->  key-padding mask  ->  online softmax  ->  O = P V_c
// in ONE pass over the cached split-bf16 K_c / V_c of the image, flash-attention style: the [S*8, M] score and
// probability matrices never touch HBM (they were 4 x 134 MB per layer-step at 16 pages x 64 instances).
// Both products run on mma.sync.m16n8k16 with the three-term split hi*hi + lo*hi + hi*lo accumulated in fp32 --
// the same fp32-class scheme as the GEMM engine; softmax statistics stay in fp32 registers.  64-query tiles are
// half a tcgen05 instruction's rows, so the warp-level MMA is the right granularity (see kernels.cu, window
// attention).  HBM floor: K_c + V_c of one decoder layer are streamed once per layer-step.
//
//   Persistent grid (2 CTAs per SM); the (image, head, query block) x key-block work list is cut into equal
//   contiguous runs, one per CTA.  4 warps, warp w owns query rows 16w..16w+15 of the block.
//   MODE 2 (optional, "xattn_wg" 2): 8 warps -- the four 16-row tiles times two key groups of 32 keys per block,
//   twice the warps per SM for the same shared memory (the 4-warp kernel is MMA-latency bound at 8 warps per SM).
//   MODE 1 / Q16 variant (<= 16 sequences per image: the pt loop has ONE): all four warps share the single 16-row query
//   tile and split every 64-key block four ways (16 keys each), so the MMA work drops 4x and the kernel is purely
//   HBM-bound on the K_c / V_c^T stream; each warp publishes its own (m, l, o) partial.
//   Key blocks of 64 are staged with cp.async into a 2-stage ring (hi/lo planes of K and V, both [key][dim],
//   144-byte pitch -> conflict-free ldmatrix).  When a run boundary cuts an (image, head, query block), the CTAs
//   sharing it publish partials (m, l, o[64]) and the last one to arrive merges them (self-resetting counter).
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


constexpr int MQ_KB = 64;                     // keys per block
constexpr int MQ_PITCH = 72;                  // bf16 per staged row (64 + 8 pad)
constexpr int MQ_PLANE = 64 * MQ_PITCH;       // one plane tile (64 rows)
constexpr int MQ_STAGE = 4 * MQ_PLANE;        // K hi, K lo, V^T hi, V^T lo
constexpr int MQ_STAGES = 2;
constexpr int MQ_SMEM = MQ_STAGES * MQ_STAGE * 2 + MQ_STAGES * MQ_KB;  // tiles + per-stage key mask bytes
constexpr int MQ_PART = 66;                   // floats per partial row: m, l, o[64]

template <int NS, int MODE>  // MODE 0: 4 warps = 4 row tiles; 1: 4 warps = 1 row tile x 4 key groups; 2: 8 warps = 4 x 2
__global__ void __launch_bounds__(MODE == 2 ? 256 : 128, MODE == 2 ? 2 : 3)
cross_attn_mq_kernel(const bf16* __restrict__ q_hi, const bf16* __restrict__ q_lo, const float* __restrict__ q_f32, int Ncap,
                     const bf16* __restrict__ kc_hi, const bf16* __restrict__ kc_lo, const bf16* __restrict__ vc_hi,
                     const bf16* __restrict__ vc_lo, const uint8_t* __restrict__ kpm, int M, int nqb,
                     int npairs, int max_parts, float* __restrict__ partial, int* __restrict__ counters,
                     bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, float* __restrict__ out_f32) {
  constexpr bool Q16 = MODE == 1;
  constexpr int NT = MODE == 2 ? 256 : 128;           // threads
  constexpr int NKG = MODE == 1 ? 4 : (MODE == 2 ? 2 : 1);  // key groups a 64-key block is split into
  constexpr int NJ = 8 / NKG;        // 8-key n-tiles of a block this warp scores
  constexpr int NKK = 4 / NKG;       // 16-key k-steps of a block this warp feeds into P.V
  extern __shared__ __align__(16) unsigned char mq_smem[];
  bf16* tiles = reinterpret_cast<bf16*>(mq_smem);
  uint8_t* smask = mq_smem + MQ_STAGES * MQ_STAGE * 2;
  __shared__ int last_flag;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nkb = (M + MQ_KB - 1) / MQ_KB;
  // ---- persistent, balanced schedule: the npairs x nkb key blocks are cut into gridDim.x equal contiguous runs;
  //      a run may cover the tail of one (image, head, query block) and the head of the next (two segments)
  const long NB = static_cast<long>(npairs) * nkb, G = gridDim.x;
  auto owner = [&](long b) { return static_cast<int>(((b + 1) * G - 1) / NB); };  // CTA whose run contains block b
  const long b_end = (static_cast<long>(blockIdx.x) + 1) * NB / G;
  for (long b = static_cast<long>(blockIdx.x) * NB / G; b < b_end;) {
  const int pair = static_cast<int>(b / nkb), qb = pair % nqb, ih = pair / nqb, h = ih & 7, img = ih >> 3;
  const int kb0 = static_cast<int>(b - static_cast<long>(pair) * nkb);
  const int kb1 = static_cast<int>(min(static_cast<long>(nkb), kb0 + (b_end - b)));
  b += kb1 - kb0;
  const int first = owner(static_cast<long>(pair) * nkb);
  const int gs = owner(static_cast<long>(pair) * nkb + nkb - 1) - first + 1;  // CTAs sharing this pair
  const int split = static_cast<int>(blockIdx.x) - first;
  const int q0 = qb * 64;
  const int kg = Q16 ? warp : (warp >> 2);  // key group of this warp (MODE 0: always 0)
  const int j0 = NJ * kg, kk0 = NKK * kg;
  const int r_lo = (Q16 ? 0 : (warp & 3) * 16) + g, r_hi = r_lo + 8;  // the two query rows (within the block) this lane holds
  const bool live_lo = q0 + r_lo < Ncap, live_hi = q0 + r_hi < Ncap;

  const long kbase = (static_cast<long>(img) * 96 + h) * M * 64;  // K_c / V_c [img][dl(base)][h][key][64]
  auto load_block = [&](int kb, int stage) {
    bf16* st = tiles + stage * MQ_STAGE;
    const int key0 = kb * MQ_KB;
#pragma unroll
    for (int i = tid; i < 512; i += NT) {
      // row r = key, 8 dims per 16-byte chunk; a warp instruction covers 4 consecutive keys = 512 contiguous bytes,
      // the whole block 8 KB contiguous per plane; keys >= M are zero-filled
      const int r = i >> 3, ch = (i & 7) * 8;
      const int key = key0 + r;
      const bool ok = key < M;
      const long src = kbase + static_cast<long>(ok ? key : 0) * 64 + ch;
      const int nb = ok ? 16 : 0;
      cp_async16(st + r * MQ_PITCH + ch, kc_hi + src, nb);
      if (NS == 3) cp_async16(st + MQ_PLANE + r * MQ_PITCH + ch, kc_lo + src, nb);
      cp_async16(st + 2 * MQ_PLANE + r * MQ_PITCH + ch, vc_hi + src, nb);
      if (NS == 3) cp_async16(st + 3 * MQ_PLANE + r * MQ_PITCH + ch, vc_lo + src, nb);
    }
    if (tid < MQ_KB) {
      const int key = key0 + tid;
      smask[stage * MQ_KB + tid] = (key < M) ? (kpm ? kpm[static_cast<long>(img) * M + key] : 0) : 1;
    }
  };

  load_block(kb0, 0);  // first key block in flight while the query fragments are fetched
  cp_async_commit();

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
  const int lrow = lane & 7, lcol = (lane >> 3) * 8;  // ldmatrix.x4: row within the 8-row group, 8-column chunk

  for (int kb = kb0; kb < kb1; ++kb) {
    const int stage = (kb - kb0) & 1;
    if (kb + 1 < kb1) {
      load_block(kb + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* sKh = tiles + stage * MQ_STAGE;
    const bf16* sKl = sKh + MQ_PLANE;
    const bf16* sVh = sKh + 2 * MQ_PLANE;
    const bf16* sVl = sKh + 3 * MQ_PLANE;
    // ---- S = q K^T over this warp's keys of the block (n-tiles of 8 keys, 4 k-steps of 16 dims)
    float s[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t bh[2][4], bl[2][4];
      const int roff = (8 * (j0 + j) + lrow) * MQ_PITCH + lcol;
      ldmatrix_x4(bh[0], sKh + roff);
      ldmatrix_x4(bh[1], sKh + roff + 32);
      if (NS == 3) {
        ldmatrix_x4(bl[0], sKl + roff);
        ldmatrix_x4(bl[1], sKl + roff + 32);
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
    const uint8_t* mk = smask + stage * MQ_KB;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const uint32_t mm = *reinterpret_cast<const uint16_t*>(mk + 8 * (j0 + j) + 2 * t);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool dead = (mm >> (8 * e)) & 0xffu;
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
        const int roff = (16 * (kk0 + kk) + ((lane >> 3) & 1) * 8 + lrow) * MQ_PITCH + (2 * np + (lane >> 4)) * 8;
        ldmatrix_x4_trans(vh, sVh + roff);
        if (NS == 3) ldmatrix_x4_trans(vl, sVl + roff);
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
    __syncthreads();  // every warp is done with this stage before the next iteration's cp.async overwrites it
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);

  if (MODE == 0 && gs == 1) {
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    const long ob_lo = (static_cast<long>(img) * Ncap + q0 + r_lo) * 512 + h * 64 + 2 * t;
    const long ob_hi = (static_cast<long>(img) * Ncap + q0 + r_hi) * 512 + h * 64 + 2 * t;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      uint32_t hh, ll;
      if (live_lo) {
        split_pack2(o[n][0] * inv_lo, o[n][1] * inv_lo, hh, ll);
        *reinterpret_cast<uint32_t*>(out_hi + ob_lo + 8 * n) = hh;
        if (out_lo) *reinterpret_cast<uint32_t*>(out_lo + ob_lo + 8 * n) = ll;
      }
      if (live_hi) {
        split_pack2(o[n][2] * inv_hi, o[n][3] * inv_hi, hh, ll);
        *reinterpret_cast<uint32_t*>(out_hi + ob_hi + 8 * n) = hh;
        if (out_lo) *reinterpret_cast<uint32_t*>(out_lo + ob_hi + 8 * n) = ll;
      }
    }
    continue;
  }
  // ---- publish (m, l, o[64]) per query row and partial; the last CTA of the pair merges.  A partial is one key
  //      split (64 rows) or, for Q16, one warp of one key split (16 rows).
  constexpr int PR = Q16 ? 16 : 64;           // rows per partial
  const int nparts = gs * NKG;
  const long pair_rows = static_cast<long>(max_parts) * NKG * PR;  // partial rows reserved per pair
  {
    const int part = split * NKG + kg;
    float* p_lo = partial + (static_cast<long>(pair) * pair_rows + part * PR + r_lo) * MQ_PART;
    float* p_hi = partial + (static_cast<long>(pair) * pair_rows + part * PR + r_hi) * MQ_PART;
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
  __syncthreads();
  if (tid == 0) last_flag = (atomicAdd(&counters[pair], 1) == gs - 1);
  __syncthreads();
  if (!last_flag) continue;
  __threadfence();
  {
    constexpr int DPT = 64 * PR / NT;         // dims per merging thread: 32 / 8 / 16
    constexpr int TPR = 64 / DPT;             // threads per row
    const int row = tid / TPR, half = (tid % TPR) * DPT;
    if (q0 + row < Ncap) {
      const float* base = partial + (static_cast<long>(pair) * pair_rows + row) * MQ_PART;
      const long sstride = static_cast<long>(PR) * MQ_PART;
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

// Launch plan: a persistent grid of `ctas_per_sm` CTAs per SM (default 2: 148 KB of the SM's shared memory, so the
// small kernels of the other in-flight decode streams still find room next to this HBM-bound one), never more
// CTAs than key blocks.
void cross_attn_mq_plan(Ctx* c, int nimg, int Ncap, int M, int* grid, int* max_parts, int* pairs) {
  const int nqb = (Ncap + 63) / 64;
  const int np = nimg * 8 * nqb;
  const int nkb = (M + MQ_KB - 1) / MQ_KB;
  const long NB = static_cast<long>(np) * nkb;
  const int G = static_cast<int>(std::min<long>(NB, static_cast<long>(c->num_sms) * std::max(1, c->xattn_ctas_per_sm)));
  const int bpc = static_cast<int>(NB / G);  // >= 1 key blocks in every run
  *grid = G;
  *max_parts = (nkb + bpc - 1) / bpc + 1;
  *pairs = np;
}

size_t cross_attn_mq_partial_floats(int pairs, int max_parts) {
  return static_cast<size_t>(pairs) * max_parts * 128 * MQ_PART;  // up to 2 key groups x 64 rows (or 4 x 16) per part
}

void cross_attn_mq(Ctx* c, const bf16* q_hi, const bf16* q_lo, const float* q_f32, int nimg, int Ncap, const bf16* kc_hi,
                   const bf16* kc_lo, const bf16* vc_hi, const bf16* vc_lo, const uint8_t* kpm, int M,
                   int grid, int max_parts, float* partial, int* counters, bf16* out_hi, bf16* out_lo,
                   float* out_f32) {
  if (c->skipped(1)) return;
  const int nqb = (Ncap + 63) / 64;
  static DeviceOnce attr;
  if (attr.need()) {
    auto prep = [](auto* k) {
      ALM_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, MQ_SMEM));
      pin_carveout(k);
    };
    prep(cross_attn_mq_kernel<3, 0>); prep(cross_attn_mq_kernel<1, 0>);
    prep(cross_attn_mq_kernel<3, 1>); prep(cross_attn_mq_kernel<1, 1>);
    prep(cross_attn_mq_kernel<3, 2>); prep(cross_attn_mq_kernel<1, 2>);
    attr.mark();
  }
  ALM_REQUIRE(q_f32 || q_hi, ALM_ERR_INVALID, "cross_attn_mq: no query operand");
  const bool q16 = Ncap <= 16;
  ALM_REQUIRE(q16 || (out_hi && !out_f32), ALM_ERR_INVALID, "cross_attn_mq: fp32 output only on the <= 16-query path");
  const bool three = c->nsplit == 3 && (q_f32 || q_lo) && kc_lo && vc_lo;
  const int npairs = nimg * 8 * nqb;
  const int mode = q16 ? 1 : (c->xattn_wg == 2 ? 2 : 0);
#define ALM_MQ_LAUNCH(NS, MODE)                                                                                          \
  cross_attn_mq_kernel<NS, MODE><<<grid, MODE == 2 ? 256 : 128, MQ_SMEM, c->stream>>>(                                   \
      q_hi, q_lo, q_f32, Ncap, kc_hi, kc_lo, vc_hi, vc_lo, kpm, M, nqb, npairs, max_parts, partial, counters, out_hi,    \
      out_lo, out_f32)
  if (three) { if (mode == 1) ALM_MQ_LAUNCH(3, 1); else if (mode == 2) ALM_MQ_LAUNCH(3, 2); else ALM_MQ_LAUNCH(3, 0); }
  else       { if (mode == 1) ALM_MQ_LAUNCH(1, 1); else if (mode == 2) ALM_MQ_LAUNCH(1, 2); else ALM_MQ_LAUNCH(1, 0); }
#undef ALM_MQ_LAUNCH
  count_launch(c); check_launch("cross_attn_mq");
}

}  // namespace alm
