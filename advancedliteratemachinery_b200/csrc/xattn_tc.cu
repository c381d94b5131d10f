// Decoder cross-attention on the 5th-gen tensor cores with a TMA operand ring ("xattn_impl" 3): the same math and the
// same persistent run / partial-merge schedule as xattn.cu (transformer.py:444-447 with Ncap live sequences per image
// sharing the image's cached K_c / V_c), rebuilt the Blackwell way:
//
//   * K_c / V_c key blocks of 128 x 64 (hi and lo planes: 64 KB per block) arrive through a 3-deep TMA ring -- 192 KB in
//     flight per SM and not one load instruction in the math warps (the cp.async kernel sits at 0.62 of the HBM peak with
//     32 KB in flight per CTA);
//   * S = Q K^T per key block: tcgen05.mma M 128 (the <= 128 query rows of one image; rows >= Ncap are zero) x N 128 x
//     K 64, 3-term split, accumulator in TMEM, double buffered so that S(i+1) runs under the softmax of block i;
//   * block softmax in registers (thread = query row = TMEM lane): key-padding mask, block max m_b, p = 2^((s - m_b) c),
//     block sum l_b; P goes back to TMEM as the A operand (tcgen05.st) and O_b = P V_b is a fresh TS-form MMA with V as
//     an MN-major operand;
//   * the per-block (m_b, l_b, O_b) are folded into a running (m, l, O[64]) held in REGISTERS with the usual flash
//     rescale -- tensor memory is never rescaled and the MMA pipe never waits for a correction step;
//   * a run boundary inside an (image, head) pair publishes the partial (m, l, O) and the last CTA to arrive merges,
//     as in xattn.cu.
//
// One kernel serves the point loop (one query per image: HBM-bound, the MMA rows are nearly all padding but the tensor
// pipe is otherwise idle) and the polygon / recognition loops (64 queries per image: the mma.sync kernel was
// MMA-latency bound at 46 % of the HMMA pipe).  Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer, warps
// 2..17 softmax / fold / merge: FOUR threads per query row (lane quarter = warp % 4 selects the 32 rows a warp may touch in
// tensor memory, key group = (warp - 2) / 4 selects 32 of the block's 128 keys and 16 of the 64 output dims), so the
// per-block softmax is a quarter of a row's work per thread and a single tensor-memory round trip.
#include <algorithm>

#include "omni.h"
#include "ptx.cuh"

namespace alm {

namespace {

constexpr int XT_THREADS = 576;            // warp 0 TMA, warp 1 MMA, warps 2..17 softmax: lane quarter = w % 4, key group = (w - 2) / 4
constexpr int XT_KB = 128;               // keys per block
constexpr int XT_TILE = XT_KB * 128;     // bytes of one 128 x 64 bf16 operand tile
constexpr int XT_STAGES = 3;
constexpr int XT_PART = 66;              // floats per partial row: m, l, o[64]

template <int NSPLIT>
struct XtSmem {
  static constexpr int NP = NSPLIT == 3 ? 2 : 1;
  static constexpr int kStage = 2 * NP * XT_TILE;          // K and V of every plane
  static constexpr int kQ = XT_STAGES * kStage;
  static constexpr int kBar = kQ + NP * XT_TILE;
  static constexpr int kXch = kBar + 256;                  // per-row exchange between the four key groups: [4][128] floats
  static constexpr int kTotal = kXch + 4 * 128 * 4;        // 231 680 B: the dynamic array is declared 1024-byte aligned
};

struct XtParams {
  int Ncap, M, nkb, nqb, npairs, max_parts, z0;
  const bf16* q_hi;
  const bf16* q_lo;
  const float* q_f32;
  const uint8_t* kpm;
  float* partial;
  int* counters;
  bf16* out_hi;
  bf16* out_lo;
  float* out_f32;
  float scale_log2e;
};

__device__ __forceinline__ float xt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NSPLIT>
__global__ void __launch_bounds__(XT_THREADS, 1)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_kh, const __grid_constant__ CUtensorMap tm_kl,
                     const __grid_constant__ CUtensorMap tm_vh, const __grid_constant__ CUtensorMap tm_vl, const XtParams p) {
  using L = XtSmem<NSPLIT>;
  constexpr int NP = L::NP;
  extern __shared__ __align__(1024) uint8_t xt_raw[];   // no static shared memory in this kernel: the window starts aligned
  uint8_t* smem = xt_raw;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();     // the 128-byte swizzle atoms need it; never expected to fire
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* kv_full = bars + 0;     // [3]
  uint64_t* kv_empty = bars + 3;    // [3]
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* p_full = bars + 8;      // [2]
  uint64_t* o_full = bars + 10;     // [2]
  uint64_t* o_empty = bars + 12;    // [2]
  uint64_t* q_ready = bars + 14;    // softmax warps -> MMA: the Q tile of the segment is in shared memory
  uint64_t* q_free = bars + 15;     // MMA -> softmax warps: every product that reads the Q tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  int* last_flag = reinterpret_cast<int*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // TMEM: S[2] at 0 / 128 (the lo halves of P overwrite S in place), P_hi[2] at 256 / 320, O[2] at 384 / 448
  constexpr uint32_t S_COL[2] = {0, 128}, P_COL[2] = {256, 320}, O_COL[2] = {384, 448};

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_kh);
    ptx::prefetch_tmap(&tm_vh);
    for (int s = 0; s < XT_STAGES; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 16);
      ptx::mbar_init(&o_full[s], 1);
      ptx::mbar_init(&o_empty[s], 16);
    }
    ptx::mbar_init(q_ready, 16);
    ptx::mbar_init(q_free, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int nkb = p.nkb, nqb = p.nqb;
  const long NB = static_cast<long>(p.npairs) * nkb, G = gridDim.x;
  const long b_begin = static_cast<long>(blockIdx.x) * NB / G, b_end = (static_cast<long>(blockIdx.x) + 1) * NB / G;

  if (warp == 0) {
    // ================================================================================= TMA producer: K / V ring
    int st = 0;
    uint32_t phase = 0;
    for (long b = b_begin; b < b_end; ++b) {
      const int pair = static_cast<int>(b / nkb), kb = static_cast<int>(b - static_cast<long>(pair) * nkb);
      const int ih = pair / nqb, h = ih & 7, img = ih >> 3;
      const int z = p.z0 + img * 96 + h;   // slice of the [slices][M][64] cache: (image, decoder-layer, head)
      ptx::mbar_wait(&kv_empty[st], phase ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&kv_full[st], 2 * NP * XT_TILE);
        uint8_t* base = smem + st * L::kStage;
        ptx::tma_load_3d(base, &tm_kh, &kv_full[st], 0, kb * XT_KB, z);
        if (NSPLIT == 3) ptx::tma_load_3d(base + XT_TILE, &tm_kl, &kv_full[st], 0, kb * XT_KB, z);
        ptx::tma_load_3d(base + NP * XT_TILE, &tm_vh, &kv_full[st], 0, kb * XT_KB, z);
        if (NSPLIT == 3) ptx::tma_load_3d(base + NP * XT_TILE + XT_TILE, &tm_vl, &kv_full[st], 0, kb * XT_KB, z);
      }
      __syncwarp();
      if (++st == XT_STAGES) { st = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_s = idesc_base | (uint32_t(XT_KB >> 3) << 17);
    const uint32_t idesc_pv = idesc_base | (1u << 16) | (uint32_t(64 >> 3) << 17);   // B (= V) MN-major
    const uint32_t q_base = ptx::smem_u32(smem + L::kQ);
    int st = 0;
    uint32_t kv_phase = 0, p_phase[2] = {0, 0}, oe_phase[2] = {0, 0}, qr_phase = 0;
    int sb = 0;       // S / P / O buffer of the segment's first block
    auto issue_s = [&](int stage, int buf) {
      const uint32_t kb = ptx::smem_u32(smem + stage * L::kStage);
#pragma unroll
      for (int pass = 0; pass < NSPLIT; ++pass)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16(tmem + S_COL[buf], ptx::make_kmajor_sw128_desc(q_base + (pass == 1 ? XT_TILE : 0) + k * 32),
                         ptx::make_kmajor_sw128_desc(kb + (pass == 2 ? XT_TILE : 0) + k * 32), idesc_s, (pass | k) != 0);
      ptx::umma_commit(&s_full[buf]);
    };
    for (long b = b_begin; b < b_end;) {
      const int pair = static_cast<int>(b / nkb);
      const int kb0 = static_cast<int>(b - static_cast<long>(pair) * nkb);
      const int n = static_cast<int>(min(static_cast<long>(nkb - kb0), b_end - b));   // key blocks of this segment
      b += n;
      ptx::mbar_wait(q_ready, qr_phase);
      qr_phase ^= 1;
      // software pipeline: S(0); then per block i: S(i+1) is issued BEFORE waiting for P(i), so it runs under softmax(i)
      int st_s = st;
      uint32_t ph_s = kv_phase;
      ptx::mbar_wait(&kv_full[st_s], ph_s);
      ptx::tc_fence_after();
      if (ptx::elect_one()) issue_s(st_s, sb);
      __syncwarp();
      for (int i = 0; i < n; ++i) {
        const int buf = (sb + i) & 1;
        if (i + 1 < n) {
          if (++st_s == XT_STAGES) { st_s = 0; ph_s ^= 1; }
          ptx::mbar_wait(&kv_full[st_s], ph_s);
          // S buffer (buf ^ 1) is free: P.V(i-1), which read the lo halves of P out of it, was issued before this point
          ptx::tc_fence_after();
          if (ptx::elect_one()) issue_s(st_s, buf ^ 1);
          __syncwarp();
        }
        ptx::mbar_wait(&p_full[buf], p_phase[buf]);
        p_phase[buf] ^= 1;
        ptx::mbar_wait(&o_empty[buf], oe_phase[buf] ^ 1);
        oe_phase[buf] ^= 1;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t vb = ptx::smem_u32(smem + st * L::kStage + NP * XT_TILE);
#pragma unroll
          for (int pass = 0; pass < NSPLIT; ++pass)
#pragma unroll
            for (int j = 0; j < XT_KB / 16; ++j)
              ptx::umma_bf16_ts(tmem + O_COL[buf], tmem + (pass == 1 ? S_COL[buf] : P_COL[buf]) + j * 8,
                                ptx::make_kmajor_sw128_desc(vb + (pass == 2 ? XT_TILE : 0) + j * 2048), idesc_pv, (pass | j) != 0);
          ptx::umma_commit(&o_full[buf]);
          ptx::umma_commit(&kv_empty[st]);   // K and V of this block are free once S(i) and P.V(i) have retired
          if (i == n - 1) ptx::umma_commit(q_free);
        }
        __syncwarp();
        if (++st == XT_STAGES) { st = 0; kv_phase ^= 1; }
      }
      sb = (sb + n) & 1;
    }
  } else {
    // ================================================================================= softmax / fold / merge warps
    const int quarter = warp & 3, cg = (warp - 2) >> 2;   // lane quarter, key / output-dim group
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const int r = quarter * 32 + lane;                 // query row of the 128-row tile
    const int tid16 = threadIdx.x - 64;                // 0..511 among the softmax warps
    float* xch = reinterpret_cast<float*>(smem + L::kXch);   // [4 key groups][128 rows]
    uint32_t s_phase[2] = {0, 0}, o_phase[2] = {0, 0}, qf_phase = 0;
    int sb = 0;
    bool first_segment = true;
    for (long b = b_begin; b < b_end;) {
      const int pair = static_cast<int>(b / nkb), qb = pair % nqb, ih = pair / nqb, h = ih & 7, img = ih >> 3;
      const int kb0 = static_cast<int>(b - static_cast<long>(pair) * nkb);
      const int n = static_cast<int>(min(static_cast<long>(nkb - kb0), b_end - b));
      b += n;
      auto owner = [&](long bb) { return static_cast<int>(((bb + 1) * G - 1) / NB); };
      const int first = owner(static_cast<long>(pair) * nkb);
      const int gs = owner(static_cast<long>(pair) * nkb + nkb - 1) - first + 1;   // CTAs sharing this pair
      const int split = static_cast<int>(blockIdx.x) - first;
      const int q0 = qb * 128;
      const bool live = q0 + r < p.Ncap;
      const bool warp_live = q0 + quarter * 32 < p.Ncap;
      // ---- Q tile of the segment -> shared memory in the K-major SW128 layout (rows >= Ncap are zero)
      if (!first_segment) {
        ptx::mbar_wait(q_free, qf_phase);   // the previous segment's products no longer read the tile
        qf_phase ^= 1;
      }
      first_segment = false;
      for (int i = tid16; i < 128 * 8; i += 512) {
        const int row = i >> 3, ch = i & 7;
        uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
        if (q0 + row < p.Ncap) {
          const long src = (static_cast<long>(img) * p.Ncap + q0 + row) * 512 + h * 64 + ch * 8;
          if (p.q_f32) {
            const float4 a = *reinterpret_cast<const float4*>(p.q_f32 + src);
            const float4 c = *reinterpret_cast<const float4*>(p.q_f32 + src + 4);
            uint32_t hh[4], ll[4];
            split_pack2_bf16(a.x, a.y, hh[0], ll[0]);
            split_pack2_bf16(a.z, a.w, hh[1], ll[1]);
            split_pack2_bf16(c.x, c.y, hh[2], ll[2]);
            split_pack2_bf16(c.z, c.w, hh[3], ll[3]);
            vh = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            vl = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          } else {
            vh = *reinterpret_cast<const uint4*>(p.q_hi + src);
            if (NSPLIT == 3) vl = *reinterpret_cast<const uint4*>(p.q_lo + src);
          }
        }
        const uint32_t off = static_cast<uint32_t>(row * 128 + ((ch ^ (row & 7)) << 4));   // 128-byte swizzle
        *reinterpret_cast<uint4*>(smem + L::kQ + off) = vh;
        if (NSPLIT == 3) *reinterpret_cast<uint4*>(smem + L::kQ + XT_TILE + off) = vl;
      }
      ptx::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(q_ready);

      // ---- running flash statistics: m is the row's (all four threads agree), l and O are this thread's share
      float m_run = -INFINITY, l_run = 0.f;
      float o_run[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) o_run[d] = 0.f;
      float m_pend = -INFINITY, l_pend = 0.f;   // block statistics of the block whose O_b is still in tensor memory
      const uint8_t* kpm = p.kpm ? p.kpm + static_cast<long>(img) * p.M : nullptr;

      auto fold = [&](int buf) {
        // (m, l, O) <- flash merge of the running state with block (m_pend, l_pend, O_b); this thread folds 16 output dims
        ptx::mbar_wait(&o_full[buf], o_phase[buf]);
        o_phase[buf] ^= 1;
        ptx::tc_fence_after();
        if (warp_live) {
          const float m_new = fmaxf(m_run, m_pend);
          const float a = m_run == -INFINITY ? 0.f : xt_ex2((m_run - m_new) * p.scale_log2e);
          const float c = m_pend == -INFINITY ? 0.f : xt_ex2((m_pend - m_new) * p.scale_log2e);
          uint32_t v[16];
          ptx::tmem_ld_32x16(lane_addr + O_COL[buf] + cg * 16, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) o_run[j] = fmaf(o_run[j], a, __uint_as_float(v[j]) * c);
          l_run = fmaf(l_run, a, l_pend * c);
          m_run = m_new;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&o_empty[buf]);
      };

      for (int i = 0; i < n; ++i) {
        const int buf = (sb + i) & 1;
        const int key0 = (kb0 + i) * XT_KB + cg * 32;   // this thread's 32 keys of the block
        ptx::mbar_wait(&s_full[buf], s_phase[buf]);
        s_phase[buf] ^= 1;
        ptx::tc_fence_after();
        uint32_t v[32];
        uint32_t mbits = 0xffffffffu;
        float m_mine = -INFINITY;
        if (warp_live) {
          ptx::tmem_ld_32x16(lane_addr + S_COL[buf] + cg * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          ptx::tmem_ld_32x16(lane_addr + S_COL[buf] + cg * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          // key-padding mask of this thread's 32 keys as one ballot (no alignment assumption on the mask rows)
          const int key = key0 + lane;
          mbits = __ballot_sync(0xffffffffu, key >= p.M || (kpm != nullptr && kpm[key < p.M ? key : 0] != 0));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!((mbits >> j) & 1u)) m_mine = fmaxf(m_mine, __uint_as_float(v[j]));
        }
        // the four threads of a row agree on the block maximum; the barrier also orders "every thread has read its S
        // columns" before anybody overwrites the S buffer with the low halves of P
        xch[cg * 128 + r] = m_mine;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float m_blk = fmaxf(fmaxf(xch[r], xch[128 + r]), fmaxf(xch[256 + r], xch[384 + r]));
        asm volatile("bar.sync 2, 512;" ::: "memory");   // xch may be rewritten for the next block only after all reads
        float l_blk = 0.f;
        if (warp_live) {
          const float mc = m_blk == -INFINITY ? 0.f : m_blk * p.scale_log2e;
          const uint32_t bits = mbits | (live ? 0u : 0xffffffffu);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t ph[8], pl[8];
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) {
              const int j = hf * 16 + jj;
              const float e0 = ((bits >> j) & 1u) ? 0.f : xt_ex2(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mc));
              const float e1 = ((bits >> (j + 1)) & 1u) ? 0.f : xt_ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mc));
              l_blk += e0 + e1;
              split_pack2_bf16(e0, e1, ph[jj >> 1], pl[jj >> 1]);
            }
            ptx::tmem_st_32x8(lane_addr + P_COL[buf] + cg * 16 + hf * 8, ph);
            if (NSPLIT == 3) ptx::tmem_st_32x8(lane_addr + S_COL[buf] + cg * 16 + hf * 8, pl);   // S is fully consumed (barrier above)
          }
          ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[buf]);
        // fold the PREVIOUS block's O_b while the tensor core works on this block's products
        if (i > 0) fold(buf ^ 1);
        m_pend = m_blk;
        l_pend = l_blk;
      }
      fold((sb + n - 1) & 1);
      sb = (sb + n) & 1;

      // ---- the row sum over the four key groups
      asm volatile("bar.sync 1, 512;" ::: "memory");
      xch[cg * 128 + r] = l_run;
      asm volatile("bar.sync 2, 512;" ::: "memory");
      const float l_row = (xch[r] + xch[128 + r]) + (xch[256 + r] + xch[384 + r]);

      // ---- result of the segment: final output when this CTA covered the whole pair, else a partial for the merge
      const long orow = (static_cast<long>(img) * p.Ncap + q0 + r) * 512 + h * 64 + cg * 16;
      auto write16 = [&](const float (&o)[16], float inv) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = o[j] * inv;
        if (p.out_f32) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(p.out_f32 + orow + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        }
        if (p.out_hi) {
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_pack2_bf16(x[2 * j], x[2 * j + 1], hh[j], ll[j]);
          *reinterpret_cast<uint4*>(p.out_hi + orow) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(p.out_hi + orow + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          if (p.out_lo) {
            *reinterpret_cast<uint4*>(p.out_lo + orow) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            *reinterpret_cast<uint4*>(p.out_lo + orow + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
          }
        }
      };
      if (gs == 1) {
        if (live) write16(o_run, 1.0f / l_row);
        continue;
      }
      float* part = p.partial + ((static_cast<long>(pair) * p.max_parts + split) * 128 + r) * XT_PART;
      if (live) {
        if (cg == 0) { part[0] = m_run; part[1] = l_row; }
#pragma unroll
        for (int d = 0; d < 16; d += 2) *reinterpret_cast<float2*>(part + 2 + cg * 16 + d) = make_float2(o_run[d], o_run[d + 1]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (tid16 == 0) *last_flag = (atomicAdd(&p.counters[pair], 1) == gs - 1);
      asm volatile("bar.sync 2, 512;" ::: "memory");
      const bool last = *last_flag != 0;
      asm volatile("bar.sync 1, 512;" ::: "memory");   // everyone has read the flag before it can be rewritten
      if (!last) continue;
      __threadfence();
      if (live) {
        const float* base = p.partial + (static_cast<long>(pair) * p.max_parts * 128 + r) * XT_PART;
        const long sstride = 128L * XT_PART;
        float mm = -INFINITY;
        for (int s_ = 0; s_ < gs; ++s_) mm = fmaxf(mm, base[s_ * sstride]);
        float ltot = 0.f, acc[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = 0.f;
        for (int s_ = 0; s_ < gs; ++s_) {
          const float* ps = base + s_ * sstride;
          const float ms = ps[0];
          const float wgt = ms == -INFINITY ? 0.f : xt_ex2((ms - mm) * p.scale_log2e);
          ltot = fmaf(wgt, ps[1], ltot);
#pragma unroll
          for (int d = 0; d < 16; d += 2) {
            const float2 vv = *reinterpret_cast<const float2*>(ps + 2 + cg * 16 + d);
            acc[d] = fmaf(wgt, vv.x, acc[d]);
            acc[d + 1] = fmaf(wgt, vv.y, acc[d + 1]);
          }
        }
        write16(acc, 1.0f / ltot);
      }
      if (tid16 == 0) p.counters[pair] = 0;   // ready for the next launch (graph replay)
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

CUtensorMap xt_plane_map(Ctx* c, const bf16* base, int M, long slices) {
  auto key = std::make_tuple(static_cast<const void*>(base), -M, slices);   // negative M: distinct from xattn_tma.cu's maps
  auto it = c->xattn_tmaps.find(key);
  if (it != c->xattn_tmaps.end()) return it->second;
  if (c->xattn_tmaps.size() > 256) c->xattn_tmaps.clear();
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, ALM_ERR_INVALID, "K/V cache plane not 16-byte aligned");
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, cuuint64_t(M), cuuint64_t(slices)};
  cuuint64_t strides[2] = {128, cuuint64_t(M) * 128};
  cuuint32_t box[3] = {64, cuuint32_t(XT_KB), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled (K/V cache plane) failed with CUresult " + std::to_string(int(r))};
  return c->xattn_tmaps.emplace(key, tm).first->second;
}

}  // namespace

// one CTA per SM; the (image, head, 128-query block) x 128-key-block list is cut into equal contiguous runs
void cross_attn_tc_plan(Ctx* c, int nimg, int Ncap, int M, int* grid, int* max_parts, int* pairs) {
  const int nqb = (Ncap + 127) / 128;
  const int np = nimg * 8 * nqb;
  const int nkb = (M + XT_KB - 1) / XT_KB;
  const long NB = static_cast<long>(np) * nkb;
  const int G = static_cast<int>(std::min<long>(NB, c->num_sms));
  const int bpc = static_cast<int>(NB / G);
  *grid = G;
  *max_parts = (nkb + bpc - 1) / bpc + 1;
  *pairs = np;
}

// kc_* / vc_* are the BASES of the whole cache ([nimg_total*96][M][64]); z0 = (first image) * 96 + (decoder-layer) * 8
void cross_attn_tc(Ctx* c, const bf16* q_hi, const bf16* q_lo, const float* q_f32, int nimg, int Ncap, const bf16* kc_hi,
                   const bf16* kc_lo, const bf16* vc_hi, const bf16* vc_lo, long slices, int z0, const uint8_t* kpm, int M,
                   int grid, int max_parts, float* partial, int* counters, bf16* out_hi, bf16* out_lo, float* out_f32) {
  if (c->skipped(1)) return;
  ALM_REQUIRE(q_f32 || q_hi, ALM_ERR_INVALID, "cross_attn_tc: no query operand");
  const bool three = c->nsplit == 3 && (q_f32 || q_lo) && kc_lo && vc_lo;
  XtParams p;
  p.Ncap = Ncap; p.M = M;
  p.nkb = (M + XT_KB - 1) / XT_KB;
  p.nqb = (Ncap + 127) / 128;
  p.npairs = nimg * 8 * p.nqb;
  p.max_parts = max_parts;
  p.z0 = z0;
  p.q_hi = q_hi; p.q_lo = q_lo; p.q_f32 = q_f32; p.kpm = kpm;
  p.partial = partial; p.counters = counters;
  p.out_hi = out_hi; p.out_lo = out_lo; p.out_f32 = out_f32;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  const CUtensorMap tkh = xt_plane_map(c, kc_hi, M, slices), tvh = xt_plane_map(c, vc_hi, M, slices);
  const CUtensorMap tkl = three ? xt_plane_map(c, kc_lo, M, slices) : tkh, tvl = three ? xt_plane_map(c, vc_lo, M, slices) : tvh;
  static DeviceOnce attr;
  if (attr.need()) {
    ALM_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, XtSmem<3>::kTotal));
    ALM_CHECK_CUDA(cudaFuncSetAttribute(cross_attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, XtSmem<1>::kTotal));
    attr.mark();
  }
  if (three) cross_attn_tc_kernel<3><<<grid, XT_THREADS, XtSmem<3>::kTotal, c->stream>>>(tkh, tkl, tvh, tvl, p);
  else cross_attn_tc_kernel<1><<<grid, XT_THREADS, XtSmem<1>::kTotal, c->stream>>>(tkh, tkl, tvh, tvl, p);
  count_launch(c);
  check_launch("cross_attn_tc");
}

}  // namespace alm
