// Swin window attention (W-MSA core, swin_transformer.py:127-148) on the 5th-gen tensor cores.
//
//     S = (q*scale) k^T            tcgen05.mma: TWO 7x7 windows packed into one M = 128 tile (98 live rows), operands are
//                                  TMA tiles of the split-bf16 qkv planes (64 channels = two 32-wide heads = one 128-byte
//                                  swizzled row), accumulator in tensor memory
//     + bias + (-100) shift mask   from TMEM in registers: thread = token row; only the 49 keys of the row's own window
//     softmax, P -> TMEM (bf16)    tcgen05.st: the probabilities are the A operand of the second MMA, straight from TMEM
//     O = P v                      tcgen05.mma with V as an MN-major shared-memory operand; cross-window blocks of P are 0
//
// Pad tokens are live keys: they are ordinary rows of the qkv planes (the gather wrote LN-free zeros, so their
// k = b_k, v = b_v).  The relative-position bias comes from the dense [heads,49,49] table; the shift mask is recomputed
// from the window position (swin_transformer.py:368-387).  q arrives pre-scaled (the scale is folded into the q rows of
// the qkv weight at load time).
//
// One persistent CTA per SM; item = (window pair, head pair).  Warp 0 = TMA producer over a 2-stage ring (q, k, v tiles of
// hi and lo planes: 6 x 98 rows x 128 B per item, no over-read), warp 1 = MMA issuer, warps 2..5 / 6..9 = softmax + epilogue of
// head 0 / head 1 of the pair (thread = token row).
// TMEM (512 columns): S of head 0 / head 1 at 0 / 128 (112 used each; the low halves of P overwrite them in place),
// P_hi at 256 / 320, O at 384 / 448 (each P.V runs 64 wide over both heads' channels; the epilogue keeps its own half).
#include <algorithm>

#include "alm_internal.h"
#include "ptx.cuh"

namespace alm {

namespace {

constexpr int WT_THREADS = 320;                // warp 0 TMA, warp 1 MMA, warps 2..5 softmax of head 0, warps 6..9 of head 1
constexpr int WT_ROWS = 98;                 // two windows
constexpr int WT_NK = 112;                  // keys padded to the UMMA N granularity
constexpr int WT_TILE = 128 * 128;          // bytes of one operand tile in shared memory (128 rows x 64 bf16)
constexpr int WT_BOXB = WT_ROWS * 128;      // bytes one TMA box delivers
constexpr int WT_STAGES = 2;

template <int NSPLIT>
struct WtSmem {
  static constexpr int NP = NSPLIT == 3 ? 2 : 1;
  static constexpr int kStage = 3 * NP * WT_TILE;       // q, k, v of every plane
  static constexpr int kBar = WT_STAGES * kStage;
  static constexpr int kReg = kBar + 128;               // region ids of the 98 tokens of the tile being soft-maxed
  static constexpr int kBias = kReg + 128 * 4;          // dense relative-position bias of the current head pair [2][49*49]
  static constexpr int kTotal = kBias + 2 * 2404 * 4 + 1024;  // + alignment slack
};

struct WtParams {
  int C, heads;        // channels, heads (head_dim 32)
  int nWh, nWw;        // windows per image
  long n_win;          // windows in the batch
  long rows;           // n_win * 49
  int shift, Hp, Wp;
  long n_pairs;        // ceil(n_win / 2)
  int n_hp;            // heads / 2
  long items;          // n_pairs * n_hp
  const float* bias;   // [heads, 49, 49]
  bf16* out_hi;
  bf16* out_lo;
  float* out_f32;
};

template <int NSPLIT>
__global__ void __launch_bounds__(WT_THREADS, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                           const WtParams p) {
  using L = WtSmem<NSPLIT>;
  constexpr int NP = L::NP;
  extern __shared__ uint8_t wt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wt_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* in_full = bars + 0;    // [2]
  uint64_t* in_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  int* sreg = reinterpret_cast<int*>(smem + L::kReg);
  float* sbias = reinterpret_cast<float*>(smem + L::kBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t S_COL[2] = {0, 128}, P_COL[2] = {256, 320}, O_COL[2] = {384, 448};

  // rows 98..127 of every operand tile are never written by TMA: zero them once (they feed discarded rows / masked keys,
  // but a stale NaN pattern in V would poison 0 * v)
  for (int i = threadIdx.x; i < WT_STAGES * 3 * NP * (30 * 128 / 16); i += WT_THREADS) {
    const int tile = i / (30 * 8), off = i % (30 * 8);
    *reinterpret_cast<uint4*>(smem + tile * WT_TILE + WT_BOXB + off * 16) = make_uint4(0, 0, 0, 0);
  }
  ptx::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hi);
    if (NSPLIT == 3) ptx::prefetch_tmap(&tm_lo);
    for (int s = 0; s < WT_STAGES; ++s) {
      ptx::mbar_init(&in_full[s], 1);
      ptx::mbar_init(&in_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 8);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 8);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // contiguous run of items per CTA, head-pair-major so that a CTA stays on one pair of bias tables for long
  const long it_begin = static_cast<long>(blockIdx.x) * p.items / gridDim.x;
  const long it_end = (static_cast<long>(blockIdx.x) + 1) * p.items / gridDim.x;

  if (warp == 0) {
    // ================================================================================= TMA producer
    int st = 0;
    uint32_t phase = 0;
    for (long it = it_begin; it < it_end; ++it) {
      const int hp = static_cast<int>(it / p.n_pairs);
      const long wp = it % p.n_pairs;
      const int row0 = static_cast<int>(wp * WT_ROWS);
      ptx::mbar_wait(&in_empty[st], phase ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&in_full[st], 3 * NP * WT_BOXB);
        uint8_t* base = smem + st * L::kStage;
        for (int pl = 0; pl < NP; ++pl) {
          const void* tm = pl ? static_cast<const void*>(&tm_lo) : static_cast<const void*>(&tm_hi);
          for (int which = 0; which < 3; ++which)   // q | k | v column blocks of the packed projection
            ptx::tma_load_2d(base + (which * NP + pl) * WT_TILE, tm, &in_full[st], which * p.C + hp * 64, row0);
        }
      }
      __syncwarp();
      if (++st == WT_STAGES) { st = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_s = idesc_base | (uint32_t(WT_NK >> 3) << 17);
    const uint32_t idesc_pv = idesc_base | (1u << 16) | (uint32_t(64 >> 3) << 17);   // B (= V) is MN-major
    int st = 0;
    uint32_t phase = 0, p_phase = 0, oe_phase = 0;
    for (long it = it_begin; it < it_end; ++it) {
      ptx::mbar_wait(&in_full[st], phase);
      ptx::tc_fence_after();
      const uint32_t base = ptx::smem_u32(smem + st * L::kStage);
      const uint32_t q_t = base, k_t = base + NP * WT_TILE, v_t = base + 2 * NP * WT_TILE;
      if (ptx::elect_one()) {
#pragma unroll
        for (int hd = 0; hd < 2; ++hd)
#pragma unroll
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint32_t qa = q_t + (pass == 1 ? WT_TILE : 0) + hd * 64;
            const uint32_t kb = k_t + (pass == 2 ? WT_TILE : 0) + hd * 64;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              ptx::umma_bf16(tmem + S_COL[hd], ptx::make_kmajor_sw128_desc(qa + k * 32),
                             ptx::make_kmajor_sw128_desc(kb + k * 32), idesc_s, (pass | k) != 0);
          }
        ptx::umma_commit(s_full);
      }
      __syncwarp();
      ptx::mbar_wait(p_full, p_phase);
      p_phase ^= 1;
      ptx::mbar_wait(o_empty, oe_phase ^ 1);
      oe_phase ^= 1;
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int hd = 0; hd < 2; ++hd)
#pragma unroll
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint32_t pa = tmem + (pass == 1 ? S_COL[hd] : P_COL[hd]);   // P_lo lives where S was
            const uint32_t vb = v_t + (pass == 2 ? WT_TILE : 0);
#pragma unroll
            for (int j = 0; j < WT_NK / 16; ++j)
              ptx::umma_bf16_ts(tmem + O_COL[hd], pa + j * 8, ptx::make_kmajor_sw128_desc(vb + j * 2048), idesc_pv,
                                (pass | j) != 0);
          }
        ptx::umma_commit(o_full);
        ptx::umma_commit(&in_empty[st]);   // q, k, v of this stage are free once everything above has retired
      }
      __syncwarp();
      if (++st == WT_STAGES) { st = 0; phase ^= 1; }
    }
  } else {
    // ================================================================================= softmax + epilogue warps
    // two warp groups, one per head of the pair: both heads' softmaxes run concurrently, two warps per scheduler
    const int quarter = warp & 3, hd = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    const uint32_t s_col = hd * 128, p_col = 256 + hd * 64, o_col = 384 + hd * 64;   // == S_COL / P_COL / O_COL [hd]
    const int r = quarter * 32 + lane;          // token row of the tile
    const int w = r >= 49 ? 1 : 0;              // which of the two windows
    const int ti = r - 49 * w;                  // token index inside the window (>= 49: idle row)
    const bool live = r < WT_ROWS;
    // 16-column chunks this warp must look at: windows 0 / 1 own key columns [0,49) / [49,98)
    const int c_lo = quarter <= 1 ? 0 : 3, c_hi = quarter == 0 ? 4 : 7;
    const int key0 = 49 * w;
    uint32_t s_phase = 0, o_phase = 0;
    int cur_hp = -1;
    const int wins_per_img = p.nWh * p.nWw;
    const uint32_t sreg_s = ptx::smem_u32(sreg), sbias_s = ptx::smem_u32(sbias);
    constexpr float kLog2e = 1.4426950408889634f;
    for (long it = it_begin; it < it_end; ++it) {
      const int hp = static_cast<int>(it / p.n_pairs);
      const long wp = it % p.n_pairs;
      // region ids of the shift mask (0 when the block is not shifted)
      int reg = 0;
      if (p.shift > 0 && live) {
        const int wi = static_cast<int>((2 * wp + w) % wins_per_img);
        const int hh = (wi / p.nWw) * 7 + ti / 7, ww = (wi % p.nWw) * 7 + ti % 7;
        const int rh = hh < p.Hp - 7 ? 0 : (hh < p.Hp - p.shift ? 1 : 2);
        const int rw = ww < p.Wp - 7 ? 0 : (ww < p.Wp - p.shift ? 1 : 2);
        reg = rh * 3 + rw;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // previous item's readers of sreg / sbias are done
      if (hd == 0) sreg[r] = reg;
      if (hp != cur_hp) {   // the CTA's run is head-pair-major: the two 49 x 49 bias tables change rarely.  Read per
        cur_hp = hp;        // token ROW from shared memory the stride (49 words) is odd, i.e. bank-conflict free (the same
                            // reads straight from global memory cost 32 L1 wavefronts per instruction).  Stored x log2(e):
                            // the softmax below works in base 2.
        const float* src = p.bias + static_cast<long>(2 * hp) * 2401;
        for (int i = threadIdx.x - 64; i < 2 * 2401; i += 256) sbias[(i / 2401) * 2404 + i % 2401] = __ldg(src + i) * kLog2e;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ptx::mbar_wait(s_full, s_phase);
      s_phase ^= 1;
      ptx::tc_fence_after();
      const uint32_t bias_row = sbias_s + 4u * static_cast<uint32_t>(hd * 2404 + (live ? ti : 0) * 49 - key0);
      const bool shifted = p.shift > 0;
      // pass 1: row maximum of x = (s + bias + mask) log2(e) over the 49 keys of the row's window
      float m = -INFINITY;
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(lane_addr + s_col + c * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c * 16 + j;
          if (live && col >= key0 && col < key0 + 49) {
            float x = fmaf(__uint_as_float(v[j]), kLog2e, ptx::lds_f32(bias_row + 4u * col));
            if (shifted && ptx::lds_s32(sreg_s + 4u * col) != reg) x += -100.0f * kLog2e;
            m = fmaxf(m, x);
          }
        }
      }
      // pass 2: p = 2^(x - m), row sum, bf16 (hi, lo) pairs back into tensor memory; keys outside the window get 0
      float l = 0.f;
      for (int c = 0; c < WT_NK / 16; ++c) {
        uint32_t ph[8], pl[8];
        if (c >= c_lo && c < c_hi) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(lane_addr + s_col + c * 16, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float e[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int col = c * 16 + j + q;
              e[q] = 0.f;
              if (live && col >= key0 && col < key0 + 49) {
                float x = fmaf(__uint_as_float(v[j + q]), kLog2e, ptx::lds_f32(bias_row + 4u * col));
                if (shifted && ptx::lds_s32(sreg_s + 4u * col) != reg) x += -100.0f * kLog2e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[q]) : "f"(x - m));
              }
            }
            l += e[0] + e[1];
            split_pack2_bf16(e[0], e[1], ph[j >> 1], pl[j >> 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) ph[j] = pl[j] = 0u;
        }
        ptx::tmem_st_32x8(lane_addr + p_col + c * 8, ph);
        if (NSPLIT == 3) ptx::tmem_st_32x8(lane_addr + s_col + c * 8, pl);  // behind this thread's read pointer
      }
      const float inv_l = live ? 1.0f / l : 0.f;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      // ---- epilogue: O / l, this head's 32 channels of the output planes
      ptx::mbar_wait(o_full, o_phase);
      o_phase ^= 1;
      ptx::tc_fence_after();
      const long grow = wp * WT_ROWS + r;
      const bool store = live && grow < p.rows;
      const long obase = grow * p.C + hp * 64 + hd * 32;
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(lane_addr + o_col + hd * 32 + c0, v);   // head hd's own channels of the 64-wide product
        ptx::tmem_ld_wait();
        if (store) {
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float x0 = __uint_as_float(v[j]) * inv_l, x1 = __uint_as_float(v[j + 1]) * inv_l;
            split_pack2_bf16(x0, x1, hh[j >> 1], ll[j >> 1]);
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + obase + c0 + j) = make_float2(x0, x1);
          }
          const long o = obase + c0;
          if (p.out_hi) {
            *reinterpret_cast<uint4*>(p.out_hi + o) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(p.out_hi + o + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          }
          if (p.out_lo) {
            *reinterpret_cast<uint4*>(p.out_lo + o) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            *reinterpret_cast<uint4*>(p.out_lo + o + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

CUtensorMap wqkv_map(Ctx* c, const bf16* base, long rows, long ld) {
  CUtensorMap tm;
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, ALM_ERR_INVALID, "window attention operand alignment");
  cuuint64_t dims[2] = {cuuint64_t(ld), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(WT_ROWS)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled (window attention qkv plane) failed with CUresult " + std::to_string(int(r))};
  return tm;
}

}  // namespace

// qkv: window-major [n_win*49, 3C] split planes, q rows pre-scaled; out: [n_win*49, C].  heads must be even.
void window_attention_tc(Ctx* c, const bf16* qkv_hi, const bf16* qkv_lo, int C, int heads, int nWh, int nWw, int B, int shift,
                         int Hp, int Wp, const float* bias_dense, bf16* out_hi, bf16* out_lo, float* out_f32) {
  ALM_REQUIRE(C == heads * 32 && heads % 2 == 0, ALM_ERR_UNSUPPORTED, "window_attention_tc: head_dim 32, even head count");
  const bool three = c->nsplit == 3;
  ALM_REQUIRE(!three || qkv_lo, ALM_ERR_INVALID, "window_attention_tc: split mode needs the lo plane");
  WtParams p;
  p.C = C; p.heads = heads; p.nWh = nWh; p.nWw = nWw;
  p.n_win = static_cast<long>(B) * nWh * nWw;
  p.rows = p.n_win * 49;
  p.shift = shift; p.Hp = Hp; p.Wp = Wp;
  p.n_pairs = (p.n_win + 1) / 2;
  p.n_hp = heads / 2;
  p.items = p.n_pairs * p.n_hp;
  p.bias = bias_dense;
  p.out_hi = out_hi; p.out_lo = out_lo; p.out_f32 = out_f32;
  const CUtensorMap th = wqkv_map(c, qkv_hi, p.rows, 3L * C);
  const CUtensorMap tl = three ? wqkv_map(c, qkv_lo, p.rows, 3L * C) : th;
  const int grid = static_cast<int>(std::min<long>(p.items, c->num_sms));
  static DeviceOnce attr;
  if (attr.need()) {
    ALM_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WtSmem<3>::kTotal));
    ALM_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WtSmem<1>::kTotal));
    attr.mark();
  }
  if (three) window_attention_tc_kernel<3><<<grid, WT_THREADS, WtSmem<3>::kTotal, c->stream>>>(th, tl, p);
  else window_attention_tc_kernel<1><<<grid, WT_THREADS, WtSmem<1>::kTotal, c->stream>>>(th, tl, p);
  count_launch(c);
  check_launch("window_attention_tc");
}

}  // namespace alm
