// Fused dense self-attention of the ViT blocks (MGP-STR: 257 tokens x 64-wide heads) on the 5th-gen tensor cores:
//
//     S = Q K^T          tcgen05.mma, operands staged by TMA into 128B-swizzled shared memory, accumulator in TMEM
//     P = softmax(S*c)   read from TMEM (tcgen05.ld), exp2 in registers, written BACK into TMEM (tcgen05.st) as bf16
//     O = P V            tcgen05.mma with the A operand in TMEM and V as an MN-major shared-memory operand
//
// Neither the scores nor the probabilities ever exist in HBM (the unfused path wrote ~5 GB of them per layer at B = 512).
// Reference arithmetic: timm-0.4.12 Attention (timm/models/vision_transformer.py, class Attention; called from
// OCR/MGP-STR/modules/mgp_str.py:73-74): attn = (q @ k^T) * head_dim^-0.5; softmax; attn @ v.
//
// One persistent CTA per SM walks the (crop, head) items.  Per item the keys / values of the head (all T <= 272 of them)
// are loaded once, then the ceil(T/128) query tiles run through
//     warp 0      TMA producer (one elected lane)
//     warp 1      tcgen05.mma issuer (one elected lane)
//     warps 2..9  softmax + output epilogue: two threads per query row (lane quarter = warp % 4, key half = (warp-2)/4)
// with mbarrier hand-offs: S(i+1) is issued right behind P(i) V, so it runs while the epilogue of tile i drains O.
// TMEM map (512 columns): S [0, TK) fp32; P_hi [TK, TK + TK/2) bf16 pairs; O [TK + TK/2, +64) fp32; in split mode the low
// halves P_lo overwrite S in place, behind the read pointer of the (single) thread that owns the row.
//
// NSPLIT = 3 carries q, k, v and p as (hi, lo) bf16 pairs and issues hi*hi + lo*hi + hi*lo into the same accumulator
// (fp32-class); NSPLIT = 1 is single-pass bf16.
#include <algorithm>

#include "alm_internal.h"
#include "ptx.cuh"

namespace alm {

namespace {

constexpr int AT_THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2..9 softmax: lane quarter = w % 4, key half = (w - 2) / 4
constexpr int AT_TKMAX = 272;             // keys padded to a multiple of 16 (UMMA N granularity at M = 128)
constexpr int AT_BOX = 136;               // rows of one TMA box: two boxes cover the keys, one covers a 128-row Q tile
constexpr int AT_BOXB = AT_BOX * 128;     // 17 408 bytes, a multiple of the 1024-byte swizzle atom
constexpr int AT_KVB = 2 * AT_BOXB;       // one K or V plane: 272 rows of 128 bytes
constexpr int AT_QSTAGES = 2;

template <int NSPLIT>
struct AtSmem {
  static constexpr int NP = NSPLIT == 3 ? 2 : 1;  // planes
  static constexpr int kK = 0;
  static constexpr int kV = kK + NP * AT_KVB;
  static constexpr int kQ = kV + NP * AT_KVB;
  static constexpr int kBar = kQ + AT_QSTAGES * NP * AT_BOXB;
  static constexpr int kXch = kBar + 256;           // per-row exchange between the two key halves: [2][128] floats
  static constexpr int kTotal = kXch + 1024 + 1024;  // + alignment slack
};

struct AtParams {
  int B, T, H, D;          // crops, tokens per crop, heads, model width (H * 64)
  int TK;                  // keys padded to 16
  int tiles;               // ceil(T / 128)
  long items;              // B * H
  float scale_log2e;       // head_dim^-0.5 * log2(e)
  bf16* out_hi;
  bf16* out_lo;
  float* out_f32;
  long ldo;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NSPLIT>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const AtParams p) {
  using L = AtSmem<NSPLIT>;
  constexpr int NP = L::NP;
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;    // [2]
  uint64_t* q_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TK = p.TK;
  const uint32_t S_COL = 0, P_COL = TK, O_COL = TK + TK / 2, PLO_COL = 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hi);
    if (NSPLIT == 3) ptx::prefetch_tmap(&tm_lo);
    ptx::mbar_init(kv_full, 1);
    ptx::mbar_init(kv_empty, 1);
    for (int s = 0; s < AT_QSTAGES; ++s) {
      ptx::mbar_init(&q_full[s], 1);
      ptx::mbar_init(&q_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 8);   // one arrival per softmax warp
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 8);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================================================================= TMA producer
    uint32_t kv_phase = 0, q_phase = 0;
    int qs = 0;
    for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const int b = static_cast<int>(it / p.H), h = static_cast<int>(it % p.H);
      const int row0 = b * p.T;
      ptx::mbar_wait(kv_empty, kv_phase ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(kv_full, 4 * NP * AT_BOXB);
        for (int pl = 0; pl < NP; ++pl) {
          const void* tm = pl ? static_cast<const void*>(&tm_lo) : static_cast<const void*>(&tm_hi);
          for (int half = 0; half < 2; ++half) {
            ptx::tma_load_2d(smem + L::kK + pl * AT_KVB + half * AT_BOXB, tm, kv_full, p.D + h * 64, row0 + half * AT_BOX);
            ptx::tma_load_2d(smem + L::kV + pl * AT_KVB + half * AT_BOXB, tm, kv_full, 2 * p.D + h * 64, row0 + half * AT_BOX);
          }
        }
      }
      __syncwarp();
      kv_phase ^= 1;
      for (int i = 0; i < p.tiles; ++i) {
        ptx::mbar_wait(&q_empty[qs], q_phase ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&q_full[qs], NP * AT_BOXB);
          for (int pl = 0; pl < NP; ++pl)
            ptx::tma_load_2d(smem + L::kQ + (qs * NP + pl) * AT_BOXB, pl ? static_cast<const void*>(&tm_lo) : static_cast<const void*>(&tm_hi),
                             &q_full[qs], h * 64, row0 + i * 128);
        }
        __syncwarp();
        if (++qs == AT_QSTAGES) { qs = 0; q_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    // instruction descriptors: D = f32, A = B = bf16; N >> 3 at [17,23), M >> 4 at [24,29); bit 16 = B is MN-major
    const int n1 = TK > 256 ? 256 : TK, n2 = TK - n1;
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_s1 = idesc_base | (uint32_t(n1 >> 3) << 17);
    const uint32_t idesc_s2 = idesc_base | (uint32_t(n2 >> 3) << 17);
    const uint32_t idesc_pv = idesc_base | (1u << 16) | (uint32_t(64 >> 3) << 17);
    uint32_t kv_phase = 0, q_phase = 0, p_phase = 0, oe_phase = 0;
    int qs = 0;
    const uint32_t k_base = ptx::smem_u32(smem + L::kK), v_base = ptx::smem_u32(smem + L::kV);
    for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
      ptx::mbar_wait_hot(kv_full, kv_phase);
      kv_phase ^= 1;
      for (int i = 0; i < p.tiles; ++i) {
        ptx::mbar_wait_hot(&q_full[qs], q_phase);
        ptx::tc_fence_after();
        const uint32_t q_base = ptx::smem_u32(smem + L::kQ + qs * NP * AT_BOXB);
        if (ptx::elect_one()) {
          // ---- S = Q K^T : passes (q_hi, k_hi), (q_lo, k_hi), (q_hi, k_lo)
#pragma unroll
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint32_t qa = q_base + (pass == 1 ? AT_BOXB : 0);
            const uint32_t kb = k_base + (pass == 2 ? AT_KVB : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = ptx::make_kmajor_sw128_desc(qa + k * 32);
              ptx::umma_bf16(tmem + S_COL, da, ptx::make_kmajor_sw128_desc(kb + k * 32), idesc_s1, (pass | k) != 0);
              if (n2 > 0)
                ptx::umma_bf16(tmem + S_COL + n1, da, ptx::make_kmajor_sw128_desc(kb + n1 * 128 + k * 32), idesc_s2,
                               (pass | k) != 0);
            }
          }
          ptx::umma_commit(&q_empty[qs]);
          ptx::umma_commit(s_full);
        }
        __syncwarp();
        if (++qs == AT_QSTAGES) { qs = 0; q_phase ^= 1; }
        // ---- O = P V once the softmax warps have written P (and the previous O has been drained)
        ptx::mbar_wait_hot(p_full, p_phase);
        p_phase ^= 1;
        ptx::mbar_wait_hot(o_empty, oe_phase ^ 1);
        oe_phase ^= 1;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const int ksteps = TK / 16;
          const int n0 = (ksteps + 1) / 2;   // key chunks of the first softmax half (see the softmax warps)
          for (int pass = 0; pass < NSPLIT; ++pass) {
            const uint32_t vb = v_base + (pass == 2 ? AT_KVB : 0);
            for (int j = 0; j < ksteps; ++j) {
              // hi halves of P: their own region, 8 columns per 16 keys.  lo halves: in place over S, each softmax half
              // inside the S columns it owns (half 0 at column 8 j, half 1 at 16 n0 + 8 (j - n0))
              const uint32_t pa = pass == 1 ? tmem + PLO_COL + (j < n0 ? j * 8 : 16 * n0 + (j - n0) * 8) : tmem + P_COL + j * 8;
              ptx::umma_bf16_ts(tmem + O_COL, pa, ptx::make_kmajor_sw128_desc(vb + j * 2048), idesc_pv, (pass | j) != 0);
            }
          }
          ptx::umma_commit(o_full);
          if (i == p.tiles - 1) ptx::umma_commit(kv_empty);  // K / V of this item are free once these MMAs retire
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================================= softmax + epilogue warps
    // TWO threads per query row (the per-tile softmax is the longest link of the tile's dependency chain): warps w and
    // w + 4 share a lane quarter; half 0 owns the first n0 16-key chunks of the row, half 1 the rest; they agree on the row
    // maximum and the row sum through shared memory; the epilogue splits the 64 output dims the same way.
    const int quarter = warp & 3, hf = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem + L::kXch);   // [2][128]
    const int nch = TK / 16, n0 = (nch + 1) / 2;
    const int c_lo = hf ? n0 : 0, c_hi = hf ? nch : n0;
    const uint32_t plo_base = PLO_COL + (hf ? 16 * n0 : 0);  // this half's in-place region for the lo halves of P
    uint32_t s_phase = 0, o_phase = 0;
    for (long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const int b = static_cast<int>(it / p.H), h = static_cast<int>(it % p.H);
      for (int i = 0; i < p.tiles; ++i) {
        const int r = quarter * 32 + lane;
        const int row = i * 128 + r;
        const bool warp_live = i * 128 + quarter * 32 < p.T;
        ptx::mbar_wait_hot(s_full, s_phase);
        s_phase ^= 1;
        ptx::tc_fence_after();
        // pass 1: maximum over this half's live keys.  Only the chunk that contains key T needs per-key predicates; the
        // chunks in front of it run a straight-line max tree (the predicated form was ~13 instructions per key)
        float m = -INFINITY;
        if (warp_live) {
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t v[16];
            ptx::tmem_ld_32x16(lane_addr + S_COL + c * 16, v);
            ptx::tmem_ld_wait();
            if ((c + 1) * 16 <= p.T) {
              float a0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), a1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
              for (int j = 4; j < 16; j += 4) {
                a0 = fmaxf(a0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                a1 = fmaxf(a1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
              }
              m = fmaxf(m, fmaxf(a0, a1));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c * 16 + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
            }
          }
        }
        xch[hf * 128 + r] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(xch[r], xch[128 + r]);
        asm volatile("bar.sync 2, 256;" ::: "memory");   // both halves have read the maxima before xch is reused
        // pass 2: p = 2^((s - m) * c); partial row sum; bf16 (hi, lo) pairs back into tensor memory
        float l = 0.f;
        if (warp_live) {
          const float sl2 = p.scale_log2e;
          const float mc = m * sl2;
          float l0 = 0.f, l1 = 0.f;
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t v[16], ph[8], pl[8];
            ptx::tmem_ld_32x16(lane_addr + S_COL + c * 16, v);
            ptx::tmem_ld_wait();
            if ((c + 1) * 16 <= p.T) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -mc));
                const float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -mc));
                l0 += e0;
                l1 += e1;
                split_pack2_bf16(e0, e1, ph[j >> 1], pl[j >> 1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float e0 = c * 16 + j < p.T ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -mc)) : 0.f;
                const float e1 = c * 16 + j + 1 < p.T ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -mc)) : 0.f;
                l0 += e0;
                l1 += e1;
                split_pack2_bf16(e0, e1, ph[j >> 1], pl[j >> 1]);
              }
            }
            ptx::tmem_st_32x8(lane_addr + P_COL + c * 8, ph);
            if (NSPLIT == 3) ptx::tmem_st_32x8(lane_addr + plo_base + (c - c_lo) * 8, pl);  // behind this thread's read pointer
          }
          l = l0 + l1;
          ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(p_full);
        // row sum over both halves
        xch[hf * 128 + r] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l = xch[r] + xch[128 + r];
        asm volatile("bar.sync 2, 256;" ::: "memory");
        // ---- epilogue: O / l -> split bf16 planes (operand of the projection GEMM), token-major [B*T, D]; this half
        //      writes output dims [32 hf, 32 hf + 32)
        ptx::mbar_wait_hot(o_full, o_phase);
        o_phase ^= 1;
        ptx::tc_fence_after();
        if (warp_live) {
          const float inv = 1.0f / l;
          const long orow = (static_cast<long>(b) * p.T + row) * p.ldo + h * 64 + hf * 32;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x16(lane_addr + O_COL + hf * 32 + c0, v);
            ptx::tmem_ld_wait();
            if (row < p.T) {
              uint32_t hh[8], ll[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float x0 = __uint_as_float(v[j]) * inv, x1 = __uint_as_float(v[j + 1]) * inv;
                split_pack2_bf16(x0, x1, hh[j >> 1], ll[j >> 1]);
                if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + orow + c0 + j) = make_float2(x0, x1);
              }
              if (p.out_hi) {
                *reinterpret_cast<uint4*>(p.out_hi + orow + c0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                *reinterpret_cast<uint4*>(p.out_hi + orow + c0 + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
              }
              if (p.out_lo) {
                *reinterpret_cast<uint4*>(p.out_lo + orow + c0) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                *reinterpret_cast<uint4*>(p.out_lo + orow + c0 + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_empty);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

CUtensorMap qkv_map(Ctx* c, const bf16* base, long rows, long ld) {
  CUtensorMap tm;
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, ALM_ERR_INVALID, "attention operand alignment");
  cuuint64_t dims[2] = {cuuint64_t(ld), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(AT_BOX)};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = c->encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw AlmError{ALM_ERR_CUDA, "cuTensorMapEncodeTiled (attention qkv plane) failed with CUresult " + std::to_string(int(r))};
  return tm;
}

}  // namespace

// qkv: token-major [B*T, 3*D] split bf16 planes (q | k | v blocks of D columns, head h at columns h*64 of its block), as
// the fused qkv GEMM writes them.  out: [B*T, D] split planes (and / or fp32).  T <= 272, head width 64.
void attention_tc(Ctx* c, const bf16* qkv_hi, const bf16* qkv_lo, long ld, int B, int T, int H, bf16* out_hi, bf16* out_lo,
                  float* out_f32, long ldo) {
  ALM_REQUIRE(T >= 1 && T <= AT_TKMAX && H >= 1 && B >= 1, ALM_ERR_UNSUPPORTED, "attention_tc: at most 272 tokens per sequence");
  ALM_REQUIRE(ld >= 3L * H * 64 && ldo % 8 == 0, ALM_ERR_INVALID, "attention_tc: leading dimensions");
  const bool three = c->nsplit == 3;
  ALM_REQUIRE(!three || qkv_lo, ALM_ERR_INVALID, "attention_tc: split mode needs the lo plane");
  AtParams p;
  p.B = B; p.T = T; p.H = H; p.D = H * 64;
  p.TK = (T + 15) & ~15;
  if (p.TK < 32) p.TK = 32;
  p.tiles = (T + 127) / 128;
  p.items = static_cast<long>(B) * H;
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  p.out_hi = out_hi; p.out_lo = three ? out_lo : nullptr; p.out_f32 = out_f32; p.ldo = ldo;
  const long rows = static_cast<long>(B) * T;
  const CUtensorMap th = qkv_map(c, qkv_hi, rows, ld);
  const CUtensorMap tl = three ? qkv_map(c, qkv_lo, rows, ld) : th;
  const int grid = static_cast<int>(std::min<long>(p.items, c->num_sms));
  static DeviceOnce attr;
  if (attr.need()) {
    ALM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem<3>::kTotal));
    ALM_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem<1>::kTotal));
    attr.mark();
  }
  if (three) attention_tc_kernel<3><<<grid, AT_THREADS, AtSmem<3>::kTotal, c->stream>>>(th, tl, p);
  else attention_tc_kernel<1><<<grid, AT_THREADS, AtSmem<1>::kTotal, c->stream>>>(th, tl, p);
  count_launch(c);
  check_launch("attention_tc");
}

}  // namespace alm
