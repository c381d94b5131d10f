// Swin W-MSA core, persistent TMA-fed variant of the warp-level tensor-core kernel (wattn_impl 3).
//
// Same arithmetic as window_attention_split_kernel (kernels.cu: mma.sync.m16n8k16 with the three-term bf16 split, fp32
// softmax in registers; swin_transformer.py:127-148) -- what changes is how the data moves:
//   * item = (window, HEAD PAIR): the q / k / v slices of two heads are one 128-byte row segment per token and plane, so
//     every DRAM access is a full 128-byte line (the per-head kernel read 64-byte segments and re-fetched the other half of
//     most lines: ncu 650 MB read for 481 MB of operands);
//   * one persistent CTA per SM, bound to one head pair: its relative-position bias (2 x 49 x 49 fp32) is staged in shared
//     memory once instead of 32 scattered L1 reads per thread and item;
//   * a dedicated producer warp streams the items through a 3-stage TMA + mbarrier ring (six 49-row x 128-byte boxes per
//     item, 128-byte swizzle, rows 49..63 of every tile stay zero), so the loads of the next two windows are in flight
//     while eight consumer warps (head = warp / 4, 16-query band = warp % 4) run the current one; no CTA-wide barrier in
//     the loop;
//   * the softmax runs in base 2 (scores carried as x log2 e, the staged bias pre-multiplied: one FFMA + MUFU.EX2 per
//     probability), the probabilities are normalised with one reciprocal per row and converted with packed cvt.rn.bf16x2
//     (the XU pipe -- accurate expf, per-element divisions, scalar conversions -- was the busiest unit of the per-head
//     kernel at 42 %).
// Measured (ncu, Swin stage-2 launch of a 16-page batch): 229 us against 310 us, DRAM read 482 MB = algorithmic.
#include <algorithm>

#include "alm_internal.h"
#include "mma.cuh"
#include "ptx.cuh"

namespace alm {

namespace {

constexpr int WM_WT = 49;                    // tokens per window
constexpr int WM_CWARPS = 8;                 // consumer warps
constexpr int WM_THREADS = 32 * (WM_CWARPS + 1);
constexpr int WM_STAGES = 3;
constexpr int WM_TILE = 64 * 128;            // one (plane, q|k|v) tile: 64 rows x 128 bytes
constexpr int WM_STAGE = 6 * WM_TILE;        // hi q,k,v then lo q,k,v
constexpr int WM_BIASP = 56;                 // floats per staged bias row: 8 rows x 8 banks per half-warp, conflict-free float2
constexpr int WM_OUTP = 36;                  // floats per row of a warp's 16 x 32 output staging block
constexpr float WM_LOG2E = 1.4426950408889634f;

struct WmSmem {
  static constexpr int kTiles = 0;
  static constexpr int kBias = kTiles + WM_STAGES * WM_STAGE;
  static constexpr int kOut = kBias + 2 * WM_WT * WM_BIASP * 4;
  static constexpr int kReg = kOut + WM_CWARPS * 16 * WM_OUTP * 4;
  static constexpr int kBar = kReg + WM_CWARPS * 64 * 4;
  static constexpr int kTotal = kBar + 64 + 1024;   // + alignment slack
};

struct WmParams {
  int C, heads, nWh, nWw;
  long n_win;
  int shift, Hp, Wp;
  int npairs, nslots;
  const float* bias;       // [heads, 49, 49]
  bf16* out_hi;
  bf16* out_lo;
  float* out_f32;
};

__device__ __forceinline__ float wm_ex2(float x) {   // 2^x, MUFU.EX2 (2 ulp); 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a 128-byte-swizzled tile (tile base 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(WM_THREADS, 1)
window_attention_ms_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const WmParams p) {
  using L = WmSmem;
  extern __shared__ uint8_t wm_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wm_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* empty = full + WM_STAGES;
  float* sbias = reinterpret_cast<float*>(smem + L::kBias);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x % p.npairs, slot = blockIdx.x / p.npairs;

  if (tid == 0) {
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
    for (int s = 0; s < WM_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], WM_CWARPS);
    }
    ptx::fence_mbar_init();
  }
  // rows 49..63 of every tile are never written by the 49-row TMA boxes: zero them once (keys / values beyond the window
  // must be finite: their probabilities are exactly 0)
  for (int i = tid; i < WM_STAGES * 6 * 15 * 8; i += WM_THREADS) {
    const int tile = i / (15 * 8), rem = i % (15 * 8);
    *reinterpret_cast<uint4*>(smem + L::kTiles + tile * WM_TILE + (WM_WT + rem / 8) * 128 + (rem % 8) * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < 2 * WM_WT * WM_WT; i += WM_THREADS) {
    const int hh = i / (WM_WT * WM_WT), rem = i % (WM_WT * WM_WT);
    sbias[(hh * WM_WT + rem / WM_WT) * WM_BIASP + rem % WM_WT] =
        p.bias[static_cast<long>(pair * 2 + hh) * WM_WT * WM_WT + rem] * WM_LOG2E;   // base-2 softmax: exp(x) = 2^(x log2 e)
  }
  __syncthreads();

  if (warp == WM_CWARPS) {
    // ================================================================================= TMA producer
    int it = 0;
    for (long w = slot; w < p.n_win; w += p.nslots, ++it) {
      const int s = it % WM_STAGES;
      const uint32_t ph = (it / WM_STAGES) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&full[s], 6 * WM_WT * 128);
        const int row = static_cast<int>(w) * WM_WT;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int which = 0; which < 3; ++which)
            ptx::tma_load_2d(smem + L::kTiles + s * WM_STAGE + (pl * 3 + which) * WM_TILE,
                             pl ? static_cast<const void*>(&tm_lo) : static_cast<const void*>(&tm_hi), &full[s],
                             which * p.C + pair * 64, row);
      }
      __syncwarp();
    }
    return;
  }

  // =================================================================================== consumer warps
  const int hh = warp >> 2, band = warp & 3;   // head of the pair, 16-row query band
  const int g = lane >> 2, t = lane & 3;
  const int lrow = lane & 7, lmat = lane >> 3;
  const int r_lo = band * 16 + g, r_hi = r_lo + 8;
  const uint32_t bias_h = ptx::smem_u32(sbias + hh * WM_WT * WM_BIASP);   // explicit ld.shared below (generic LD otherwise)
  float* so = reinterpret_cast<float*>(smem + L::kOut) + warp * 16 * WM_OUTP;
  int* sreg = reinterpret_cast<int*>(smem + L::kReg) + warp * 64;
  const uint32_t sreg_a = ptx::smem_u32(sreg);
  const uint32_t tiles = ptx::smem_u32(smem + L::kTiles);
  const int head = pair * 2 + hh;

  int it = 0;
  for (long w = slot; w < p.n_win; w += p.nslots, ++it) {
    const int s = it % WM_STAGES;
    const uint32_t ph = (it / WM_STAGES) & 1;
    // shift-mask region of every token of this window (warp-private copy; swin_transformer.py:225-243)
    if (p.shift > 0) {
      const int wi = static_cast<int>(w % (static_cast<long>(p.nWh) * p.nWw));
      for (int k = lane; k < 64; k += 32) {
        int reg = 0;
        if (k < WM_WT) {
          const int y = (wi / p.nWw) * 7 + k / 7, x = (wi % p.nWw) * 7 + k % 7;
          const int rh = y < p.Hp - 7 ? 0 : (y < p.Hp - p.shift ? 1 : 2);
          const int rw = x < p.Wp - 7 ? 0 : (x < p.Wp - p.shift ? 1 : 2);
          reg = rh * 3 + rw;
        }
        sreg[k] = reg;
      }
      __syncwarp();
    }
    ptx::mbar_wait_hot(&full[s], ph);
    const uint32_t st = tiles + s * WM_STAGE;
    const uint32_t q_hi = st, k_hi = st + WM_TILE, v_hi = st + 2 * WM_TILE;
    const uint32_t q_lo = st + 3 * WM_TILE, k_lo = st + 4 * WM_TILE, v_lo = st + 5 * WM_TILE;

    // ---- S = q k^T (q carries the 32^-0.5 scale): A fragments of q (hi, lo) for both k-steps
    uint32_t aq[2][2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t off = sw128(band * 16 + (lane & 15), hh * 4 + ks * 2 + (lane >> 4));
      ldsm_x4(aq[0][ks], q_hi + off);
      ldsm_x4(aq[1][ks], q_lo + off);
    }
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 7; ++j) {   // key block 7 (keys 56..63) is all padding
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      uint32_t bh[4], bl[4];  // (ks0 b0, ks0 b1, ks1 b0, ks1 b1)
      const uint32_t off = sw128(8 * j + lrow, hh * 4 + lmat);
      ldsm_x4(bh, k_hi + off);
      ldsm_x4(bl, k_lo + off);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        mma_bf16_16816(sc[j], aq[0][ks], bh[2 * ks], bh[2 * ks + 1]);
        mma_bf16_16816(sc[j], aq[1][ks], bh[2 * ks], bh[2 * ks + 1]);
        mma_bf16_16816(sc[j], aq[0][ks], bl[2 * ks], bl[2 * ks + 1]);
      }
    }
    // ---- + relative-position bias + shift mask; padding columns excluded; softmax per row in base 2 (fp32): the scores
    //      are carried as x log2(e) (bias pre-multiplied in shared memory), so each probability is one FADD + MUFU.EX2
    const int reg_lo = p.shift > 0 ? ptx::lds_s32(sreg_a + r_lo * 4) : 0, reg_hi = p.shift > 0 ? ptx::lds_s32(sreg_a + r_hi * 4) : 0;
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int col = 8 * j + 2 * t;
      float2 b_lo = make_float2(0.f, 0.f), b_hi = make_float2(0.f, 0.f);
      if (r_lo < WM_WT) b_lo = ptx::lds_f32x2(bias_h + (r_lo * WM_BIASP + col) * 4);
      if (r_hi < WM_WT) b_hi = ptx::lds_f32x2(bias_h + (r_hi * WM_BIASP + col) * 4);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (col + e < WM_WT) {
          sc[j][e] = fmaf(sc[j][e], WM_LOG2E, e ? b_lo.y : b_lo.x);
          sc[j][2 + e] = fmaf(sc[j][2 + e], WM_LOG2E, e ? b_hi.y : b_hi.x);
          if (p.shift > 0) {
            const int rc = ptx::lds_s32(sreg_a + (col + e) * 4);
            if (rc != reg_lo) sc[j][e] += -100.0f * WM_LOG2E;
            if (rc != reg_hi) sc[j][2 + e] += -100.0f * WM_LOG2E;
          }
        } else {
          sc[j][e] = -INFINITY;
          sc[j][2 + e] = -INFINITY;
        }
        m_lo = fmaxf(m_lo, sc[j][e]);
        m_hi = fmaxf(m_hi, sc[j][2 + e]);
      }
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1)); m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1)); m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[j][e] = wm_ex2(sc[j][e] - m_lo);
        sc[j][2 + e] = wm_ex2(sc[j][2 + e] - m_hi);
        sum_lo += sc[j][e];
        sum_hi += sc[j][2 + e];
      }
    sc[7][0] = sc[7][1] = sc[7][2] = sc[7][3] = 0.f;
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
    // ---- O = P v : P (normalised, split) is the A operand straight from the accumulator fragments; v [key][dim] is read
    //      through ldmatrix.trans: matrix i = (key half i & 1, dim tile 2 np + (i >> 1))
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph4[4], pl4[4];
      split_pack2(sc[2 * kk][0] * inv_lo, sc[2 * kk][1] * inv_lo, ph4[0], pl4[0]);
      split_pack2(sc[2 * kk][2] * inv_hi, sc[2 * kk][3] * inv_hi, ph4[1], pl4[1]);
      split_pack2(sc[2 * kk + 1][0] * inv_lo, sc[2 * kk + 1][1] * inv_lo, ph4[2], pl4[2]);
      split_pack2(sc[2 * kk + 1][2] * inv_hi, sc[2 * kk + 1][3] * inv_hi, ph4[3], pl4[3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vh[4], vl[4];  // (b0, b1) of dim tile 2np, (b0, b1) of dim tile 2np+1
        const uint32_t off = sw128(16 * kk + (lmat & 1) * 8 + lrow, hh * 4 + 2 * np + (lmat >> 1));
        ldsm_x4_trans(vh, v_hi + off);
        ldsm_x4_trans(vl, v_lo + off);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          mma_bf16_16816(o[2 * np + q], ph4, vh[2 * q], vh[2 * q + 1]);
          mma_bf16_16816(o[2 * np + q], pl4, vh[2 * q], vh[2 * q + 1]);
          mma_bf16_16816(o[2 * np + q], ph4, vl[2 * q], vl[2 * q + 1]);
        }
      }
    }
    // the stage is free for the producer once every lane of this warp has issued its last ldmatrix
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);
    // ---- store: the warp's 16 x 32 block through its private staging block, then 16-byte row segments
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<float2*>(so + g * WM_OUTP + 8 * n + 2 * t) = make_float2(o[n][0], o[n][1]);
      *reinterpret_cast<float2*>(so + (g + 8) * WM_OUTP + 8 * n + 2 * t) = make_float2(o[n][2], o[n][3]);
    }
    __syncwarp();
    const long row0 = w * WM_WT;
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = i >> 3, c4 = (i & 7) * 4;
      const int row = band * 16 + r;
      if (row < WM_WT) {
        const float4 y = *reinterpret_cast<const float4*>(so + r * WM_OUTP + c4);
        const long ooff = (row0 + row) * p.C + head * 32 + c4;
        if (p.out_hi) {
          uint32_t h01, l01, h23, l23;
          split_pack2_bf16(y.x, y.y, h01, l01);
          split_pack2_bf16(y.z, y.w, h23, l23);
          *reinterpret_cast<uint2*>(p.out_hi + ooff) = make_uint2(h01, h23);
          if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + ooff) = make_uint2(l01, l23);
        }
        if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + ooff) = y;
      }
    }
    __syncwarp();   // the staging block and the region table are rewritten by the next item
  }
}

CUtensorMap wm_qkv_map(Ctx* c, const bf16* base, long rows, long ld) {
  CUtensorMap tm;
  ALM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, ALM_ERR_INVALID, "window attention operand alignment");
  cuuint64_t dims[2] = {cuuint64_t(ld), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(WM_WT)};
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
void window_attention_ms(Ctx* c, const bf16* qkv_hi, const bf16* qkv_lo, int C, int heads, int nWh, int nWw, int B, int shift,
                         int Hp, int Wp, const float* bias_dense, bf16* out_hi, bf16* out_lo, float* out_f32) {
  ALM_REQUIRE(C == heads * 32 && heads % 2 == 0, ALM_ERR_UNSUPPORTED, "window_attention_ms: head_dim 32, even head count");
  ALM_REQUIRE(qkv_lo, ALM_ERR_INVALID, "window_attention_ms: needs the lo plane");
  WmParams p;
  p.C = C; p.heads = heads; p.nWh = nWh; p.nWw = nWw;
  p.n_win = static_cast<long>(B) * nWh * nWw;
  p.shift = shift; p.Hp = Hp; p.Wp = Wp;
  p.npairs = heads / 2;
  p.nslots = static_cast<int>(std::max<long>(1, std::min<long>(p.n_win, c->num_sms / p.npairs)));
  p.bias = bias_dense;
  p.out_hi = out_hi; p.out_lo = out_lo; p.out_f32 = out_f32;
  const long rows = p.n_win * WM_WT;
  ALM_REQUIRE(rows < (1L << 31), ALM_ERR_UNSUPPORTED, "window_attention_ms: too many rows for one tensor map");
  const CUtensorMap th = wm_qkv_map(c, qkv_hi, rows, 3L * C);
  const CUtensorMap tl = wm_qkv_map(c, qkv_lo, rows, 3L * C);
  static DeviceOnce attr;
  if (attr.need()) {
    ALM_CHECK_CUDA(cudaFuncSetAttribute(window_attention_ms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WmSmem::kTotal));
    attr.mark();
  }
  window_attention_ms_kernel<<<p.npairs * p.nslots, WM_THREADS, WmSmem::kTotal, c->stream>>>(th, tl, p);
  count_launch(c);
  check_launch("window_attention_ms");
}

}  // namespace alm
