// Per-output-pixel bodies of the pre-processing kernels (preproc.cu).  Plain index arithmetic, integer MACs and two
// IEEE float32 operations -- no warp primitives, no shared memory -- so the very same functions also compile as host
// code: tests/emu/preproc_emu.cpp runs them in a loop over the "grid" and checks the result bit for bit against the
// reference fixtures on a machine without a GPU (test infrastructure only; the library itself never runs them on
// the host).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ALM_HD __host__ __device__ __forceinline__
#else
#define ALM_HD inline
#endif

namespace alm {
namespace pre_core {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

struct Norm {
  float mean[3], sd[3];
  int on;
};

// IEEE round-to-nearest float32 ops that the compiler must not contract or approximate
ALM_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
ALM_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}

ALM_HD int clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift: negative sums (bicubic undershoot) clip to 0 like Pillow's lookup table
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// output pixel i = (y, xx) of the horizontal pass: src [h][w][3] -> dst [h][ow][3]
ALM_HD void resample_h_px(long i, const uint8_t* src, int w, int ow, const int* bounds, const int* coefs, int ksize,
                          uint8_t* dst) {
  const int y = static_cast<int>(i / ow), xx = static_cast<int>(i - static_cast<long>(y) * ow);
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* k = coefs + static_cast<long>(xx) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const uint8_t* p = src + (static_cast<long>(y) * w + xmin) * 3;
  for (int x = 0; x < n; ++x) {
    const int kk = k[x];
    s0 += p[3 * x] * kk; s1 += p[3 * x + 1] * kk; s2 += p[3 * x + 2] * kk;
  }
  uint8_t* o = dst + i * 3;
  o[0] = static_cast<uint8_t>(clip8(s0)); o[1] = static_cast<uint8_t>(clip8(s1)); o[2] = static_cast<uint8_t>(clip8(s2));
}

// output pixel i = (yy, xx) of the vertical pass + F.to_tensor (uint8 -> float32, .div(255)) + F.normalize
// (.sub_(mean).div_(std)): src [h][ow][3] -> dst[c][yy][xx] (row stride Wc, plane stride `plane`)
ALM_HD void resample_v_norm_px(long i, const uint8_t* src, int ow, const int* bounds, const int* coefs, int ksize, float* dst,
                               long plane, int Wc, const Norm& nm) {
  const int yy = static_cast<int>(i / ow), xx = static_cast<int>(i - static_cast<long>(yy) * ow);
  const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int* k = coefs + static_cast<long>(yy) * ksize;
  int s[3];
  s[0] = s[1] = s[2] = 1 << (kPrecisionBits - 1);
  const uint8_t* p = src + (static_cast<long>(ymin) * ow + xx) * 3;
  for (int y = 0; y < n; ++y) {
    const int kk = k[y];
    const uint8_t* q = p + static_cast<long>(y) * ow * 3;
    s[0] += q[0] * kk; s[1] += q[1] * kk; s[2] += q[2] * kk;
  }
  const long o = static_cast<long>(yy) * Wc + xx;
  for (int ch = 0; ch < 3; ++ch) {
    float f = fdiv(static_cast<float>(clip8(s[ch])), 255.0f);
    if (nm.on) f = fdiv(fsub(f, nm.mean[ch]), nm.sd[ch]);
    dst[ch * plane + o] = f;
  }
}

// mask element i of [n][Hc][Wc]: 1 = padding (nested_tensor.py:47-51); sizes [n][2] = resized (h, w)
ALM_HD void pad_mask_px(long i, uint8_t* mask, int Hc, int Wc, const int* sizes) {
  const long per = static_cast<long>(Hc) * Wc;
  const int b = static_cast<int>(i / per);
  const long r = i - b * per;
  const int y = static_cast<int>(r / Wc), x = static_cast<int>(r - static_cast<long>(y) * Wc);
  mask[i] = (y >= sizes[2 * b] || x >= sizes[2 * b + 1]) ? 1 : 0;
}

}  // namespace pre_core
}  // namespace alm
