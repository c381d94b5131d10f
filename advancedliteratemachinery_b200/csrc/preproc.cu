// Test-time image pipeline on the GPU: the step right before the forward.
//
//   OmniParser  dataset/__init__.py:109-113, dataset/transforms.py:249-298 (RandomResize with one min size =
//               shorter-side resize with a long-side cap, PIL bilinear), :312-322 (ToTensor, Normalize with the
//               ImageNet statistics), utils/nested_tensor.py:37-54 (zero-pad to the batch maximum + bool pad mask)
//   MGP-STR     demo.py:126-132, dataset.py:459-461 (PIL bicubic resize to imgW x imgH, ToTensor, no normalise)
//
// The resize is Pillow's two-pass separable resampling of 8-bit images (third-party dependency of the reference,
// pillow==8.1.0 in OCR/MGP-STR/requirements.txt; algorithm of src/libImaging/Resample.c, unchanged since Pillow 3):
// per output pixel a window of support*max(scale,1) source pixels, triangle (bilinear) or Keys a=-0.5 (bicubic)
// weights normalised in double, quantised to 22-bit fixed point, horizontal pass rounded and clipped to uint8, then
// the vertical pass.  The weights are computed on the host exactly like Pillow computes them (a few hundred doubles
// per image); the two passes run as integer kernels, so the result is bit-identical to PIL by construction, and
// ToTensor / Normalize are the same two IEEE float32 operations per element fused into the vertical pass, which
// writes straight into the padded NCHW batch.  HBM-bound: reads 3 B and writes 3 B (+12 B of float output) per pixel.
#include <math.h>

#include <algorithm>
#include <map>

#include "alm_internal.h"
#include "pre.h"
#include "preproc_core.h"

namespace alm {
namespace {

constexpr int kPrecisionBits = pre_core::kPrecisionBits;  // Pillow: PRECISION_BITS

double filt_bilinear(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}
double filt_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct Coefs {
  int ksize = 0;
  std::vector<int> bounds;  // [out][2] = first source index, tap count
  std::vector<int> k;       // [out][ksize] fixed-point weights
};

// precompute_coeffs + normalize_coeffs_8bpc (Resample.c), same operations in the same order, in double
Coefs precompute(int in_size, int out_size, int filter) {
  double (*f)(double) = filter == PRE_BICUBIC ? filt_bicubic : filt_bilinear;
  const double support0 = filter == PRE_BICUBIC ? 2.0 : 1.0;
  double scale = static_cast<double>(in_size) / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = support0 * filterscale;
  Coefs c;
  c.ksize = static_cast<int>(ceil(support)) * 2 + 1;
  c.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  c.k.assign(static_cast<size_t>(out_size) * c.ksize, 0);
  std::vector<double> w(static_cast<size_t>(c.ksize));
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      w[x] = f((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      c.k[static_cast<size_t>(xx) * c.ksize + x] =
          v < 0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits)) : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
    }
    c.bounds[2 * xx] = xmin;
    c.bounds[2 * xx + 1] = xmax;
  }
  return c;
}

// horizontal pass: src [h][w][3] u8 -> dst [h][ow][3] u8; one thread per output pixel (body in preproc_core.h)
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, int h, int w, int ow, const int* __restrict__ bounds,
                                  const int* __restrict__ coefs, int ksize, uint8_t* __restrict__ dst) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long>(h) * ow) pre_core::resample_h_px(i, src, w, ow, bounds, coefs, ksize, dst);
}

// vertical pass fused with ToTensor (/255) and Normalize ((x - mean) / std), writing the planar float canvas:
// src [h][ow][3] u8 -> dst[c][yy][xx] (row stride Wc, plane stride Hc*Wc); one thread per output pixel
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ src, int ow, int oh, const int* __restrict__ bounds,
                                       const int* __restrict__ coefs, int ksize, float* __restrict__ dst, long plane, int Wc,
                                       pre_core::Norm nm) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long>(oh) * ow) pre_core::resample_v_norm_px(i, src, ow, bounds, coefs, ksize, dst, plane, Wc, nm);
}

__global__ void pad_mask_kernel(uint8_t* __restrict__ mask, int n, int Hc, int Wc, const int* __restrict__ sizes) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long>(n) * Hc * Wc) pre_core::pad_mask_px(i, mask, Hc, Wc, sizes);
}

}  // namespace

// weights of one resampling pass as the kernels consume them (host only; lets the CPU tests pin the double
// arithmetic of `precompute` against Pillow's through the oracle)
int pre_coeffs(int in_size, int out_size, int filter, int* ksize, int* bounds, int* coefs, size_t cap_ints) {
  if (in_size <= 0 || out_size <= 0 || (filter != PRE_BILINEAR && filter != PRE_BICUBIC) || !ksize) return ALM_ERR_INVALID;
  const Coefs c = precompute(in_size, out_size, filter);
  *ksize = c.ksize;
  if (!bounds || !coefs || cap_ints < c.k.size()) return ALM_ERR_INVALID;
  std::copy(c.bounds.begin(), c.bounds.end(), bounds);
  std::copy(c.k.begin(), c.k.end(), coefs);
  return ALM_OK;
}

// RandomResize.get_size_with_aspect_ratio (transforms.py:275-296) with the reference's float / int conversions
void pre_omni_size(int h, int w, int min_size, int max_size, int* oh, int* ow) {
  int size = min_size;
  if (max_size > 0) {
    const double mn = static_cast<double>(h < w ? h : w), mx = static_cast<double>(h < w ? w : h);
    if (mx / mn * size > max_size) size = static_cast<int>(nearbyint(max_size * mn / mx));  // Python round(): half to even
  }
  if ((w <= h && w == size) || (h <= w && h == size)) {
    *oh = h; *ow = w;
    return;
  }
  if (w < h) {
    *ow = size;
    *oh = static_cast<int>(static_cast<double>(static_cast<long>(size) * h) / w);  // int(size * h / w)
  } else {
    *oh = size;
    *ow = static_cast<int>(static_cast<double>(static_cast<long>(size) * w) / h);
  }
}

void pre_resize_batch(Ctx* c, const std::vector<PreImage>& imgs, int filter, bool normalize, float* out, int Hc, int Wc,
                      uint8_t* mask) {
  const int n = static_cast<int>(imgs.size());
  if (n == 0) return;
  c->ensure_ws();
  Arena& ws = c->ws;
  const size_t mark = ws.mark();
  const long plane = static_cast<long>(Hc) * Wc;
  ALM_CHECK_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(n) * 3 * plane * sizeof(float), c->stream));  // zero pad
  // ---- weights of every distinct (source, target) length, one host array, one upload
  std::map<std::pair<int, int>, std::pair<size_t, int>> where;  // (in, out) -> (offset of bounds in ints, ksize)
  std::vector<int> host;
  auto add = [&](int in_size, int out_size) {
    const auto key = std::make_pair(in_size, out_size);
    if (where.count(key)) return;
    const Coefs cf = precompute(in_size, out_size, filter);
    where[key] = {host.size(), cf.ksize};
    host.insert(host.end(), cf.bounds.begin(), cf.bounds.end());
    host.insert(host.end(), cf.k.begin(), cf.k.end());
  };
  std::vector<int> sizes(static_cast<size_t>(n) * 2);
  for (int b = 0; b < n; ++b) {
    const PreImage& im = imgs[static_cast<size_t>(b)];
    ALM_REQUIRE(im.h > 0 && im.w > 0 && im.oh > 0 && im.ow > 0 && im.oh <= Hc && im.ow <= Wc && im.dev_rgb, ALM_ERR_INVALID,
                "pre_resize_batch: image geometry");
    add(im.w, im.ow);
    add(im.h, im.oh);
    sizes[2 * b] = im.oh; sizes[2 * b + 1] = im.ow;
  }
  const size_t ncoef = host.size();
  host.insert(host.end(), sizes.begin(), sizes.end());
  int* dev = ws.get<int>(host.size());
  ALM_CHECK_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));  // `host` is pageable and goes out of scope with this call
  // ---- two integer passes per image (Pillow: horizontal first, then vertical)
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};  // transforms.py:318
  for (int b = 0; b < n; ++b) {
    const PreImage& im = imgs[static_cast<size_t>(b)];
    const auto wh = where[{im.w, im.ow}], wv = where[{im.h, im.oh}];
    uint8_t* tmp = ws.get<uint8_t>(static_cast<size_t>(im.h) * im.ow * 3);
    const long nh = static_cast<long>(im.h) * im.ow, nv = static_cast<long>(im.oh) * im.ow;
    resample_h_kernel<<<static_cast<unsigned>((nh + 255) / 256), 256, 0, c->stream>>>(
        im.dev_rgb, im.h, im.w, im.ow, dev + wh.first, dev + wh.first + 2 * static_cast<size_t>(im.ow), wh.second, tmp);
    count_launch(c); check_launch("resample_h");
    const pre_core::Norm nm{{mean[0], mean[1], mean[2]}, {sd[0], sd[1], sd[2]}, normalize ? 1 : 0};
    resample_v_norm_kernel<<<static_cast<unsigned>((nv + 255) / 256), 256, 0, c->stream>>>(
        tmp, im.ow, im.oh, dev + wv.first, dev + wv.first + 2 * static_cast<size_t>(im.oh), wv.second,
        out + static_cast<long>(b) * 3 * plane, plane, Wc, nm);
    count_launch(c); check_launch("resample_v_norm");
  }
  if (mask) {
    const long tot = static_cast<long>(n) * plane;
    pad_mask_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, c->stream>>>(mask, n, Hc, Wc, dev + ncoef);
    count_launch(c); check_launch("pad_mask");
  }
  ws.release(mark);  // stream-ordered: later allocations are only touched by later launches
}

}  // namespace alm
