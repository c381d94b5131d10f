// Test-time image pipeline (preproc.cu): shared declarations.
#pragma once
#include <vector>

#include "alm_internal.h"

namespace alm {

enum { PRE_BILINEAR = 0, PRE_BICUBIC = 1 };

struct PreImage {
  const uint8_t* dev_rgb;  // device, [h][w][3] uint8
  int h, w;                // source size
  int oh, ow;              // resized size
};

// RandomResize.get_size_with_aspect_ratio for one (h, w) (host only)
void pre_omni_size(int h, int w, int min_size, int max_size, int* oh, int* ow);
// fixed-point weights of one pass: bounds [out][2], coefs [out][*ksize] (call with coefs == nullptr to size the buffers)
int pre_coeffs(int in_size, int out_size, int filter, int* ksize, int* bounds, int* coefs, size_t cap_ints);
// Resize every image with Pillow's 8-bit resampling, ToTensor (+ Normalize), write it into the top-left corner of its
// [3, Hc, Wc] slot of `out` (zero elsewhere) and, optionally, the pad mask [n, Hc, Wc] (1 = padding).
void pre_resize_batch(Ctx* c, const std::vector<PreImage>& imgs, int filter, bool normalize, float* out, int Hc, int Wc,
                      uint8_t* mask);

}  // namespace alm
