// TEST INFRASTRUCTURE: host emulation of the pre-processing kernels.  Compiles the per-output-pixel bodies shared with
// the CUDA kernels (advancedliteratemachinery_b200/csrc/preproc_core.h) as plain C++ and runs them over the whole
// "grid" in the launch order preproc.cu uses, so that the kernels' index arithmetic and integer / float32 math is checked
// bit for bit against the reference fixtures without a GPU.  Built on the fly by tests/test_preprocess.py
// (g++ -O1 -ffp-contract=off); never part of libalm_ocr.so.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../advancedliteratemachinery_b200/csrc/preproc_core.h"

using namespace alm::pre_core;

extern "C" {

// one image: [h][w][3] u8 -> its slot in the planar canvas (dst points at the image's first plane)
void emu_resize_norm(const uint8_t* src, int h, int w, int oh, int ow, const int* hb, const int* hk, int hks, const int* vb,
                     const int* vk, int vks, float* dst, long plane, int Wc, const float* mean, const float* sd, int on) {
  std::vector<uint8_t> tmp(static_cast<size_t>(h) * ow * 3);
  for (long i = 0; i < static_cast<long>(h) * ow; ++i) resample_h_px(i, src, w, ow, hb, hk, hks, tmp.data());
  Norm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean[c]; nm.sd[c] = sd[c]; }
  nm.on = on;
  for (long i = 0; i < static_cast<long>(oh) * ow; ++i) resample_v_norm_px(i, tmp.data(), ow, vb, vk, vks, dst, plane, Wc, nm);
}

void emu_pad_mask(uint8_t* mask, int n, int Hc, int Wc, const int* sizes) {
  for (long i = 0; i < static_cast<long>(n) * Hc * Wc; ++i) pad_mask_px(i, mask, Hc, Wc, sizes);
}

}  // extern "C"
