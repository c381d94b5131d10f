// MGP-STR (ViT recogniser + A^3 heads) model state and entry points.
#pragma once
#include "alm_internal.h"

namespace alm {
struct MgpModel;
void mgp_load(Ctx* c, const std::map<std::string, HostTensor>& t);
void mgp_forward(Ctx* c, const float* img_dev, int B, float* attn, float* char_logits, float* bpe_logits,
                 float* wp_logits, int32_t* ids, float* prob);
void mgp_free(MgpModel* m);
// dim, depth, heads, number of A^3 heads (3 = MGP-STR, 1 = CHAR-STR), class counts of the heads
void mgp_info(const MgpModel* m, int* dim, int* depth, int* heads, int* n_a3, int* vocab3);
MgpModel* mgp_share(const MgpModel* owner);  // weights only: a plain copy of the pointer table
}  // namespace alm
