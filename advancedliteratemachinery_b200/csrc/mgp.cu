// MGP-STR forward on the GPU: ViT encoder (257 tokens) + three A^3 token learners + char/BPE/WordPiece heads.
// Host-side graph; the arithmetic is the shared tcgen05 GEMM engine plus the row kernels in kernels.cu.
//
// Reference being replaced (relative to /root/reference/OCR/MGP-STR/): modules/mgp_str.py:64-101,
// modules/token_learner.py:21-32, timm==0.4.12 vision_transformer Block/Attention/Mlp (block LN eps 1e-6,
// scores scaled after q k^T, final norm not applied), demo.py:36-60 (top-1 id + max softmax prob).
#include "mgp.h"

#include <string.h>

#include <algorithm>

#include "omni.h"

#include "ptx.cuh"

namespace alm {

struct MgpLin {
  SplitW w;
  float* b = nullptr;
};
struct MgpLN {
  float* g = nullptr;
  float* b = nullptr;
};
struct MgpBlock {
  MgpLN n1, n2;
  MgpLin qkv, proj, fc1, fc2;  // qkv = the packed [3D, D] projection (rows q | k | v), as in the checkpoint
  MgpLin qk, v;                // row views [0, 2D) and [2D, 3D) of qkv for the unfused attention path
};
struct MgpA3 {
  MgpLN token_norm, norm;
  SplitW w0, w1, feat;  // grouped conv [D, D/8], select conv [27, D], grouped feat conv [D, D/8]
  MgpLin head;
  int V = 0, Vpad = 0;
};
struct MgpModel {
  int D = 768, depth = 12, heads = 12;
  MgpLin patch;
  float* cls = nullptr;   // [D]
  float* pos = nullptr;   // [257, D]
  std::vector<MgpBlock> blocks;
  MgpA3 a3[3];
  int n_a3 = 3;  // 3 = MGP-STR (char, bpe, wp); 1 = the char-only CHAR-STR ablation (modules/char_str.py:43-81)
};

void mgp_free(MgpModel* m) { delete m; }
void mgp_info(const MgpModel* m, int* dim, int* depth, int* heads, int* n_a3, int* vocab3) {
  if (dim) *dim = m->D;
  if (depth) *depth = m->depth;
  if (heads) *heads = m->heads;
  if (n_a3) *n_a3 = m->n_a3;
  if (vocab3)
    for (int a = 0; a < 3; ++a) vocab3[a] = a < m->n_a3 ? m->a3[a].V : 0;
}
MgpModel* mgp_share(const MgpModel* owner) { return new MgpModel(*owner); }

namespace {

constexpr int T = 257, TP = 264, NTOK = 27, NPATCH = 256;

MgpLN ld_ln(Ctx* c, const std::map<std::string, HostTensor>& t, const std::string& p, int dim) {
  const HostTensor& g = need(t, p + ".weight");
  const HostTensor& b = need(t, p + ".bias");
  ALM_REQUIRE(static_cast<int>(g.numel()) == dim && static_cast<int>(b.numel()) == dim, ALM_ERR_INVALID,
              "LayerNorm shape mismatch at " + p);
  return MgpLN{upload_f32(c, g.f32.data(), dim), upload_f32(c, b.f32.data(), dim)};
}
MgpLin ld_lin(Ctx* c, const std::map<std::string, HostTensor>& t, const std::string& p, int N, int K, int Kpad = 0) {
  const HostTensor& w = need(t, p + ".weight");
  const HostTensor& b = need(t, p + ".bias");
  ALM_REQUIRE(static_cast<long>(w.numel()) == static_cast<long>(N) * K && static_cast<int>(b.numel()) == N,
              ALM_ERR_INVALID, "Linear shape mismatch at " + p);
  return MgpLin{upload_split(c, w.f32.data(), N, K, Kpad), upload_f32(c, b.f32.data(), N)};
}

__global__ void mgp_patch_maps_kernel(int* out_map, int* resid_map, int B) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * NPATCH) return;
  out_map[r] = (r / NPATCH) * T + 1 + r % NPATCH;
  resid_map[r] = 1 + r % NPATCH;
}
// x[b*257 + 0, :] = cls + pos_embed[0]   (mgp_str.py:68-70)
__global__ void mgp_cls_kernel(float* x, const float* cls, const float* pos, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  x[static_cast<long>(b) * T * D + d] = cls[d] + pos[d];
}
// per row: argmax id and max softmax probability over n classes (demo.py:36-60)
__global__ void __launch_bounds__(256)
argmax_prob_kernel(const float* __restrict__ logits, long ld, int n, int* __restrict__ ids, float* __restrict__ prob) {
  __shared__ float sf[8];
  __shared__ int si[8];
  const long r = blockIdx.x;
  const float* x = logits + r * ld;
  const int t = threadIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = t; j < n; j += 256) {
    const float v = x[j];
    if (v > best) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((t & 31) == 0) { sf[t >> 5] = best; si[t >> 5] = bi; }
  __syncthreads();
  best = sf[0]; bi = si[0];
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (sf[i] > best || (sf[i] == best && si[i] < bi)) { best = sf[i]; bi = si[i]; }
  __syncthreads();
  float sum = 0.f;
  for (int j = t; j < n; j += 256) sum += expf(x[j] - best);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((t & 31) == 0) sf[t >> 5] = sum;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += sf[i];
    ids[r] = bi;
    prob[r] = 1.0f / tot;  // exp(best - best) / sum
  }
}

struct SB {
  bf16* hi;
  bf16* lo;
};
SB sb(Ctx* c, size_t n) { return SB{c->ws.get<bf16>(n), c->ws.get<bf16>(n)}; }
Operand opnd(const bf16* hi, const bf16* lo, int rows, int K, long ld) {
  Operand o;
  o.hi = hi; o.lo = lo; o.rows = rows; o.K = K; o.ld = ld;
  return o;
}
void lin(Ctx* c, const SB& a, long rows, const MgpLin& l, int act, float* out_f32, SB* out_split, const float* resid) {
  Epilogue e;
  e.out_f32 = out_f32;
  if (out_split) { e.out_hi = out_split->hi; e.out_lo = out_split->lo; }
  e.ldo = l.w.N;
  e.bias = l.b; e.bias_mode = l.b ? BIAS_COL : BIAS_NONE;
  e.act = act;
  e.resid = resid; e.ldr = l.w.N;
  gemm(c, opnd(a.hi, a.lo, static_cast<int>(rows), l.w.K, l.w.K), l.w.op(), e);
}

}  // namespace

void mgp_load(Ctx* c, const std::map<std::string, HostTensor>& tin) {
  // accept 'module.mgp_str.' (DataParallel + Model wrapper, test_final.py:348-356), 'mgp_str.' or bare keys
  std::map<std::string, HostTensor> t;
  for (auto& kv : tin) {
    std::string k = kv.first;
    for (const char* pre : {"module.", "mgp_str."})
      if (k.rfind(pre, 0) == 0) k = k.substr(strlen(pre));
    t.emplace(k, kv.second);
  }
  MgpModel* m = new MgpModel();
  const HostTensor& pe = need(t, "pos_embed");
  ALM_REQUIRE(pe.shape.size() == 3 && pe.shape[1] == T, ALM_ERR_INVALID, "pos_embed must be [1,257,D] (32x128 input, patch 4)");
  m->D = static_cast<int>(pe.shape[2]);
  const int D = m->D;
  // tiny 192 / small 384 / base 768 / large 1024 (mgp_str.py:176-230); every variant has 64-wide heads
  ALM_REQUIRE(D % 64 == 0 && (D % 128 == 0 || D == 192), ALM_ERR_UNSUPPORTED,
              "embed dim must be 192 or a multiple of 128 (tiny=192, small=384, base=768, large=1024)");
  m->heads = D / 64;
  ALM_REQUIRE((D / 8) % 8 == 0, ALM_ERR_UNSUPPORTED, "A^3 group width must be a multiple of 8");
  m->depth = 0;
  while (t.count("blocks." + std::to_string(m->depth) + ".norm1.weight")) ++m->depth;
  ALM_REQUIRE(m->depth > 0, ALM_ERR_INVALID, "no transformer blocks in the state dict");
  m->patch = ld_lin(c, t, "patch_embed.proj", D, 48, 64);
  m->cls = upload_f32(c, need(t, "cls_token").f32.data(), D);
  m->pos = upload_f32(c, pe.f32.data(), static_cast<size_t>(T) * D);
  m->blocks.resize(m->depth);
  for (int b = 0; b < m->depth; ++b) {
    const std::string p = "blocks." + std::to_string(b) + ".";
    MgpBlock& w = m->blocks[b];
    w.n1 = ld_ln(c, t, p + "norm1", D);
    w.n2 = ld_ln(c, t, p + "norm2", D);
    const HostTensor& qw = need(t, p + "attn.qkv.weight");
    const HostTensor& qb = need(t, p + "attn.qkv.bias");
    ALM_REQUIRE(qw.numel() == static_cast<size_t>(3) * D * D && static_cast<int>(qb.numel()) == 3 * D, ALM_ERR_INVALID,
                "qkv shape mismatch at " + p);
    w.qkv.w = upload_split(c, qw.f32.data(), 3 * D, D, 0);
    w.qkv.b = upload_f32(c, qb.f32.data(), 3 * D);
    w.qk = w.qkv; w.qk.w.N = 2 * D;
    w.v = w.qkv; w.v.w.N = D;
    w.v.w.hi += static_cast<size_t>(2) * D * w.qkv.w.ld; w.v.w.lo += static_cast<size_t>(2) * D * w.qkv.w.ld;
    w.v.b += 2 * D;
    w.proj = ld_lin(c, t, p + "attn.proj", D, D);
    w.fc1 = ld_lin(c, t, p + "mlp.fc1", 4 * D, D);
    w.fc2 = ld_lin(c, t, p + "mlp.fc2", D, 4 * D);
  }
  const char* names[3] = {"char", "bpe", "wp"};
  // CHAR-STR (modules/char_str.py:43-81): only the character A^3 module, and its logits come from timm's own
  // classifier `head` (char_str.py:70 calls self.head, not the char_head that reset_classifier creates)
  m->n_a3 = t.count("bpe_tokenLearner.token_norm.weight") ? 3 : 1;
  for (int a = 0; a < m->n_a3; ++a) {
    const std::string p = std::string(names[a]) + "_tokenLearner.";
    MgpA3& w = m->a3[a];
    w.token_norm = ld_ln(c, t, p + "token_norm", D);
    w.norm = ld_ln(c, t, p + "norm", D);
    const HostTensor& w0 = need(t, p + "tokenLearner.0.weight");
    const HostTensor& w1 = need(t, p + "tokenLearner.1.weight");
    const HostTensor& wf = need(t, p + "feat.weight");
    ALM_REQUIRE(w0.numel() == static_cast<size_t>(D) * (D / 8) && wf.numel() == w0.numel() &&
                    w1.numel() == static_cast<size_t>(NTOK) * D, ALM_ERR_INVALID, "A^3 conv shape mismatch at " + p);
    w.w0 = upload_split(c, w0.f32.data(), D, D / 8, 0);
    w.w1 = upload_split(c, w1.f32.data(), NTOK, D, 0);
    w.feat = upload_split(c, wf.f32.data(), D, D / 8, 0);
    const std::string hname = m->n_a3 == 1 ? std::string("head") : std::string(names[a]) + "_head";
    const HostTensor& hw = need(t, hname + ".weight");
    ALM_REQUIRE(hw.shape.size() == 2 && hw.shape[1] == D, ALM_ERR_INVALID, "head shape mismatch");
    w.V = static_cast<int>(hw.shape[0]);
    w.Vpad = (w.V + 7) & ~7;
    w.head = ld_lin(c, t, hname, w.V, D);
  }
  mgp_free(c->mgp);
  c->mgp = m;
}

void mgp_forward(Ctx* c, const float* img, int B, float* attn_out, float* char_logits, float* bpe_logits,
                 float* wp_logits, int32_t* ids_out, float* prob_out) {
  MgpModel* m = c->mgp;
  ALM_REQUIRE(m != nullptr, ALM_ERR_STATE, "alm_mgpstr_forward before alm_load_weights");
  c->ensure_ws();
  Arena& ws = c->ws;
  ws.off = 0;
  if (c->omni) c->omni->encoded = false;  // the arena is shared
  const int D = m->D, H = m->heads;
  const long R = static_cast<long>(B) * T;

  float* x = ws.get<float>(R * D);
  {  // patch embed + cls + pos (mgp_str.py:66-70)
    const size_t mk = ws.mark();
    SB a = sb(c, static_cast<size_t>(B) * NPATCH * 64);
    int* omap = ws.get<int>(static_cast<size_t>(B) * NPATCH);
    int* rmap = ws.get<int>(static_cast<size_t>(B) * NPATCH);
    im2col_patch4(c, img, B, 32, 128, 8, 32, a.hi, a.lo);
    mgp_patch_maps_kernel<<<(B * NPATCH + 255) / 256, 256, 0, c->stream>>>(omap, rmap, B);
    mgp_cls_kernel<<<(B * D + 255) / 256, 256, 0, c->stream>>>(x, m->cls, m->pos, B, D);
    count_launch(c, 2);
    check_launch("mgp patch maps");
    Epilogue e;
    e.out_f32 = x; e.ldo = D;
    e.bias = m->patch.b; e.bias_mode = BIAS_COL;
    e.resid = m->pos; e.ldr = D; e.out_map = omap; e.resid_map = rmap;
    gemm(c, opnd(a.hi, a.lo, B * NPATCH, 64, 64), m->patch.w.op(), e);
    ws.release(mk);
  }
  const bool fused_attn = c->attn_impl == 0;
  // single-pass bf16 with the fused attention: every consumer (GEMM, attention) reads only the hi planes, so the
  // producers skip computing and writing the lo halves (half of the operand-plane traffic of the ViT)
  struct LoGuard {
    Ctx* c;
    ~LoGuard() { c->lo_unused = false; }
  } lo_guard{c};
  c->lo_unused = (c->nsplit == 1) && fused_attn;
  SB ln = sb(c, R * D);
  SB qk{nullptr, nullptr}, vt{nullptr, nullptr}, P{nullptr, nullptr}, qkv{nullptr, nullptr};
  float* S = nullptr;
  if (fused_attn) {
    qkv = sb(c, R * 3 * D);
  } else {
    qk = sb(c, R * 2 * D);
    vt = sb(c, static_cast<size_t>(B) * D * TP);
    S = ws.get<float>(static_cast<size_t>(B) * H * T * TP);
    P = sb(c, static_cast<size_t>(B) * H * T * TP);
  }
  SB o = sb(c, R * D);
  SB hid = sb(c, R * 4 * D);
  for (int b = 0; b < m->depth; ++b) {
    const MgpBlock& w = m->blocks[b];
    gather_ln(c, x, D, nullptr, 1, D, R, w.n1.g, w.n1.b, 1e-6f, false, nullptr, 0, nullptr, 0, ln.hi, ln.lo, D, nullptr,
              nullptr);
    if (fused_attn) {
      // one packed q|k|v projection, then scores + softmax + P.V on tcgen05 with S and P in tensor memory (attn_tc.cu)
      lin(c, ln, R, w.qkv, ACT_NONE, nullptr, &qkv, nullptr);
      attention_tc(c, qkv.hi, qkv.lo, 3L * D, B, T, H, o.hi, o.lo, nullptr, D);
    } else {
    lin(c, ln, R, w.qk, ACT_NONE, nullptr, &qk, nullptr);
    {  // V^T[b, f, t] = Wv ln_b^T + bv  (feature-major: P.V becomes a K-major GEMM)
      Operand bop = opnd(ln.hi, ln.lo, T, D, D);
      bop.nb1 = B; bop.bs1 = static_cast<long>(T) * D;
      Epilogue e;
      e.out_hi = vt.hi; e.out_lo = vt.lo; e.ldo = TP; e.obs1 = static_cast<long>(D) * TP;
      e.bias = w.v.b; e.bias_mode = BIAS_ROW;
      gemm(c, w.v.w.op(), bop, e);
    }
    {  // scores = (q k^T) * 64^-0.5
      Operand q = opnd(qk.hi, qk.lo, T, 64, 2 * D);
      q.nb0 = H; q.bs0 = 64; q.nb1 = B; q.bs1 = static_cast<long>(T) * 2 * D;
      Operand k = q;
      k.hi = qk.hi + D; k.lo = qk.lo + D;
      Epilogue e;
      e.out_f32 = S; e.ldo = TP; e.obs0 = static_cast<long>(T) * TP; e.obs1 = static_cast<long>(H) * T * TP;
      e.alpha = 0.125f;
      gemm(c, q, k, e);
    }
    softmax_rows(c, S, TP, static_cast<long>(B) * H * T, T, nullptr, 1, 0, nullptr, P.hi, P.lo, TP);
    {
      Operand p = opnd(P.hi, P.lo, T, T, TP);
      p.nb0 = H; p.bs0 = static_cast<long>(T) * TP; p.nb1 = B; p.bs1 = static_cast<long>(H) * T * TP;
      Operand v = opnd(vt.hi, vt.lo, 64, T, TP);
      v.nb0 = H; v.bs0 = static_cast<long>(64) * TP; v.nb1 = B; v.bs1 = static_cast<long>(D) * TP;
      Epilogue e;
      e.out_hi = o.hi; e.out_lo = o.lo; e.ldo = D; e.obs0 = 64; e.obs1 = static_cast<long>(T) * D;
      gemm(c, p, v, e);
    }
    }  // unfused attention
    lin(c, o, R, w.proj, ACT_NONE, x, nullptr, x);
    gather_ln(c, x, D, nullptr, 1, D, R, w.n2.g, w.n2.b, 1e-6f, false, nullptr, 0, nullptr, 0, ln.hi, ln.lo, D, nullptr,
              nullptr);
    lin(c, ln, R, w.fc1, ACT_GELU, nullptr, &hid, nullptr);
    lin(c, hid, R, w.fc2, ACT_NONE, x, nullptr, x);
  }
  // ---- three A^3 modules + heads (token_learner.py:21-32, mgp_str.py:79-92)
  const size_t a3_mark = ws.mark();
  float* logit_dst[3] = {char_logits, bpe_logits, wp_logits};
  const int G = 8, Dg = D / 8;
  for (int a = 0; a < m->n_a3; ++a) {
    ws.release(a3_mark);
    const MgpA3& w = m->a3[a];
    SB tn = sb(c, R * D);
    SB s0 = sb(c, R * D);
    float* sel = ws.get<float>(static_cast<size_t>(B) * NTOK * TP);
    float* attn = ws.get<float>(static_cast<size_t>(B) * NTOK * TP);
    SB Pa = sb(c, static_cast<size_t>(B) * NTOK * TP);
    SB ft = sb(c, static_cast<size_t>(B) * D * TP);
    float* xa = ws.get<float>(static_cast<size_t>(B) * NTOK * D);
    SB xn = sb(c, static_cast<size_t>(B) * NTOK * D);
    float* logits = ws.get<float>(static_cast<size_t>(B) * NTOK * w.Vpad);
    int* ids = ws.get<int>(static_cast<size_t>(B) * NTOK);
    float* prob = ws.get<float>(static_cast<size_t>(B) * NTOK);
    gather_ln(c, x, D, nullptr, 1, D, R, w.token_norm.g, w.token_norm.b, 1e-5f, false, nullptr, 0, nullptr, 0, tn.hi,
              tn.lo, D, nullptr, nullptr);
    {  // grouped 1x1 conv (8 groups): per group a [R, Dg] x [Dg, Dg]^T GEMM
      Operand aop = opnd(tn.hi, tn.lo, static_cast<int>(R), Dg, D);
      aop.nb0 = G; aop.bs0 = Dg;
      Operand wop = opnd(w.w0.hi, w.w0.lo, Dg, Dg, Dg);
      wop.nb0 = G; wop.bs0 = static_cast<long>(Dg) * Dg;
      Epilogue e;
      e.out_hi = s0.hi; e.out_lo = s0.lo; e.ldo = D; e.obs0 = Dg;
      gemm(c, aop, wop, e);
    }
    {  // selected^T[b, s, t] = W1 s0_b^T
      Operand bop = opnd(s0.hi, s0.lo, T, D, D);
      bop.nb1 = B; bop.bs1 = static_cast<long>(T) * D;
      Epilogue e;
      e.out_f32 = sel; e.ldo = TP; e.obs1 = static_cast<long>(NTOK) * TP;
      gemm(c, w.w1.op(), bop, e);
    }
    softmax_rows(c, sel, TP, static_cast<long>(B) * NTOK, T, nullptr, 1, 0, attn, Pa.hi, Pa.lo, TP);
    {  // feat^T[b, g*Dg + f, t] = Wf_g tn_b[:, g]^T
      Operand aop = opnd(w.feat.hi, w.feat.lo, Dg, Dg, Dg);
      aop.nb0 = G; aop.bs0 = static_cast<long>(Dg) * Dg;
      Operand bop = opnd(tn.hi, tn.lo, T, Dg, D);
      bop.nb0 = G; bop.bs0 = Dg; bop.nb1 = B; bop.bs1 = static_cast<long>(T) * D;
      Epilogue e;
      e.out_hi = ft.hi; e.out_lo = ft.lo; e.ldo = TP; e.obs0 = static_cast<long>(Dg) * TP; e.obs1 = static_cast<long>(D) * TP;
      gemm(c, aop, bop, e);
    }
    {  // x_a[b] = selected[b] (27 x 257) . feat[b] (257 x D)
      Operand aop = opnd(Pa.hi, Pa.lo, NTOK, T, TP);
      aop.nb1 = B; aop.bs1 = static_cast<long>(NTOK) * TP;
      Operand bop = opnd(ft.hi, ft.lo, D, T, TP);
      bop.nb1 = B; bop.bs1 = static_cast<long>(D) * TP;
      Epilogue e;
      e.out_f32 = xa; e.ldo = D; e.obs1 = static_cast<long>(NTOK) * D;
      gemm(c, aop, bop, e);
    }
    gather_ln(c, xa, D, nullptr, 1, D, static_cast<long>(B) * NTOK, w.norm.g, w.norm.b, 1e-5f, false, nullptr, 0, nullptr,
              0, xn.hi, xn.lo, D, nullptr, nullptr);
    {
      Epilogue e;
      e.out_f32 = logits; e.ldo = w.Vpad;
      e.bias = w.head.b; e.bias_mode = BIAS_COL;
      gemm(c, opnd(xn.hi, xn.lo, B * NTOK, D, D), w.head.w.op(), e);
    }
    argmax_prob_kernel<<<B * NTOK, 256, 0, c->stream>>>(logits, w.Vpad, w.V, ids, prob);
    count_launch(c);
    check_launch("argmax_prob");
    const size_t rows = static_cast<size_t>(B) * NTOK;
    if (attn_out)
      ALM_CHECK_CUDA(cudaMemcpy2DAsync(attn_out + static_cast<size_t>(a) * rows * T, T * sizeof(float), attn,
                                       TP * sizeof(float), T * sizeof(float), rows, cudaMemcpyDeviceToHost, c->stream));
    if (logit_dst[a])
      ALM_CHECK_CUDA(cudaMemcpy2DAsync(logit_dst[a], w.V * sizeof(float), logits, w.Vpad * sizeof(float),
                                       w.V * sizeof(float), rows, cudaMemcpyDeviceToHost, c->stream));
    if (ids_out)
      ALM_CHECK_CUDA(cudaMemcpyAsync(ids_out + a * rows, ids, rows * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (prob_out)
      ALM_CHECK_CUDA(cudaMemcpyAsync(prob_out + a * rows, prob, rows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ALM_CHECK_CUDA(cudaStreamSynchronize(c->stream));  // buffers are reused by the next head
  }
}

}  // namespace alm
