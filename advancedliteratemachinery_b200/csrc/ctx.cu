// Context, weight slab and host->device weight conversion.
#include <string.h>

#include "alm_internal.h"

namespace alm {

void* Ctx::walloc(size_t bytes) {
  size_t a = (woff + 255) & ~size_t(255);
  if (wbase == nullptr || a + bytes > wcap) {
    const size_t slab = std::max<size_t>(size_t(256) << 20, bytes);
    void* p = nullptr;
    if (cudaMalloc(&p, slab) != cudaSuccess)
      throw AlmError{ALM_ERR_OOM, "cudaMalloc of a " + std::to_string(slab >> 20) + " MiB weight slab failed"};
    if (!wstore) {
      wstore = std::make_shared<WeightStore>();
      wstore->device = device;
    }
    wstore->slabs.push_back(p);
    wstore->used.push_back(0);
    wbase = static_cast<char*>(p);
    wcap = slab;
    a = 0;
  }
  woff = a + bytes;
  wstore->used.back() = woff;
  return wbase + a;
}

void Ctx::ensure_ws() {
  if (ws.base) return;
  void* p = nullptr;
  if (cudaMalloc(&p, ws_bytes) != cudaSuccess)
    throw AlmError{ALM_ERR_OOM, "cudaMalloc of the " + std::to_string(ws_bytes >> 20) + " MiB workspace failed"};
  ws.base = static_cast<char*>(p);
  ws.cap = ws_bytes;
  ws.off = 0;
}

const HostTensor& need(const std::map<std::string, HostTensor>& m, const std::string& k) {
  auto it = m.find(k);
  if (it == m.end()) throw AlmError{ALM_ERR_INVALID, "missing tensor in state dict: " + k};
  return it->second;
}

static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

float* upload_f32(Ctx* c, const float* h, size_t n) {
  float* d = static_cast<float*>(c->walloc(n * sizeof(float)));
  ALM_CHECK_CUDA(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

int* upload_i32(Ctx* c, const std::vector<int>& v) {
  int* d = static_cast<int*>(c->walloc(v.size() * sizeof(int)));
  ALM_CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return d;
}

SplitW upload_split(Ctx* c, const float* w, int N, int K, int Kpad) {
  if (Kpad <= 0) Kpad = (K + 7) & ~7;
  std::vector<uint16_t> hi(static_cast<size_t>(N) * Kpad, 0), lo(static_cast<size_t>(N) * Kpad, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float x = w[static_cast<size_t>(n) * K + k];
      const uint16_t h = f2bf(x);
      hi[static_cast<size_t>(n) * Kpad + k] = h;
      lo[static_cast<size_t>(n) * Kpad + k] = f2bf(x - bf2f(h));
    }
  SplitW s;
  s.N = N; s.K = Kpad; s.ld = Kpad;
  // one allocation for both planes: the GEMM fetches hi and lo with a single 5-D TMA box, so lo must sit at a
  // fixed positive distance after hi
  const size_t plane = (hi.size() * 2 + 255) & ~size_t(255);
  s.hi = static_cast<bf16*>(c->walloc(2 * plane));
  s.lo = reinterpret_cast<bf16*>(reinterpret_cast<char*>(s.hi) + plane);
  ALM_CHECK_CUDA(cudaMemcpy(s.hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
  ALM_CHECK_CUDA(cudaMemcpy(s.lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  return s;
}

}  // namespace alm
