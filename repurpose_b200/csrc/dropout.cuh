// Counter-based dropout masks for the training step: the nn.Dropout(0.1) sites of the reference's train
// graph — nn.TransformerEncoderLayer's dropout1 / dropout / dropout2 and the attention-weight dropout of its
// nn.MultiheadAttention (models/MMCTransformer.py:41-49), feature_map[3], cls_head[3] / [6], reg_head[3] / [6]
// (:63-93) — under model.train() (main.py:285).
//
// A mask is a pure function of (site keys, element index): nothing is stored for the element-wise sites (the forward
// epilogue and the backward kernel evaluate the same function), and the attention-weight mask is materialised once per
// layer as one bit per (query, key) because the dK/dV kernel walks it transposed.
//
//   pair k = element >> 1;   h = fmix32(k * 0x9E3779B1 + a) ^ b;   element 2k is KEPT iff (h & 0xffff) >= thr,
//   element 2k+1 iff (h >> 16) >= thr;   thr = round(p * 65536)   (p = 0.1 -> 6554: P(drop) = 0.100006)
//
// fmix32 is MurmurHash3's 32-bit finaliser.  (a, b) are derived on the host from (seed, step, site) with the same
// function (repurpose_b200/train.py), so every site and every step draws from its own stream.  torch's own Philox
// stream cannot be reproduced (its offsets depend on ATen's launch geometry): same distribution, different draws —
// the parity tests feed OUR masks to the autograd reference.
#pragma once
#include <stdint.h>

namespace rp {

struct DropKey {
  uint32_t a = 0, b = 0;
  uint32_t thr = 0;     // 0 = dropout off
  float scale = 1.0f;   // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t k, uint32_t a, uint32_t b) {
  uint32_t x = k * 0x9E3779B1u + a;
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x ^ b;
}

// v0, v1 = elements 2k, 2k+1 of the site: scaled if kept, 0 if dropped
__device__ __forceinline__ void drop_pair(uint32_t k, const DropKey& d, float& v0, float& v1) {
  const uint32_t h = drop_hash(k, d.a, d.b);
  v0 = (h & 0xffffu) >= d.thr ? v0 * d.scale : 0.0f;
  v1 = (h >> 16) >= d.thr ? v1 * d.scale : 0.0f;
}

inline DropKey make_drop_key(uint32_t a, uint32_t b, float p) {
  DropKey d;
  if (p > 0.0f) {
    d.a = a;
    d.b = b;
    long t = long(double(p) * 65536.0 + 0.5);
    d.thr = uint32_t(t < 1 ? 1 : (t > 65535 ? 65535 : t));
    d.scale = 1.0f / (1.0f - p);
  }
  return d;
}

}  // namespace rp
