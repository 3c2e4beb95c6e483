// Counter-based dropout masks (Philox-4x32, 7 rounds: the cheapest variant that still passes BigCrush). A mask element is a pure function of (seed, stream id, element index):
// the backward kernels regenerate it instead of reading a stored mask.
#pragma once
#include <stdint.h>

namespace dvgr {

template <int kRounds>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < kRounds; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

struct DropoutCfg {
  unsigned long long seed;   // per-forward seed (host side)
  unsigned int stream;       // distinguishes the dropout sites of one step
  float p;                   // drop probability; 0 disables
  const unsigned long long* seed_off;   // optional device counter added to the seed: lets a captured CUDA graph draw
                                        // fresh masks on every replay (dvgr_set_seed_offset)
};

// Keep-scale (0 or 1/(1-p)) for 4 consecutive elements starting at index 4*quad.
__device__ __forceinline__ void dropout_scale4(const DropoutCfg& c, unsigned long long quad, float (&s)[4]) {
  if (c.p <= 0.f) {
    s[0] = s[1] = s[2] = s[3] = 1.f;
    return;
  }
  const unsigned long long seed = c.seed + (c.seed_off != nullptr ? *c.seed_off : 0ull);
  uint4 r = philox4x32<7>(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), c.stream, 0x2545F491u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float inv = 1.f / (1.f - c.p);
  const uint32_t thr = (uint32_t)(c.p * 4294967296.0f);
  s[0] = r.x >= thr ? inv : 0.f;
  s[1] = r.y >= thr ? inv : 0.f;
  s[2] = r.z >= thr ? inv : 0.f;
  s[3] = r.w >= thr ? inv : 0.f;
}

// Keep-scale for 8 consecutive elements starting at index 8*oct: ONE Philox call, 16 random bits per element
// (p is quantised to 1/65536, irrelevant for dropout). This is the form the streaming kernels use.
__device__ __forceinline__ void dropout_scale8(const DropoutCfg& c, unsigned long long oct, float (&s)[8]) {
  if (c.p <= 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 1.f;
    return;
  }
  const unsigned long long seed = c.seed + (c.seed_off != nullptr ? *c.seed_off : 0ull);
  uint4 r = philox4x32<7>(make_uint4((uint32_t)oct, (uint32_t)(oct >> 32), c.stream, 0x8F1BBCDCu),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float inv = 1.f / (1.f - c.p);
  const uint32_t thr = (uint32_t)(c.p * 65536.0f);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i] = (w[i] & 0xFFFFu) >= thr ? inv : 0.f;
    s[2 * i + 1] = (w[i] >> 16) >= thr ? inv : 0.f;
  }
}

// Cheap per-element mask for sites where one thread needs ONE element at a scattered index (the N x N attention
// dropout): murmur3 finaliser over (seed, stream, index) — 10 integer instructions instead of a Philox call.
__device__ __forceinline__ float dropout_scale1_hash(const DropoutCfg& c, unsigned long long idx) {
  if (c.p <= 0.f) return 1.f;
  const unsigned long long seed = c.seed + (c.seed_off != nullptr ? *c.seed_off : 0ull);
  uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u) ^ (c.stream * 0x85EBCA77u) ^
               ((uint32_t)idx * 0xC2B2AE3Du) ^ ((uint32_t)(idx >> 32) * 0x27D4EB2Fu);
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h >= (uint32_t)(c.p * 4294967296.0f) ? 1.f / (1.f - c.p) : 0.f;
}

// Two mask elements (an adjacent column pair) from ONE murmur hash, 16 random bits each (p quantised to 1/65536): the form
// the graph-attention kernels use for their output dropout, where a thread owns column pairs of scattered rows.
__device__ __forceinline__ float2 dropout_scale2_hash(const DropoutCfg& c, unsigned long long pair_idx) {
  if (c.p <= 0.f) return make_float2(1.f, 1.f);
  const unsigned long long seed = c.seed + (c.seed_off != nullptr ? *c.seed_off : 0ull);
  uint32_t h = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u) ^ (c.stream * 0x85EBCA77u) ^
               ((uint32_t)pair_idx * 0xC2B2AE3Du) ^ ((uint32_t)(pair_idx >> 32) * 0x27D4EB2Fu);
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  const uint32_t thr = (uint32_t)(c.p * 65536.0f);
  const float inv = 1.f / (1.f - c.p);
  return make_float2((h & 0xFFFFu) >= thr ? inv : 0.f, (h >> 16) >= thr ? inv : 0.f);
}

// The same two hash masks with everything that does not depend on the element hoisted out of the inner loops (the device
// seed-offset load, the seed / stream mixing, the thresholds): build ONE HashMask per dropout site and thread, then call
// keep1 / keep2 per element. Bit-identical to dropout_scale1_hash / dropout_scale2_hash.
struct HashMask {
  uint32_t base, thr32, thr16;
  float inv;
  bool on;
  __device__ __forceinline__ explicit HashMask(const DropoutCfg& c) {
    on = c.p > 0.f;
    const unsigned long long seed = c.seed + ((on && c.seed_off != nullptr) ? *c.seed_off : 0ull);
    base = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u) ^ (c.stream * 0x85EBCA77u);
    thr32 = (uint32_t)(c.p * 4294967296.0f);
    thr16 = (uint32_t)(c.p * 65536.0f);
    inv = on ? 1.f / (1.f - c.p) : 1.f;
  }
  __device__ __forceinline__ uint32_t mix(unsigned long long idx) const {
    uint32_t h = base ^ ((uint32_t)idx * 0xC2B2AE3Du) ^ ((uint32_t)(idx >> 32) * 0x27D4EB2Fu);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
  }
  __device__ __forceinline__ float keep1(unsigned long long idx) const {
    if (!on) return 1.f;
    return mix(idx) >= thr32 ? inv : 0.f;
  }
  __device__ __forceinline__ float2 keep2(unsigned long long pair_idx) const {
    if (!on) return make_float2(1.f, 1.f);
    const uint32_t h = mix(pair_idx);
    return make_float2((h & 0xFFFFu) >= thr16 ? inv : 0.f, (h >> 16) >= thr16 ? inv : 0.f);
  }
};

// Keep-scale for a single element index (costs a full Philox call; use dropout_scale4 in streaming kernels).
__device__ __forceinline__ float dropout_scale1(const DropoutCfg& c, unsigned long long idx) {
  if (c.p <= 0.f) return 1.f;
  float s[4];
  dropout_scale4(c, idx >> 2, s);
  const int k = (int)(idx & 3);
  return k == 0 ? s[0] : k == 1 ? s[1] : k == 2 ? s[2] : s[3];
}

}  // namespace dvgr
