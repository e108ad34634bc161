// abn_drop.cuh -- the dropout keep mask as a pure function of (seed, step, layer, row, col)
// (include/abnet3_b200.h, abn_dropout): evaluated in the epilogues of the forward kernels
// (z * keep / (1 - p) before the activation, abnet3/model.py:136-141) and of the backward
// kernels (dz * keep / (1 - p)), never stored.
#pragma once
#include <stdint.h>

#include "../../include/abnet3_b200.h"

namespace abn {

struct DropArgs {
    const unsigned long long *state;     // device {seed, step}; nullptr: no dropout
    unsigned thresh;                     // keep iff the element's 16-bit draw >= thresh (= p * 65536)
    float inv_keep;                      // 1 / (1 - p)
    int layer;
};

inline DropArgs drop_args(const abn_dropout *d) {
    DropArgs a;
    a.state = nullptr; a.thresh = 0; a.inv_keep = 1.f; a.layer = 0;
    if (d && d->state && d->p > 0.f) {
        a.state = d->state;
        float t = d->p * 65536.f + 0.5f;
        a.thresh = t >= 65535.f ? 65535u : (unsigned)t;
        a.inv_keep = 1.f / (1.f - d->p);
        a.layer = d->layer;
    }
    return a;
}

__host__ __device__ __forceinline__ unsigned long long drop_mix(unsigned long long x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
// the key of one (seed, step, layer): read once per kernel / tile
__device__ __forceinline__ unsigned long long drop_key(const DropArgs &d) {
    const unsigned long long seed = d.state[0], step = d.state[1];
    return drop_mix(seed ^ drop_mix(step * 0x9e3779b97f4a7c15ull + (unsigned long long)(d.layer + 1)));
}
// 64 bits for the four elements (row, 4 q .. 4 q + 3)
__device__ __forceinline__ unsigned long long drop_bits4(unsigned long long key, long long row, int q) {
    return drop_mix(key + (unsigned long long)row * 0x9e3779b97f4a7c15ull + (unsigned long long)(unsigned)q);
}
__device__ __forceinline__ bool drop_keep_of(unsigned long long bits, int r, unsigned thresh) {
    return ((unsigned)(bits >> (16 * r)) & 0xffffu) >= thresh;
}
__device__ __forceinline__ bool drop_keep(unsigned long long key, long long row, int col, unsigned thresh) {
    return drop_keep_of(drop_bits4(key, row, col >> 2), col & 3, thresh);
}

}  // namespace abn
