// lfs_b200 -- device-wide exclusive scan and stable LSD radix sort (u32 keys, u32 values), hand written
// for sm_100a (no CUB).  Both take their element count either from the host (`n_cap`) or from device
// memory (`n_dev`, clamped to n_cap) so the trainer's step needs no host synchronisation.
//
// Radix sort: classic three-kernel pass (block histogram -> per-digit scan -> stable scatter) with up to
// 11-bit digits; ranks inside a block are computed with warp-level __match_any_sync multisplit, which
// keeps every pass stable (ties keep their input order -- the property that makes the two-level
// depth-then-tile sort reproduce the reference's single 46-bit CUB sort bit for bit,
// reference gsplat/IntersectTile.cu:290-328).
#pragma once
#include "common.cuh"

namespace lfs {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems; // 4096 keys per block
constexpr int kRsMaxBits = 11;

struct RadixPlan {
    int n_pass;
    int bits[8];
    int shift[8];
};
// split `total_bits` (starting at bit `begin`) into passes of at most kRsMaxBits bits
static inline RadixPlan make_radix_plan(int begin, int total_bits) {
    RadixPlan p;
    p.n_pass = (total_bits + kRsMaxBits - 1) / kRsMaxBits;
    if (p.n_pass < 1)
        p.n_pass = 1;
    int done = 0;
    for (int i = 0; i < p.n_pass; ++i) {
        int b = (total_bits - done + (p.n_pass - i) - 1) / (p.n_pass - i);
        if (b < 1)
            b = 1;
        p.bits[i] = b;
        p.shift[i] = begin + done;
        done += b;
    }
    return p;
}

// scratch needed by radix_sort_pairs for up to n_cap elements: histogram table + digit bases
static inline size_t radix_scratch_bytes(uint32_t n_cap) {
    const size_t nblk = (n_cap + kRsTile - 1) / kRsTile + 1;
    return align_up(sizeof(uint32_t) * ((size_t)(1 << kRsMaxBits) * nblk + 2 * (1 << kRsMaxBits)), 256);
}

// Sorts (keys, vals) by bits [begin_bit, begin_bit + n_bits) of the key, stable.
// keys_a/vals_a hold the input; keys_b/vals_b are the alternate buffers.  Returns (through *result_in_b)
// whether the sorted data ended in the b buffers.  `scratch` must hold radix_scratch_bytes(n_cap).
void set_sort_variant(int v); // 0: hist / scan / scatter chain, 1 / 2: onesweep (all sorts / short keys only), 3: chain with
                              // ballot ranking in the scatter passes of <= 8 bits
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n_cap,
                     const uint32_t* n_dev, int begin_bit, int n_bits, void* scratch, int* result_in_b,
                     cudaStream_t stream);

// Exclusive scan of uint32 values: out[i] = sum_{j<i} f(j), f(j) = gather ? in[gather[j]] : in[j];
// *total_out (device) receives the grand total.  n from host (n_cap) or device (n_dev, clamped).
// scratch: scan_scratch_bytes(n_cap).
static inline size_t scan_scratch_bytes(uint32_t n_cap) {
    return align_up(sizeof(uint32_t) * ((size_t)(n_cap + 1023) / 1024 + 2), 256);
}
int exclusive_scan_u32(const uint32_t* in, const uint32_t* gather, uint32_t* out, uint32_t* total_out,
                       uint32_t n_cap, const uint32_t* n_dev, void* scratch, cudaStream_t stream);

} // namespace lfs
