// lfs_b200 -- device-wide exclusive scan + stable LSD radix sort (see sort_scan.cuh).
#include "sort_scan.cuh"

namespace lfs {

__device__ __forceinline__ uint32_t resolve_n(uint32_t n_cap, const uint32_t* n_dev) {
    if (n_dev) {
        const uint32_t n = *n_dev;
        return n < n_cap ? n : n_cap;
    }
    return n_cap;
}

// inclusive scan of one value per thread across a block of `blockDim.x` (<= 1024) threads.
// s_warp must hold 33 uint32.  Returns inclusive prefix; *block_total gets the block sum.
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (blockDim.x + 31) >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
            x += y;
    }
    if (lane == 31)
        s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < n_warps) ? s_warp[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o)
                w += y;
        }
        s_warp[lane] = w; // inclusive over warps
        if (lane == 31)
            s_warp[32] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (warp == 0) ? 0u : s_warp[warp - 1];
    *block_total = s_warp[32];
    const uint32_t r = x + warp_off;
    __syncthreads(); // s_warp may be reused by the caller
    return r;
}

// ------------------------------------------------------------------------------------------------- scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t scan_fetch(const uint32_t* in, const uint32_t* gather, uint32_t i) {
    return gather ? __ldg(in + __ldg(gather + i)) : __ldg(in + i);
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_block_sums(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather, uint32_t n_cap,
                      const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_warp[33];
    const uint32_t n = resolve_n(n_cap, n_dev);
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n)
            s += scan_fetch(in, gather, base + k);
    uint32_t total;
    block_inclusive_scan(s, s_warp, &total);
    if (threadIdx.x == 0)
        sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the block sums in place, writes the grand total
__global__ void __launch_bounds__(1024)
    k_scan_sums(uint32_t* __restrict__ sums, uint32_t n_cap, const uint32_t* __restrict__ n_dev,
                uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[33];
    const uint32_t n = resolve_n(n_cap, n_dev);
    const uint32_t nblk = (n + kScanTile - 1) / kScanTile;
    uint32_t carry = 0;
    for (uint32_t start = 0; start < nblk; start += blockDim.x) {
        const uint32_t i = start + threadIdx.x;
        const uint32_t v = (i < nblk) ? sums[i] : 0u;
        uint32_t total;
        const uint32_t incl = block_inclusive_scan(v, s_warp, &total);
        if (i < nblk)
            sums[i] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out)
        *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_apply(const uint32_t* __restrict__ in, const uint32_t* __restrict__ gather, uint32_t* __restrict__ out,
                 const uint32_t* __restrict__ sums, uint32_t n_cap, const uint32_t* __restrict__ n_dev) {
    __shared__ uint32_t s_warp[33];
    const uint32_t n = resolve_n(n_cap, n_dev);
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? scan_fetch(in, gather, base + k) : 0u;
        s += v[k];
    }
    uint32_t total;
    const uint32_t incl = block_inclusive_scan(s, s_warp, &total);
    uint32_t run = sums[blockIdx.x] + incl - s;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n)
            out[base + k] = run;
        run += v[k];
    }
}

int exclusive_scan_u32(const uint32_t* in, const uint32_t* gather, uint32_t* out, uint32_t* total_out,
                       uint32_t n_cap, const uint32_t* n_dev, void* scratch, cudaStream_t stream) {
    uint32_t* sums = static_cast<uint32_t*>(scratch);
    const unsigned nblk = div_up(n_cap, kScanTile);
    if (n_cap == 0) {
        if (total_out)
            LFS_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream));
        return LFS_OK;
    }
    k_scan_block_sums<<<nblk, kScanThreads, 0, stream>>>(in, gather, n_cap, n_dev, sums);
    LFS_LAUNCH_OK("k_scan_block_sums");
    k_scan_sums<<<1, 1024, 0, stream>>>(sums, n_cap, n_dev, total_out);
    LFS_LAUNCH_OK("k_scan_sums");
    k_scan_apply<<<nblk, kScanThreads, 0, stream>>>(in, gather, out, sums, n_cap, n_dev);
    LFS_LAUNCH_OK("k_scan_apply");
    return LFS_OK;
}

// ------------------------------------------------------------------------------------------- radix sort
__global__ void __launch_bounds__(kRsThreads)
    k_rs_hist(const uint32_t* __restrict__ keys, uint32_t n_cap, const uint32_t* __restrict__ n_dev, int shift,
              int ndig, uint32_t nblk, uint32_t* __restrict__ table) {
    extern __shared__ uint32_t sh[]; // [kHistCopies][ndig] privatised copies (fewer same-address conflicts)
    constexpr int kHistCopies = 4;
    const uint32_t n = resolve_n(n_cap, n_dev);
    for (int d = threadIdx.x; d < kHistCopies * ndig; d += kRsThreads)
        sh[d] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kRsTile;
    if (base < n) {
        uint32_t* mine = sh + ((threadIdx.x >> 5) & (kHistCopies - 1)) * ndig;
        const uint32_t mask = (uint32_t)(ndig - 1);
        if (base + kRsTile <= n && (reinterpret_cast<uintptr_t>(keys + base) & 15u) == 0) { // full tile: 128-bit loads
            const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
#pragma unroll
            for (int r = 0; r < kRsItems / 4; ++r) {
                const uint4 v = __ldg(k4 + r * kRsThreads + threadIdx.x);
                atomicAdd(&mine[(v.x >> shift) & mask], 1u);
                atomicAdd(&mine[(v.y >> shift) & mask], 1u);
                atomicAdd(&mine[(v.z >> shift) & mask], 1u);
                atomicAdd(&mine[(v.w >> shift) & mask], 1u);
            }
        } else {
            for (int r = 0; r < kRsItems; ++r) {
                const uint32_t i = base + r * kRsThreads + threadIdx.x;
                if (i < n)
                    atomicAdd(&mine[(__ldg(keys + i) >> shift) & mask], 1u);
            }
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < ndig; d += kRsThreads)
        table[(size_t)d * nblk + blockIdx.x] = sh[d] + sh[ndig + d] + sh[2 * ndig + d] + sh[3 * ndig + d];
}

// grid = ndig CTAs: exclusive scan of row d over the blocks, totals[d] = row sum
__global__ void __launch_bounds__(256)
    k_rs_scan_rows(uint32_t* __restrict__ table, uint32_t nblk, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[33];
    uint32_t* row = table + (size_t)blockIdx.x * nblk;
    uint32_t carry = 0;
    for (uint32_t start = 0; start < nblk; start += blockDim.x) {
        const uint32_t i = start + threadIdx.x;
        const uint32_t v = (i < nblk) ? row[i] : 0u;
        uint32_t total;
        const uint32_t incl = block_inclusive_scan(v, s_warp, &total);
        if (i < nblk)
            row[i] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0)
        totals[blockIdx.x] = carry;
}

// one CTA of 1024 threads: base[d] = exclusive scan of totals (ndig <= 2048)
__global__ void __launch_bounds__(1024)
    k_rs_scan_totals(const uint32_t* __restrict__ totals, uint32_t* __restrict__ base, int ndig) {
    __shared__ uint32_t s_warp[33];
    const int d0 = 2 * threadIdx.x, d1 = d0 + 1;
    const uint32_t a = d0 < ndig ? totals[d0] : 0u;
    const uint32_t b = d1 < ndig ? totals[d1] : 0u;
    uint32_t total;
    const uint32_t incl = block_inclusive_scan(a + b, s_warp, &total);
    const uint32_t excl = incl - (a + b);
    if (d0 < ndig)
        base[d0] = excl;
    if (d1 < ndig)
        base[d1] = excl + a;
}

// MODE 0: peers of a lane by MATCH.ANY, (key, value) pairs held in registers.
// MODE 1: peers by one ballot per digit bit -- a fixed cost, where MATCH.ANY iterates once per
// DISTINCT value in the warp, and the tile keys of 32 consecutive instances (one Gaussian's run of neighbouring tiles) are
// nearly all distinct -- and the values are read again when the pairs are reordered, which takes 16 registers out of the
// ranking loop (3 CTAs per SM instead of 2).
template <int MODE>
__global__ void __launch_bounds__(kRsThreads, MODE == 1 ? 3 : 2)
    k_rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                 uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n_cap,
                 const uint32_t* __restrict__ n_dev, int shift, int ndig, uint32_t nblk,
                 const uint32_t* __restrict__ table, const uint32_t* __restrict__ base) {
    extern __shared__ uint32_t sh[];
    constexpr int kWarps = kRsThreads / 32;
    __shared__ uint32_t s_scan[33];
    uint32_t* whist = sh;                 // [kWarps][ndig]
    uint32_t* sbase = sh + kWarps * ndig; // [ndig] global start of this block's run of digit d
    uint32_t* dstart = sbase + ndig;      // [ndig] block-local start of digit d
    uint32_t* skey = dstart + ndig;       // [kRsTile]
    uint32_t* sval = skey + kRsTile;      // [kRsTile]
    const uint32_t n = resolve_n(n_cap, n_dev);
    const uint32_t blk_base = blockIdx.x * kRsTile;
    if (blk_base >= n)
        return;
    for (int d = threadIdx.x; d < kWarps * ndig; d += kRsThreads)
        whist[d] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t wbase = blk_base + warp * (kRsItems * 32);
    uint32_t* wh = whist + warp * ndig;
    uint32_t key[kRsItems], val[MODE == 0 ? kRsItems : 1], lrank[kRsItems];
    const int nbits = 31 - __clz(ndig);
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? __ldg(keys_in + i) : 0u;
        if (MODE == 0)
            val[r] = valid ? __ldg(vals_in + i) : 0u;
        const uint32_t d = valid ? ((key[r] >> shift) & (uint32_t)(ndig - 1)) : (uint32_t)ndig;
        uint32_t peers;
        if (MODE == 0) {
            peers = __match_any_sync(0xffffffffu, d);
        } else {
            peers = __ballot_sync(0xffffffffu, valid); // invalid lanes (tail of the last tile) only match each other
            if (!valid)
                peers = ~peers;
#pragma unroll
            for (int b = 0; b < kRsMaxBits; ++b)
                if (b < nbits) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                    peers &= ((d >> b) & 1u) ? m : ~m;
                }
        }
        const uint32_t rank = __popc(peers & lt_mask);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (valid && rank == 0) // one shared atomic per distinct digit; same-address atomics of one warp retire in
            before = atomicAdd(&wh[d], (uint32_t)__popc(peers)); // program order -> ranks stay stable across rounds
        before = __shfl_sync(0xffffffffu, before, leader);
        lrank[r] = before + rank;
    }
    __syncthreads();
    // warp-private counts -> exclusive prefix over the warps; digit totals of this block into dtot
    for (int d = threadIdx.x; d < ndig; d += kRsThreads) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = whist[w * ndig + d];
            whist[w * ndig + d] = run;
            run += t;
        }
        dstart[d] = run;
        sbase[d] = base[d] + table[(size_t)d * nblk + blockIdx.x];
    }
    __syncthreads();
    { // exclusive scan of the digit totals (block-local start of every digit's run)
        const int per = ndig > kRsThreads ? ndig / kRsThreads : 1;
        const int d0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (int k = 0; k < per; ++k)
            if (d0 + k < ndig)
                sum += dstart[d0 + k];
        uint32_t total;
        const uint32_t incl = block_inclusive_scan(sum, s_scan, &total);
        uint32_t run = incl - sum;
        for (int k = 0; k < per; ++k)
            if (d0 + k < ndig) {
                const uint32_t t = dstart[d0 + k];
                dstart[d0 + k] = run;
                run += t;
            }
    }
    __syncthreads();
    // block-local reorder through shared memory, then every digit's run leaves as one contiguous, coalesced segment
    // (before: 32 lanes of a store hit up to 32 different runs -> one 32-B sector per 4-B element)
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & (uint32_t)(ndig - 1);
            const uint32_t lp = dstart[d] + wh[d] + lrank[r];
            skey[lp] = key[r];
            sval[lp] = MODE == 0 ? val[r] : __ldg(vals_in + i);
        }
    }
    __syncthreads();
    const uint32_t n_here = n - blk_base < (uint32_t)kRsTile ? n - blk_base : (uint32_t)kRsTile;
    for (uint32_t t = threadIdx.x; t < n_here; t += kRsThreads) {
        const uint32_t k = skey[t];
        const uint32_t d = (k >> shift) & (uint32_t)(ndig - 1);
        const uint32_t pos = sbase[d] + (t - dstart[d]);
        keys_out[pos] = k;
        vals_out[pos] = sval[t];
    }
}

// ------------------------------------------------------------------------------------------- onesweep
// Single-sweep LSD passes (Adinets & Merrill, "Onesweep", 2022) with 8-bit digits: ONE histogram kernel reads the keys
// once and counts the digits of every pass (the global digit histogram does not depend on the element order), then each
// pass is ONE kernel: a tile of 4096 pairs is ranked in shared memory, publishes its digit counts, obtains its global
// offsets by DECOUPLED LOOK-BACK over the preceding tiles' published counts, and scatters.  Compared with the
// hist / scan-rows / scan-totals / scatter chain above this removes one full read of the keys and three small launches
// per pass.  Tiles take their index from an atomic ticket, so a tile only ever waits for tiles whose CTAs are already
// running (forward progress without relying on the block scheduler's order); the wait is bounded and a timeout raises
// a device-side error flag instead of hanging.
constexpr int kOsBits = 8, kOsDig = 1 << kOsBits, kOsMaxPass = 4;
constexpr uint32_t kOsFlagAgg = 1u << 30, kOsFlagInc = 2u << 30, kOsValMask = (1u << 30) - 1u;

struct OsPlan {
    int n_pass;
    int shift[kOsMaxPass];
    int bits[kOsMaxPass];
};
static inline OsPlan make_os_plan(int begin, int total_bits) {
    OsPlan p;
    p.n_pass = (total_bits + kOsBits - 1) / kOsBits;
    int done = 0;
    for (int i = 0; i < p.n_pass; ++i) { // balanced digits: 13 bits -> 7 + 6
        const int b = (total_bits - done + (p.n_pass - i) - 1) / (p.n_pass - i);
        p.bits[i] = b, p.shift[i] = begin + done;
        done += b;
    }
    return p;
}

// ghist [n_pass][256] digit counts of the whole input (zeroed by the caller)
__global__ void __launch_bounds__(kRsThreads)
    k_os_hist(const uint32_t* __restrict__ keys, uint32_t n_cap, const uint32_t* __restrict__ n_dev, const OsPlan plan,
              uint32_t* __restrict__ ghist) {
    __shared__ uint32_t sh[kOsMaxPass * kOsDig];
    const uint32_t n = resolve_n(n_cap, n_dev);
    for (int d = threadIdx.x; d < plan.n_pass * kOsDig; d += kRsThreads)
        sh[d] = 0;
    __syncthreads();
    const uint32_t n4 = n >> 2;
    const bool al = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;
    auto count = [&](const uint32_t k) {
#pragma unroll
        for (int p = 0; p < kOsMaxPass; ++p)
            if (p < plan.n_pass)
                atomicAdd(&sh[p * kOsDig + ((k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
    };
    if (al) {
        const uint4* k4 = reinterpret_cast<const uint4*>(keys);
        for (uint32_t i = blockIdx.x * kRsThreads + threadIdx.x; i < n4; i += gridDim.x * kRsThreads) {
            const uint4 v = __ldg(k4 + i);
            count(v.x), count(v.y), count(v.z), count(v.w);
        }
        for (uint32_t i = (n4 << 2) + blockIdx.x * kRsThreads + threadIdx.x; i < n; i += gridDim.x * kRsThreads)
            count(__ldg(keys + i));
    } else {
        for (uint32_t i = blockIdx.x * kRsThreads + threadIdx.x; i < n; i += gridDim.x * kRsThreads)
            count(__ldg(keys + i));
    }
    __syncthreads();
    for (int d = threadIdx.x; d < plan.n_pass * kOsDig; d += kRsThreads)
        if (sh[d])
            atomicAdd(&ghist[d], sh[d]);
}

// One pass.  status [n_tiles][256]: (flag << 30) | count, zeroed by the caller; ticket: tile counter of this pass (zeroed);
// err: set to 1 on a look-back timeout.
template <int MODE> // ranking as in k_rs_scatter<MODE>
__global__ void __launch_bounds__(kRsThreads, MODE == 1 ? 3 : 2)
    k_os_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
              uint32_t* __restrict__ vals_out, uint32_t n_cap, const uint32_t* __restrict__ n_dev, const int shift,
              const int bits, const uint32_t* __restrict__ ghist /* [256] of this pass */, uint32_t* __restrict__ status,
              uint32_t* __restrict__ ticket, uint32_t* __restrict__ err) {
    constexpr int kWarps = kRsThreads / 32;
    __shared__ uint32_t whist[kWarps * kOsDig]; // warp-private digit counts -> exclusive prefix over the warps
    __shared__ uint32_t sbase[kOsDig];          // global position of this tile's run of digit d
    __shared__ uint32_t dstart[kOsDig];         // tile-local start of digit d
    __shared__ uint32_t skey[kRsTile], sval[kRsTile];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tile;
    const uint32_t n = resolve_n(n_cap, n_dev);
    if (threadIdx.x == 0)
        s_tile = atomicAdd(ticket, 1u);
    for (int d = threadIdx.x; d < kWarps * kOsDig; d += kRsThreads)
        whist[d] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t blk_base = tile * kRsTile;
    if (blk_base >= n)
        return;
    const uint32_t dmask = (1u << bits) - 1u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t wbase = blk_base + warp * (kRsItems * 32);
    uint32_t* wh = whist + warp * kOsDig;
    uint32_t key[kRsItems], val[MODE == 0 ? kRsItems : 1], lrank[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? __ldg(keys_in + i) : 0u;
        if (MODE == 0)
            val[r] = valid ? __ldg(vals_in + i) : 0u;
        const uint32_t d = valid ? ((key[r] >> shift) & dmask) : (uint32_t)kOsDig;
        uint32_t peers;
        if (MODE == 0) {
            peers = __match_any_sync(0xffffffffu, d);
        } else {
            peers = __ballot_sync(0xffffffffu, valid);
            if (!valid)
                peers = ~peers;
#pragma unroll
            for (int b = 0; b < kOsBits; ++b)
                if (b < bits) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                    peers &= ((d >> b) & 1u) ? m : ~m;
                }
        }
        const uint32_t rank = __popc(peers & lt_mask);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (valid && rank == 0) // same-address atomics of one warp retire in program order -> ranks stay stable
            before = atomicAdd(&wh[d], (uint32_t)__popc(peers));
        before = __shfl_sync(0xffffffffu, before, leader);
        lrank[r] = before + rank;
    }
    __syncthreads();
    // one thread per digit: counts of this tile, publish, look back, global base
    {
        const int d = threadIdx.x; // kRsThreads == kOsDig
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = whist[w * kOsDig + d];
            whist[w * kOsDig + d] = run;
            run += t;
        }
        dstart[d] = run; // count; scanned below
        volatile uint32_t* st = status + (size_t)tile * kOsDig + d;
        if (tile == 0) {
            *st = kOsFlagInc | run;
        } else {
            *st = kOsFlagAgg | run;
        }
        __threadfence();
        uint32_t excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            uint32_t spins = 0;
            while (true) {
                const uint32_t v = *(volatile uint32_t*)(status + (size_t)t * kOsDig + d);
                const uint32_t flag = v & ~kOsValMask;
                if (flag == 0) {
                    if (++spins > (1u << 26)) { // ~ seconds: a predecessor never published -> report, do not hang
                        *err = 1u;
                        break;
                    }
                    continue;
                }
                excl += v & kOsValMask;
                if (flag == kOsFlagInc || t == 0)
                    break;
                --t;
            }
            __threadfence();
            *st = kOsFlagInc | ((excl + run) & kOsValMask);
        }
        // exclusive scan of the global digit histogram (256 values, every tile redundantly: cheaper than a launch)
        uint32_t total;
        const uint32_t gh = __ldg(ghist + d);
        const uint32_t gincl = block_inclusive_scan(gh, s_scan, &total);
        sbase[d] = gincl - gh + excl;
        // tile-local starts
        const uint32_t lincl = block_inclusive_scan(run, s_scan, &total);
        dstart[d] = lincl - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & dmask;
            const uint32_t lp = dstart[d] + wh[d] + lrank[r];
            skey[lp] = key[r];
            sval[lp] = MODE == 0 ? val[r] : __ldg(vals_in + i);
        }
    }
    __syncthreads();
    const uint32_t n_here = n - blk_base < (uint32_t)kRsTile ? n - blk_base : (uint32_t)kRsTile;
    for (uint32_t t = threadIdx.x; t < n_here; t += kRsThreads) {
        const uint32_t k = skey[t];
        const uint32_t d = (k >> shift) & dmask;
        const uint32_t pos = sbase[d] + (t - dstart[d]);
        keys_out[pos] = k;
        vals_out[pos] = sval[t];
    }
}

static int g_sort_variant = 0;
void set_sort_variant(int v) { g_sort_variant = v; }

static int radix_sort_pairs_onesweep(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n_cap,
                                     const uint32_t* n_dev, int begin_bit, int n_bits, void* scratch, int* result_in_b,
                                     cudaStream_t stream) {
    static_assert(kRsThreads == kOsDig, "one thread per digit");
    const OsPlan plan = make_os_plan(begin_bit, n_bits);
    const uint32_t ntiles = div_up(n_cap, kRsTile);
    // scratch layout: ghist [4][256] | ticket [4] | err [1] | pad | status [n_pass][ntiles][256]
    uint32_t* ghist = static_cast<uint32_t*>(scratch);
    uint32_t* ticket = ghist + kOsMaxPass * kOsDig;
    uint32_t* err = ticket + kOsMaxPass;
    uint32_t* status = ghist + kOsMaxPass * kOsDig + 64;
    const size_t zero_bytes = sizeof(uint32_t) * (kOsMaxPass * kOsDig + 64 + (size_t)plan.n_pass * ntiles * kOsDig);
    LFS_CUDA_OK(cudaMemsetAsync(scratch, 0, zero_bytes, stream));
    const unsigned hgrid = ntiles < (unsigned)(num_sms() * 8) ? ntiles : (unsigned)(num_sms() * 8);
    k_os_hist<<<hgrid, kRsThreads, 0, stream>>>(keys_a, n_cap, n_dev, plan, ghist);
    LFS_LAUNCH_OK("k_os_hist");
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < plan.n_pass; ++p) {
if (g_sort_variant == 4)
                    k_os_pass<1><<<ntiles, kRsThreads, 0, stream>>>(kin, vin, kout, vout, n_cap, n_dev, plan.shift[p], plan.bits[p],
                                                     ghist + p * kOsDig, status + (size_t)p * ntiles * kOsDig, ticket + p, err);
        else
                    k_os_pass<0><<<ntiles, kRsThreads, 0, stream>>>(kin, vin, kout, vout, n_cap, n_dev, plan.shift[p], plan.bits[p],
                                                     ghist + p * kOsDig, status + (size_t)p * ntiles * kOsDig, ticket + p, err);
        LFS_LAUNCH_OK("k_os_pass");
        uint32_t* t = kin;
        kin = kout, kout = t;
        t = vin;
        vin = vout, vout = t;
    }
    *result_in_b = (plan.n_pass & 1);
    return LFS_OK;
}

static size_t rs_scatter_smem(int ndig) {
    return sizeof(uint32_t) * ((size_t)(kRsThreads / 32 + 2) * ndig + 2 * (size_t)kRsTile);
}

int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t n_cap,
                     const uint32_t* n_dev, int begin_bit, int n_bits, void* scratch, int* result_in_b,
                     cudaStream_t stream) {
    *result_in_b = 0;
    if (n_cap == 0 || n_bits <= 0)
        return LFS_OK;
    // variant 1: onesweep for every sort; variant 2: only where it does not add a pass (<= 16 key bits: the tile sort; the
    // 32-bit depth sort keeps three 11-bit passes instead of four 8-bit ones)
    if ((g_sort_variant == 1 || ((g_sort_variant == 2 || g_sort_variant == 4) && n_bits <= 2 * kOsBits)) &&
        n_bits <= kOsBits * kOsMaxPass &&
        n_cap < (1u << 30))
        return radix_sort_pairs_onesweep(keys_a, vals_a, keys_b, vals_b, n_cap, n_dev, begin_bit, n_bits, scratch,
                                         result_in_b, stream);
    // per-device attribute; cheap enough to set on every call (one process may drive several devices)
    LFS_CUDA_OK(cudaFuncSetAttribute(k_rs_scatter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rs_scatter_smem(1 << kRsMaxBits)));
    LFS_CUDA_OK(cudaFuncSetAttribute(k_rs_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rs_scatter_smem(1 << kRsMaxBits)));
    const RadixPlan plan = make_radix_plan(begin_bit, n_bits);
    const uint32_t nblk = div_up(n_cap, kRsTile);
    uint32_t* table = static_cast<uint32_t*>(scratch);
    uint32_t* totals = table + (size_t)(1 << kRsMaxBits) * (nblk + 1);
    uint32_t* base = totals + (1 << kRsMaxBits);
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < plan.n_pass; ++p) {
        const int ndig = 1 << plan.bits[p];
        k_rs_hist<<<nblk, kRsThreads, sizeof(uint32_t) * 4 * ndig, stream>>>(kin, n_cap, n_dev, plan.shift[p], ndig, nblk,
                                                                         table);
        LFS_LAUNCH_OK("k_rs_hist");
        k_rs_scan_rows<<<ndig, 256, 0, stream>>>(table, nblk, totals);
        LFS_LAUNCH_OK("k_rs_scan_rows");
        k_rs_scan_totals<<<1, 1024, 0, stream>>>(totals, base, ndig);
        LFS_LAUNCH_OK("k_rs_scan_totals");
        if (g_sort_variant == 3 || g_sort_variant == 4)
            k_rs_scatter<1><<<nblk, kRsThreads, rs_scatter_smem(ndig), stream>>>(
                kin, vin, kout, vout, n_cap, n_dev, plan.shift[p], ndig, nblk, table, base);
        else
            k_rs_scatter<0><<<nblk, kRsThreads, rs_scatter_smem(ndig), stream>>>(
                kin, vin, kout, vout, n_cap, n_dev, plan.shift[p], ndig, nblk, table, base);
        LFS_LAUNCH_OK("k_rs_scatter");
        uint32_t* t = kin;
        kin = kout;
        kout = t;
        t = vin;
        vin = vout;
        vout = t;
    }
    *result_in_b = (plan.n_pass & 1);
    return LFS_OK;
}

} // namespace lfs
