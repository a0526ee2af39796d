// Evidence for profiles/r02_ref_fastgs_root_cause.txt: cub::DeviceRadixSort::SortPairs(DoubleBuffer<ushort>,
// DoubleBuffer<uint>) temp-storage size as QUERIED by the reference (size_t num_items, buffer_utils.h:101-104) and as
// NEEDED by its sort call (int num_items, forward.cu:141-146), on the device this runs on.
//   nvcc -std=c++20 -gencode arch=compute_100a,code=sm_100a -o cub_query_sizes cub_query_sizes.cu      (fails above ~470 k)
//   nvcc -std=c++20 -gencode arch=compute_90,code=compute_90 -o cub_query_sizes_ptx90 cub_query_sizes.cu (query >= need)
#include <cstdio>
#include <cub/cub.cuh>
template <typename N> size_t query(N n) {
    cub::DoubleBuffer<unsigned short> k(nullptr, nullptr);
    cub::DoubleBuffer<unsigned> v(nullptr, nullptr);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, n);
    return bytes;
}
int main() {
    int ptx = 0;
    cub::PtxVersion(ptx);
    for (long long n : {38943LL, 155106LL, 390059LL, 775065LL, 7377576LL}) {
        const size_t q = query<size_t>((size_t)n), need = query<int>((int)n);
        printf("{\"cub_ptx_version\":%d,\"num_items\":%lld,\"queried_as_size_t\":%zu,\"needed_as_int\":%zu,\"enough\":%s}\n", ptx,
               n, q, need, q >= need ? "true" : "false");
    }
    return 0;
}
