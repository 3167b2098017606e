// TEST INFRASTRUCTURE ONLY (see ../../cuda_runtime.h): the one cub entry point the EXACT voxel mode calls, as a stable
// host sort with DeviceRadixSort::SortPairs' calling convention (first call: size query with d_temp_storage == NULL).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
struct DeviceRadixSort {
    template <typename K, typename V>
    static cudaError_t SortPairs(void* d_temp_storage, size_t& temp_storage_bytes, const K* keys_in, K* keys_out, const V* values_in,
                                 V* values_out, int num_items, int begin_bit, int end_bit, cudaStream_t) {
        if (d_temp_storage == nullptr) {
            temp_storage_bytes = 256;
            return cudaSuccess;
        }
        const int bits = end_bit - begin_bit;
        const unsigned long long mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
        std::vector<int> order(static_cast<size_t>(num_items));
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return ((static_cast<unsigned long long>(keys_in[a]) >> begin_bit) & mask) <
                   ((static_cast<unsigned long long>(keys_in[b]) >> begin_bit) & mask);
        });
        for (int i = 0; i < num_items; ++i) {
            keys_out[i] = keys_in[order[i]];
            values_out[i] = values_in[order[i]];
        }
        return cudaSuccess;
    }
};
}  // namespace cub
