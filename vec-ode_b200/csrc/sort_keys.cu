// sort_keys.cu — the radix sort behind the dynamic grouping of the exponential integrators (exp.cu: exp_dynamic_order). CUB's device
// radix sort over (float key, int index) pairs; kept in its own translation unit because CUB's headers pull <unistd.h>, whose pipe()
// collides with namespace pipe of tile_pipe.cuh.
#include <cub/device/device_radix_sort.cuh>

#include <cstddef>

size_t vo_sort_pairs_tmp_bytes(int n) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const float*)nullptr, (float*)nullptr, (const int*)nullptr, (int*)nullptr, n, 16, 32, (cudaStream_t)0);
    return tmp;
}

cudaError_t vo_sort_pairs(void* tmp, size_t tmp_bytes, const float* key_in, float* key_out, const int* idx_in, int* idx_out, int n, cudaStream_t stream) {
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, idx_in, idx_out, n, 16, 32, stream);  // the top 16 bits of the key (sign, exponent, 7 mantissa bits) order the systems finely enough: 2 passes instead of 4
}
