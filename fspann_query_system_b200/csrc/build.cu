// build.cu -- Setup-side index build on the device.
//
// Replaces, for bulk Setup, PartitionedIndexService.finalizeForSearch's per-division build (index/.../paper/
// PartitionedIndexService.java:412-425, 789-845) -> GreedyPartitioner.build (index/.../paper/GreedyPartitioner.java:37-76):
//   entries   = HashMap<String, BitSet>(N) iteration order of the staged (id, code) pairs (PIS:413-420):
//               bucket = (h ^ h>>>16) & (cap-1) of String.hashCode(decimal id), ascending; insertion order inside a bucket;
//   sort      = List.sort by computeKey(code), stable (GP:51, GP:87-96);
//   partition = consecutive blocks of 64; minKey / maxKey = keys of the block's first / last entry; repCode = code of the
//               entry at i + ((end - i - 1) >>> 1) (GP:55-72).
// Both sorts are stable LSD radix sorts (cub::DeviceRadixSort -- library code, Setup only, not on the query path): one over
// the Java buckets (shared by all T*D divisions), one per division over the 63-bit keys carrying the iteration position.
#include <cub/device/device_radix_sort.cuh>

#include "fspann_internal.cuh"

namespace fsp {

__global__ void build_bucket_kernel(const int32_t *__restrict__ staged, int64_t n, uint32_t cap_mask, uint32_t *__restrict__ bucket,
                                    int32_t *__restrict__ pos) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bucket[i] = java_hash_decimal(staged[i]) & cap_mask;
        pos[i] = (int32_t)i;
    }
}
// ids_it[i] = staged[order[i]]; also flags a bin of >= 9 entries (java.util.HashMap would treeify it: iteration order unspecified)
__global__ void build_iter_kernel(const int32_t *__restrict__ staged, const int32_t *__restrict__ order, const uint32_t *__restrict__ bucket_sorted,
                                  int64_t n, int32_t *__restrict__ ids_it, int32_t *__restrict__ treeified) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ids_it[i] = staged[order[i]];
        if (i + 8 < n && bucket_sorted[i] == bucket_sorted[i + 8]) *treeified = 1;
    }
}
// key of every entry in iteration order for division g: GP:87-96 (code bit i < 63 -> key bit 62 - i)
__global__ void build_keys_kernel(const uint64_t *__restrict__ codes, const int32_t *__restrict__ ids_it, int64_t n, int TD, int W, int g,
                                  uint64_t *__restrict__ keys, int32_t *__restrict__ pos) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t c0 = codes[((size_t)ids_it[i] * TD + g) * W];
        keys[i] = __brevll(c0) >> 1;
        pos[i] = (int32_t)i;
    }
}
// ids of division g in partition order + per-partition (minKey, maxKey) and representative code
__global__ void build_emit_kernel(const uint64_t *__restrict__ codes, const int32_t *__restrict__ ids_it, const uint64_t *__restrict__ keys_sorted,
                                  const int32_t *__restrict__ order, int64_t n, int TD, int W, int g, int64_t P, int32_t *__restrict__ ids_out,
                                  int64_t *__restrict__ keys_out /* [P][2] */, uint64_t *__restrict__ rep_out /* [P][W] */) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t id = ids_it[order[i]];
        ids_out[i] = id;
        const int64_t p = i / kBlock, start = p * kBlock, end = min(n, start + (int64_t)kBlock);
        if (i == start) keys_out[2 * p] = (int64_t)keys_sorted[i];
        if (i == end - 1) keys_out[2 * p + 1] = (int64_t)keys_sorted[i];
        if (i == start + ((end - start - 1) >> 1))
            for (int w = 0; w < W; w++) rep_out[(size_t)p * W + w] = codes[((size_t)id * TD + g) * W + w];
    }
}

// staged[i] = the facade's insertion order into the staging map: ids 999 .. n-1, then 0 .. 998 (PIS:280-298, 821-831); iota: out[i] = i
__global__ void staged_order_kernel(int32_t *__restrict__ staged, int64_t n, int iota) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        staged[i] = iota ? (int32_t)i : (int32_t)(i < n - 999 ? i + 999 : i - (n - 999));
}
int launch_staged_order(cudaStream_t s, int32_t *staged, int64_t n) {
    if (n <= 0) return 0;
    staged_order_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)cur_sm_count() * 16), 256, 0, s>>>(staged, n, 0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_iota(cudaStream_t s, int32_t *out, int64_t n) {
    if (n <= 0) return 0;
    staged_order_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)cur_sm_count() * 16), 256, 0, s>>>(out, n, 1);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

static int bits_for(uint64_t maxv) { int b = 1; while (b < 64 && (maxv >> b)) b++; return b; }

// codes [N][TD][W] (by id), staged [N] on the device -> ids [TD][N], keys [TD][P][2], rep [TD][P][W] on the device.
// scratch: bucket/keys (8N), keys_alt (8N), pos (4N), pos_alt (4N), ids_it (4N), cub temp.  Returns kernels launched or -1.
int launch_partition_build(cudaStream_t s, const uint64_t *codes, const int32_t *staged, int64_t n, int TD, int W, uint32_t cap, int code_bits,
                           int32_t *ids_out, int64_t *keys_out, uint64_t *rep_out, void *scratch, size_t scratch_bytes, int32_t *treeified) {
    const int64_t P = (n + kBlock - 1) / kBlock;
    unsigned char *sp = (unsigned char *)scratch;
    auto take = [&](size_t bytes) { unsigned char *p = sp; sp += (bytes + 255) / 256 * 256; return (void *)p; };
    uint64_t *k_a = (uint64_t *)take(8 * (size_t)n), *k_b = (uint64_t *)take(8 * (size_t)n);
    int32_t *p_a = (int32_t *)take(4 * (size_t)n), *p_b = (int32_t *)take(4 * (size_t)n), *ids_it = (int32_t *)take(4 * (size_t)n);
    size_t tmp64 = 0, tmp32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp64, k_a, k_b, p_a, p_b, (int)n, 0, 64, s);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp32, (uint32_t *)k_a, (uint32_t *)k_b, p_a, p_b, (int)n, 0, 32, s);
    const size_t tmp = std::max(tmp64, tmp32);
    void *cub_tmp = take(tmp);
    if ((size_t)(sp - (unsigned char *)scratch) > scratch_bytes) return -1;
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)cur_sm_count() * 16);
    int launched = 0;
    // 1. HashMap iteration order (shared by every division)
    build_bucket_kernel<<<grid, 256, 0, s>>>(staged, n, cap - 1u, (uint32_t *)k_a, p_a);
    size_t t = tmp;
    if (cub::DeviceRadixSort::SortPairs(cub_tmp, t, (uint32_t *)k_a, (uint32_t *)k_b, p_a, p_b, (int)n, 0, bits_for(cap - 1u), s) != cudaSuccess) return -1;
    build_iter_kernel<<<grid, 256, 0, s>>>(staged, p_b, (const uint32_t *)k_b, n, ids_it, treeified);
    launched += 3;
    // 2. per division: stable sort by key, cut into blocks
    const int key_bits = std::min(63, std::max(1, code_bits));
    const int begin_bit = 63 - key_bits;                       // the key keeps the code's low `code_bits` bits in its top bits
    for (int g = 0; g < TD; g++) {
        build_keys_kernel<<<grid, 256, 0, s>>>(codes, ids_it, n, TD, W, g, k_a, p_a);
        t = tmp;
        if (cub::DeviceRadixSort::SortPairs(cub_tmp, t, k_a, k_b, p_a, p_b, (int)n, begin_bit, 63, s) != cudaSuccess) return -1;
        build_emit_kernel<<<grid, 256, 0, s>>>(codes, ids_it, k_b, p_b, n, TD, W, g, P, ids_out + (size_t)g * n, keys_out + (size_t)g * P * 2,
                                               rep_out + (size_t)g * P * W);
        launched += 3;
    }
    return cudaGetLastError() == cudaSuccess ? launched : -1;
}

size_t partition_build_scratch_bytes(int64_t n) {
    size_t tmp64 = 0, tmp32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp64, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 64);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp32, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, 32);
    return 2 * (8 * (size_t)n + 256) + 3 * (4 * (size_t)n + 256) + std::max(tmp64, tmp32) + 1024;
}

}  // namespace fsp
