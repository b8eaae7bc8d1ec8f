// fspann_internal.cuh -- context layout and kernel launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/fspann_gpu.h"
#include "aes_gcm.cuh"

namespace fsp {

constexpr int kMaxKeys = FSPANN_MAX_KEYS;
constexpr int kBlock = FSPANN_BLOCK;  // ids per partition (PIS:92)

// Device-visible key ring: AES-256 round keys per live version (KRS:82-88 getVersion -> key).
struct DevKeyRing {
    int32_t n;
    int32_t version[kMaxKeys];
    uint32_t rk[kMaxKeys][60];
};

// Routing index as the kernels see it.
struct RoutingView {
    int32_t dim, T, D, m, lambda, W, TD;
    int64_t n_ids, P;
    const double *alpha, *r, *omega;  // [TD][m][dim], [TD][m], [TD][m]
    const float *alpha_f32;           // [TD][m][dim] alpha rounded to FP32 (TokenGen pre-filter)
    const float *alpha_norm;          // [TD][m] upper bounds of ||alpha_j||_2
    const uint16_t *alpha_tc;         // alpha as three BF16 pieces in UMMA tile layout (tokengen_tc.cu), or nullptr
    const int64_t *keys;              // [TD][P][2] interleaved (minKey, maxKey)
    const uint64_t *rep;              // [TD][P][W]
    const int32_t *ids;               // [TD][n_ids]
    const uint8_t *deleted;           // [n_deleted] or nullptr
    int64_t n_deleted;
};

// Record store as the kernels see it.  One record = [iv 12 B | key_version 4 B | ciphertext 8*dim B | tag 16 B | pad],
// rec_stride a multiple of 16 so every record starts 128-bit aligned.
struct StoreView {
    int64_t N;            // records held by this context: global ids [id_base, id_base + N)
    int64_t id_base;      // first global id of this shard (0 for an unsharded store)
    int64_t n_global;     // global id space [0, n_global); ids outside it do not exist
    int32_t dim;
    int64_t rec_stride;
    const uint8_t *rec;
    const DevKeyRing *keys;
    const DevKeyRing *keys_host;  // host copy of the same ring (passed by value to kernels that want it in the constant bank)
    const u128 *hpow;     // [kMaxKeys][npow+1] GHASH key powers H^1..H^npow (index 0 unused)
    int32_t npow;         // c + 4, c = ceil(8*dim/16)
    const uint32_t *te0;  // [256]
    const uint4 *shoup;   // [kMaxKeys][16][256] GHASH multiply-by-H tables (one 64 KB table per live version)
    const uint8_t *deleted;
    int64_t n_deleted;
};

struct RouteParams {
    int64_t Q;
    const uint64_t *codes;  // [Q][TD][W]
    int32_t probes;
    int64_t hard_cap;
    int32_t B;
    int32_t *cand_ids, *cand_scores;  // [Q][B]
    int32_t *n_cand, *raw_seen, *unique;
    // per-CTA scratch
    int32_t cap0;        // Java initial table size = tableSizeFor(min(hard_cap, 65536)) (PIS:619)
    int32_t max_nodes;   // upper bound on distinct candidates of one query
    int32_t *scratch;    // [grid][scratch_ints]
    int64_t scratch_ints;
    int32_t *chain_overflow;  // set when a bucket chain reached 9 (Java would treeify; order then unspecified)
    const int32_t *qlist, *qlist_n;   // general kernel: serve these queries only (nullptr = all)
};

struct RouteFastExtra {
    const int32_t *vis_part;
    const uint8_t *vis_score;
    const uint8_t *vis_n;
    int n_raw;    // T*D*probes*64
    int tbl;      // open-addressing slots (power of two)
    int sort_n;   // bitonic sort width (power of two >= B + slack)
    int wl_extra; // dedicated worklist entries (the worklist continues over skey / sid)
    // route_fast2_kernel (two CTAs per SM): 0 = not eligible
    size_t v2_smem;
    int v2_region, v2_cls_cap, v2_wl_cap;
    int v2_big;                           // refinementLimit > 1024: the selected keys go to big_keys, route_sort_big_kernel orders them
    int v1_ok, tbl1;                      // route_fast_kernel is eligible (non-binding HARD_CAP, state fits) with tbl1 table slots
    unsigned long long *big_keys;         // [Q][B] (score | Java bucket | position)
    int32_t *big_seg;                     // [Q][5] segment starts of big_keys (segments are cut at score-class boundaries)
    int32_t *ovf_n, *ovf_list;            // queries route_fast2_kernel hands to route_fast_kernel
    const int32_t *qlist, *qlist_n;       // route_fast_kernel: serve these queries only (nullptr = all)
};

struct RefineParams {
    int64_t Q;
    const double *queries;    // [Q][dim]
    const float *queries_f32; // [Q][dim] compact copy (or nullptr); valid for distances only when *f32_exact != 0
    const uint8_t *queries_u8; // [Q][dim] uint8 copy; valid when f32_exact[1] != 0 (all values are integers 0..255)
    const int32_t *f32_exact;  // [3]: {FP32 copy exact, uint8 copy exact, every value finite}
    const uint8_t *qfinite;    // [Q] 1 = the query row is all-finite (QSI:137 / QSI:407-413); nullptr = not checked
    const int32_t *cand_ids;  // [Q][stride]
    const int32_t *n_cand;    // [Q]
    int32_t stride, k;
    uint32_t div_magic, div_shift;  // pair / stride = (umulhi(pair, div_magic) + pair) >> div_shift for pair < 2^31 (set_stride_divisor)
    // grouping scratch
    int32_t *cnt;       // [N+1]  pairs per record -> exclusive offsets after the scan
    int32_t *flag_pref; // [N+1]  exclusive prefix of (cnt>0)
    int32_t *fill;      // [N]
    int32_t *uniq;      // [<= min(N, Q*stride)]  distinct records named by the batch, ascending
    int32_t *uoff;      // [n_uniq + 1] where the pairs of distinct record #u start in pairs[] (uoff[n_uniq] = number of pairs)
    uint32_t *pairs;    // [Q*stride]
    int32_t *block_sums;  // scan scratch
    int32_t *totals;    // [4]: n_pairs, n_uniq, work counter, spare
    double *dist;       // [Q*stride]
    uint8_t *verdict;   // [Q*stride]
    uint32_t *touched;  // [ceil(N/32)]
    uint8_t *rec_verdict;  // [n_uniq] authentication verdict per distinct record (verify kernel -> decrypt kernel)
    int32_t *vorder;       // [n_uniq] positions of uniq[] bucketed by key-version slot, or nullptr (one live version)
    int32_t *voff;         // [kMaxKeys + 1] bucket offsets
    int32_t *vcnt;         // [2 * kMaxKeys] bucket counts / scatter cursors
    // outputs
    int32_t *topk_ids;  // [Q][k]
    double *topk_dist;  // [Q][k]
    int32_t *n_ret, *n_dec;
    int32_t *topk_rank;  // [Q][k] candidate position of each result (for the cross-shard merge), or nullptr
    const int32_t *rank_map;  // [Q][stride] original position of every (shard-compacted) candidate slot, or nullptr: positions are the slots
};

// Opt a kernel in to the device's full dynamic shared memory (opt-in limit minus the kernel's static shared memory) on the CURRENT
// device.  The attribute is per device: fspann_ctx_create runs the configure_*_kernels() functions after cudaSetDevice, so one
// process may drive any number of GPUs (one context each).
template <class K>
inline int opt_in_smem(K kernel) {
    cudaFuncAttributes a;
    int dev = 0, optin = 0;
    if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    const int dyn = optin - (int)a.sharedSizeBytes;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) == cudaSuccess ? 0 : -1;
}
// SM count of the current device (grid caps of the grid-stride helper kernels)
inline int cur_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 1;
    return n;
}
constexpr int kMaxDynSmem = 227 * 1024;   // shared memory one CTA can opt in to on sm_100a (static + dynamic)
int configure_tokengen_kernels();
int configure_route_kernels();
int configure_refine_kernels();

inline void set_stride_divisor(RefineParams &p) {
    uint32_t l = 0;
    while ((1ull << l) < (uint64_t)p.stride) l++;
    p.div_shift = l;
    p.div_magic = (uint32_t)(((((uint64_t)1 << l) - (uint64_t)p.stride) << 32) / (uint64_t)p.stride + 1);
}

// launchers (each returns the number of kernels it launched, or -1 after setting a CUDA error)
int launch_tokengen(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, uint64_t *codes, int32_t *work,
                    unsigned long long *list, int64_t list_cap, float *qf, float *qnorm, int mode, int sm_count, int *path_out);
// tokengen_tc.cu: the pre-filter's contraction on the tensor cores (tcgen05.mma, accumulator in TMEM)
size_t tokengen_tc_alpha_bytes(const RoutingView &rv);
int launch_alpha_tc_prepare(cudaStream_t s, const RoutingView &rv, uint16_t *out);
int launch_tokengen_tc(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, const uint16_t *alpha_tc, uint64_t *codes, int32_t *work,
                       unsigned long long *list, int64_t list_cap, int sm_count);
int configure_tokengen_tc_kernels();
int64_t tokengen_list_capacity(const RoutingView &rv, int64_t Q);
int launch_alpha_prepare(cudaStream_t s, const double *alpha, int64_t rows, int dim, float *alpha_f, float *norm);
int launch_route(cudaStream_t s, const RoutingView &rv, const RouteParams &p, int grid);
int route_grid(int64_t Q, int sm_count);
bool route_fast_eligible(const RoutingView &rv, int probes, int64_t hard_cap, int B, RouteFastExtra &x, size_t &smem);
int launch_route_fast(cudaStream_t s, const RoutingView &rv, const RouteParams &p, RouteFastExtra x, size_t smem, int sm_count,
                      int32_t *vis_part, uint8_t *vis_score, uint8_t *vis_n, const RouteParams *pg, int grid_g);
int64_t route_scratch_ints(int32_t cap0, int32_t max_nodes);
int launch_refine_group(cudaStream_t s, const StoreView &sv, const RefineParams &p);
int launch_refine_verify(cudaStream_t s, const StoreView &sv, const RefineParams &p, int sm_count);
int launch_version_bucket(cudaStream_t s, const StoreView &sv, const RefineParams &p, int64_t n_upper);
int launch_refine_decrypt(cudaStream_t s, const StoreView &sv, const RefineParams &p, int sm_count);
int launch_queries_to_f32(cudaStream_t s, const double *q, float *out, uint8_t *out8, int64_t Q, int dim, int32_t *exact, uint8_t *qfinite);
int launch_retry_select(cudaStream_t s, int64_t Q, int k, const int32_t *n_ret, const int32_t *n_dec, const int32_t *exact, int32_t *rows, int32_t *out);
int launch_refine_topk(cudaStream_t s, const RefineParams &p);
int launch_shard_compact(cudaStream_t s, int64_t Q, int stride, const int32_t *cand, const int32_t *n_cand, int64_t id_lo, int64_t id_hi,
                         int32_t *out_ids, int32_t *out_rank, int32_t *out_n);
int launch_counters(cudaStream_t s, int64_t Q, const int32_t *raw, const int32_t *uniq, const int32_t *n_dec,
                    const int32_t *n_ret, const int32_t *n_cand, int32_t retried, int64_t *counters, const uint8_t *qfinite);
int launch_gather_rows(cudaStream_t s, const void *src, void *dst, const int32_t *rows, int64_t n_rows, int64_t row_bytes, bool scatter);
int launch_store_pack(cudaStream_t s, uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *ids /* or null */,
                      const uint8_t *iv, const uint8_t *ct, const int32_t *ver);
int launch_gcm_tag(cudaStream_t s, const StoreView &sv, const int32_t *list, const int32_t *gid, int n_list, uint8_t *verdict,
                   const uint8_t *write_flag, int sm_count);
int launch_migrate_xcrypt(cudaStream_t s, const StoreView &sv, int n, const int32_t *list, const uint8_t *fresh_iv, int32_t target_version,
                          const uint8_t *verdict, uint8_t *flag, int sm_count);
int launch_encrypt_xcrypt(cudaStream_t s, const StoreView &sv, int n, const double *vectors, const uint8_t *ivs, int32_t version, uint8_t *flag,
                          int sm_count);
int launch_store_unpack(cudaStream_t s, const uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *rows, uint8_t *iv,
                        uint8_t *ct, int32_t *ver);
int launch_partition_build(cudaStream_t s, const uint64_t *codes, const int32_t *staged, int64_t n, int TD, int W, uint32_t cap, int code_bits,
                           int32_t *ids_out, int64_t *keys_out, uint64_t *rep_out, void *scratch, size_t scratch_bytes, int32_t *treeified);
size_t partition_build_scratch_bytes(int64_t n);
int launch_staged_order(cudaStream_t s, int32_t *staged, int64_t n);
int launch_iota(cudaStream_t s, int32_t *out, int64_t n);
int launch_gt_chunk(cudaStream_t s, const float *base, int64_t N, int dim, const float *queries, int Qc, int K, double *dist, int32_t *out_ids,
                    double *out_d2, int32_t *tie_overflow, int sm_count);
int launch_recall(cudaStream_t s, int Q, int K, const int32_t *gt, int gt_stride, const int32_t *res, int res_stride, const int32_t *n_ret,
                  double *recall);
int gt_max_k();
int launch_merge_topk(cudaStream_t s, int S, int64_t Q, int k, const double *dist, const int32_t *rank, const int32_t *ids, int32_t *out_ids,
                      double *out_dist, int32_t *out_nret);
int launch_debug_decrypt(cudaStream_t s, const StoreView &sv, int64_t n, const int32_t *ids, double *pt, uint8_t *verdict);

}  // namespace fsp
