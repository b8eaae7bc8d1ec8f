// abi_internal.cuh -- the context behind the C ABI and the host-side helpers shared by abi.cu and sharded.cu.
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>

#include "fspann_internal.cuh"

using fsp::DevKeyRing;
using fsp::RoutingView;
using fsp::StoreView;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct fspann_ctx {
    int device = 0;
    int sm_count = 1;                            // multiProcessorCount of ctx->device (set by fspann_ctx_create)
    cudaStream_t stream = nullptr;
    // host-pointer search entries: the query upload runs on `copy_stream` in up to 4 chunks (h2d_ev), so TokenGen + Route of chunk c overlap
    // the PCIe copy of chunk c+1; with supplied codes the whole query upload overlaps Route (h2d_late).  Consumed by search_pass.
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t h2d_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int h2d_chunks = 0;            // > 0: chunk c = queries [h2d_q0[c], h2d_q0[c+1]) arrives with h2d_ev[c]
    int64_t h2d_q0[5] = {0, 0, 0, 0, 0};
    bool h2d_late = false;         // the queries arrive with h2d_ev[4] (needed by Refine only)
    int opt_h2d_overlap = 2;       // chunks of the overlapped query upload (0 = one copy on the main stream)
    int32_t *h_pin = nullptr;                    // pinned host staging for the per-batch retry decision (n_ret, n_decrypted, flags)
    size_t h_pin_ints = 0;
    std::string err;
    int64_t launches = 0;

    // routing state
    bool routing_ready = false;
    RoutingView rv{};
    DevBuf d_alpha, d_r, d_omega, d_keys, d_rep, d_ids, d_deleted, d_alpha_f, d_alpha_norm, s_tg_work, s_tg_list, s_tg_qf, s_tg_norm;
    DevBuf d_alpha_tc;           // alpha as BF16 pieces in UMMA tile layout (tensor-core TokenGen)
    int opt_tokengen_mode = 0;   // 0 automatic (tensor cores when the shape allows), 2 = FP32 pre-filter only
    int last_tokengen_path = 0;  // 1 exact, 2 FP32 pre-filter, 3 tensor-core pre-filter
    int opt_tokengen_exact = 0;  // run the exact FP64 TokenGen kernel alone (no FP32 pre-filter)
    int64_t opt_tg_list_cap = 0; // test hook: clamp the re-check list (forces the overflow -> exact-kernel fallback)

    // store
    bool store_ready = false;
    StoreView sv{};
    DevBuf d_rec, d_keyring, d_hpow, d_shoup, d_te0, d_touched;
    std::map<int32_t, std::vector<uint8_t>> keys;  // live versions -> raw key

    // scratch (grow only)
    DevBuf s_vis_part, s_vis_score, s_vis_n;
    int opt_route_general = 0;   // force the general (sequential, cap-exact) Route kernel
    int opt_shard_compact = 1;   // sharded search: compact the gathered candidate lists to this shard's ids before Refine
    int opt_route_small_v1 = 1;  // batches of at most one query per SM take the one-CTA Route kernel (0: keep the two-CTA kernel, tests)
    int opt_route_v1 = 0;        // use the one-CTA-per-SM fast Route kernel only (A/B switch)
    int last_route_v2 = 0;
    int opt_route_wl_extra = -1; // test hook: clamp the fast path's dedicated worklist (forces the no-worklist fallback when exceeded)
    int last_route_path = 0;     // 1 = shared-memory fast path, 2 = general path
    bool last_queries_finite = true;
    DevBuf s_queries, s_codes, s_cand_ids, s_cand_sc, s_ncand, s_raw, s_uniq_cnt, s_route_scratch, s_overflow, s_route_ovf, s_route_big;
    DevBuf s_rec_verdict, s_qf32, s_qu8, s_f32_exact, s_vorder, s_voff, s_qfinite, s_retry_out, s_codes_in;
    DevKeyRing ring_host{};
    int32_t *want_rank = nullptr;
    const int32_t *rank_map = nullptr;   // set around do_refine by the sharded search (shard-compacted candidate lists)
    DevBuf s_cnt, s_flag, s_fill, s_uniq, s_uoff, s_pairs, s_bsums, s_totals, s_dist, s_verdict;
    DevBuf s_topk_ids, s_topk_dist, s_topk_rank, s_nret, s_ndec, s_counters, sh_c_ids, sh_c_rank, sh_c_n;
    DevBuf s_stage_a, s_stage_b, s_stage_c;  // upload staging
    DevBuf g_base, g_q, g_dist, g_ids, g_d2, g_flag, g_res, g_nret, g_rec;  // ground truth / recall
    DevBuf b_codes, b_staged, b_scratch, b_ids, b_keys, b_rep, b_flag;  // device index build
    int last_build_treeified = 0;
    int64_t build_n = 0, build_added = 0;       // fspann_routing_build_begin / add / finish
    DevBuf m_list, m_gid, m_iv, m_verdict, m_flag, m_rec, m_vec, m_out_iv, m_out_ct, m_out_ver;  // Migrate / bulk encryption
    DevBuf r_rows, r_queries, r_codes, r_topk_ids, r_topk_dist, r_nret, r_counters;  // retry subset
    DevBuf t_cand_ids, t_cand_sc, t_ncand, t_raw, t_uniq_cnt, t_ndec;               // retry subset route outputs

    // database-sharded search (sharded.cu): NCCL communicator over the contexts that hold the shards of ONE store
    void *nccl_comm = nullptr;                   // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    DevBuf sh_cand, sh_ncand, sh_raw, sh_uniq, sh_cand_all, sh_ncand_all, sh_raw_all, sh_uniq_all;
    DevBuf sh_loc_ids, sh_loc_dist, sh_loc_rank, sh_loc_nret, sh_ndec, sh_all_ids, sh_all_dist, sh_all_rank;
    DevBuf sh_queries, sh_out_ids, sh_out_dist, sh_out_nret, sh_out_cnt, sh_r_queries, sh_r_ids, sh_r_dist, sh_r_nret, sh_r_cnt;
    cudaEvent_t sh_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool sh_ev_valid = false;
    int64_t sh_gather_bytes = 0;                 // bytes this rank RECEIVED through the collectives of the last sharded pass

    // CUDA graphs for small batches: the whole first pass (TokenGen .. counters, retry decision) of a (Q, k, probes, cap, B, buffers) combination
    // is captured on its second use and replayed afterwards; `epoch` changes whenever a pointer or a by-value kernel argument may have changed
    struct GraphEntry { std::array<int64_t, 12> key; cudaGraphExec_t exec; int64_t launches; };
    std::vector<GraphEntry> graphs;
    std::array<int64_t, 12> graph_seen{};        // key of the last eager small-batch pass (capture happens on its repetition)
    int64_t epoch = 1;
    int opt_graphs = 1;
    bool capturing = false;
    int64_t graph_replays = 0, graph_captures = 0;

    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    int64_t last_call_launches = 0;
};


namespace fspabi {
int fail(fspann_ctx *c, int code, const char *fmt, ...);
int ensure(fspann_ctx *ctx, DevBuf &b, size_t bytes);
void release(DevBuf &b);
int check_routing(fspann_ctx *ctx);
int check_store(fspann_ctx *ctx);
int run_tokengen(fspann_ctx *ctx, int64_t Q, const double *d_queries, uint64_t *d_codes);
int do_route(fspann_ctx *ctx, int64_t Q, const uint64_t *d_codes, int probes, int64_t hard_cap, int B, int32_t *d_cand_ids, int32_t *d_cand_sc,
             int32_t *d_ncand, int32_t *d_raw, int32_t *d_uniq);
int do_refine(fspann_ctx *ctx, int64_t Q, const double *d_queries, const int32_t *d_cand_ids, const int32_t *d_ncand, int stride, int k,
              int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_nret, int32_t *d_ndec, bool stage_events);
bool all_finite(const double *v, int64_t n);
int record_ev(fspann_ctx *ctx, int i);
void sharded_release(fspann_ctx *ctx);           // sharded.cu: frees the communicator, events and scratch of the sharded path
}  // namespace fspabi

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(ctx, e__ == cudaErrorMemoryAllocation ? FSPANN_E_NOMEM : FSPANN_E_CUDA, \
                                            "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LAUNCHED(expr)                                                                              \
    do {                                                                                            \
        int n__ = (expr);                                                                           \
        if (n__ < 0) return fail(ctx, FSPANN_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(cudaGetLastError()), __FILE__, __LINE__); \
        ctx->launches += n__;                                                                       \
    } while (0)

#define ENSURE(buf, bytes) do { int rc__ = ensure(ctx, buf, bytes); if (rc__) return rc__; } while (0)
