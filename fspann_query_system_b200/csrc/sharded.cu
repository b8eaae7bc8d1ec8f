// sharded.cu -- database-sharded search (BASELINE config 4) THROUGH THE C ABI: the NCCL collectives live inside
// libfspann_gpu.so, so a Java host (one thread / one context per GPU) runs config 4 without any Python.
//
// The routing index is replicated on every GPU and the encrypted store is sharded by contiguous global-id range
// (fspann_store_upload_shard).  One batch, W ranks:
//   1. Route is QUERY-parallel: rank r codes + routes rows [r*per, (r+1)*per) of the batch (TokenGen + Route kernels), producing the
//      reference's ordered candidate lists for its slice (QTF:98-131, PIS:592-715, QSI:153-214);
//   2. ncclAllGather of the candidate lists (Q*B*4 bytes in total) and of the per-query route counters: every rank now holds the
//      IDENTICAL ordered lists, exactly what a single-index reference would have computed;
//   3. Refine is DATA-parallel: every rank authenticates + decrypts + scores, for ALL queries, the candidates its shard holds (the
//      others get verdict 0xFD) and keeps its local stable top-k with the candidates' positions in the global lists (QSI:238-322);
//   4. ncclAllGather of the per-shard top-k (distance, candidate rank, id) and ncclAllReduce of the decrypted counts, then
//      merge_topk_kernel orders on (distance, rank) -- the rank reproduces the reference's stable sort (QSI:298), so the result is
//      bit-identical to the unsharded search;
//   5. the adaptive retry (QSI:327-337) is decided on the device from the merged counts, identically on every rank.
// NCCL is resolved at run time with dlopen (libnccl.so.2: the copy already loaded in the process if there is one, else the system's),
// so the library loads -- and everything unsharded works -- on a box without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "abi_internal.cuh"

using namespace fsp;
using namespace fspabi;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string error;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("FSPANN_NCCL_LIB");
        void *h = nullptr;
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the host process already uses (e.g. PyTorch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found") + " (set FSPANN_NCCL_LIB)"; return; }
        api.handle = h;
#define FSP_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); if (!api.field) api.error = std::string("libnccl lacks ") + name;
        FSP_SYM(GetUniqueId, "ncclGetUniqueId") FSP_SYM(CommInitRank, "ncclCommInitRank") FSP_SYM(CommDestroy, "ncclCommDestroy")
        FSP_SYM(AllGather, "ncclAllGather") FSP_SYM(AllReduce, "ncclAllReduce") FSP_SYM(GroupStart, "ncclGroupStart")
        FSP_SYM(GroupEnd, "ncclGroupEnd") FSP_SYM(GetErrorString, "ncclGetErrorString") FSP_SYM(GetVersion, "ncclGetVersion")
#undef FSP_SYM
    });
    return api;
}

#define NC(call)                                                                                                         \
    do {                                                                                                                 \
        ncclResult_t r__ = (call);                                                                                       \
        if (r__ != ncclSuccess) return fail(ctx, FSPANN_E_CUDA, "%s failed: %s (%s:%d)", #call, nccl().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

// One pass of the sharded search over Q device-resident queries (identical on every rank).  Outputs are device arrays.
int sharded_pass(fspann_ctx *ctx, int64_t Q, const double *d_queries, int k, int probes, int64_t hard_cap, int B, int retried, int32_t *d_out_ids,
                 double *d_out_dist, int32_t *d_out_nret, int64_t *d_out_cnt, bool main_pass) {
    const int W = ctx->comm_size, r = ctx->comm_rank;
    const RoutingView &rv = ctx->rv;
    const int64_t per = (Q + W - 1) / W;                       // rows per rank; the last ranks may own fewer (padded for the gathers)
    const int64_t lo = std::min<int64_t>(Q, r * per), hi = std::min<int64_t>(Q, lo + per);
    const size_t slice_c = sizeof(int32_t) * (size_t)per * B, slice_m = sizeof(int32_t) * (size_t)per;
    ENSURE(ctx->sh_cand, slice_c); ENSURE(ctx->sh_ncand, slice_m); ENSURE(ctx->sh_raw, slice_m); ENSURE(ctx->sh_uniq, slice_m);
    ENSURE(ctx->s_codes, sizeof(uint64_t) * (size_t)std::max<int64_t>(hi - lo, 1) * rv.TD * rv.W);
    ENSURE(ctx->s_cand_sc, sizeof(int32_t) * (size_t)per * B);
    int32_t *cand_all = (int32_t *)ctx->sh_cand.p, *ncand_all = (int32_t *)ctx->sh_ncand.p, *raw_all = (int32_t *)ctx->sh_raw.p,
            *uniq_all = (int32_t *)ctx->sh_uniq.p;
    if (W > 1) {
        ENSURE(ctx->sh_cand_all, slice_c * W); ENSURE(ctx->sh_ncand_all, slice_m * W); ENSURE(ctx->sh_raw_all, slice_m * W); ENSURE(ctx->sh_uniq_all, slice_m * W);
        cand_all = (int32_t *)ctx->sh_cand_all.p; ncand_all = (int32_t *)ctx->sh_ncand_all.p; raw_all = (int32_t *)ctx->sh_raw_all.p;
        uniq_all = (int32_t *)ctx->sh_uniq_all.p;
    }
    int rc;
    if (main_pass) CK(cudaEventRecord(ctx->sh_ev[0], ctx->stream));
    // ---- 1. TokenGen + Route on my slice of the batch (replicated routing index)
    if (hi - lo < per) {                                        // padding rows of a short slice: empty lists
        CK(cudaMemsetAsync(ctx->sh_cand.p, 0xff, slice_c, ctx->stream));
        CK(cudaMemsetAsync(ctx->sh_ncand.p, 0, slice_m, ctx->stream)); CK(cudaMemsetAsync(ctx->sh_raw.p, 0, slice_m, ctx->stream));
        CK(cudaMemsetAsync(ctx->sh_uniq.p, 0, slice_m, ctx->stream));
    }
    if (hi > lo) {
        rc = run_tokengen(ctx, hi - lo, d_queries + (size_t)lo * rv.dim, (uint64_t *)ctx->s_codes.p); if (rc) return rc;
        rc = do_route(ctx, hi - lo, (const uint64_t *)ctx->s_codes.p, probes, hard_cap, B, (int32_t *)ctx->sh_cand.p, (int32_t *)ctx->s_cand_sc.p,
                      (int32_t *)ctx->sh_ncand.p, (int32_t *)ctx->sh_raw.p, (int32_t *)ctx->sh_uniq.p);
        if (rc) return rc;
    }
    if (main_pass) CK(cudaEventRecord(ctx->sh_ev[1], ctx->stream));
    // ---- 2. every rank gets the identical ordered candidate lists
    int64_t gathered = 0;
    if (W > 1) {
        NcclApi &n = nccl();
        ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
        NC(n.GroupStart());
        NC(n.AllGather(ctx->sh_cand.p, cand_all, (size_t)per * B, ncclInt32, comm, ctx->stream));
        NC(n.AllGather(ctx->sh_ncand.p, ncand_all, (size_t)per, ncclInt32, comm, ctx->stream));
        NC(n.AllGather(ctx->sh_raw.p, raw_all, (size_t)per, ncclInt32, comm, ctx->stream));
        NC(n.AllGather(ctx->sh_uniq.p, uniq_all, (size_t)per, ncclInt32, comm, ctx->stream));
        NC(n.GroupEnd());
        gathered += (int64_t)(W - 1) * (int64_t)(slice_c + 3 * slice_m);
    }
    if (main_pass) CK(cudaEventRecord(ctx->sh_ev[2], ctx->stream));
    // ---- 3. Refine on my shard, for all queries
    const size_t tk_i = sizeof(int32_t) * (size_t)Q * k, tk_d = sizeof(double) * (size_t)Q * k, qi = sizeof(int32_t) * (size_t)Q;
    ENSURE(ctx->sh_loc_ids, tk_i); ENSURE(ctx->sh_loc_dist, tk_d); ENSURE(ctx->sh_loc_rank, tk_i); ENSURE(ctx->sh_loc_nret, qi); ENSURE(ctx->sh_ndec, qi);
    ctx->want_rank = (int32_t *)ctx->sh_loc_rank.p;
    if (main_pass) { rc = record_ev(ctx, 0); if (rc) return rc; rc = record_ev(ctx, 1); if (rc) return rc; rc = record_ev(ctx, 2); if (rc) return rc; }
    const int32_t *ref_cand = cand_all, *ref_n = ncand_all;
    if (W > 1 && ctx->opt_shard_compact) {
        // shard-local lists: this shard's candidates only, in order, with their positions in the full lists (the merge orders on those)
        ENSURE(ctx->sh_c_ids, sizeof(int32_t) * (size_t)Q * B); ENSURE(ctx->sh_c_rank, sizeof(int32_t) * (size_t)Q * B); ENSURE(ctx->sh_c_n, qi);
        LAUNCHED(launch_shard_compact(ctx->stream, Q, B, cand_all, ncand_all, ctx->sv.id_base, ctx->sv.id_base + ctx->sv.N, (int32_t *)ctx->sh_c_ids.p,
                                      (int32_t *)ctx->sh_c_rank.p, (int32_t *)ctx->sh_c_n.p));
        ref_cand = (const int32_t *)ctx->sh_c_ids.p; ref_n = (const int32_t *)ctx->sh_c_n.p;
        ctx->rank_map = (const int32_t *)ctx->sh_c_rank.p;
    }
    rc = do_refine(ctx, Q, d_queries, ref_cand, ref_n, B, k, (int32_t *)ctx->sh_loc_ids.p, (double *)ctx->sh_loc_dist.p, (int32_t *)ctx->sh_loc_nret.p,
                   (int32_t *)ctx->sh_ndec.p, main_pass);
    ctx->want_rank = nullptr; ctx->rank_map = nullptr;
    if (rc) return rc;
    if (main_pass) { ctx->ev_valid = true; CK(cudaEventRecord(ctx->sh_ev[3], ctx->stream)); }
    // ---- 4. per-shard top-k of every rank + total decrypted count, then the global stable top-k
    const int32_t *all_ids = (const int32_t *)ctx->sh_loc_ids.p, *all_rank = (const int32_t *)ctx->sh_loc_rank.p;
    const double *all_dist = (const double *)ctx->sh_loc_dist.p;
    if (W > 1) {
        ENSURE(ctx->sh_all_ids, tk_i * W); ENSURE(ctx->sh_all_dist, tk_d * W); ENSURE(ctx->sh_all_rank, tk_i * W);
        NcclApi &n = nccl();
        ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
        NC(n.GroupStart());
        NC(n.AllGather(ctx->sh_loc_dist.p, ctx->sh_all_dist.p, (size_t)Q * k, ncclFloat64, comm, ctx->stream));
        NC(n.AllGather(ctx->sh_loc_rank.p, ctx->sh_all_rank.p, (size_t)Q * k, ncclInt32, comm, ctx->stream));
        NC(n.AllGather(ctx->sh_loc_ids.p, ctx->sh_all_ids.p, (size_t)Q * k, ncclInt32, comm, ctx->stream));
        NC(n.AllReduce(ctx->sh_ndec.p, ctx->sh_ndec.p, (size_t)Q, ncclInt32, ncclSum, comm, ctx->stream));
        NC(n.GroupEnd());
        all_ids = (const int32_t *)ctx->sh_all_ids.p; all_rank = (const int32_t *)ctx->sh_all_rank.p; all_dist = (const double *)ctx->sh_all_dist.p;
        gathered += (int64_t)(W - 1) * (int64_t)(2 * tk_i + tk_d) + (int64_t)qi;
    }
    LAUNCHED(launch_merge_topk(ctx->stream, W, Q, k, all_dist, all_rank, all_ids, d_out_ids, d_out_dist, d_out_nret));
    if (d_out_cnt)
        LAUNCHED(launch_counters(ctx->stream, Q, raw_all, uniq_all, (const int32_t *)ctx->sh_ndec.p, d_out_nret, ncand_all, retried, d_out_cnt,
                                 (const uint8_t *)ctx->s_qfinite.p));
    if (main_pass) { CK(cudaEventRecord(ctx->sh_ev[4], ctx->stream)); ctx->sh_ev_valid = true; ctx->sh_gather_bytes = gathered; }
    return 0;
}

int sharded_search_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int k, int probes, int64_t hard_cap, int B, bool allow_retry,
                       int32_t *d_ids, double *d_dist, int32_t *d_nret, int64_t *d_cnt) {
    if (probes <= 0) probes = 5;
    for (int i = 0; i < 5; i++) if (!ctx->sh_ev[i]) CK(cudaEventCreate(&ctx->sh_ev[i]));
    const int64_t l0 = ctx->launches;
    int rc = sharded_pass(ctx, Q, d_queries, k, probes, hard_cap, B, 0, d_ids, d_dist, d_nret, d_cnt, true);
    if (rc) return rc;
    if (allow_retry) {
        // identical inputs on every rank (merged n_ret, all-reduced decrypted counts) => identical, identically ORDERED rows everywhere
        if (!ctx->h_pin) { CK(cudaHostAlloc((void **)&ctx->h_pin, sizeof(int32_t) * 16, cudaHostAllocDefault)); ctx->h_pin_ints = 16; }
        ENSURE(ctx->r_rows, sizeof(int32_t) * (size_t)Q);
        ENSURE(ctx->s_retry_out, sizeof(int32_t) * 4);
        LAUNCHED(launch_retry_select(ctx->stream, Q, k, d_nret, (const int32_t *)ctx->sh_ndec.p, (const int32_t *)ctx->s_f32_exact.p, (int32_t *)ctx->r_rows.p,
                                     (int32_t *)ctx->s_retry_out.p));
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->s_retry_out.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int64_t R = ctx->h_pin[0];
        if (R > 0) {
            const int dim = ctx->rv.dim;
            ENSURE(ctx->sh_r_queries, sizeof(double) * (size_t)R * dim);
            ENSURE(ctx->sh_r_ids, sizeof(int32_t) * (size_t)R * k); ENSURE(ctx->sh_r_dist, sizeof(double) * (size_t)R * k);
            ENSURE(ctx->sh_r_nret, sizeof(int32_t) * (size_t)R); ENSURE(ctx->sh_r_cnt, sizeof(int64_t) * FSPANN_COUNTERS * (size_t)R);
            const int32_t *rows = (const int32_t *)ctx->r_rows.p;
            LAUNCHED(launch_gather_rows(ctx->stream, d_queries, ctx->sh_r_queries.p, rows, R, sizeof(double) * dim, false));
            rc = sharded_pass(ctx, R, (const double *)ctx->sh_r_queries.p, k, 10, hard_cap, B, 1, (int32_t *)ctx->sh_r_ids.p, (double *)ctx->sh_r_dist.p,
                              (int32_t *)ctx->sh_r_nret.p, d_cnt ? (int64_t *)ctx->sh_r_cnt.p : nullptr, false);
            if (rc) return rc;
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->sh_r_ids.p, d_ids, rows, R, sizeof(int32_t) * k, true));
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->sh_r_dist.p, d_dist, rows, R, sizeof(double) * k, true));
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->sh_r_nret.p, d_nret, rows, R, sizeof(int32_t), true));
            if (d_cnt) LAUNCHED(launch_gather_rows(ctx->stream, ctx->sh_r_cnt.p, d_cnt, rows, R, sizeof(int64_t) * FSPANN_COUNTERS, true));
        }
    }
    ctx->last_call_launches = ctx->launches - l0;
    return FSPANN_OK;
}

int check_sharded_args(fspann_ctx *ctx, int64_t Q, int k, int B) {
    int rc = check_routing(ctx); if (rc) return rc;
    rc = check_store(ctx); if (rc) return rc;
    if (ctx->rv.n_ids <= 1 && ctx->rv.P <= 1) return fail(ctx, FSPANN_E_STATE, "Index not finalized: only GFunctions uploaded");
    if (ctx->comm_size > 1 && !ctx->nccl_comm) return fail(ctx, FSPANN_E_STATE, "no communicator: call fspann_comm_init first");
    if (Q < 0) return fail(ctx, FSPANN_E_ARG, "negative batch size");
    if (k <= 0) return fail(ctx, FSPANN_E_ARG, "topK must be > 0 (QTF:65)");
    if (B <= 0) return fail(ctx, FSPANN_E_ARG, "refinementLimit must be > 0");
    return 0;
}

}  // namespace

namespace fspabi {
void sharded_release(fspann_ctx *ctx) {
    if (ctx->nccl_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->comm_size = 1; ctx->comm_rank = 0;
    DevBuf *bufs[] = {&ctx->sh_cand, &ctx->sh_ncand, &ctx->sh_raw, &ctx->sh_uniq, &ctx->sh_cand_all, &ctx->sh_ncand_all, &ctx->sh_raw_all, &ctx->sh_uniq_all,
                      &ctx->sh_loc_ids, &ctx->sh_loc_dist, &ctx->sh_loc_rank, &ctx->sh_loc_nret, &ctx->sh_ndec, &ctx->sh_all_ids, &ctx->sh_all_dist,
                      &ctx->sh_all_rank, &ctx->sh_queries, &ctx->sh_out_ids, &ctx->sh_out_dist, &ctx->sh_out_nret, &ctx->sh_out_cnt, &ctx->sh_r_queries,
                      &ctx->sh_r_ids, &ctx->sh_r_dist, &ctx->sh_r_nret, &ctx->sh_r_cnt};
    for (DevBuf *b : bufs) release(*b);
    for (int i = 0; i < 5; i++) if (ctx->sh_ev[i]) { cudaEventDestroy(ctx->sh_ev[i]); ctx->sh_ev[i] = nullptr; }
}
}  // namespace fspabi

extern "C" {

int fspann_comm_unique_id(uint8_t id_out[FSPANN_COMM_ID_BYTES]) {
    if (!id_out) return FSPANN_E_ARG;
    static_assert(sizeof(ncclUniqueId) <= FSPANN_COMM_ID_BYTES, "ncclUniqueId larger than FSPANN_COMM_ID_BYTES");
    NcclApi &n = nccl();
    if (!n.handle || !n.error.empty()) return FSPANN_E_STATE;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return FSPANN_E_CUDA;
    memset(id_out, 0, FSPANN_COMM_ID_BYTES);
    memcpy(id_out, &id, sizeof id);
    return FSPANN_OK;
}

int fspann_comm_init(fspann_ctx *ctx, int32_t n_ranks, int32_t rank, const uint8_t id[FSPANN_COMM_ID_BYTES]) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, FSPANN_E_ARG, "rank %d of %d", rank, n_ranks);
    if (ctx->nccl_comm) { nccl().CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ctx->comm_size = n_ranks; ctx->comm_rank = rank;
    if (n_ranks == 1) return FSPANN_OK;                          // a single shard needs no collective
    if (!id) return fail(ctx, FSPANN_E_ARG, "null communicator id");
    NcclApi &n = nccl();
    if (!n.handle || !n.error.empty()) { ctx->comm_size = 1; ctx->comm_rank = 0; return fail(ctx, FSPANN_E_STATE, "NCCL unavailable: %s", n.error.c_str()); }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    CK(cudaStreamSynchronize(ctx->stream));
    const ncclResult_t r = n.CommInitRank(&comm, n_ranks, uid, rank);
    if (r != ncclSuccess) { ctx->comm_size = 1; ctx->comm_rank = 0; return fail(ctx, FSPANN_E_CUDA, "ncclCommInitRank failed: %s", n.GetErrorString(r)); }
    ctx->nccl_comm = comm;
    return FSPANN_OK;
}

int fspann_comm_destroy(fspann_ctx *ctx) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->nccl_comm) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->comm_size = 1; ctx->comm_rank = 0;
    return FSPANN_OK;
}

int fspann_sharded_search_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                                    int32_t allow_retry, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_n_ret, int64_t *d_counters) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_sharded_args(ctx, Q, k, B); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (!d_queries || !d_topk_ids || !d_topk_dist || !d_n_ret) return fail(ctx, FSPANN_E_ARG, "null array");
    return sharded_search_dev(ctx, Q, d_queries, k, probes, hard_cap, B, allow_retry != 0, d_topk_ids, d_topk_dist, d_n_ret, d_counters);
}

int fspann_sharded_search_batch(fspann_ctx *ctx, int64_t Q, const double *queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                                int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_sharded_args(ctx, Q, k, B); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (!queries || !topk_ids_out || !topk_dist_out || !n_ret_out) return fail(ctx, FSPANN_E_ARG, "null array");
    const int dim = ctx->rv.dim;
    // every rank must take the same decision: the finite check runs on the full (replicated) batch before any collective
    if (!all_finite(queries, Q * dim)) return fail(ctx, FSPANN_E_ARG, "Vector contains NaN/Inf (Coding:357-359)");
    ENSURE(ctx->sh_queries, sizeof(double) * (size_t)Q * dim);
    ENSURE(ctx->sh_out_ids, sizeof(int32_t) * (size_t)Q * k); ENSURE(ctx->sh_out_dist, sizeof(double) * (size_t)Q * k);
    ENSURE(ctx->sh_out_nret, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->sh_out_cnt, sizeof(int64_t) * FSPANN_COUNTERS * (size_t)Q);
    CK(cudaMemcpyAsync(ctx->sh_queries.p, queries, sizeof(double) * (size_t)Q * dim, cudaMemcpyHostToDevice, ctx->stream));
    rc = sharded_search_dev(ctx, Q, (const double *)ctx->sh_queries.p, k, probes, hard_cap, B, true, (int32_t *)ctx->sh_out_ids.p, (double *)ctx->sh_out_dist.p,
                            (int32_t *)ctx->sh_out_nret.p, (int64_t *)ctx->sh_out_cnt.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(topk_ids_out, ctx->sh_out_ids.p, sizeof(int32_t) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(topk_dist_out, ctx->sh_out_dist.p, sizeof(double) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(n_ret_out, ctx->sh_out_nret.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    if (counters_out) CK(cudaMemcpyAsync(counters_out, ctx->sh_out_cnt.p, sizeof(int64_t) * FSPANN_COUNTERS * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

int64_t fspann_sharded_last_stage_ms(fspann_ctx *ctx, float out[4], int64_t *gather_bytes_out) {
    if (!ctx || !out) return -1;
    for (int i = 0; i < 4; i++) out[i] = 0.f;
    if (gather_bytes_out) *gather_bytes_out = 0;
    if (!ctx->sh_ev_valid) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    for (int i = 0; i < 4; i++) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->sh_ev[i], ctx->sh_ev[i + 1]) == cudaSuccess) out[i] = ms; }
    if (gather_bytes_out) *gather_bytes_out = ctx->sh_gather_bytes;
    return ctx->last_call_launches;
}

}  // extern "C"
