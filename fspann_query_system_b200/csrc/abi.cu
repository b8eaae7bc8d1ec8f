// abi.cu -- the C ABI of libfspann_gpu.so (include/fspann_gpu.h): context, uploads, batched entry points.
// Host-side orchestration only; all arithmetic of the hot path runs in the kernels of tokengen.cu / route.cu /
// refine.cu.  There is no CPU fallback: every compute entry point fails with FSPANN_E_CUDA if the device path fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "abi_internal.cuh"

using namespace fsp;
using namespace fspabi;

namespace fspabi {

int fail(fspann_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

int ensure(fspann_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return 0;
    if (ctx) ctx->epoch++;                                         // a buffer moves: captured graphs hold stale pointers
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
    if (e != cudaSuccess) { b.p = nullptr; return fail(ctx, FSPANN_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
    b.bytes = want;
    return 0;
}

void release(DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }

uint32_t table_size_for(int64_t cap) {  // java.util.HashMap.tableSizeFor
    if (cap <= 1) return 1;
    uint32_t n = 1;
    while ((int64_t)n < cap && n < (1u << 30)) n <<= 1;
    return n;
}

struct TeHost { const uint32_t *t; uint32_t operator()(uint32_t x) const { return t[x]; } };
struct RkHost { const uint32_t *r; uint32_t operator()(int i) const { return r[i]; } };

// Rebuilds the device key ring (round keys) and GHASH key-power tables from ctx->keys.
int rebuild_keys(fspann_ctx *ctx) {
    static uint32_t te0[256];
    static std::once_flag te_once;                 // contexts may be created from several host threads (one per GPU)
    std::call_once(te_once, [] { aes_make_te0(te0); });
    DevKeyRing ring;
    memset(&ring, 0, sizeof ring);
    const int dim = ctx->sv.dim > 0 ? ctx->sv.dim : 0;
    const int c = (8 * dim + 15) / 16;
    const int npow = c + 4;
    std::vector<u128> hp((size_t)kMaxKeys * (npow + 1));
    std::vector<uint32_t> shoup((size_t)std::max<size_t>(ctx->keys.size(), 1) * 4096 * 4);
    int n = 0;
    for (auto &kv : ctx->keys) {
        ring.version[n] = kv.first;
        aes256_expand_key(kv.second.data(), ring.rk[n]);
        uint32_t h[4];
        aes256_encrypt(TeHost{te0}, RkHost{ring.rk[n]}, 0u, 0u, 0u, 0u, h);   // H = E_K(0^128)
        u128 H{((uint64_t)h[0] << 32) | h[1], ((uint64_t)h[2] << 32) | h[3]};
        u128 *row = hp.data() + (size_t)n * (npow + 1);
        row[0] = u128{0, 0};
        row[1] = H;
        for (int p = 2; p <= npow; p++) row[p] = gf128_mul_ref(row[p - 1], H);
        ghash_make_shoup8(H, shoup.data() + (size_t)n * 4096 * 4);
        n++;
    }
    ring.n = n;
    ctx->ring_host = ring;
    ctx->epoch++;                                                  // the key ring is a by-value kernel argument
    ctx->sv.keys_host = &ctx->ring_host;
    ENSURE(ctx->d_keyring, sizeof ring);
    ENSURE(ctx->d_hpow, sizeof(u128) * hp.size());
    ENSURE(ctx->d_shoup, sizeof(uint32_t) * shoup.size());
    CK(cudaMemcpyAsync(ctx->d_shoup.p, shoup.data(), sizeof(uint32_t) * shoup.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_keyring.p, &ring, sizeof ring, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_hpow.p, hp.data(), sizeof(u128) * hp.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    ctx->sv.keys = (const DevKeyRing *)ctx->d_keyring.p;
    ctx->sv.hpow = (const u128 *)ctx->d_hpow.p;
    ctx->sv.npow = npow;
    ctx->sv.shoup = (const uint4 *)ctx->d_shoup.p;
    return 0;
}

int check_routing(fspann_ctx *ctx) {
    if (!ctx->routing_ready) return fail(ctx, FSPANN_E_STATE, "Index not finalized: no routing state uploaded (PIS:594)");
    return 0;
}
int check_store(fspann_ctx *ctx) {
    if (!ctx->store_ready) return fail(ctx, FSPANN_E_STATE, "record store not uploaded");
    if (ctx->routing_ready && ctx->rv.dim != ctx->sv.dim) return fail(ctx, FSPANN_E_STATE, "dimension mismatch: routing dim=%d store dim=%d", ctx->rv.dim, ctx->sv.dim);
    return 0;
}

int record_ev(fspann_ctx *ctx, int i) {
    if (ctx->capturing) return 0;
    CK(cudaEventRecord(ctx->ev[i], ctx->stream));
    return 0;
}

// TokenGen for Q device-resident queries: FP32 pre-filter + exact re-check (or the exact kernel alone), see tokengen.cu.
int run_tokengen(fspann_ctx *ctx, int64_t Q, const double *d_queries, uint64_t *d_codes) {
    int64_t cap = tokengen_list_capacity(ctx->rv, Q);
    if (ctx->opt_tg_list_cap > 0) cap = std::min(cap, ctx->opt_tg_list_cap);
    ENSURE(ctx->s_tg_work, 4 * sizeof(int32_t));
    ENSURE(ctx->s_tg_list, sizeof(unsigned long long) * (size_t)cap);
    ENSURE(ctx->s_tg_qf, sizeof(float) * (size_t)Q * (((size_t)ctx->rv.dim + 3) / 4 * 4));
    ENSURE(ctx->s_tg_norm, sizeof(float) * (size_t)Q);
    LAUNCHED(launch_tokengen(ctx->stream, ctx->rv, Q, d_queries, d_codes, (int32_t *)ctx->s_tg_work.p, (unsigned long long *)ctx->s_tg_list.p, cap,
                             (float *)ctx->s_tg_qf.p, (float *)ctx->s_tg_norm.p, ctx->opt_tokengen_exact ? 1 : ctx->opt_tokengen_mode, ctx->sm_count,
                             &ctx->last_tokengen_path));
    return 0;
}

// Route for Q queries whose codes are on the device.
int do_route(fspann_ctx *ctx, int64_t Q, const uint64_t *d_codes, int probes, int64_t hard_cap, int B, int32_t *d_cand_ids,
             int32_t *d_cand_sc, int32_t *d_ncand, int32_t *d_raw, int32_t *d_uniq) {
    const RoutingView &rv = ctx->rv;
    if (probes < 0) probes = 0;
    if ((int64_t)rv.TD * std::max(probes, 1) >= 262144) return fail(ctx, FSPANN_E_ARG, "T*D*probes too large (%d*%d)", rv.TD, probes);
    RouteParams p{};
    p.Q = Q; p.codes = d_codes; p.probes = probes; p.hard_cap = hard_cap; p.B = B;
    p.cand_ids = d_cand_ids; p.cand_scores = d_cand_sc; p.n_cand = d_ncand; p.raw_seen = d_raw; p.unique = d_uniq;
    p.cap0 = (int32_t)table_size_for(std::min<int64_t>(hard_cap, 65536));           // PIS:619
    int64_t max_nodes = (int64_t)rv.TD * probes * kBlock;
    if (max_nodes > hard_cap + kBlock) max_nodes = hard_cap + kBlock;
    if (max_nodes < kBlock) max_nodes = kBlock;
    p.max_nodes = (int32_t)max_nodes;
    CK(cudaMemsetAsync(d_cand_ids, 0xff, sizeof(int32_t) * (size_t)Q * B, ctx->stream));
    CK(cudaMemsetAsync(d_cand_sc, 0xff, sizeof(int32_t) * (size_t)Q * B, ctx->stream));
    RouteFastExtra fx{};
    size_t fsmem = 0;
    if (!ctx->opt_route_general && route_fast_eligible(rv, probes, hard_cap, B, fx, fsmem)) {
        if (ctx->opt_route_wl_extra >= 0 && fx.wl_extra > ctx->opt_route_wl_extra) fx.wl_extra = ctx->opt_route_wl_extra & ~7;   // smem size kept
        if (ctx->opt_route_wl_extra >= 0 && fx.v2_wl_cap > ctx->opt_route_wl_extra) fx.v2_wl_cap = std::max(16, ctx->opt_route_wl_extra & ~15);
        if (ctx->opt_route_v1 && fx.v1_ok) fx.v2_smem = 0;
        // at most one query per SM: the one-CTA kernel gives a single query 1024 threads and needs no overflow pass (latency of small batches)
        if (fx.v1_ok && Q <= ctx->sm_count && ctx->opt_route_wl_extra < 0 && ctx->opt_route_small_v1) fx.v2_smem = 0;
        ENSURE(ctx->s_route_ovf, sizeof(int32_t) * ((size_t)Q + 4));
        fx.ovf_n = (int32_t *)ctx->s_route_ovf.p; fx.ovf_list = fx.ovf_n + 4;
        ENSURE(ctx->s_vis_part, sizeof(int32_t) * (size_t)Q * rv.TD * probes);
        ENSURE(ctx->s_vis_score, (size_t)Q * rv.TD * probes);
        ENSURE(ctx->s_vis_n, (size_t)Q * rv.TD);
        if (fx.v2_smem && fx.v2_big) {
            ENSURE(ctx->s_route_big, sizeof(unsigned long long) * (size_t)Q * B + sizeof(int32_t) * 8 * (size_t)Q);
            fx.big_keys = (unsigned long long *)ctx->s_route_big.p;
            fx.big_seg = (int32_t *)(fx.big_keys + (size_t)Q * B);
        }
        RouteParams pg = p;
        int grid_g = 0;
        if (fx.v2_smem && !fx.v1_ok) {                                      // the general kernel serves what the two-CTA kernel hands back
            pg.scratch_ints = route_scratch_ints(pg.cap0, pg.max_nodes);
            grid_g = std::min(route_grid(Q, ctx->sm_count), 64);           // few queries overflow: bound the scratch
            ENSURE(ctx->s_route_scratch, sizeof(int32_t) * (size_t)pg.scratch_ints * grid_g);
            ENSURE(ctx->s_overflow, sizeof(int32_t));
            CK(cudaMemsetAsync(ctx->s_overflow.p, 0, sizeof(int32_t), ctx->stream));
            pg.scratch = (int32_t *)ctx->s_route_scratch.p;
            pg.chain_overflow = (int32_t *)ctx->s_overflow.p;
        }
        LAUNCHED(launch_route_fast(ctx->stream, rv, p, fx, fsmem, ctx->sm_count, (int32_t *)ctx->s_vis_part.p, (uint8_t *)ctx->s_vis_score.p,
                                   (uint8_t *)ctx->s_vis_n.p, &pg, grid_g));
        ctx->last_route_v2 = fx.v2_smem != 0;
        ctx->last_route_path = 1;
        return 0;
    }
    p.scratch_ints = route_scratch_ints(p.cap0, p.max_nodes);
    const int grid = route_grid(Q, ctx->sm_count);
    ENSURE(ctx->s_route_scratch, sizeof(int32_t) * (size_t)p.scratch_ints * grid);
    ENSURE(ctx->s_overflow, sizeof(int32_t));
    CK(cudaMemsetAsync(ctx->s_overflow.p, 0, sizeof(int32_t), ctx->stream));     // per call; read back by fspann_get_info("route_treeified")
    p.scratch = (int32_t *)ctx->s_route_scratch.p;
    p.chain_overflow = (int32_t *)ctx->s_overflow.p;
    LAUNCHED(launch_route(ctx->stream, rv, p, grid));
    ctx->last_route_path = 2;
    return 0;
}

// Refine for Q queries; everything on the device.
int do_refine(fspann_ctx *ctx, int64_t Q, const double *d_queries, const int32_t *d_cand_ids, const int32_t *d_ncand, int stride, int k,
              int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_nret, int32_t *d_ndec, bool stage_events) {
    const StoreView &sv = ctx->sv;
    const int64_t total = Q * (int64_t)stride;
    if (total >= (1LL << 31)) return fail(ctx, FSPANN_E_ARG, "Q*B too large for one batch (%lld)", (long long)total);
    const int64_t n1 = sv.N + 1;
    const int nblocks = (int)((n1 + 2047) / 2048);
    ENSURE(ctx->s_cnt, sizeof(int32_t) * (size_t)n1);
    ENSURE(ctx->s_fill, sizeof(int32_t) * (size_t)n1);
    ENSURE(ctx->s_uniq, sizeof(int32_t) * (size_t)std::min<int64_t>(n1, total + 1));
    ENSURE(ctx->s_uoff, sizeof(int32_t) * ((size_t)std::min<int64_t>(n1, total + 1) + 2));
    ENSURE(ctx->s_pairs, sizeof(uint32_t) * (size_t)(total + 1));
    ENSURE(ctx->s_bsums, sizeof(int32_t) * 2 * (size_t)(nblocks + 1));
    ENSURE(ctx->s_totals, sizeof(int32_t) * 4);
    ENSURE(ctx->s_dist, sizeof(double) * (size_t)(total + 1));
    ENSURE(ctx->s_verdict, (size_t)(total + 1));
    ENSURE(ctx->s_rec_verdict, (size_t)std::min<int64_t>(n1, total + 1) + 64);
    RefineParams p{};
    p.Q = Q; p.queries = d_queries; p.cand_ids = d_cand_ids; p.n_cand = d_ncand; p.stride = stride; p.k = k;
    set_stride_divisor(p);
    p.cnt = (int32_t *)ctx->s_cnt.p; p.fill = (int32_t *)ctx->s_fill.p; p.uniq = (int32_t *)ctx->s_uniq.p; p.uoff = (int32_t *)ctx->s_uoff.p;
    p.pairs = (uint32_t *)ctx->s_pairs.p; p.block_sums = (int32_t *)ctx->s_bsums.p; p.totals = (int32_t *)ctx->s_totals.p;
    p.dist = (double *)ctx->s_dist.p; p.verdict = (uint8_t *)ctx->s_verdict.p; p.touched = (uint32_t *)ctx->d_touched.p;
    p.rec_verdict = (uint8_t *)ctx->s_rec_verdict.p;
    p.topk_ids = d_topk_ids; p.topk_dist = d_topk_dist; p.n_ret = d_nret; p.n_dec = d_ndec; p.topk_rank = ctx->want_rank; p.rank_map = ctx->rank_map;
    ENSURE(ctx->s_qf32, sizeof(float) * (size_t)Q * sv.dim);
    ENSURE(ctx->s_qu8, (size_t)Q * sv.dim + 16);
    ENSURE(ctx->s_f32_exact, 4 * sizeof(int32_t));
    ENSURE(ctx->s_qfinite, (size_t)Q + 16);
    p.queries_f32 = (const float *)ctx->s_qf32.p; p.queries_u8 = (const uint8_t *)ctx->s_qu8.p; p.f32_exact = (const int32_t *)ctx->s_f32_exact.p;
    p.qfinite = (const uint8_t *)ctx->s_qfinite.p;
    LAUNCHED(launch_queries_to_f32(ctx->stream, d_queries, (float *)ctx->s_qf32.p, (uint8_t *)ctx->s_qu8.p, Q, sv.dim, (int32_t *)ctx->s_f32_exact.p,
                                   (uint8_t *)ctx->s_qfinite.p));
    // unknown / retired key version is the default verdict (KRS:82-88); the verify kernel overwrites it per live version
    CK(cudaMemsetAsync(p.rec_verdict, FSPANN_V_NO_KEY, (size_t)std::min<int64_t>(n1, total + 1), ctx->stream));
    LAUNCHED(launch_refine_group(ctx->stream, sv, p));
    if (stage_events) { int rc = record_ev(ctx, 3); if (rc) return rc; }
    if (ctx->ring_host.n > 1) {      // several live key versions (after a Rotate): every verify pass walks a dense list of its own records
        const int64_t n_upper = std::min<int64_t>(n1, total + 1);
        ENSURE(ctx->s_vorder, sizeof(int32_t) * (size_t)n_upper);
        ENSURE(ctx->s_voff, sizeof(int32_t) * (3 * kMaxKeys + 4));
        p.vorder = (int32_t *)ctx->s_vorder.p; p.voff = (int32_t *)ctx->s_voff.p; p.vcnt = p.voff + kMaxKeys + 1;
        LAUNCHED(launch_version_bucket(ctx->stream, sv, p, n_upper));
    }
    LAUNCHED(launch_refine_verify(ctx->stream, sv, p, ctx->sm_count));
    if (stage_events) { int rc = record_ev(ctx, 4); if (rc) return rc; }
    LAUNCHED(launch_refine_decrypt(ctx->stream, sv, p, ctx->sm_count));
    if (stage_events) { int rc = record_ev(ctx, 5); if (rc) return rc; }
    LAUNCHED(launch_refine_topk(ctx->stream, p));
    if (stage_events) { int rc = record_ev(ctx, 6); if (rc) return rc; }
    return 0;
}

bool all_finite(const double *v, int64_t n) {
    for (int64_t i = 0; i < n; i++) if (!std::isfinite(v[i])) return false;
    return true;
}

}  // namespace fspabi

extern "C" {

int fspann_ctx_create(int device, fspann_ctx **out) {
    if (!out) return FSPANN_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return FSPANN_E_CUDA;   // no GPU, no service: there is no CPU path
    if (device < 0 || device >= ndev) return FSPANN_E_ARG;
    fspann_ctx *ctx = new fspann_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FSPANN_E_CUDA; }
    ctx->sm_count = cur_sm_count();
    // > 48 KB of dynamic shared memory is a per-device opt-in of each kernel: configure THIS device (a process may hold one context per GPU)
    if (configure_tokengen_kernels() || configure_route_kernels() || configure_refine_kernels()) { delete ctx; return FSPANN_E_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FSPANN_E_CUDA; }
    for (int i = 0; i < 7; i++) cudaEventCreate(&ctx->ev[i]);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { fspann_ctx_destroy(ctx); return FSPANN_E_CUDA; }
    for (int i = 0; i < 5; i++) cudaEventCreateWithFlags(&ctx->h2d_ev[i], cudaEventDisableTiming);
    uint32_t te0[256];
    aes_make_te0(te0);
    if (ensure(ctx, ctx->d_te0, sizeof te0) || cudaMemcpy(ctx->d_te0.p, te0, sizeof te0, cudaMemcpyHostToDevice) != cudaSuccess) {
        fspann_ctx_destroy(ctx);
        return FSPANN_E_CUDA;
    }
    ctx->sv.te0 = (const uint32_t *)ctx->d_te0.p;
    if (rebuild_keys(ctx)) { fspann_ctx_destroy(ctx); return FSPANN_E_CUDA; }
    *out = ctx;
    return FSPANN_OK;
}

void fspann_ctx_destroy(fspann_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_alpha_f, &ctx->d_alpha_norm, &ctx->d_alpha_tc, &ctx->s_vorder, &ctx->s_voff, &ctx->s_qfinite, &ctx->s_retry_out, &ctx->s_codes_in, &ctx->s_tg_work, &ctx->s_tg_list, &ctx->s_tg_qf, &ctx->s_tg_norm, &ctx->d_alpha, &ctx->d_r, &ctx->d_omega, &ctx->d_keys, &ctx->d_rep, &ctx->d_ids, &ctx->d_deleted, &ctx->d_rec,
                      &ctx->d_keyring, &ctx->d_hpow, &ctx->d_shoup, &ctx->s_rec_verdict, &ctx->s_qf32, &ctx->s_qu8, &ctx->s_f32_exact, &ctx->d_te0, &ctx->d_touched, &ctx->s_queries, &ctx->s_codes, &ctx->s_cand_ids,
                      &ctx->s_cand_sc, &ctx->s_ncand, &ctx->s_raw, &ctx->s_uniq_cnt, &ctx->s_route_scratch, &ctx->s_overflow, &ctx->s_route_ovf, &ctx->s_route_big, &ctx->s_cnt,
                      &ctx->s_flag, &ctx->s_fill, &ctx->s_uniq, &ctx->s_uoff, &ctx->s_pairs, &ctx->s_bsums, &ctx->s_totals, &ctx->s_dist, &ctx->s_verdict,
                      &ctx->s_topk_ids, &ctx->s_topk_dist, &ctx->s_topk_rank, &ctx->sh_c_ids, &ctx->sh_c_rank, &ctx->sh_c_n, &ctx->s_nret, &ctx->s_ndec, &ctx->s_counters, &ctx->s_stage_a, &ctx->s_stage_b,
                      &ctx->s_stage_c, &ctx->s_vis_part, &ctx->s_vis_score, &ctx->s_vis_n, &ctx->r_rows, &ctx->r_queries, &ctx->r_codes, &ctx->r_topk_ids, &ctx->r_topk_dist, &ctx->r_nret,
                      &ctx->r_counters, &ctx->t_cand_ids, &ctx->t_cand_sc, &ctx->t_ncand, &ctx->t_raw, &ctx->t_uniq_cnt, &ctx->t_ndec,
                      &ctx->g_base, &ctx->g_q, &ctx->g_dist, &ctx->g_ids, &ctx->g_d2, &ctx->g_flag, &ctx->g_res, &ctx->g_nret, &ctx->g_rec,
                      &ctx->b_codes, &ctx->b_staged, &ctx->b_scratch, &ctx->b_ids, &ctx->b_keys, &ctx->b_rep, &ctx->b_flag,
                      &ctx->m_list, &ctx->m_gid, &ctx->m_iv, &ctx->m_verdict, &ctx->m_flag, &ctx->m_rec, &ctx->m_vec, &ctx->m_out_iv, &ctx->m_out_ct, &ctx->m_out_ver};
    for (DevBuf *b : bufs) release(*b);
    for (auto &g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
    sharded_release(ctx);
    for (int i = 0; i < 7; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 5; i++) if (ctx->h2d_ev[i]) cudaEventDestroy(ctx->h2d_ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    delete ctx;
}

const char *fspann_last_error(const fspann_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *fspann_ctx_stream(fspann_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int fspann_ctx_sync(fspann_ctx *ctx) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}
int64_t fspann_ctx_launch_count(const fspann_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fspann_set_option(fspann_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return FSPANN_E_ARG;
    ctx->epoch++;
    if (!strcmp(name, "graphs")) { ctx->opt_graphs = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "route_general")) { ctx->opt_route_general = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "route_wl_extra")) { ctx->opt_route_wl_extra = (int)value; return FSPANN_OK; }
    if (!strcmp(name, "route_small_v1")) { ctx->opt_route_small_v1 = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "shard_compact")) { ctx->opt_shard_compact = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "route_v1")) { ctx->opt_route_v1 = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "h2d_overlap")) { ctx->opt_h2d_overlap = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 4); return FSPANN_OK; }
    if (!strcmp(name, "tokengen_exact")) { ctx->opt_tokengen_exact = value != 0; return FSPANN_OK; }
    if (!strcmp(name, "tokengen_mode")) { ctx->opt_tokengen_mode = (int)value; return FSPANN_OK; }
    if (!strcmp(name, "tokengen_list_cap")) { ctx->opt_tg_list_cap = value; return FSPANN_OK; }
    return fail(ctx, FSPANN_E_ARG, "unknown option %s", name);
}
int64_t fspann_get_info(fspann_ctx *ctx, const char *name) {
    if (!ctx || !name) return -1;
    if (!strcmp(name, "last_route_path")) return ctx->last_route_path;
    if (!strcmp(name, "last_tokengen_path")) return ctx->last_tokengen_path;
    if (!strcmp(name, "route_overflowed")) {   // fast path, two-CTA kernel: queries of the last call that went to the one-CTA kernel
        int32_t f = 0;
        if (ctx->last_route_path != 1 || !ctx->last_route_v2 || !ctx->s_route_ovf.p) return 0;
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
            cudaMemcpy(&f, ctx->s_route_ovf.p, sizeof f, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return f;
    }
    if (!strcmp(name, "last_route_v2")) return ctx->last_route_v2;
    if (!strcmp(name, "route_treeified")) {   // general Route kernel: did a bestScore bin of the last call reach 9 entries (the JDK would treeify it)?
        int32_t f = 0;
        if (ctx->last_route_path != 2 || !ctx->s_overflow.p) return 0;
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
            cudaMemcpy(&f, ctx->s_overflow.p, sizeof f, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return f;
    }
    if (!strcmp(name, "sm_count")) return ctx->sm_count;
    if (!strcmp(name, "graph_replays")) return ctx->graph_replays;
    if (!strcmp(name, "graph_captures")) return ctx->graph_captures;
    if (!strcmp(name, "build_treeified")) return ctx->last_build_treeified;
    if (!strcmp(name, "tokengen_rechecked") || !strcmp(name, "tokengen_overflow")) {   // of the last TokenGen launch on this context
        int32_t w[2] = {0, 0};
        if (ctx->last_tokengen_path == 1) return 0;                  // the exact kernel alone ran: nothing was listed
        if (!ctx->s_tg_work.p || cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
            cudaMemcpy(w, ctx->s_tg_work.p, sizeof w, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return !strcmp(name, "tokengen_rechecked") ? w[0] : w[1];
    }
    return -1;
}

int fspann_routing_upload(fspann_ctx *ctx, int32_t dim, int32_t T, int32_t D, int32_t m, int32_t lambda, const double *alpha,
                          const double *r, const double *omega, int64_t n_ids, const int64_t *min_key, const int64_t *max_key,
                          const uint64_t *rep_code, const int32_t *ids) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!alpha || !r || !omega || !min_key || !max_key || !rep_code || !ids) return fail(ctx, FSPANN_E_ARG, "null array");
    if (dim <= 0 || T <= 0 || D <= 0 || m <= 0 || lambda <= 0 || n_ids <= 0) return fail(ctx, FSPANN_E_ARG, "non-positive parameter");
    if (lambda > 32) return fail(ctx, FSPANN_E_ARG, "lambda > 32");
    if ((int64_t)m * lambda > 255) return fail(ctx, FSPANN_E_ARG, "m*lambda = %d > 255 code bits not supported", m * lambda);
    const int W = (m * lambda + 63) / 64, TD = T * D;
    const size_t tg_smem = sizeof(double) * ((size_t)64 * (((dim + 1) & ~1) + 4) + (size_t)m * ((dim + 1) & ~1) + 2 * (size_t)m) + 4 * (size_t)64 * 2 * W;
    if (tg_smem > 227 * 1024) return fail(ctx, FSPANN_E_ARG, "dim=%d, m=%d need %zu B of shared memory for TokenGen (> 227 KB)", dim, m, tg_smem);
    for (int64_t i = 0; i < (int64_t)TD * m; i++)
        if (!(omega[i] > 0.0)) return fail(ctx, FSPANN_E_ARG, "omega_j <= 0 (Coding:85-87)");
    const int64_t P = (n_ids + kBlock - 1) / kBlock;
    ctx->routing_ready = false;
    ENSURE(ctx->d_alpha, sizeof(double) * (size_t)TD * m * dim);
    ENSURE(ctx->d_r, sizeof(double) * (size_t)TD * m);
    ENSURE(ctx->d_omega, sizeof(double) * (size_t)TD * m);
    ENSURE(ctx->d_keys, sizeof(int64_t) * 2 * (size_t)TD * P);
    ENSURE(ctx->d_rep, sizeof(uint64_t) * (size_t)TD * P * W);
    ENSURE(ctx->d_ids, sizeof(int32_t) * (size_t)TD * n_ids);
    std::vector<int64_t> inter(2 * (size_t)TD * P);
    for (size_t i = 0; i < (size_t)TD * P; i++) { inter[2 * i] = min_key[i]; inter[2 * i + 1] = max_key[i]; }
    CK(cudaMemcpy(ctx->d_alpha.p, alpha, sizeof(double) * (size_t)TD * m * dim, cudaMemcpyHostToDevice));
    ENSURE(ctx->d_alpha_f, sizeof(float) * (size_t)TD * m * dim);
    ENSURE(ctx->d_alpha_norm, sizeof(float) * (size_t)TD * m);
    LAUNCHED(launch_alpha_prepare(ctx->stream, (const double *)ctx->d_alpha.p, (int64_t)TD * m, dim, (float *)ctx->d_alpha_f.p, (float *)ctx->d_alpha_norm.p));
    const uint16_t *alpha_tc = nullptr;
    {   // alpha as BF16 pieces in tensor-core tile layout (tokengen_tc.cu); shapes it does not cover keep alpha_tc = nullptr
        RoutingView tv{};
        tv.dim = dim; tv.T = T; tv.D = D; tv.m = m; tv.lambda = lambda; tv.W = W; tv.TD = TD; tv.alpha = (const double *)ctx->d_alpha.p;
        const size_t tc_bytes = tokengen_tc_alpha_bytes(tv);
        if (tc_bytes) {
            ENSURE(ctx->d_alpha_tc, tc_bytes);
            LAUNCHED(launch_alpha_tc_prepare(ctx->stream, tv, (uint16_t *)ctx->d_alpha_tc.p));
            alpha_tc = (const uint16_t *)ctx->d_alpha_tc.p;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->d_r.p, r, sizeof(double) * (size_t)TD * m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_omega.p, omega, sizeof(double) * (size_t)TD * m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_keys.p, inter.data(), sizeof(int64_t) * inter.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_rep.p, rep_code, sizeof(uint64_t) * (size_t)TD * P * W, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_ids.p, ids, sizeof(int32_t) * (size_t)TD * n_ids, cudaMemcpyHostToDevice));
    RoutingView &rv = ctx->rv;
    rv.dim = dim; rv.T = T; rv.D = D; rv.m = m; rv.lambda = lambda; rv.W = W; rv.TD = TD; rv.n_ids = n_ids; rv.P = P;
    rv.alpha = (const double *)ctx->d_alpha.p; rv.r = (const double *)ctx->d_r.p; rv.omega = (const double *)ctx->d_omega.p;
    rv.alpha_f32 = (const float *)ctx->d_alpha_f.p; rv.alpha_norm = (const float *)ctx->d_alpha_norm.p; rv.alpha_tc = alpha_tc;
    rv.keys = (const int64_t *)ctx->d_keys.p; rv.rep = (const uint64_t *)ctx->d_rep.p; rv.ids = (const int32_t *)ctx->d_ids.p;
    ctx->routing_ready = true;
    return FSPANN_OK;
}

// GFunctions only (Setup-side bulk coding of the base set runs TokenGen before any partition exists).
int fspann_gfunctions_upload(fspann_ctx *ctx, int32_t dim, int32_t T, int32_t D, int32_t m, int32_t lambda, const double *alpha,
                             const double *r, const double *omega) {
    const int64_t one_min = 0, one_max = 0; const uint64_t rep[4] = {0, 0, 0, 0}; const int32_t id0 = 0;
    if (!ctx) return FSPANN_E_ARG;
    // a single dummy partition per (t,d) keeps the routing view well-formed; Route on it is meaningless and rejected below
    std::vector<int64_t> mn((size_t)T * D, one_min), mx((size_t)T * D, one_max);
    std::vector<uint64_t> rp((size_t)T * D * 4, rep[0]);
    std::vector<int32_t> ii((size_t)T * D, id0);
    int rc = fspann_routing_upload(ctx, dim, T, D, m, lambda, alpha, r, omega, 1, mn.data(), mx.data(), rp.data(), ii.data());
    return rc;
}

// ---- Setup-side index build on the device: coding loop of PIS.insert / finalizeForSearch (PIS:331-346, 789-845)
// + GreedyPartitioner.build (GP:37-76) for every (table, division), installed as the routing state.  Chunked form: begin -> add* ->
// finish, so a base set that does not fit the host (config 4: 100 M x 96) streams through in pieces; fspann_routing_build is the
// one-call wrapper ------------------------------------------------------------------------------------------------------------
int fspann_routing_build_begin(fspann_ctx *ctx, int64_t N) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->routing_ready) return fail(ctx, FSPANN_E_STATE, "GFunctionRegistry not initialized: upload the GFunctions first (PIS:812-819)");
    if (N < 1000) return fail(ctx, FSPANN_E_STATE, "Cannot finalize index: only %lld samples collected (< MIN_SAMPLE_SIZE) (PIS:803-808)", (long long)N);
    if (N >= (1LL << 31) - 1) return fail(ctx, FSPANN_E_ARG, "N too large for int32 ids");
    ENSURE(ctx->b_codes, sizeof(uint64_t) * (size_t)N * ctx->rv.TD * ctx->rv.W);
    ctx->build_n = N; ctx->build_added = 0;
    return FSPANN_OK;
}

static int build_add_dev(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *d_vectors) {
    if (ctx->build_n <= 0) return fail(ctx, FSPANN_E_STATE, "fspann_routing_build_begin has not been called");
    if (first_id < 0 || n < 0 || first_id + n > ctx->build_n) return fail(ctx, FSPANN_E_ARG, "ids [%lld, %lld) outside 0..N-1", (long long)first_id, (long long)(first_id + n));
    const RoutingView &rv = ctx->rv;
    const int64_t chunk = 1 << 20;                               // the TokenGen re-check list is sized per launch
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        const int64_t c = std::min(chunk, n - s0);
        int rc = run_tokengen(ctx, c, d_vectors + (size_t)s0 * rv.dim, (uint64_t *)ctx->b_codes.p + (size_t)(first_id + s0) * rv.TD * rv.W);
        if (rc) return rc;
    }
    ctx->build_added += n;
    return FSPANN_OK;
}

int fspann_routing_build_add_dev(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *d_vectors) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n > 0 && !d_vectors) return fail(ctx, FSPANN_E_ARG, "null array");
    return build_add_dev(ctx, first_id, n, d_vectors);
}

int fspann_routing_build_add(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *vectors) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n > 0 && !vectors) return fail(ctx, FSPANN_E_ARG, "null array");
    const int dim = ctx->rv.dim;
    if (!all_finite(vectors, n * dim)) return fail(ctx, FSPANN_E_ARG, "Vector contains NaN/Inf (Coding:357-359)");
    const int64_t chunk = 131072;
    ENSURE(ctx->s_queries, sizeof(double) * (size_t)std::min(chunk, std::max<int64_t>(n, 1)) * dim);
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        const int64_t c = std::min(chunk, n - s0);
        CK(cudaMemcpyAsync(ctx->s_queries.p, vectors + (size_t)s0 * dim, sizeof(double) * (size_t)c * dim, cudaMemcpyHostToDevice, ctx->stream));
        int rc = build_add_dev(ctx, first_id + s0, c, (const double *)ctx->s_queries.p);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));                 // the staging buffer is reused
    }
    return FSPANN_OK;
}

int fspann_routing_build_finish(fspann_ctx *ctx, const int32_t *staged_ids, int64_t *min_key_out, int64_t *max_key_out, uint64_t *rep_code_out,
                                int32_t *ids_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int64_t N = ctx->build_n;
    if (N <= 0) return fail(ctx, FSPANN_E_STATE, "fspann_routing_build_begin has not been called");
    if (ctx->build_added != N) return fail(ctx, FSPANN_E_STATE, "only %lld of %lld vectors were added", (long long)ctx->build_added, (long long)N);
    RoutingView &rv = ctx->rv;
    const int TD = rv.TD, W = rv.W;
    const int64_t P = (N + kBlock - 1) / kBlock;
    ENSURE(ctx->b_staged, sizeof(int32_t) * (size_t)N);
    if (staged_ids) {   // ids are the ordinals 0..N-1 (FSA:501,515), each staged exactly once
        std::vector<uint8_t> seen((size_t)N, 0);
        for (int64_t i = 0; i < N; i++) {
            const int32_t id = staged_ids[i];
            if (id < 0 || id >= N || seen[(size_t)id]) return fail(ctx, FSPANN_E_ARG, "staged ids must be a permutation of 0..N-1 (bad id %d)", id);
            seen[(size_t)id] = 1;
        }
        CK(cudaMemcpyAsync(ctx->b_staged.p, staged_ids, sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    } else {            // the facade's insertion order: the 1000th and later vectors first, then the first 999 (PIS:280-298, 821-831)
        LAUNCHED(launch_staged_order(ctx->stream, (int32_t *)ctx->b_staged.p, N));
    }
    // partitions: HashMap iteration order -> stable sort by key -> blocks of 64
    uint32_t cap = table_size_for(N);                                   // new HashMap<>(staged.size()) (PIS:413) ...
    while ((double)N > 0.75 * (double)cap && cap < (1u << 30)) cap <<= 1;   // ... doubled while size > 0.75 * cap
    const size_t scratch = partition_build_scratch_bytes(N);
    ENSURE(ctx->b_scratch, scratch);
    ENSURE(ctx->b_ids, sizeof(int32_t) * (size_t)TD * N);
    ENSURE(ctx->b_keys, sizeof(int64_t) * 2 * (size_t)TD * P);
    ENSURE(ctx->b_rep, sizeof(uint64_t) * (size_t)TD * P * W);
    ENSURE(ctx->b_flag, sizeof(int32_t));
    CK(cudaMemsetAsync(ctx->b_flag.p, 0, sizeof(int32_t), ctx->stream));
    LAUNCHED(launch_partition_build(ctx->stream, (const uint64_t *)ctx->b_codes.p, (const int32_t *)ctx->b_staged.p, N, TD, W, cap, rv.m * rv.lambda,
                                    (int32_t *)ctx->b_ids.p, (int64_t *)ctx->b_keys.p, (uint64_t *)ctx->b_rep.p, ctx->b_scratch.p, ctx->b_scratch.bytes,
                                    (int32_t *)ctx->b_flag.p));
    int32_t tree = 0;
    CK(cudaMemcpyAsync(&tree, ctx->b_flag.p, sizeof tree, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->last_build_treeified = tree;
    // freeze: the built arrays become the routing state (PIS:842)
    ctx->routing_ready = false;
    std::swap(ctx->d_ids, ctx->b_ids);
    std::swap(ctx->d_keys, ctx->b_keys);
    std::swap(ctx->d_rep, ctx->b_rep);
    rv.n_ids = N; rv.P = P;
    rv.keys = (const int64_t *)ctx->d_keys.p; rv.rep = (const uint64_t *)ctx->d_rep.p; rv.ids = (const int32_t *)ctx->d_ids.p;
    ctx->epoch++;
    ctx->routing_ready = true;
    ctx->build_n = 0; ctx->build_added = 0;
    // optional copies for the host's own persistence, in fspann_routing_upload's layout
    if (ids_out) CK(cudaMemcpy(ids_out, ctx->d_ids.p, sizeof(int32_t) * (size_t)TD * N, cudaMemcpyDeviceToHost));
    if (rep_code_out) CK(cudaMemcpy(rep_code_out, ctx->d_rep.p, sizeof(uint64_t) * (size_t)TD * P * W, cudaMemcpyDeviceToHost));
    if (min_key_out || max_key_out) {
        std::vector<int64_t> inter(2 * (size_t)TD * P);
        CK(cudaMemcpy(inter.data(), ctx->d_keys.p, sizeof(int64_t) * inter.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < (size_t)TD * P; i++) { if (min_key_out) min_key_out[i] = inter[2 * i]; if (max_key_out) max_key_out[i] = inter[2 * i + 1]; }
    }
    release(ctx->b_codes); release(ctx->b_scratch); release(ctx->b_ids); release(ctx->b_keys); release(ctx->b_rep);
    return FSPANN_OK;
}

int fspann_routing_build(fspann_ctx *ctx, int64_t N, const double *vectors, const int32_t *staged_ids, int64_t *min_key_out,
                         int64_t *max_key_out, uint64_t *rep_code_out, int32_t *ids_out) {
    if (!ctx) return FSPANN_E_ARG;
    if (!vectors || !staged_ids) return fail(ctx, FSPANN_E_ARG, "null array");
    int rc = fspann_routing_build_begin(ctx, N); if (rc) return rc;
    rc = fspann_routing_build_add(ctx, 0, N, vectors); if (rc) return rc;
    return fspann_routing_build_finish(ctx, staged_ids, min_key_out, max_key_out, rep_code_out, ids_out);
}

int fspann_deleted_set(fspann_ctx *ctx, const uint8_t *flags, int64_t n) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->epoch++;
    if (!flags || n <= 0) {
        ctx->rv.deleted = nullptr; ctx->rv.n_deleted = 0; ctx->sv.deleted = nullptr; ctx->sv.n_deleted = 0;
        return FSPANN_OK;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ENSURE(ctx->d_deleted, (size_t)n);
    CK(cudaMemcpy(ctx->d_deleted.p, flags, (size_t)n, cudaMemcpyHostToDevice));
    ctx->rv.deleted = (const uint8_t *)ctx->d_deleted.p; ctx->rv.n_deleted = n;
    ctx->sv.deleted = (const uint8_t *)ctx->d_deleted.p; ctx->sv.n_deleted = n;
    return FSPANN_OK;
}

static int store_write(fspann_ctx *ctx, int64_t n, const int32_t *ids, const uint8_t *iv, const uint8_t *ct, const int32_t *ver) {
    const int dim = ctx->sv.dim;
    const size_t ct_row = 8 * (size_t)dim + 16;
    const int64_t chunk = 65536;
    ENSURE(ctx->s_stage_a, (size_t)chunk * 12);
    ENSURE(ctx->s_stage_b, (size_t)chunk * ct_row);
    ENSURE(ctx->s_stage_c, (size_t)chunk * 8);
    for (int64_t s = 0; s < n; s += chunk) {
        const int64_t c = std::min(chunk, n - s);
        CK(cudaMemcpyAsync(ctx->s_stage_a.p, iv + (size_t)s * 12, (size_t)c * 12, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->s_stage_b.p, ct + (size_t)s * ct_row, (size_t)c * ct_row, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->s_stage_c.p, ver + s, (size_t)c * 4, cudaMemcpyHostToDevice, ctx->stream));
        int32_t *d_ids = nullptr;
        if (ids) {
            d_ids = (int32_t *)ctx->s_stage_c.p + chunk;
            CK(cudaMemcpyAsync(d_ids, ids + s, (size_t)c * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        uint8_t *base = (uint8_t *)ctx->d_rec.p + (ids ? 0 : (size_t)s * ctx->sv.rec_stride);
        LAUNCHED(launch_store_pack(ctx->stream, base, ctx->sv.rec_stride, dim, c, d_ids, (const uint8_t *)ctx->s_stage_a.p,
                                   (const uint8_t *)ctx->s_stage_b.p, (const int32_t *)ctx->s_stage_c.p));
        CK(cudaStreamSynchronize(ctx->stream));   // staging buffers are reused
    }
    return 0;
}

int fspann_store_upload_shard(fspann_ctx *ctx, int64_t id_base, int64_t N, int64_t n_global, int32_t dim, const uint8_t *iv, const uint8_t *ct,
                              const int32_t *key_version) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!iv || !ct || !key_version) return fail(ctx, FSPANN_E_ARG, "null array");
    if (N <= 0 || dim <= 0) return fail(ctx, FSPANN_E_ARG, "non-positive N/dim");
    if (id_base < 0 || n_global < id_base + N) return fail(ctx, FSPANN_E_ARG, "shard [%lld, %lld) outside the global id space %lld",
                                                           (long long)id_base, (long long)(id_base + N), (long long)n_global);
    if (n_global >= (1LL << 31) - 1) return fail(ctx, FSPANN_E_ARG, "N too large for int32 ids");
    ctx->store_ready = false;
    const int64_t stride = ((32 + 8LL * dim) + 15) / 16 * 16;
    ENSURE(ctx->d_rec, (size_t)N * stride + 64);
    ENSURE(ctx->d_touched, sizeof(uint32_t) * (size_t)((N + 31) / 32 + 1));
    CK(cudaMemsetAsync(ctx->d_rec.p, 0, (size_t)N * stride + 64, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_touched.p, 0, sizeof(uint32_t) * (size_t)((N + 31) / 32 + 1), ctx->stream));
    const bool dim_changed = ctx->sv.dim != dim;
    ctx->epoch++;
    ctx->sv.N = N; ctx->sv.id_base = id_base; ctx->sv.n_global = n_global;
    ctx->sv.dim = dim; ctx->sv.rec_stride = stride; ctx->sv.rec = (const uint8_t *)ctx->d_rec.p;
    if (dim_changed || !ctx->sv.hpow) { int rc = rebuild_keys(ctx); if (rc) return rc; }
    int rc = store_write(ctx, N, nullptr, iv, ct, key_version);
    if (rc) return rc;
    ctx->store_ready = true;
    return FSPANN_OK;
}

// Empty (zeroed) shard whose records arrive later through fspann_store_encrypt_dev / fspann_store_update: Setup of a store that is
// produced on the device and never fits the host in one piece (config 4).
int fspann_store_alloc_shard(fspann_ctx *ctx, int64_t id_base, int64_t N, int64_t n_global, int32_t dim) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (N <= 0 || dim <= 0) return fail(ctx, FSPANN_E_ARG, "non-positive N/dim");
    if (id_base < 0 || n_global < id_base + N) return fail(ctx, FSPANN_E_ARG, "shard [%lld, %lld) outside the global id space %lld",
                                                           (long long)id_base, (long long)(id_base + N), (long long)n_global);
    if (n_global >= (1LL << 31) - 1) return fail(ctx, FSPANN_E_ARG, "N too large for int32 ids");
    ctx->store_ready = false;
    const int64_t stride = ((32 + 8LL * dim) + 15) / 16 * 16;
    ENSURE(ctx->d_rec, (size_t)N * stride + 64);
    ENSURE(ctx->d_touched, sizeof(uint32_t) * (size_t)((N + 31) / 32 + 1));
    CK(cudaMemsetAsync(ctx->d_rec.p, 0, (size_t)N * stride + 64, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_touched.p, 0, sizeof(uint32_t) * (size_t)((N + 31) / 32 + 1), ctx->stream));
    const bool dim_changed = ctx->sv.dim != dim;
    ctx->epoch++;
    ctx->sv.N = N; ctx->sv.id_base = id_base; ctx->sv.n_global = n_global;
    ctx->sv.dim = dim; ctx->sv.rec_stride = stride; ctx->sv.rec = (const uint8_t *)ctx->d_rec.p;
    if (dim_changed || !ctx->sv.hpow) { int rc = rebuild_keys(ctx); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->store_ready = true;      // records of version 0 with a zero tag never authenticate: an unwritten row is a TAG_FAIL / NO_KEY candidate
    return FSPANN_OK;
}

// encryptToPoint (AGC:55-112) for the n vectors of ids first_id .. first_id+n-1, written DIRECTLY into this context's HBM store
// (device-resident FP64 vectors [n][dim] and IVs [n][12]; AAD id:<id>|v:<version>|d:<dim>).  The ids must lie inside the shard.
int fspann_store_encrypt_dev(fspann_ctx *ctx, int64_t first_id, int64_t n, const double *d_vectors, const uint8_t *d_ivs, int32_t version) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->store_ready) return fail(ctx, FSPANN_E_STATE, "record store not allocated (fspann_store_alloc_shard / fspann_store_upload_shard)");
    if (n == 0) return FSPANN_OK;
    if (n < 0 || !d_vectors || !d_ivs) return fail(ctx, FSPANN_E_ARG, "null array");
    if (first_id < ctx->sv.id_base || first_id + n > ctx->sv.id_base + ctx->sv.N) return fail(ctx, FSPANN_E_ARG, "ids [%lld, %lld) are not held by this context",
                                                                                            (long long)first_id, (long long)(first_id + n));
    if (!ctx->keys.count(version)) return fail(ctx, FSPANN_E_ARG, "key version %d is not live (KRS:82-88)", version);
    const int64_t chunk = 1 << 20;
    StoreView tv = ctx->sv;
    tv.deleted = nullptr; tv.n_deleted = 0;
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        const int c = (int)std::min<int64_t>(chunk, n - s0);
        ENSURE(ctx->m_list, sizeof(int32_t) * (size_t)c);
        ENSURE(ctx->m_flag, (size_t)c);
        tv.rec = ctx->sv.rec + (size_t)(first_id + s0 - ctx->sv.id_base) * ctx->sv.rec_stride;
        tv.N = c; tv.id_base = first_id + s0;                     // the AAD binds the GLOBAL id (EP:80-83)
        LAUNCHED(launch_iota(ctx->stream, (int32_t *)ctx->m_list.p, c));
        LAUNCHED(launch_encrypt_xcrypt(ctx->stream, tv, c, d_vectors + (size_t)s0 * tv.dim, d_ivs + (size_t)s0 * 12, version, (uint8_t *)ctx->m_flag.p, ctx->sm_count));
        LAUNCHED(launch_gcm_tag(ctx->stream, tv, (const int32_t *)ctx->m_list.p, nullptr, c, nullptr, (const uint8_t *)ctx->m_flag.p, ctx->sm_count));
    }
    return FSPANN_OK;
}

int fspann_store_upload(fspann_ctx *ctx, int64_t N, int32_t dim, const uint8_t *iv, const uint8_t *ct, const int32_t *key_version) {
    return fspann_store_upload_shard(ctx, 0, N, N, dim, iv, ct, key_version);
}

int fspann_store_update(fspann_ctx *ctx, int64_t n, const int32_t *ids, const uint8_t *iv, const uint8_t *ct, const int32_t *key_version) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->store_ready) return fail(ctx, FSPANN_E_STATE, "record store not uploaded");
    if (n == 0) return FSPANN_OK;
    if (!ids || !iv || !ct || !key_version || n < 0) return fail(ctx, FSPANN_E_ARG, "null array");
    std::vector<int32_t> local((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        if (ids[i] < ctx->sv.id_base || ids[i] >= ctx->sv.id_base + ctx->sv.N) return fail(ctx, FSPANN_E_ARG, "id %d is not held by this context", ids[i]);
        local[(size_t)i] = (int32_t)(ids[i] - ctx->sv.id_base);
    }
    return store_write(ctx, n, local.data(), iv, ct, key_version);
}

// ---- Migrate on the device: KeyRotationServiceImpl.reencryptTouched (KRS:215-289) -------------------------------------------
int fspann_migrate(fspann_ctx *ctx, int64_t n, const int32_t *ids, const uint8_t *fresh_ivs, int32_t target_version,
                   uint8_t *reencrypted_out, uint8_t *iv_out, uint8_t *ct_out, int64_t *n_reencrypted_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_reencrypted_out) *n_reencrypted_out = 0;
    if (!ctx->store_ready) return fail(ctx, FSPANN_E_STATE, "record store not uploaded");
    if (n < 0 || (n > 0 && (!ids || !fresh_ivs))) return fail(ctx, FSPANN_E_ARG, "null array");
    if (!ctx->keys.count(target_version)) return fail(ctx, FSPANN_E_ARG, "target key version %d is not live (KRS:82-88)", target_version);
    if (reencrypted_out) memset(reencrypted_out, 0, (size_t)n);
    if (n == 0) return FSPANN_OK;                                                      // KRS:220-223
    // new LinkedHashSet<>(touchedIds) (KRS:232): first occurrence of every id, in order; ids this store does not hold are
    // skipped like a null loadEncryptedPoint (KRS:243).  pos[] = position of the kept entry in the caller's list.
    std::vector<int32_t> local, pos;
    {
        std::vector<uint8_t> seen((size_t)ctx->sv.N, 0);
        for (int64_t i = 0; i < n; i++) {
            const int64_t li = (int64_t)ids[i] - ctx->sv.id_base;
            if (li < 0 || li >= ctx->sv.N || seen[(size_t)li]) continue;
            seen[(size_t)li] = 1;
            local.push_back((int32_t)li); pos.push_back((int32_t)i);
        }
    }
    const int dim = ctx->sv.dim;
    const size_t ct_row = 8 * (size_t)dim + 16;
    const int64_t chunk = 131072;
    int64_t total = 0;
    std::vector<uint8_t> h_iv, h_flag, h_oiv, h_oct;
    for (int64_t s0 = 0; s0 < (int64_t)local.size(); s0 += chunk) {
        const int c = (int)std::min<int64_t>(chunk, (int64_t)local.size() - s0);
        ENSURE(ctx->m_list, sizeof(int32_t) * (size_t)c);
        ENSURE(ctx->m_iv, (size_t)c * 12);
        ENSURE(ctx->m_verdict, (size_t)c);
        ENSURE(ctx->m_flag, (size_t)c);
        h_iv.resize((size_t)c * 12);
        for (int j = 0; j < c; j++) memcpy(&h_iv[(size_t)j * 12], fresh_ivs + (size_t)pos[(size_t)(s0 + j)] * 12, 12);
        CK(cudaMemcpyAsync(ctx->m_list.p, local.data() + s0, sizeof(int32_t) * (size_t)c, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->m_iv.p, h_iv.data(), (size_t)c * 12, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->m_verdict.p, FSPANN_V_NO_KEY, (size_t)c, ctx->stream));   // getVersion(old) throws -> skip (KRS:277)
        // decryptFromPoint under the stored version: authenticate first ...
        LAUNCHED(launch_gcm_tag(ctx->stream, ctx->sv, (const int32_t *)ctx->m_list.p, nullptr, c, (uint8_t *)ctx->m_verdict.p, nullptr, ctx->sm_count));
        // ... then decrypt + re-encrypt under the target version with the fresh IV, in place, and seal with the new tag
        LAUNCHED(launch_migrate_xcrypt(ctx->stream, ctx->sv, c, (const int32_t *)ctx->m_list.p, (const uint8_t *)ctx->m_iv.p, target_version,
                                       (const uint8_t *)ctx->m_verdict.p, (uint8_t *)ctx->m_flag.p, ctx->sm_count));
        LAUNCHED(launch_gcm_tag(ctx->stream, ctx->sv, (const int32_t *)ctx->m_list.p, nullptr, c, nullptr, (const uint8_t *)ctx->m_flag.p, ctx->sm_count));
        h_flag.resize((size_t)c);
        CK(cudaMemcpyAsync(h_flag.data(), ctx->m_flag.p, (size_t)c, cudaMemcpyDeviceToHost, ctx->stream));
        if (iv_out || ct_out) {
            ENSURE(ctx->m_out_iv, (size_t)c * 12);
            ENSURE(ctx->m_out_ct, (size_t)c * ct_row);
            ENSURE(ctx->m_out_ver, sizeof(int32_t) * (size_t)c);
            LAUNCHED(launch_store_unpack(ctx->stream, ctx->sv.rec, ctx->sv.rec_stride, dim, c, (const int32_t *)ctx->m_list.p, (uint8_t *)ctx->m_out_iv.p,
                                         (uint8_t *)ctx->m_out_ct.p, (int32_t *)ctx->m_out_ver.p));
            h_oiv.resize((size_t)c * 12); h_oct.resize((size_t)c * ct_row);
            CK(cudaMemcpyAsync(h_oiv.data(), ctx->m_out_iv.p, (size_t)c * 12, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(h_oct.data(), ctx->m_out_ct.p, (size_t)c * ct_row, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < c; j++) {
            if (!h_flag[(size_t)j]) continue;
            const size_t at = (size_t)pos[(size_t)(s0 + j)];
            total++;
            if (reencrypted_out) reencrypted_out[at] = 1;
            if (iv_out) memcpy(iv_out + at * 12, &h_oiv[(size_t)j * 12], 12);
            if (ct_out) memcpy(ct_out + at * ct_row, &h_oct[(size_t)j * ct_row], ct_row);
        }
    }
    if (n_reencrypted_out) *n_reencrypted_out = total;
    return FSPANN_OK;
}

// ---- Setup-side bulk encryption: AesGcmCryptoService.encryptToPoint (AGC:55-112) for n vectors -----------------------------
int fspann_encrypt_batch(fspann_ctx *ctx, int64_t n, int32_t dim, const int32_t *ids, const double *vectors, const uint8_t *ivs,
                         int32_t version, uint8_t *ct_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n < 0 || dim <= 0) return fail(ctx, FSPANN_E_ARG, "non-positive n/dim");
    if (n == 0) return FSPANN_OK;
    if (!ids || !vectors || !ivs || !ct_out) return fail(ctx, FSPANN_E_ARG, "id, plaintext, and key cannot be null");   // AGC:57-59
    if (!ctx->keys.count(version)) return fail(ctx, FSPANN_E_ARG, "key version %d is not live (KRS:82-88)", version);
    for (int64_t i = 0; i < n; i++) if (ids[i] < 0) return fail(ctx, FSPANN_E_ARG, "negative id %d", ids[i]);
    const size_t ct_row = 8 * (size_t)dim + 16;
    const int64_t stride = ((32 + 8LL * dim) + 15) / 16 * 16;
    const int64_t chunk = 131072;
    StoreView tv = ctx->sv;                       // key ring / tables of the context, records in a scratch buffer
    tv.dim = dim; tv.rec_stride = stride; tv.id_base = 0; tv.deleted = nullptr; tv.n_deleted = 0;
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
        const int c = (int)std::min<int64_t>(chunk, n - s0);
        ENSURE(ctx->m_rec, (size_t)c * stride + 64);
        ENSURE(ctx->m_vec, sizeof(double) * (size_t)c * dim);
        ENSURE(ctx->m_iv, (size_t)c * 12);
        ENSURE(ctx->m_gid, sizeof(int32_t) * (size_t)c);
        ENSURE(ctx->m_list, sizeof(int32_t) * (size_t)c);
        ENSURE(ctx->m_flag, (size_t)c);
        ENSURE(ctx->m_out_iv, (size_t)c * 12);
        ENSURE(ctx->m_out_ct, (size_t)c * ct_row);
        ENSURE(ctx->m_out_ver, sizeof(int32_t) * (size_t)c);
        tv.rec = (const uint8_t *)ctx->m_rec.p; tv.N = c; tv.n_global = c;
        std::vector<int32_t> iota((size_t)c);
        for (int j = 0; j < c; j++) iota[(size_t)j] = j;
        CK(cudaMemsetAsync(ctx->m_rec.p, 0, (size_t)c * stride + 64, ctx->stream));
        CK(cudaMemcpyAsync(ctx->m_vec.p, vectors + (size_t)s0 * dim, sizeof(double) * (size_t)c * dim, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->m_iv.p, ivs + (size_t)s0 * 12, (size_t)c * 12, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->m_gid.p, ids + s0, sizeof(int32_t) * (size_t)c, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->m_list.p, iota.data(), sizeof(int32_t) * (size_t)c, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCHED(launch_encrypt_xcrypt(ctx->stream, tv, c, (const double *)ctx->m_vec.p, (const uint8_t *)ctx->m_iv.p, version, (uint8_t *)ctx->m_flag.p,
                                       ctx->sm_count));
        LAUNCHED(launch_gcm_tag(ctx->stream, tv, (const int32_t *)ctx->m_list.p, (const int32_t *)ctx->m_gid.p, c, nullptr, (const uint8_t *)ctx->m_flag.p,
                                ctx->sm_count));
        LAUNCHED(launch_store_unpack(ctx->stream, tv.rec, stride, dim, c, nullptr, (uint8_t *)ctx->m_out_iv.p, (uint8_t *)ctx->m_out_ct.p,
                                     (int32_t *)ctx->m_out_ver.p));
        CK(cudaMemcpyAsync(ct_out + (size_t)s0 * ct_row, ctx->m_out_ct.p, (size_t)c * ct_row, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // iota / staging reuse
    }
    // the plaintext staging buffer is wiped: vectors never outlive the call on the device
    CK(cudaMemsetAsync(ctx->m_vec.p, 0, ctx->m_vec.bytes, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

// ---- exact ground truth and recall@K on the device: GroundtruthPrecompute.run (api/.../GroundtruthPrecompute.java:218-276)
// and the recall of computeMetricsAtK (FSA:785-794) --------------------------------------------------------------------------
static bool all_finite_f32(const float *v, int64_t n) {
    for (int64_t i = 0; i < n; i++) if (!std::isfinite(v[i])) return false;
    return true;
}
int fspann_groundtruth(fspann_ctx *ctx, int64_t N, int32_t dim, const float *base, int64_t Q, const float *queries, int32_t K,
                       int32_t *gt_ids_out, double *gt_d2_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!base || !queries || !gt_ids_out) return fail(ctx, FSPANN_E_ARG, "null array");
    if (N <= 0 || Q <= 0 || dim <= 0) return fail(ctx, FSPANN_E_ARG, "Empty or malformed vector files (zero records).");   // GroundtruthPrecompute:241-243
    if (N >= (1LL << 31) - 1) return fail(ctx, FSPANN_E_ARG, "N too large for int32 ids");
    if (K < 1 || K > N || K > gt_max_k()) return fail(ctx, FSPANN_E_ARG, "K=%d must be in [1, min(N, %d)] (the caller clamps like kFinal, :238)", K, gt_max_k());
    if (!all_finite_f32(base, N * dim) || !all_finite_f32(queries, Q * dim)) return fail(ctx, FSPANN_E_ARG, "vector contains NaN/Inf");
    int64_t Qc = (int64_t)(2e9 / (8.0 * (double)N));
    Qc = std::max<int64_t>(64, Qc / 64 * 64);
    Qc = std::min<int64_t>(Qc, (Q + 63) / 64 * 64);
    ENSURE(ctx->g_base, sizeof(float) * (size_t)N * dim);
    ENSURE(ctx->g_q, sizeof(float) * (size_t)Qc * dim);
    ENSURE(ctx->g_dist, sizeof(double) * (size_t)Qc * N);
    ENSURE(ctx->g_ids, sizeof(int32_t) * (size_t)Qc * K);
    ENSURE(ctx->g_d2, sizeof(double) * (size_t)Qc * K);
    ENSURE(ctx->g_flag, sizeof(int32_t));
    CK(cudaMemcpyAsync(ctx->g_base.p, base, sizeof(float) * (size_t)N * dim, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->g_flag.p, 0, sizeof(int32_t), ctx->stream));
    for (int64_t q0 = 0; q0 < Q; q0 += Qc) {
        const int c = (int)std::min(Qc, Q - q0);
        CK(cudaMemcpyAsync(ctx->g_q.p, queries + (size_t)q0 * dim, sizeof(float) * (size_t)c * dim, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCHED(launch_gt_chunk(ctx->stream, (const float *)ctx->g_base.p, N, dim, (const float *)ctx->g_q.p, c, K, (double *)ctx->g_dist.p,
                                 (int32_t *)ctx->g_ids.p, (double *)ctx->g_d2.p, (int32_t *)ctx->g_flag.p, ctx->sm_count));
        CK(cudaMemcpyAsync(gt_ids_out + (size_t)q0 * K, ctx->g_ids.p, sizeof(int32_t) * (size_t)c * K, cudaMemcpyDeviceToHost, ctx->stream));
        if (gt_d2_out) CK(cudaMemcpyAsync(gt_d2_out + (size_t)q0 * K, ctx->g_d2.p, sizeof(double) * (size_t)c * K, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    release(ctx->g_dist);
    return FSPANN_OK;
}

int fspann_recall_batch(fspann_ctx *ctx, int64_t Q, int32_t K, const int32_t *gt_ids, int32_t gt_stride, const int32_t *result_ids,
                        int32_t result_stride, const int32_t *n_ret, double *recall_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !gt_ids || !result_ids || !recall_out) return fail(ctx, FSPANN_E_ARG, "null array");
    if (K < 1 || gt_stride < K || result_stride < 1) return fail(ctx, FSPANN_E_ARG, "groundtruth row shorter than K=%d (FSA:779-782 returns NaN metrics)", K);
    ENSURE(ctx->g_ids, sizeof(int32_t) * (size_t)Q * gt_stride);
    ENSURE(ctx->g_res, sizeof(int32_t) * (size_t)Q * result_stride);
    ENSURE(ctx->g_nret, sizeof(int32_t) * (size_t)Q);
    ENSURE(ctx->g_rec, sizeof(double) * (size_t)Q);
    CK(cudaMemcpyAsync(ctx->g_ids.p, gt_ids, sizeof(int32_t) * (size_t)Q * gt_stride, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->g_res.p, result_ids, sizeof(int32_t) * (size_t)Q * result_stride, cudaMemcpyHostToDevice, ctx->stream));
    if (n_ret) CK(cudaMemcpyAsync(ctx->g_nret.p, n_ret, sizeof(int32_t) * (size_t)Q, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCHED(launch_recall(ctx->stream, (int)Q, K, (const int32_t *)ctx->g_ids.p, gt_stride, (const int32_t *)ctx->g_res.p, std::min(result_stride, result_stride),
                           n_ret ? (const int32_t *)ctx->g_nret.p : nullptr, (double *)ctx->g_rec.p));
    CK(cudaMemcpyAsync(recall_out, ctx->g_rec.p, sizeof(double) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

int fspann_keys_set(fspann_ctx *ctx, int32_t version, const uint8_t key[32]) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!key) return fail(ctx, FSPANN_E_ARG, "null key");
    if (!ctx->keys.count(version) && (int)ctx->keys.size() >= kMaxKeys) return fail(ctx, FSPANN_E_STATE, "more than %d live key versions", kMaxKeys);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->keys[version] = std::vector<uint8_t>(key, key + 32);
    return rebuild_keys(ctx);
}
int fspann_keys_retire(fspann_ctx *ctx, int32_t version) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    auto it = ctx->keys.find(version);
    if (it == ctx->keys.end()) return FSPANN_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    std::fill(it->second.begin(), it->second.end(), 0);   // SecureKeyDeletion analogue: wipe before dropping
    ctx->keys.erase(it);
    return rebuild_keys(ctx);
}

int fspann_tokengen_batch(fspann_ctx *ctx, int64_t Q, const double *queries, uint64_t *codes_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_routing(ctx); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !queries || !codes_out) return fail(ctx, FSPANN_E_ARG, "null array");
    const RoutingView &rv = ctx->rv;
    if (!all_finite(queries, Q * rv.dim)) return fail(ctx, FSPANN_E_ARG, "Vector contains NaN/Inf (Coding:357-359)");
    const size_t cb = sizeof(uint64_t) * (size_t)Q * rv.TD * rv.W;
    ENSURE(ctx->s_queries, sizeof(double) * (size_t)Q * rv.dim);
    ENSURE(ctx->s_codes, cb);
    CK(cudaMemcpyAsync(ctx->s_queries.p, queries, sizeof(double) * (size_t)Q * rv.dim, cudaMemcpyHostToDevice, ctx->stream));
    { int rc_ = run_tokengen(ctx, Q, (const double *)ctx->s_queries.p, (uint64_t *)ctx->s_codes.p); if (rc_) return rc_; }
    CK(cudaMemcpyAsync(codes_out, ctx->s_codes.p, cb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

int fspann_tokengen_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, uint64_t *d_codes) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_routing(ctx); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !d_queries || !d_codes) return fail(ctx, FSPANN_E_ARG, "null array");
    const int64_t chunk = 1 << 20;                               // the re-check list is sized per launch
    for (int64_t s0 = 0; s0 < Q; s0 += chunk) {
        rc = run_tokengen(ctx, std::min(chunk, Q - s0), d_queries + (size_t)s0 * ctx->rv.dim, d_codes + (size_t)s0 * ctx->rv.TD * ctx->rv.W);
        if (rc) return rc;
    }
    return FSPANN_OK;
}

int fspann_route_batch(fspann_ctx *ctx, int64_t Q, const uint64_t *codes, int32_t probes, int64_t hard_cap, int32_t ham_threshold,
                       int32_t B, int32_t *cand_ids_out, int32_t *cand_scores_out, int32_t *n_cand_out, int32_t *raw_seen_out,
                       int32_t *unique_out) {
    (void)ham_threshold;  // the list is score-sorted, so QSI:176-199's two passes select exactly the first B entries
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_routing(ctx); if (rc) return rc;
    if (ctx->rv.n_ids <= 1 && ctx->rv.P <= 1) return fail(ctx, FSPANN_E_STATE, "Index not finalized: only GFunctions uploaded");
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !codes || !cand_ids_out || !n_cand_out) return fail(ctx, FSPANN_E_ARG, "null array");
    if (B <= 0) return fail(ctx, FSPANN_E_ARG, "refinementLimit must be > 0");
    const RoutingView &rv = ctx->rv;
    const size_t cb = sizeof(uint64_t) * (size_t)Q * rv.TD * rv.W, ib = sizeof(int32_t) * (size_t)Q * B;
    ENSURE(ctx->s_codes, cb); ENSURE(ctx->s_cand_ids, ib); ENSURE(ctx->s_cand_sc, ib);
    ENSURE(ctx->s_ncand, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->s_raw, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->s_uniq_cnt, sizeof(int32_t) * (size_t)Q);
    CK(cudaMemcpyAsync(ctx->s_codes.p, codes, cb, cudaMemcpyHostToDevice, ctx->stream));
    rc = do_route(ctx, Q, (const uint64_t *)ctx->s_codes.p, probes, hard_cap, B, (int32_t *)ctx->s_cand_ids.p, (int32_t *)ctx->s_cand_sc.p,
                  (int32_t *)ctx->s_ncand.p, (int32_t *)ctx->s_raw.p, (int32_t *)ctx->s_uniq_cnt.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(cand_ids_out, ctx->s_cand_ids.p, ib, cudaMemcpyDeviceToHost, ctx->stream));
    if (cand_scores_out) CK(cudaMemcpyAsync(cand_scores_out, ctx->s_cand_sc.p, ib, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(n_cand_out, ctx->s_ncand.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    if (raw_seen_out) CK(cudaMemcpyAsync(raw_seen_out, ctx->s_raw.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    if (unique_out) CK(cudaMemcpyAsync(unique_out, ctx->s_uniq_cnt.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

int fspann_refine_batch(fspann_ctx *ctx, int64_t Q, const double *queries, const int32_t *cand_ids, const int32_t *n_cand,
                        int32_t cand_stride, int32_t k, int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out,
                        uint8_t *verdict_out, int32_t *n_decrypted_out) {
    return fspann_refine_batch_ex(ctx, Q, queries, cand_ids, n_cand, cand_stride, k, topk_ids_out, topk_dist_out, nullptr, n_ret_out, verdict_out,
                                  n_decrypted_out);
}

int fspann_refine_batch_ex(fspann_ctx *ctx, int64_t Q, const double *queries, const int32_t *cand_ids, const int32_t *n_cand,
                           int32_t cand_stride, int32_t k, int32_t *topk_ids_out, double *topk_dist_out, int32_t *topk_rank_out,
                           int32_t *n_ret_out, uint8_t *verdict_out, int32_t *n_decrypted_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_store(ctx); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !queries || !cand_ids || !n_cand || !topk_ids_out || !topk_dist_out || !n_ret_out) return fail(ctx, FSPANN_E_ARG, "null array");
    if (k <= 0 || cand_stride <= 0) return fail(ctx, FSPANN_E_ARG, "topK and cand_stride must be > 0");
    for (int64_t q = 0; q < Q; q++)
        if (n_cand[q] < 0 || n_cand[q] > cand_stride) return fail(ctx, FSPANN_E_ARG, "n_cand[%lld] out of range", (long long)q);
    const int dim = ctx->sv.dim;
    const size_t ib = sizeof(int32_t) * (size_t)Q * cand_stride;
    ENSURE(ctx->s_queries, sizeof(double) * (size_t)Q * dim); ENSURE(ctx->s_cand_ids, ib); ENSURE(ctx->s_ncand, sizeof(int32_t) * (size_t)Q);
    ENSURE(ctx->s_topk_ids, sizeof(int32_t) * (size_t)Q * k); ENSURE(ctx->s_topk_dist, sizeof(double) * (size_t)Q * k);
    ENSURE(ctx->s_nret, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->s_ndec, sizeof(int32_t) * (size_t)Q);
    CK(cudaMemcpyAsync(ctx->s_queries.p, queries, sizeof(double) * (size_t)Q * dim, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->s_cand_ids.p, cand_ids, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->s_ncand.p, n_cand, sizeof(int32_t) * (size_t)Q, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t l0 = ctx->launches;
    rc = record_ev(ctx, 0); if (rc) return rc;
    rc = record_ev(ctx, 1); if (rc) return rc;
    rc = record_ev(ctx, 2); if (rc) return rc;
    ENSURE(ctx->s_topk_rank, sizeof(int32_t) * (size_t)Q * k);
    ctx->want_rank = topk_rank_out ? (int32_t *)ctx->s_topk_rank.p : nullptr;
    rc = do_refine(ctx, Q, (const double *)ctx->s_queries.p, (const int32_t *)ctx->s_cand_ids.p, (const int32_t *)ctx->s_ncand.p, cand_stride, k,
                   (int32_t *)ctx->s_topk_ids.p, (double *)ctx->s_topk_dist.p, (int32_t *)ctx->s_nret.p, (int32_t *)ctx->s_ndec.p, true);
    ctx->want_rank = nullptr;
    if (rc) return rc;
    if (topk_rank_out) CK(cudaMemcpyAsync(topk_rank_out, ctx->s_topk_rank.p, sizeof(int32_t) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->ev_valid = true; ctx->last_call_launches = ctx->launches - l0;
    CK(cudaMemcpyAsync(topk_ids_out, ctx->s_topk_ids.p, sizeof(int32_t) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(topk_dist_out, ctx->s_topk_dist.p, sizeof(double) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(n_ret_out, ctx->s_nret.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    if (verdict_out) CK(cudaMemcpyAsync(verdict_out, ctx->s_verdict.p, (size_t)Q * cand_stride, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_decrypted_out) CK(cudaMemcpyAsync(n_decrypted_out, ctx->s_ndec.p, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

// One full pass (tokengen -> route -> refine -> counters) over device-resident queries.  d_codes_in != nullptr: the token's own
// codes are routed on (PIS:600 token.getBitCodes()) and TokenGen does not run.
namespace fspabi {
int search_pass(fspann_ctx *ctx, int64_t Q, const double *d_queries, const uint64_t *d_codes_in, int k, int probes, int64_t hard_cap, int B,
                       int retried, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_nret, int64_t *d_counters, bool main_pass) {
    const RoutingView &rv = ctx->rv;
    DevBuf &b_codes = main_pass ? ctx->s_codes : ctx->r_codes;
    DevBuf &b_cid = main_pass ? ctx->s_cand_ids : ctx->t_cand_ids;
    DevBuf &b_csc = main_pass ? ctx->s_cand_sc : ctx->t_cand_sc;
    DevBuf &b_nc = main_pass ? ctx->s_ncand : ctx->t_ncand;
    DevBuf &b_raw = main_pass ? ctx->s_raw : ctx->t_raw;
    DevBuf &b_un = main_pass ? ctx->s_uniq_cnt : ctx->t_uniq_cnt;
    DevBuf &b_nd = main_pass ? ctx->s_ndec : ctx->t_ndec;
    if (!d_codes_in) ENSURE(b_codes, sizeof(uint64_t) * (size_t)Q * rv.TD * rv.W);
    ENSURE(b_cid, sizeof(int32_t) * (size_t)Q * B); ENSURE(b_csc, sizeof(int32_t) * (size_t)Q * B);
    ENSURE(b_nc, sizeof(int32_t) * (size_t)Q); ENSURE(b_raw, sizeof(int32_t) * (size_t)Q); ENSURE(b_un, sizeof(int32_t) * (size_t)Q);
    ENSURE(b_nd, sizeof(int32_t) * (size_t)Q);
    int rc;
    // upload plan of the host-pointer entries (consumed here, once)
    const int chunks = main_pass && !d_codes_in ? ctx->h2d_chunks : 0;
    const bool late = main_pass && ctx->h2d_late;
    int64_t cq[5];
    for (int c = 0; c < 5; c++) cq[c] = ctx->h2d_q0[c];
    ctx->h2d_chunks = 0; ctx->h2d_late = false;
    if (main_pass) { rc = record_ev(ctx, 0); if (rc) return rc; }
    if (chunks > 0) {
        // TokenGen + Route per arriving chunk (both are independent per query); Refine below groups the whole batch
        const size_t crow = (size_t)rv.TD * rv.W;
        for (int c = 0; c < chunks; c++) {
            const int64_t q0 = cq[c], n = cq[c + 1] - cq[c];
            if (n <= 0) continue;
            CK(cudaStreamWaitEvent(ctx->stream, ctx->h2d_ev[c], 0));
            { int rc_ = run_tokengen(ctx, n, d_queries + (size_t)q0 * rv.dim, (uint64_t *)b_codes.p + (size_t)q0 * crow); if (rc_) return rc_; }
            if (c == 0) { rc = record_ev(ctx, 1); if (rc) return rc; }
            rc = do_route(ctx, n, (const uint64_t *)b_codes.p + (size_t)q0 * crow, probes, hard_cap, B, (int32_t *)b_cid.p + (size_t)q0 * B,
                          (int32_t *)b_csc.p + (size_t)q0 * B, (int32_t *)b_nc.p + q0, (int32_t *)b_raw.p + q0, (int32_t *)b_un.p + q0);
            if (rc) return rc;
        }
    } else {
        if (!d_codes_in) { int rc_ = run_tokengen(ctx, Q, d_queries, (uint64_t *)b_codes.p); if (rc_) return rc_; }
        if (main_pass) { rc = record_ev(ctx, 1); if (rc) return rc; }
        rc = do_route(ctx, Q, d_codes_in ? d_codes_in : (const uint64_t *)b_codes.p, probes, hard_cap, B, (int32_t *)b_cid.p, (int32_t *)b_csc.p,
                      (int32_t *)b_nc.p, (int32_t *)b_raw.p, (int32_t *)b_un.p);
        if (rc) return rc;
    }
    if (late) CK(cudaStreamWaitEvent(ctx->stream, ctx->h2d_ev[4], 0));     // the queries themselves: Refine is their first reader
    if (main_pass) { rc = record_ev(ctx, 2); if (rc) return rc; }
    rc = do_refine(ctx, Q, d_queries, (const int32_t *)b_cid.p, (const int32_t *)b_nc.p, B, k, d_topk_ids, d_topk_dist, d_nret, (int32_t *)b_nd.p, main_pass);
    if (rc) return rc;
    if (d_counters)
        LAUNCHED(launch_counters(ctx->stream, Q, (const int32_t *)b_raw.p, (const int32_t *)b_un.p, (const int32_t *)b_nd.p, d_nret,
                                 (const int32_t *)b_nc.p, retried, d_counters, (const uint8_t *)ctx->s_qfinite.p));
    return 0;
}
}  // namespace fspabi

namespace {
struct HostOut {            // host result buffers of the host-pointer entry points (nullptr members are skipped)
    int32_t *topk_ids; double *topk_dist; int32_t *n_ret; int64_t *counters;
};

int copy_results(fspann_ctx *ctx, int64_t Q, int k, const HostOut &h, const int32_t *d_ids, const double *d_dist, const int32_t *d_nret, const int64_t *d_cnt) {
    CK(cudaMemcpyAsync(h.topk_ids, d_ids, sizeof(int32_t) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h.topk_dist, d_dist, sizeof(double) * (size_t)Q * k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h.n_ret, d_nret, sizeof(int32_t) * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    if (h.counters && d_cnt) CK(cudaMemcpyAsync(h.counters, d_cnt, sizeof(int64_t) * FSPANN_COUNTERS * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

// createToken + QueryServiceImpl.search for a batch whose queries (and optionally codes) are resident in HBM.
//  * The adaptive retry (QSI:327-337, 444-447) is decided ON THE DEVICE (retry_select_kernel): the host reads back two integers
//    (how many queries retry, whether a query held NaN/Inf) together with the results, so the common no-retry batch costs ONE
//    stream synchronisation; only a batch that does retry runs the second pass (10 probes, those rows only) and copies again.
//  * host != nullptr: results are copied to the host buffers (the host-pointer entry points); otherwise nothing is copied and,
//    without allow_retry, nothing synchronises.
//  * reject_nonfinite: fail the call with FSPANN_E_ARG when a query holds NaN/Inf (createToken throws, Coding:357-359); otherwise such
//    a query just returns empty and touches nothing (QSI:137).
int search_core(fspann_ctx *ctx, int64_t Q, const double *d_queries, const uint64_t *d_codes_in, int k, int probes, int64_t hard_cap, int B,
                bool allow_retry, bool reject_nonfinite, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_n_ret, int64_t *d_counters,
                const HostOut *host) {
    if (probes <= 0) probes = 5;   // DEFAULT_MAX_PROBES (PIS:93) when no override is configured (PIS:880-888)
    const int64_t l0 = ctx->launches;
    const bool need_flags = allow_retry || reject_nonfinite;
    int rc;
    // first pass + retry decision: eager, or (small batches, from the second use of the same shape and buffers on) one CUDA-graph launch
    auto first_pass = [&]() -> int {
        int rc_ = search_pass(ctx, Q, d_queries, d_codes_in, k, probes, hard_cap, B, 0, d_topk_ids, d_topk_dist, d_n_ret, d_counters, true);
        if (rc_) return rc_;
        if (need_flags) {
            ENSURE(ctx->r_rows, sizeof(int32_t) * (size_t)Q);
            ENSURE(ctx->s_retry_out, sizeof(int32_t) * 4);
            LAUNCHED(launch_retry_select(ctx->stream, Q, k, d_n_ret, (const int32_t *)ctx->s_ndec.p, (const int32_t *)ctx->s_f32_exact.p,
                                         (int32_t *)ctx->r_rows.p, (int32_t *)ctx->s_retry_out.p));
        }
        return 0;
    };
    bool done = false;
    if (ctx->opt_graphs && Q <= 64) {
        const std::array<int64_t, 12> key = {Q, k, probes, hard_cap, B, (int64_t)(intptr_t)d_queries, (int64_t)(intptr_t)d_codes_in, (int64_t)(intptr_t)d_topk_ids,
                                             (int64_t)(intptr_t)d_topk_dist, (int64_t)(intptr_t)d_n_ret, (int64_t)(intptr_t)d_counters ^ (need_flags ? 1 : 0), ctx->epoch};
        for (auto &g : ctx->graphs)
            if (g.key == key) {
                CK(cudaGraphLaunch(g.exec, ctx->stream));
                ctx->launches += g.launches;
                ctx->graph_replays++;
                ctx->ev_valid = false;
                done = true;
                break;
            }
        if (!done && ctx->graph_seen == key) {                     // second use, nothing moved in between: capture, then launch
            const int64_t lc = ctx->launches;
            ctx->capturing = true;
            cudaError_t ce = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
            int rc_ = ce == cudaSuccess ? first_pass() : FSPANN_E_CUDA;
            cudaGraph_t graph = nullptr;
            if (ce == cudaSuccess) ce = cudaStreamEndCapture(ctx->stream, &graph);
            ctx->capturing = false;
            cudaGraphExec_t exec = nullptr;
            if (rc_ == 0 && ce == cudaSuccess && graph && ctx->epoch == key[11] && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                if (ctx->graphs.size() >= 16) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
                ctx->graphs.push_back({key, exec, ctx->launches - lc});
                ctx->graph_captures++;
                CK(cudaGraphLaunch(exec, ctx->stream));
                ctx->ev_valid = false;
                done = true;
            } else {
                cudaGetLastError();                                // a failed capture leaves a sticky-looking (but cleared) error: fall back to eager
                ctx->launches = lc;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (!done) {
            rc = first_pass(); if (rc) return rc;
            ctx->ev_valid = true;
            done = true;
            std::array<int64_t, 12> seen = key;
            seen[11] = ctx->epoch;                                 // the eager pass may have grown a buffer
            ctx->graph_seen = seen;
        }
    }
    if (!done) { rc = first_pass(); if (rc) return rc; ctx->ev_valid = true; }
    if (need_flags) {
        if (!ctx->h_pin) { CK(cudaHostAlloc((void **)&ctx->h_pin, sizeof(int32_t) * 16, cudaHostAllocDefault)); ctx->h_pin_ints = 16; }
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->s_retry_out.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (host) { rc = copy_results(ctx, Q, k, *host, d_topk_ids, d_topk_dist, d_n_ret, d_counters); if (rc) return rc; }
    if (need_flags || host) CK(cudaStreamSynchronize(ctx->stream));
    if (need_flags) {
        ctx->last_queries_finite = ctx->h_pin[1] == 0;
        if (reject_nonfinite && !ctx->last_queries_finite) return fail(ctx, FSPANN_E_ARG, "Vector contains NaN/Inf (Coding:357-359)");
        const int64_t R = allow_retry ? ctx->h_pin[0] : 0;
        if (R > 0) {
            const int dim = ctx->rv.dim;
            const size_t code_row = sizeof(uint64_t) * (size_t)ctx->rv.TD * ctx->rv.W;
            const size_t Rc = (size_t)(Q <= 64 ? Q : R);               // small batches: size once, so a later, larger retry set moves nothing (graphs)
            ENSURE(ctx->r_queries, sizeof(double) * Rc * dim);
            ENSURE(ctx->r_topk_ids, sizeof(int32_t) * Rc * k); ENSURE(ctx->r_topk_dist, sizeof(double) * Rc * k);
            ENSURE(ctx->r_nret, sizeof(int32_t) * Rc); ENSURE(ctx->r_counters, sizeof(int64_t) * FSPANN_COUNTERS * Rc);
            const int32_t *rows = (const int32_t *)ctx->r_rows.p;
            LAUNCHED(launch_gather_rows(ctx->stream, d_queries, ctx->r_queries.p, rows, R, sizeof(double) * dim, false));
            const uint64_t *codes2 = nullptr;
            if (d_codes_in) {
                ENSURE(ctx->r_codes, code_row * Rc);
                LAUNCHED(launch_gather_rows(ctx->stream, d_codes_in, ctx->r_codes.p, rows, R, (int64_t)code_row, false));
                codes2 = (const uint64_t *)ctx->r_codes.p;
            }
            rc = search_pass(ctx, R, (const double *)ctx->r_queries.p, codes2, k, 10, hard_cap, B, 1, (int32_t *)ctx->r_topk_ids.p,
                             (double *)ctx->r_topk_dist.p, (int32_t *)ctx->r_nret.p, d_counters ? (int64_t *)ctx->r_counters.p : nullptr, false);
            if (rc) return rc;
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->r_topk_ids.p, d_topk_ids, rows, R, sizeof(int32_t) * k, true));
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->r_topk_dist.p, d_topk_dist, rows, R, sizeof(double) * k, true));
            LAUNCHED(launch_gather_rows(ctx->stream, ctx->r_nret.p, d_n_ret, rows, R, sizeof(int32_t), true));
            if (d_counters)
                LAUNCHED(launch_gather_rows(ctx->stream, ctx->r_counters.p, d_counters, rows, R, sizeof(int64_t) * FSPANN_COUNTERS, true));
            if (host) { rc = copy_results(ctx, Q, k, *host, d_topk_ids, d_topk_dist, d_n_ret, d_counters); if (rc) return rc; }
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    ctx->last_call_launches = ctx->launches - l0;
    return FSPANN_OK;
}

int check_search_args(fspann_ctx *ctx, int64_t Q, int k, int B) {
    int rc = check_routing(ctx); if (rc) return rc;
    rc = check_store(ctx); if (rc) return rc;
    if (ctx->rv.n_ids <= 1 && ctx->rv.P <= 1) return fail(ctx, FSPANN_E_STATE, "Index not finalized: only GFunctions uploaded");
    if (Q < 0) return fail(ctx, FSPANN_E_ARG, "negative batch size");
    if (k <= 0) return fail(ctx, FSPANN_E_ARG, "topK must be > 0 (QTF:65)");
    if (B <= 0) return fail(ctx, FSPANN_E_ARG, "refinementLimit must be > 0");
    return 0;
}
}  // namespace

int fspann_search_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                            int32_t ham_threshold, int32_t allow_retry, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_n_ret,
                            int64_t *d_counters) {
    (void)ham_threshold;
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_search_args(ctx, Q, k, B); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (!d_queries || !d_topk_ids || !d_topk_dist || !d_n_ret) return fail(ctx, FSPANN_E_ARG, "null array");
    return search_core(ctx, Q, d_queries, nullptr, k, probes, hard_cap, B, allow_retry != 0, allow_retry != 0, d_topk_ids, d_topk_dist, d_n_ret,
                       d_counters, nullptr);
}

int fspann_search_tokens_dev(fspann_ctx *ctx, int64_t Q, const uint64_t *d_codes, const double *d_queries, int32_t k, int32_t probes,
                             int64_t hard_cap, int32_t B, int32_t ham_threshold, int32_t allow_retry, int32_t *d_topk_ids, double *d_topk_dist,
                             int32_t *d_n_ret, int64_t *d_counters) {
    (void)ham_threshold;
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_search_args(ctx, Q, k, B); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (!d_codes || !d_queries || !d_topk_ids || !d_topk_dist || !d_n_ret) return fail(ctx, FSPANN_E_ARG, "null array");
    return search_core(ctx, Q, d_queries, d_codes, k, probes, hard_cap, B, allow_retry != 0, false, d_topk_ids, d_topk_dist, d_n_ret, d_counters,
                       nullptr);
}

static int search_host(fspann_ctx *ctx, int64_t Q, const uint64_t *codes, const double *queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                       int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_search_args(ctx, Q, k, B); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (!queries || !topk_ids_out || !topk_dist_out || !n_ret_out) return fail(ctx, FSPANN_E_ARG, "null array");
    const int dim = ctx->rv.dim;
    // createToken (no codes supplied) throws on NaN/Inf (Coding:357-359): small batches are checked on the host before any work, large
    // ones by the device pass that scans every query value anyway (flag read back with the retry decision).  With codes supplied the
    // token exists already and QSI:137 applies instead: that query returns empty, the others are unaffected.
    if (!codes && Q * dim <= 65536 && !all_finite(queries, Q * dim)) return fail(ctx, FSPANN_E_ARG, "Vector contains NaN/Inf (Coding:357-359)");
    ENSURE(ctx->s_queries, sizeof(double) * (size_t)Q * dim);
    ENSURE(ctx->s_topk_ids, sizeof(int32_t) * (size_t)Q * k); ENSURE(ctx->s_topk_dist, sizeof(double) * (size_t)Q * k);
    ENSURE(ctx->s_nret, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->s_counters, sizeof(int64_t) * FSPANN_COUNTERS * (size_t)Q);
    const uint64_t *d_codes = nullptr;
    ctx->h2d_chunks = 0; ctx->h2d_late = false;
    const bool overlap = ctx->opt_h2d_overlap > 0 && Q >= 4096;
    if (overlap) {                                                          // nothing of an earlier call still reads s_queries (host entries end synchronised)
        CK(cudaEventRecord(ctx->h2d_ev[4], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->h2d_ev[4], 0));
    }
    if (codes) {
        const size_t cb = sizeof(uint64_t) * (size_t)Q * ctx->rv.TD * ctx->rv.W;
        ENSURE(ctx->s_codes_in, cb);
        CK(cudaMemcpyAsync(ctx->s_codes_in.p, codes, cb, cudaMemcpyHostToDevice, ctx->stream));
        d_codes = (const uint64_t *)ctx->s_codes_in.p;
        if (overlap) {                                                      // Route needs only the codes: the query upload runs beside it
            CK(cudaMemcpyAsync(ctx->s_queries.p, queries, sizeof(double) * (size_t)Q * dim, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->h2d_ev[4], ctx->copy_stream));
            ctx->h2d_late = true;
        } else {
            CK(cudaMemcpyAsync(ctx->s_queries.p, queries, sizeof(double) * (size_t)Q * dim, cudaMemcpyHostToDevice, ctx->stream));
        }
    } else if (overlap) {
        // up to 4 chunks, each a multiple of the Route grid (2 CTAs per SM) so no chunk ends in a partial wave
        const int64_t wave = 2 * (int64_t)ctx->sm_count;
        const int nch = std::max(2, ctx->opt_h2d_overlap);
        const int64_t per = ((Q + 3) / 4 + wave - 1) / wave * wave;         // nch = 2: a quarter first, then the rest
        int n = 0;
        for (int64_t q0 = 0; q0 < Q && n < nch; q0 += per) {
            const int64_t q1 = n == nch - 1 ? Q : std::min<int64_t>(Q, q0 + per);
            ctx->h2d_q0[n] = q0; ctx->h2d_q0[n + 1] = q1;
            CK(cudaMemcpyAsync((double *)ctx->s_queries.p + (size_t)q0 * dim, queries + (size_t)q0 * dim, sizeof(double) * (size_t)(q1 - q0) * dim,
                               cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->h2d_ev[n], ctx->copy_stream));
            n++;
            if (q1 >= Q) break;
        }
        ctx->h2d_chunks = n;
    } else {
        CK(cudaMemcpyAsync(ctx->s_queries.p, queries, sizeof(double) * (size_t)Q * dim, cudaMemcpyHostToDevice, ctx->stream));
    }
    const HostOut h{topk_ids_out, topk_dist_out, n_ret_out, counters_out};
    rc = search_core(ctx, Q, (const double *)ctx->s_queries.p, d_codes, k, probes, hard_cap, B, true, codes == nullptr, (int32_t *)ctx->s_topk_ids.p,
                     (double *)ctx->s_topk_dist.p, (int32_t *)ctx->s_nret.p, (int64_t *)ctx->s_counters.p, &h);
    if (ctx->h2d_chunks || ctx->h2d_late) {                                 // the call failed before its first pass consumed the upload plan
        ctx->h2d_chunks = 0; ctx->h2d_late = false;
        cudaStreamSynchronize(ctx->copy_stream);
    }
    return rc;
}

int fspann_search_batch(fspann_ctx *ctx, int64_t Q, const double *queries, int32_t k, int32_t probes, int64_t hard_cap, int32_t B,
                        int32_t ham_threshold, int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out) {
    (void)ham_threshold;
    return search_host(ctx, Q, nullptr, queries, k, probes, hard_cap, B, topk_ids_out, topk_dist_out, n_ret_out, counters_out);
}

int fspann_search_tokens(fspann_ctx *ctx, int64_t Q, const uint64_t *codes, const double *queries, int32_t k, int32_t probes, int64_t hard_cap,
                         int32_t B, int32_t ham_threshold, int32_t *topk_ids_out, double *topk_dist_out, int32_t *n_ret_out, int64_t *counters_out) {
    (void)ham_threshold;
    if (ctx && Q > 0 && !codes) return fail(ctx, FSPANN_E_STATE, "MSANNP violation: QueryToken missing BitSet codes (PIS:604-606)");
    return search_host(ctx, Q, codes, queries, k, probes, hard_cap, B, topk_ids_out, topk_dist_out, n_ret_out, counters_out);
}

// ---- device-resident building blocks of the database-sharded deployment (BASELINE config 4) ---------------------
// Route is query-parallel (every rank routes its slice of the batch on the replicated index), Refine is data-parallel (every rank
// refines, for ALL queries, the candidates its store shard holds); two all-gathers (candidate lists, per-shard top-k) connect them.
int fspann_route_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, int32_t probes, int64_t hard_cap, int32_t B,
                           int32_t *d_cand_ids, int32_t *d_n_cand, int32_t *d_raw_seen, int32_t *d_unique) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_routing(ctx); if (rc) return rc;
    if (ctx->rv.n_ids <= 1 && ctx->rv.P <= 1) return fail(ctx, FSPANN_E_STATE, "Index not finalized: only GFunctions uploaded");
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !d_queries || !d_cand_ids || !d_n_cand) return fail(ctx, FSPANN_E_ARG, "null array");
    if (B <= 0) return fail(ctx, FSPANN_E_ARG, "refinementLimit must be > 0");
    if (probes <= 0) probes = 5;
    const RoutingView &rv = ctx->rv;
    ENSURE(ctx->s_codes, sizeof(uint64_t) * (size_t)Q * rv.TD * rv.W);
    ENSURE(ctx->s_cand_sc, sizeof(int32_t) * (size_t)Q * B);
    ENSURE(ctx->s_raw, sizeof(int32_t) * (size_t)Q); ENSURE(ctx->s_uniq_cnt, sizeof(int32_t) * (size_t)Q);
    { int rc_ = run_tokengen(ctx, Q, d_queries, (uint64_t *)ctx->s_codes.p); if (rc_) return rc_; }
    return do_route(ctx, Q, (const uint64_t *)ctx->s_codes.p, probes, hard_cap, B, d_cand_ids, (int32_t *)ctx->s_cand_sc.p, d_n_cand,
                    d_raw_seen ? d_raw_seen : (int32_t *)ctx->s_raw.p, d_unique ? d_unique : (int32_t *)ctx->s_uniq_cnt.p);
}

int fspann_refine_batch_dev(fspann_ctx *ctx, int64_t Q, const double *d_queries, const int32_t *d_cand_ids, const int32_t *d_n_cand,
                            int32_t cand_stride, int32_t k, int32_t *d_topk_ids, double *d_topk_dist, int32_t *d_topk_rank,
                            int32_t *d_n_ret, int32_t *d_n_decrypted) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = check_store(ctx); if (rc) return rc;
    if (Q == 0) return FSPANN_OK;
    if (Q < 0 || !d_queries || !d_cand_ids || !d_n_cand || !d_topk_ids || !d_topk_dist || !d_n_ret) return fail(ctx, FSPANN_E_ARG, "null array");
    if (k <= 0 || cand_stride <= 0) return fail(ctx, FSPANN_E_ARG, "topK and cand_stride must be > 0");
    ENSURE(ctx->s_ndec, sizeof(int32_t) * (size_t)Q);
    ctx->want_rank = d_topk_rank;
    rc = do_refine(ctx, Q, d_queries, d_cand_ids, d_n_cand, cand_stride, k, d_topk_ids, d_topk_dist, d_n_ret,
                   d_n_decrypted ? d_n_decrypted : (int32_t *)ctx->s_ndec.p, false);
    ctx->want_rank = nullptr;
    return rc;
}

int fspann_merge_topk_dev(fspann_ctx *ctx, int32_t n_shards, int64_t Q, int32_t k, const double *d_dist, const int32_t *d_rank,
                          const int32_t *d_ids, int32_t *d_out_ids, double *d_out_dist, int32_t *d_out_n_ret) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (Q == 0) return FSPANN_OK;
    if (n_shards <= 0 || Q < 0 || k <= 0 || !d_dist || !d_rank || !d_ids || !d_out_ids || !d_out_dist || !d_out_n_ret) return fail(ctx, FSPANN_E_ARG, "null array");
    LAUNCHED(launch_merge_topk(ctx->stream, n_shards, Q, k, d_dist, d_rank, d_ids, d_out_ids, d_out_dist, d_out_n_ret));
    return FSPANN_OK;
}

int fspann_touched_fetch(fspann_ctx *ctx, uint32_t *bitmap_out, int64_t n_words, int32_t clear) {
    if (!ctx) return FSPANN_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->store_ready) return fail(ctx, FSPANN_E_STATE, "record store not uploaded");
    const int64_t have = (ctx->sv.N + 31) / 32;
    if (n_words > have) n_words = have;
    if (bitmap_out && n_words > 0) CK(cudaMemcpyAsync(bitmap_out, ctx->d_touched.p, sizeof(uint32_t) * (size_t)n_words, cudaMemcpyDeviceToHost, ctx->stream));
    if (clear) CK(cudaMemsetAsync(ctx->d_touched.p, 0, sizeof(uint32_t) * (size_t)have, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
}

int64_t fspann_last_stage_ms(fspann_ctx *ctx, float out[6]) {
    if (!ctx || !out) return -1;
    for (int i = 0; i < 6; i++) out[i] = 0.f;
    if (!ctx->ev_valid) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    for (int i = 0; i < 6; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) out[i] = ms;
    }
    return ctx->last_call_launches;
}

int fspann_debug_decrypt(fspann_ctx *ctx, int64_t n, const int32_t *ids, double *plaintext_out, uint8_t *verdict_out) {
    if (!ctx) return FSPANN_E_ARG;
#ifndef FSPANN_DEBUG_TAP
    (void)n; (void)ids; (void)plaintext_out; (void)verdict_out;
    return fail(ctx, FSPANN_E_STATE, "production build: plaintext never leaves the SM (rebuild with -DFSPANN_DEBUG_TAP for parity tests)");
#else
    CK(cudaSetDevice(ctx->device));
    int rc = check_store(ctx); if (rc) return rc;
    if (n <= 0) return FSPANN_OK;
    if (!ids || !plaintext_out || !verdict_out) return fail(ctx, FSPANN_E_ARG, "null array");
    const int dim = ctx->sv.dim;
    ENSURE(ctx->s_cand_ids, sizeof(int32_t) * (size_t)n); ENSURE(ctx->s_dist, sizeof(double) * (size_t)n * dim); ENSURE(ctx->s_verdict, (size_t)n);
    CK(cudaMemcpyAsync(ctx->s_cand_ids.p, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCHED(launch_debug_decrypt(ctx->stream, ctx->sv, n, (const int32_t *)ctx->s_cand_ids.p, (double *)ctx->s_dist.p, (uint8_t *)ctx->s_verdict.p));
    CK(cudaMemcpyAsync(plaintext_out, ctx->s_dist.p, sizeof(double) * (size_t)n * dim, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(verdict_out, ctx->s_verdict.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSPANN_OK;
#endif
}

}  // extern "C"
