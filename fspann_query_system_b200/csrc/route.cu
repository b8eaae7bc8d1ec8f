// route.cu -- stage 2: partition probe, candidate gather, min-score de-duplication and ordered cut at B.
//
// Replaces PartitionedIndexService.lookupCandidatesWithScores (index/.../paper/PartitionedIndexService.java:592-715),
// collectPartitionOrdered (PIS:726-753), GreedyPartitioner.computeKey / findNearestPartition / hamming
// (index/.../paper/GreedyPartitioner.java:87-96, 101-130, 78-82) and the first-B cut of QueryServiceImpl stage A/A.5
// (query/.../service/QueryServiceImpl.java:153-214).
//
// What has to be reproduced exactly (the cut at B falls inside large tie classes of equal Hamming score):
//  * probe order per (table, division): best-first expansion of a contiguous interval around the centre partition,
//    java.util.PriorityQueue tie rule = the entry enqueued earlier wins, left is enqueued before right (PIS:657-685);
//  * every id of a probed partition inherits that partition's Hamming distance to its representative code; bestScore
//    keeps the minimum, and lastRawVisited counts first insertions plus strict improvements in visit order (PIS:736-751);
//  * the "bestScore.size() < HARD_CAP" early-outs before every (t,d) and before every poll (PIS:624,628,657-659);
//  * the result order = stable sort by score over java.util.HashMap<String,Long> iteration order (PIS:690-696), i.e.
//    by (score, bucket of String.hashCode(decimal id) in the final table size, first-insertion order).
//
// One CTA per query (persistent, grid-stride).  The CTA keeps a chained hash table with the SAME bucket function as the
// Java map (at its initial capacity) in an L2-resident scratch area: heads[cap0] + a node pool.  The T*D groups are
// processed in order (the algorithm is sequential across groups through the cap and the improvement counter), the
// <= probes*64 ids of one group in parallel (they are distinct, so no two threads touch the same map entry).
#include <algorithm>

#include "fspann_internal.cuh"

namespace fsp {

constexpr int RT_THREADS = 256;
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int RT_MAX_SCORE = 255;  // m*lambda <= 255 bits supported (W <= 4); score is packed into 8 bits

__device__ __forceinline__ int64_t code_key(const uint64_t *code, int W) {  // GP:87-96
    // code bit i (i < 63) -> key bit 62-i : reverse the low 63 bits of word 0
    (void)W;
    return (int64_t)(__brevll(code[0]) >> 1);
}
__device__ __forceinline__ int hamming_w(const uint64_t *a, const uint64_t *b, int W) {  // GP:78-82
    int c = 0;
    for (int w = 0; w < W; w++) c += __popcll(a[w] ^ b[w]);
    return c;
}

// GP:101-130 over the interleaved (minKey, maxKey) array of one (t,d).
__device__ int64_t find_nearest(const int64_t *__restrict__ keys, int64_t P, int64_t q) {
    int64_t lo = 0, hi = P - 1;
    while (lo <= hi) {
        const int64_t mid = (lo + hi) >> 1;
        const longlong2 mm = *reinterpret_cast<const longlong2 *>(keys + 2 * mid);
        if (q < mm.x) hi = mid - 1;
        else if (q > mm.y) lo = mid + 1;
        else return mid;
    }
    if (lo <= 0) return 0;
    if (lo >= P) return P - 1;
    const longlong2 l = *reinterpret_cast<const longlong2 *>(keys + 2 * (lo - 1));
    const longlong2 r = *reinterpret_cast<const longlong2 *>(keys + 2 * lo);
    const int64_t dl = q < l.x ? l.x - q : (q > l.y ? q - l.y : 0);
    const int64_t dr = q < r.x ? r.x - q : (q > r.y ? q - r.y : 0);
    return dl <= dr ? lo - 1 : lo;
}

// scratch is written with L2 atomics / .cg stores and re-used across queries: always read it through L2
__device__ __forceinline__ int32_t ld(const int32_t *p) { return __ldcg(p); }

struct RouteScratch {
    int32_t *head;     // [cap0]
    int32_t *nid;      // [max_nodes]
    int32_t *nval;     // [max_nodes]  score << 24 | seq
    int32_t *nnext;    // [max_nodes]
    int32_t *cl_id;    // [max_nodes]  compacted qualifying entries in HashMap iteration order
    int32_t *cl_sc;    // [max_nodes]
};

int64_t route_scratch_ints(int32_t cap0, int32_t max_nodes) { return (int64_t)cap0 + 5LL * max_nodes + 32; }

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

__global__ void __launch_bounds__(RT_THREADS) route_kernel(RoutingView rv, RouteParams p) {
    extern __shared__ int32_t sm_dyn[];                 // visits: part[TD*probes], score[TD*probes], nvis[TD]
    int32_t *v_part = sm_dyn;
    int32_t *v_score = v_part + rv.TD * p.probes;
    int32_t *v_n = v_score + rv.TD * p.probes;
    __shared__ int32_t s_count, s_raw, s_overflow;
    __shared__ int32_t s_hist[RT_MAX_SCORE + 2];
    __shared__ int32_t s_warp_tot[RT_WARPS];
    __shared__ int32_t s_cls[RT_WARPS][RT_MAX_SCORE + 1];   // per-warp running class offsets
    __shared__ int32_t s_misc[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = rv.W, TD = rv.TD, probes = p.probes;
    const int64_t P = rv.P, n_ids = rv.n_ids;
    const uint32_t cap0 = (uint32_t)p.cap0, mask0 = cap0 - 1;

    RouteScratch sc;
    int32_t *base = p.scratch + (size_t)blockIdx.x * p.scratch_ints;
    sc.head = base;
    sc.nid = sc.head + cap0;
    sc.nval = sc.nid + p.max_nodes;
    sc.nnext = sc.nval + p.max_nodes;
    sc.cl_id = sc.nnext + p.max_nodes;
    sc.cl_sc = sc.cl_id + p.max_nodes;

    const int64_t n_q = p.qlist ? (int64_t)*p.qlist_n : p.Q;          // all queries, or the list the fast kernels could not hold
    for (int64_t qi = blockIdx.x; qi < n_q; qi += gridDim.x) {
        const int64_t q = p.qlist ? (int64_t)p.qlist[qi] : qi;
        const uint64_t *qcodes = p.codes + (size_t)q * TD * W;
        __syncthreads();
        // ---- 0. reset ----
        for (uint32_t i = tid; i < cap0; i += RT_THREADS) __stcg(&sc.head[i], -1);
        for (int i = tid; i < RT_MAX_SCORE + 2; i += RT_THREADS) s_hist[i] = 0;
        if (tid == 0) { s_count = 0; s_raw = 0; s_overflow = 0; }

        // ---- 1. probe sequence of every (t,d): PIS:640-685 with the PriorityQueue reduced to its two frontier entries ----
        for (int g = tid; g < TD; g += RT_THREADS) {
            int nv = 0;
            if (P > 0 && probes > 0) {
                uint64_t qc[4];
                for (int w = 0; w < W; w++) qc[w] = qcodes[(size_t)g * W + w];
                const int64_t *keys = rv.keys + (size_t)g * P * 2;
                const uint64_t *rep = rv.rep + (size_t)g * P * W;
                const int64_t center = find_nearest(keys, P, code_key(qc, W));
                int64_t lo = center, hi = center;                 // visited interval [lo, hi]
                // frontier entries: valid flag, distance, enqueue order
                bool lv = false, rvd = false; int ldist = 0, rdist = 0; int lseq = 0, rseq = 0, seq = 0;
                int64_t cur = center; int curd = hamming_w(qc, rep + (size_t)center * W, W);
                for (;;) {
                    v_part[g * probes + nv] = (int32_t)cur;
                    v_score[g * probes + nv] = curd;
                    nv++;
                    // enqueue unvisited neighbours of `cur`: left first, then right (PIS:672-684)
                    if (cur == lo && lo - 1 >= 0 && !lv) { lv = true; ldist = hamming_w(qc, rep + (size_t)(lo - 1) * W, W); lseq = seq++; }
                    if (cur == hi && hi + 1 < P && !rvd) { rvd = true; rdist = hamming_w(qc, rep + (size_t)(hi + 1) * W, W); rseq = seq++; }
                    if (nv >= probes) break;
                    if (!lv && !rvd) break;                        // queue empty
                    bool take_left;
                    if (lv && rvd) take_left = (ldist < rdist) || (ldist == rdist && lseq < rseq);   // tie: older entry wins
                    else take_left = lv;
                    if (take_left) { lo -= 1; cur = lo; curd = ldist; lv = false; }
                    else { hi += 1; cur = hi; curd = rdist; rvd = false; }
                }
            }
            v_n[g] = nv;
        }
        __syncthreads();

        // ---- 2. gather + de-duplicate, group by group ----
        int my_raw = 0;
        for (int g = 0; g < TD; g++) {
            const int cnt_g = s_count;
            __syncthreads();                                        // everyone latched the size before anyone grows it
            if (cnt_g >= p.hard_cap) break;                         // PIS:624,628
            const int nv = v_n[g];
            const int32_t *gids = rv.ids + (size_t)g * n_ids;
            // fast path: the cap cannot bind before the last poll of this group, so all its visits run in parallel;
            // slow path: one visit at a time so the size check before every poll is exact (PIS:657-659)
            const bool fast = (int64_t)cnt_g + (int64_t)(nv > 0 ? nv - 1 : 0) * kBlock < p.hard_cap;
            for (int vs = 0; vs < (fast ? 1 : nv); vs++) {
                if (!fast) {
                    const int cnt_v = s_count;
                    __syncthreads();
                    if (cnt_v >= p.hard_cap) break;
                }
                const int v_lo = fast ? 0 : vs, v_hi = fast ? nv : vs + 1;
                for (int e = v_lo * kBlock + tid; e < v_hi * kBlock; e += RT_THREADS) {
                    const int v = e / kBlock, pos = e - v * kBlock;
                    const int64_t slot = (int64_t)v_part[g * probes + v] * kBlock + pos;
                    if (slot >= n_ids) continue;
                    const int32_t id = gids[slot];
                    if (rv.deleted && id >= 0 && id < rv.n_deleted && rv.deleted[id]) continue;   // PIS:739
                    const uint32_t score = (uint32_t)v_score[g * probes + v];
                    const uint32_t seq = (uint32_t)((g * probes + v) * kBlock + pos);
                    const uint32_t b = java_hash_decimal(id) & mask0;
                    // search the chain (another thread may be pushing a DIFFERENT id concurrently)
                    int32_t n = atomicAdd(&sc.head[b], 0);
                    int found = -1, len = 0;
                    while (n >= 0) {
                        if (ld(&sc.nid[n]) == id) { found = n; break; }
                        n = ld(&sc.nnext[n]);
                        len++;
                    }
                    if (found >= 0) {
                        const uint32_t val = (uint32_t)ld(&sc.nval[found]);
                        if (score < (val >> 24)) { __stcg(&sc.nval[found], (int32_t)((score << 24) | (val & 0xffffffu))); my_raw++; }  // improvement (PIS:747-750)
                    } else {
                        const int32_t nn = atomicAdd(&s_count, 1);
                        __stcg(&sc.nid[nn], id);
                        __stcg(&sc.nval[nn], (int32_t)((score << 24) | seq));
                        int32_t old = atomicAdd(&sc.head[b], 0);
                        for (;;) {
                            __stcg(&sc.nnext[nn], old);
                            __threadfence_block();
                            const int32_t prev = atomicCAS(&sc.head[b], old, nn);
                            if (prev == old) break;
                            old = prev;
                        }
                        my_raw++;
                        // 9th entry in a bin: Java would treeify.  `len` was counted before pushes that raced with this one, so
                        // re-count behind the node now that it is linked: the last node pushed into a bin sees the whole chain.
                        int behind = 0;
                        for (int32_t c = old; c >= 0 && behind < 8; c = ld(&sc.nnext[c])) behind++;
                        if (behind >= 8) s_overflow = 1;
                    }
                }
                __syncthreads();
            }
        }
        atomicAdd(&s_raw, my_raw);
        __syncthreads();
        const int n_unique = s_count;

        // ---- 3. score histogram -> cut score s* (smallest s with #(score <= s) >= min(B, n)) ----
        for (int i = tid; i < n_unique; i += RT_THREADS) atomicAdd(&s_hist[(uint32_t)ld(&sc.nval[i]) >> 24], 1);
        __syncthreads();
        const int want = min(p.B, n_unique);
        if (tid == 0) {
            int cum = 0, sstar = RT_MAX_SCORE;
            for (int s = 0; s <= RT_MAX_SCORE; s++) { cum += s_hist[s]; if (cum >= want) { sstar = s; break; } }
            s_misc[0] = sstar;
            // class bases for the output
            int run = 0;
            for (int s = 0; s <= RT_MAX_SCORE; s++) { const int c = s_hist[s]; s_hist[s] = run; run += c; }
        }
        __syncthreads();
        const int sstar = s_misc[0];

        // ---- 4. compaction of entries with score <= s* in java.util.HashMap iteration order ----
        // final Java table size: doubled while size > 0.75 * cap (HashMap.resize); a split keeps relative order, so the
        // order is: for hi in [0, ratio): for bucket b at cap0: chain entries whose next hash bits == hi, oldest first.
        uint32_t capF = cap0;
        while ((double)n_unique > 0.75 * (double)capF && capF < (1u << 30)) capF <<= 1;
        const uint32_t ratio = capF / cap0;
        int shift0 = 0; while ((1u << shift0) < cap0) shift0++;
        const uint32_t per_warp = (cap0 + RT_WARPS - 1) / RT_WARPS;
        const uint32_t wb_lo = warp * per_warp, wb_hi = min(cap0, wb_lo + per_warp);
        int cl_total = 0;
        for (uint32_t hi = 0; hi < ratio; hi++) {
            // pass A: count per warp
            int cnt = 0;
            for (uint32_t b = wb_lo + lane; b < wb_hi; b += 32) {
                for (int32_t n = ld(&sc.head[b]); n >= 0; n = ld(&sc.nnext[n])) {
                    if ((int)((uint32_t)ld(&sc.nval[n]) >> 24) > sstar) continue;
                    if (ratio > 1 && ((java_hash_decimal(ld(&sc.nid[n])) >> shift0) & (ratio - 1)) != hi) continue;
                    cnt++;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) s_warp_tot[warp] = cnt;
            __syncthreads();
            int wbase = cl_total, all = 0;
            for (int w = 0; w < RT_WARPS; w++) { if (w < warp) wbase += s_warp_tot[w]; all += s_warp_tot[w]; }
            // pass B: write, bucket order inside the warp's range, insertion (seq) order inside a bucket
            int run = wbase;
            for (uint32_t b0 = wb_lo; b0 < wb_hi; b0 += 32) {
                const uint32_t b = b0 + lane;
                int c = 0;
                if (b < wb_hi)
                    for (int32_t n = ld(&sc.head[b]); n >= 0; n = ld(&sc.nnext[n])) {
                        if ((int)((uint32_t)ld(&sc.nval[n]) >> 24) > sstar) continue;
                        if (ratio > 1 && ((java_hash_decimal(ld(&sc.nid[n])) >> shift0) & (ratio - 1)) != hi) continue;
                        c++;
                    }
                int tot;
                const int off = run + warp_excl_scan(c, lane, tot);
                if (c > 0) {
                    // emit the qualifying chain entries by increasing seq (selection by repeated minimum; chains are short)
                    int last_seq = -1;
                    for (int k = 0; k < c; k++) {
                        int best = -1, best_seq = 0x7fffffff;
                        for (int32_t n = ld(&sc.head[b]); n >= 0; n = ld(&sc.nnext[n])) {
                            const uint32_t val = (uint32_t)ld(&sc.nval[n]);
                            if ((int)(val >> 24) > sstar) continue;
                            if (ratio > 1 && ((java_hash_decimal(ld(&sc.nid[n])) >> shift0) & (ratio - 1)) != hi) continue;
                            const int sq = (int)(val & 0xffffffu);
                            if (sq > last_seq && sq < best_seq) { best_seq = sq; best = n; }
                        }
                        __stcg(&sc.cl_id[off + k], ld(&sc.nid[best]));
                        __stcg(&sc.cl_sc[off + k], (int32_t)((uint32_t)ld(&sc.nval[best]) >> 24));
                        last_seq = best_seq;
                    }
                }
                run += tot;
            }
            cl_total += all;
            __syncthreads();
        }

        // ---- 5. stable counting sort of the compact list by score -> first B entries in the reference's order ----
        for (int i = tid; i < RT_WARPS * (RT_MAX_SCORE + 1); i += RT_THREADS) (&s_cls[0][0])[i] = 0;
        __syncthreads();
        const int per_w = (cl_total + RT_WARPS - 1) / RT_WARPS;
        const int l_lo = min(cl_total, warp * per_w), l_hi = min(cl_total, l_lo + per_w);
        for (int i = l_lo + lane; i < l_hi; i += 32) atomicAdd(&s_cls[warp][ld(&sc.cl_sc[i])], 1);
        __syncthreads();
        // exclusive prefix over warps for each class, plus the class base
        for (int s = tid; s <= RT_MAX_SCORE; s += RT_THREADS) {
            int run = s_hist[s];
            for (int w = 0; w < RT_WARPS; w++) { const int c = s_cls[w][s]; s_cls[w][s] = run; run += c; }
        }
        __syncthreads();
        int32_t *out_id = p.cand_ids + (size_t)q * p.B, *out_sc = p.cand_scores + (size_t)q * p.B;
        for (int i0 = l_lo; i0 < l_hi; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < l_hi;
            const int s = act ? ld(&sc.cl_sc[i]) : -1 - lane;            // inactive lanes get unique keys
            const unsigned peers = __match_any_sync(0xffffffffu, s);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (act) {
                const int pos = s_cls[warp][s] + rank;
                if (pos < p.B) { out_id[pos] = ld(&sc.cl_id[i]); out_sc[pos] = s; }
            }
            __syncwarp();
            if (act && rank == 0) s_cls[warp][s] += __popc(peers);   // leader advances the class cursor
            __syncwarp();
        }
        if (tid == 0) {
            p.n_cand[q] = want;
            p.unique[q] = n_unique;
            p.raw_seen[q] = s_raw;
            if (s_overflow) *p.chain_overflow = 1;
        }
    }
}

int route_grid(int64_t Q, int sm_count) {
    int64_t g = (int64_t)sm_count * 2;
    if (g > Q) g = Q;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_route(cudaStream_t s, const RoutingView &rv, const RouteParams &p, int grid) {
    if (p.Q <= 0) return 0;
    const size_t smem = sizeof(int32_t) * ((size_t)2 * rv.TD * p.probes + rv.TD);
    route_kernel<<<grid, RT_THREADS, smem, s>>>(rv, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// =====================================================================================================================
// Fast path, first generation (one CTA per SM; today it serves the batches of at most one query per SM and the queries the two-CTA
// kernel further down hands back).  Used when (a) the HARD_CAP can never bind (T*D*probes*64 - 64 < hard_cap, so every poll of
// PIS:657-659 proceeds; route_fast2_kernel lifts this) and (b) the per-query working set fits in shared memory.  Then nothing in
// the algorithm is sequential any more except the per-id improvement count, which is resolved per id from its (short) occurrence chain:
//   kernel 1 (route_probe_kernel): one thread per (query, table, division): centre partition + probe order + scores;
//   kernel 2 (route_fast_kernel) : one 1024-thread CTA per query, everything in shared memory:
//        gather the <= T*D*probes*64 ids (coalesced 256-byte partition rows), insert them into an open-addressing table
//        of 16-bit positions (all occurrences of an id are chained through next_s), finalise per id (first position =
//        Java insertion order, min score, #strict improvements in visit order), radix-select the B smallest
//        (score, Java bucket, insertion order) keys and bitonic-sort just those.
// =====================================================================================================================
constexpr int RQ_THREADS = 1024;
constexpr uint32_t RQ_EMPTY = 0xffffu;

__global__ void route_probe_kernel(RoutingView rv, int64_t Q, const uint64_t *__restrict__ codes, int probes, int32_t *__restrict__ vis_part,
                                   uint8_t *__restrict__ vis_score, uint8_t *__restrict__ vis_n) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= Q * rv.TD) return;
    const int g = (int)(idx % rv.TD);
    const int W = rv.W;
    const int64_t P = rv.P;
    int nv = 0;
    if (P > 0 && probes > 0) {
        uint64_t qc[4];
        for (int w = 0; w < W; w++) qc[w] = codes[idx * W + w];
        const int64_t *keys = rv.keys + (size_t)g * P * 2;
        const uint64_t *rep = rv.rep + (size_t)g * P * W;
        const int64_t center = find_nearest(keys, P, code_key(qc, W));
        int64_t lo = center, hi = center;
        bool lv = false, rvd = false; int ldist = 0, rdist = 0, lseq = 0, rseq = 0, seq = 0;
        int64_t cur = center; int curd = hamming_w(qc, rep + (size_t)center * W, W);
        for (;;) {
            vis_part[idx * probes + nv] = (int32_t)cur;
            vis_score[idx * probes + nv] = (uint8_t)curd;
            nv++;
            if (cur == lo && lo - 1 >= 0 && !lv) { lv = true; ldist = hamming_w(qc, rep + (size_t)(lo - 1) * W, W); lseq = seq++; }
            if (cur == hi && hi + 1 < P && !rvd) { rvd = true; rdist = hamming_w(qc, rep + (size_t)(hi + 1) * W, W); rseq = seq++; }
            if (nv >= probes) break;
            if (!lv && !rvd) break;
            bool take_left;
            if (lv && rvd) take_left = (ldist < rdist) || (ldist == rdist && lseq < rseq);
            else take_left = lv;
            if (take_left) { lo -= 1; cur = lo; curd = ldist; lv = false; }
            else { hi += 1; cur = hi; curd = rdist; rvd = false; }
        }
    }
    vis_n[idx] = (uint8_t)nv;
}

// 16-bit compare-and-swap on a shared-memory table of uint16 (tw = the table viewed as 32-bit words).
__device__ __forceinline__ uint32_t cas16(uint32_t *tw, uint32_t slot, uint32_t expect, uint32_t val) {
    uint32_t *w = tw + (slot >> 1);
    const int shift = (slot & 1u) ? 16 : 0;
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
    for (;;) {
        const uint32_t cur = (old >> shift) & 0xffffu;
        if (cur != expect) return cur;
        const uint32_t neu = (old & ~(0xffffu << shift)) | (val << shift);
        const uint32_t prev = atomicCAS(w, old, neu);
        if (prev == old) return expect;
        old = prev;
    }
}

size_t route_fast_smem(int TD, int probes, int n_raw, int tbl, int sort_n, int wl_extra) {
    const size_t nvis = (size_t)TD * probes, nvis16 = (nvis + 15) / 16 * 16;
    size_t s = 0;
    s += sizeof(int32_t) * n_raw;                   // ids_s
    s += sizeof(uint16_t) * n_raw;                  // next_s (later: low 16 bits of the Java hash)
    s += (size_t)n_raw;                             // best_s
    s = (s + 15) / 16 * 16;
    s += sizeof(uint16_t) * tbl;                    // filter / exact table / class + selection lists
    s += sizeof(int64_t) * nvis;                    // vbase_s
    s += 2 * nvis16;                                // vs_s, vlen_s
    s += sizeof(uint16_t) * nvis16;                 // lowvis_s
    s += sizeof(uint16_t) * nvis16;                 // inv_cnt (involved / deleted positions per visit)
    s += sizeof(uint32_t) * (((size_t)n_raw + 31) / 32 + 3) / 4 * 4;   // inv_bm (one bit per position: involved)
    s += sizeof(uint16_t) * wl_extra;               // worklist head (continues into skey / sid)
    s += sizeof(uint64_t) * sort_n;                 // skey
    s += sizeof(int32_t) * sort_n;                  // sid
    return s + 64;
}

// Finds the first bin d (of nd <= 256) with base + sum(hist[0..d]) >= want; warp 0 only.  Writes s_out = {d, cum_before_d, hist[d]}.
__device__ __forceinline__ void warp_pick_digit(const int32_t *hist, int nd, int base, int want, int lane, int32_t *s_out) {
    int loc[8], sum = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) { const int d = lane * 8 + u; loc[u] = d < nd ? hist[d] : 0; sum += loc[u]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int excl = base + incl - sum;
    const unsigned hit = __ballot_sync(0xffffffffu, excl + sum >= want);
    const int owner = hit ? __ffs(hit) - 1 : 31;
    if (lane == owner) {
        int c = excl, d = lane * 8, u = 0;
        for (u = 0; u < 8; u++) { if (c + loc[u] >= want) break; c += loc[u]; }
        if (u == 8) { u = 7; c -= loc[7]; }
        d += u;
        if (d >= nd) { d = nd - 1; }
        s_out[0] = d; s_out[1] = c; s_out[2] = hist[d];
    }
}

// warp-aggregated histogram increment: one shared atomic per distinct bin per warp
__device__ __forceinline__ void hist_add(int32_t *hist, bool active, int bin, int lane) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(act, bin);
    if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

// warp-aggregated append to a shared list: returns this lane's slot (or -1), one shared atomic per warp
__device__ __forceinline__ int list_slot(int32_t *counter, bool in, int lane) {
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(counter, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    return in ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
}

// Bitonic sort of one 32-bit key per thread (element i lives in thread i): strides < 32 are exchanged with warp shuffles, larger
// strides through `buf` in shared memory.  N > 0: the network for exactly N elements, fully unrolled; N == 0: n at run time.
template <int N>
__device__ __forceinline__ uint32_t bitonic_sort32(uint32_t a, int i, uint32_t *buf, int n_rt = 0) {
    const int n = N > 0 ? N : n_rt;
#pragma unroll
    for (int k = 2; k <= n; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool asc = (i & k) == 0;
            uint32_t b;
            if (j >= 32) {
                __syncthreads();
                if (i < n) buf[i] = a;
                __syncthreads();
                b = i < n ? buf[i ^ j] : 0xffffffffu;
            } else {
                b = __shfl_xor_sync(0xffffffffu, a, j);
            }
            const bool lower = (i & j) == 0;                // this thread keeps the smaller of the pair when ascending
            a = (lower == asc) ? min(a, b) : max(a, b);     // (a select between VIMNMX results; the if/else form compiles to branches)
        }
    }
    return a;
}

// best_s is valid only for INVOLVED positions (bit set in inv_bm; written by the exact path): 255 = a later occurrence of a duplicated
// id, 0x80 | s = first occurrence of an id whose minimum score over all occurrences is s.  Every other valid position is the only
// occurrence of its id and its score is its visit's score vs_s[e >> 6].
constexpr uint32_t RQ_DUP = 0x80u;

__global__ void __launch_bounds__(RQ_THREADS, 1) route_fast_kernel(RoutingView rv, RouteParams p, RouteFastExtra x) {
    extern __shared__ __align__(16) unsigned char rq_smem[];
    const int TD = rv.TD, probes = p.probes, n_raw = x.n_raw, tbl = x.tbl, sort_n = x.sort_n, nvis = TD * probes;
    const int nvis16 = (nvis + 15) / 16 * 16;
    int32_t *ids_s = reinterpret_cast<int32_t *>(rq_smem);
    uint16_t *next_s = reinterpret_cast<uint16_t *>(ids_s + n_raw);
    uint8_t *best_s = reinterpret_cast<uint8_t *>(next_s + n_raw);
    uint32_t *table_w = reinterpret_cast<uint32_t *>(rq_smem + ((size_t)7 * n_raw + 15) / 16 * 16);
    uint16_t *table = reinterpret_cast<uint16_t *>(table_w);
    int64_t *vbase_s = reinterpret_cast<int64_t *>(table + tbl);                // per visit: offset of its partition row in rv.ids
    uint8_t *vs_s = reinterpret_cast<uint8_t *>(vbase_s + nvis);               // per visit: Hamming score of the partition
    uint8_t *vlen_s = vs_s + nvis16;                                           // valid ids in the visited partition row
    uint16_t *lowvis_s = reinterpret_cast<uint16_t *>(vlen_s + nvis16);
    uint16_t *inv_cnt = lowvis_s + nvis16;                                     // per visit: positions that are NOT singles (involved, or deleted ids)
    uint32_t *inv_bm = reinterpret_cast<uint32_t *>(inv_cnt + nvis16);         // bit e: position e is involved (its score lives in best_s[e])
    const int nbm = ((n_raw + 31) / 32 + 3) / 4 * 4;
    uint16_t *wl = reinterpret_cast<uint16_t *>(inv_bm + nbm);                 // worklist: wl_extra entries, then over skey / sid
    uint64_t *skey = reinterpret_cast<uint64_t *>(wl + x.wl_extra);
    int32_t *sid = reinterpret_cast<int32_t *>(skey + sort_n);
    const int wl_cap = x.wl_extra + 6 * sort_n;
    uint16_t *cls = table;                                                     // after the exact path the table is dead: cut-class list ...
    uint16_t *sel = table + n_raw;                                             // ... and the list of selected positions (sort_n entries)
    __shared__ int32_t s_hist[256];
    __shared__ int32_t s_raw, s_uniq, s_m, s_ncls, s_nwl, s_nlow, s_pick[3], s_wsum[32];

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t n_ids = rv.n_ids;
    const int fshift = 32 - (31 - __clz(tbl * 4 - 1) + 1);                             // buckets per filter array = tbl * 4 (a power of two; two arrays share the table region)

    const int64_t n_q = x.qlist ? (int64_t)*x.qlist_n : p.Q;          // all queries, or the overflow list of route_fast2_kernel
    for (int64_t qi = blockIdx.x; qi < n_q; qi += gridDim.x) {
        const int64_t q = x.qlist ? (int64_t)x.qlist[qi] : qi;
        __syncthreads();
        // ---- 1. stage visits, reset the duplicate filter (it lives in the table region: 8*tbl two-bit buckets) ----
        for (int v = tid; v < nvis; v += RQ_THREADS) {
            const int g = v / probes, j = v - g * probes;
            const int64_t vi = (q * TD + g) * probes + j;
            const bool valid = j < (int)x.vis_n[q * TD + g];
            const int64_t row = valid ? (int64_t)x.vis_part[vi] * kBlock : 0;
            vbase_s[v] = (int64_t)g * n_ids + row;
            vlen_s[v] = valid ? (uint8_t)min((int64_t)kBlock, n_ids - row) : 0;
            vs_s[v] = valid ? x.vis_score[vi] : 255;
        }
        for (int i = tid; i < tbl / 8; i += RQ_THREADS) reinterpret_cast<uint4 *>(table_w)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < nbm; i += RQ_THREADS) inv_bm[i] = 0u;
        for (int i = tid; i < nvis16 / 2; i += RQ_THREADS) reinterpret_cast<uint32_t *>(inv_cnt)[i] = 0u;
        if (tid < 256) s_hist[tid] = 0;
        if (tid == 0) { s_raw = 0; s_uniq = 0; s_m = 0; s_ncls = 0; s_nwl = 0; s_nlow = 0; }
        __syncthreads();
        // ---- 2. gather ids and set the duplicate filter.  A thread owns 4 consecutive positions (one 16-byte slice of a 256-byte
        //         partition row; two slices in flight).  The filter is TWO arrays of 2-bit buckets (different multiplicative
        //         hashes): bit0 = "bucket taken", bit1 = "taken twice".  ~95 % of the visited ids occur once; an id whose bucket
        //         was taken once in EITHER array is the only occurrence of its id, everything else takes the exact path below.
        uint32_t *F = table_w;
        const uint32_t f2_off = (uint32_t)tbl >> 2;                        // second array: the upper half of the region (words)
        const bool vec_ok = (n_ids & 3) == 0;
        const int n4 = n_raw >> 2;
        for (int q0 = tid; q0 < n4; q0 += 2 * RQ_THREADS) {
            int4 idv[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int e = (q0 + u * RQ_THREADS) << 2;
                idv[u] = make_int4(-1, -1, -1, -1);
                if (e < n_raw) {
                    const int v = e >> 6, pos = e & 63, len = (int)vlen_s[v];
                    const int32_t *src = rv.ids + vbase_s[v] + pos;
                    if (vec_ok && pos + 4 <= len) idv[u] = __ldg(reinterpret_cast<const int4 *>(src));
                    else {
                        if (pos + 0 < len) idv[u].x = __ldg(src + 0);
                        if (pos + 1 < len) idv[u].y = __ldg(src + 1);
                        if (pos + 2 < len) idv[u].z = __ldg(src + 2);
                        if (pos + 3 < len) idv[u].w = __ldg(src + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int e = (q0 + u * RQ_THREADS) << 2;
                if (e < n_raw) {
                    int32_t id4[4] = {idv[u].x, idv[u].y, idv[u].z, idv[u].w};
                    if (rv.deleted) {
                        int dead = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) if (id4[j] >= 0 && id4[j] < rv.n_deleted && rv.deleted[id4[j]]) { id4[j] = -1; dead++; }     // PIS:739
                        if (dead) atomicAdd(reinterpret_cast<unsigned int *>(inv_cnt) + (e >> 7), (unsigned)dead << (((e >> 6) & 1) * 16));
                    }
                    *reinterpret_cast<int4 *>(ids_s + e) = make_int4(id4[0], id4[1], id4[2], id4[3]);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (id4[j] >= 0) {
                            const uint32_t h1 = ((uint32_t)id4[j] * 0x9E3779B1u) >> fshift, h2 = ((uint32_t)id4[j] * 0x85EBCA6Bu) >> fshift;
                            const uint32_t s1 = (h1 & 15u) * 2u, s2 = (h2 & 15u) * 2u;
                            const uint32_t o1 = atomicOr(&F[h1 >> 4], 1u << s1);
                            if ((o1 >> s1) & 1u) atomicOr(&F[h1 >> 4], 2u << s1);
                            const uint32_t o2 = atomicOr(&F[f2_off + (h2 >> 4)], 1u << s2);
                            if ((o2 >> s2) & 1u) atomicOr(&F[f2_off + (h2 >> 4)], 2u << s2);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- 3. classify every position: single (nothing to do: its score is its visit's score) or involved (-> bitmap + worklist for
        //         the exact path, ~12 % of the positions).  Singles are counted per VISIT afterwards: valid - deleted - involved.
        int my_raw = 0, my_uniq = 0;
        bool wl_ok = true;
        for (int q4 = tid; q4 < n4; q4 += RQ_THREADS) {
            const int e = q4 << 2;
            const int4 idq = *reinterpret_cast<const int4 *>(ids_s + e);
            const int32_t id4[4] = {idq.x, idq.y, idq.z, idq.w};
            uint32_t invmask = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (id4[j] >= 0) {
                    const uint32_t h1 = ((uint32_t)id4[j] * 0x9E3779B1u) >> fshift;
                    if ((F[h1 >> 4] >> ((h1 & 15u) * 2u)) & 2u) {
                        const uint32_t h2 = ((uint32_t)id4[j] * 0x85EBCA6Bu) >> fshift;
                        if ((F[f2_off + (h2 >> 4)] >> ((h2 & 15u) * 2u)) & 2u) invmask |= 1u << j;
                    }
                }
            }
            if (invmask) {
                const int c = __popc(invmask);
                int base = atomicAdd(&s_nwl, c);
                atomicAdd(reinterpret_cast<unsigned int *>(inv_cnt) + (e >> 7), (unsigned)c << (((e >> 6) & 1) * 16));
                atomicOr(&inv_bm[e >> 5], invmask << (e & 31));
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if ((invmask >> j) & 1u) { if (base < wl_cap) wl[base] = (uint16_t)(e + j); else wl_ok = false; base++; }
                }
            }
        }
        wl_ok = __syncthreads_and(wl_ok ? 1 : 0) != 0;
        const int n_inv = s_nwl;
        if (tid < nvis) {                                                  // the singles of visit `tid` share its score
            const int singles = (int)vlen_s[tid] - (int)inv_cnt[tid];
            if (singles > 0) { atomicAdd(&s_hist[vs_s[tid]], singles); my_raw += singles; my_uniq += singles; }
        }
        // ---- 4. exact path on the involved positions: open addressing (double hashing) keyed by id, table[slot] = newest position
        //         holding that id, every position links to the previous newest one, the FIRST arrival's link is the terminator
        //         0x8000|slot (so the slot of any position is found by following its short chain).  Then every occurrence walks
        //         its id's chain once: (a) did it strictly improve on all earlier visits (PIS:747 -> lastRawVisited), (b) is it the
        //         first occurrence (= HashMap insertion order); the first one records the id's min score.
        if (n_inv > 0) {
            int tslots = 2 * n_inv <= 1024 ? 1024 : 1 << (32 - __clz(2 * n_inv - 1));
            if (tslots > tbl) tslots = tbl;
            const uint32_t smask = (uint32_t)tslots - 1u;
            const int sshift = 32 - (31 - __clz(tslots));
            for (int i = tid; i < tslots / 2; i += RQ_THREADS) table_w[i] = 0xffffffffu;
            __syncthreads();
            const int n_dom = wl_ok ? n_inv : n_raw;
            for (int i = tid; i < n_dom; i += RQ_THREADS) {
                const int e = wl_ok ? (int)wl[i] : i;
                if (!wl_ok && !((inv_bm[e >> 5] >> (e & 31)) & 1u)) continue;
                const int32_t id = ids_s[e];
                const uint32_t h = (uint32_t)id * 0x9E3779B1u;
                uint32_t slot = h >> sshift;
                const uint32_t step = ((h >> 7) | 1u) & smask;         // odd => visits every slot of the power-of-two table
                for (;;) {
                    uint32_t cur = *reinterpret_cast<volatile uint16_t *>(&table[slot]);
                    if (cur == RQ_EMPTY) {
                        next_s[e] = (uint16_t)(0x8000u | slot);
                        cur = cas16(table_w, slot, RQ_EMPTY, (uint32_t)e);
                        if (cur == RQ_EMPTY) break;                   // claimed an empty slot: first arrival
                    }
                    if (ids_s[cur] == id) {                           // same id: become the newest element of its chain
                        for (;;) {
                            next_s[e] = (uint16_t)cur;
                            const uint32_t prev = cas16(table_w, slot, cur, (uint32_t)e);
                            if (prev == cur) break;
                            cur = prev;                               // still the same id (slots never change owner)
                        }
                        break;
                    }
                    slot = (slot + step) & smask;
                }
            }
            __syncthreads();
            for (int i = tid; i < n_dom; i += RQ_THREADS) {
                const int e = wl_ok ? (int)wl[i] : i;
                if (!wl_ok && !((inv_bm[e >> 5] >> (e & 31)) & 1u)) continue;
                const uint32_t sc = vs_s[e >> 6];
                uint32_t lk = next_s[e];
                while (!(lk & 0x8000u)) lk = next_s[lk];               // terminator carries the slot
                const uint32_t head = table[lk & 0x7fffu];
                uint32_t first = head, best = 255; bool low = true;
                for (uint32_t y = head;;) {
                    const uint32_t sy = vs_s[y >> 6];
                    first = min(first, y); best = min(best, sy);
                    if (y < (uint32_t)e && sy <= sc) low = false;
                    const uint32_t ny = next_s[y];
                    if (ny & 0x8000u) break;
                    y = ny;
                }
                my_raw += low;
                const bool is_rep = first == (uint32_t)e;
                if (is_rep) { atomicAdd(&s_hist[best], 1); my_uniq++; }
                best_s[e] = (uint8_t)(is_rep ? (RQ_DUP | best) : 255u);                          // chain walks never read best_s
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { my_raw += __shfl_xor_sync(0xffffffffu, my_raw, o); my_uniq += __shfl_xor_sync(0xffffffffu, my_uniq, o); }
        if (lane == 0) { atomicAdd(&s_raw, my_raw); atomicAdd(&s_uniq, my_uniq); }
        __syncthreads();
        const int n_unique = s_uniq;
        const int want = min(p.B, n_unique);
        // Java's final table size for this many entries (HashMap.resize doubles while size > 0.75*cap); <= 65536 here
        uint32_t capF = (uint32_t)p.cap0;
        while ((double)n_unique > 0.75 * (double)capF && capF < (1u << 30)) capF <<= 1;
        const int cb = capF <= 1u ? 0 : 32 - __clz(capF - 1u);
        // ---- 5. level 0 of the radix select = the score histogram ----
        if (tid < 32) warp_pick_digit(s_hist, 256, 0, want, lane, s_pick);
        __syncthreads();
        const uint32_t sstar = (uint32_t)s_pick[0];    // cut score class
        int cum = s_pick[1];                           // entries with score < sstar (all selected)
        const int cls_n_expected = s_pick[2];
        const bool need_levels = cum + cls_n_expected != want;
        // ---- 6. one scan of the positions that can hold a score <= s*: the visits whose own score is <= s* (singles), and the
        //         worklist (first occurrences of duplicated ids carry their minimum over all visits).  Scores below s* are selected
        //         outright; the cut class is listed for the exact select on (Java bucket, first position). ----
        {
            bool low = false;
            if (tid < nvis) low = vlen_s[tid] != 0 && (uint32_t)vs_s[tid] <= sstar;
            const int slot = list_slot(&s_nlow, low, lane);
            if (low) lowvis_s[slot] = (uint16_t)tid;
        }
        __syncthreads();                                               // also: every thread has read s_pick
        {
            const int n_low = s_nlow * kBlock;
            const int n_dup = n_inv == 0 ? 0 : (wl_ok ? n_inv : n_raw);
            for (int i0 = 0; i0 < n_low + n_dup; i0 += RQ_THREADS) {
                const int i = i0 + tid;
                int e = -1; uint32_t sc = 255;
                if (i < n_low) {
                    const int v = (int)lowvis_s[i >> 6];
                    e = v * kBlock + (i & 63);
                    const bool single = ids_s[e] >= 0 && !((inv_bm[e >> 5] >> (e & 31)) & 1u);
                    sc = single ? (uint32_t)vs_s[v] : 255u;             // singles only; duplicates are taken from the worklist
                } else if (i < n_low + n_dup) {
                    e = wl_ok ? (int)wl[i - n_low] : i - n_low;
                    const uint32_t b = (wl_ok || ((inv_bm[e >> 5] >> (e & 31)) & 1u)) ? best_s[e] : 255u;
                    sc = (b >= RQ_DUP && b < 255u) ? (b & 0x7fu) : 255u;
                }
                const bool below = sc < sstar || (sc == sstar && !need_levels && sc != 255u);
                const bool incls = need_levels && sc == sstar && sc != 255u;
                const int a = list_slot(&s_m, below, lane);
                if (below && a < sort_n) sel[a] = (uint16_t)e;
                const int c = list_slot(&s_ncls, incls, lane);
                if (incls) cls[c] = (uint16_t)e;
            }
        }
        __syncthreads();
        // ---- 7. exact select inside the cut class on key2 = (Java bucket << 16) | first position, 8 bits per level ----
        if (need_levels) {
            const int bits2 = cb + 16;
            uint32_t prefix = 0; int used = 0;
            const int ncls = s_ncls;
            for (int i = tid; i < ncls; i += RQ_THREADS) { const uint32_t e = cls[i]; next_s[e] = (uint16_t)(java_hash_decimal(ids_s[e]) & 0xffffu); }
            __syncthreads();
            for (;;) {
                const int take = min(8, bits2 - used);
                const int shift = bits2 - used - take;
                if (tid < 256) s_hist[tid] = 0;
                __syncthreads();
                for (int i0 = 0; i0 < ncls; i0 += RQ_THREADS) {
                    const int i = i0 + tid;
                    bool in = false; int bin = 0;
                    if (i < ncls) {
                        const uint32_t e = cls[i];
                        const uint32_t key2 = (((uint32_t)next_s[e] & (capF - 1u)) << 16) | e;
                        in = used == 0 || (key2 >> (shift + take)) == prefix;
                        bin = (int)((key2 >> shift) & ((1u << take) - 1u));
                    }
                    hist_add(s_hist, in, bin, lane);
                }
                __syncthreads();
                if (tid < 32) warp_pick_digit(s_hist, 1 << take, cum, want, lane, s_pick);
                __syncthreads();
                prefix = (prefix << take) | (uint32_t)s_pick[0];
                used += take;
                cum = s_pick[1];
                const int bin = s_pick[2];
                __syncthreads();
                if (cum + bin == want || used >= bits2) break;
            }
            const int sel_shift = bits2 - used;
            for (int i0 = 0; i0 < ncls; i0 += RQ_THREADS) {
                const int i = i0 + tid;
                bool take_it = false; uint32_t e = 0;
                if (i < ncls) {
                    e = cls[i];
                    const uint32_t key2 = (((uint32_t)next_s[e] & (capF - 1u)) << 16) | e;
                    take_it = (key2 >> sel_shift) <= prefix;
                }
                const int a = list_slot(&s_m, take_it, lane);
                if (take_it && a < sort_n) sel[a] = (uint16_t)e;
            }
            __syncthreads();
        }
        const int m = min(s_m, sort_n);
        int32_t *out_id = p.cand_ids + (size_t)q * p.B, *out_sc = p.cand_scores + (size_t)q * p.B;
        const int sb = 32 - __clz(sstar);                               // bits that hold every selected score (<= s*)
        const int rb = sort_n <= 1 ? 0 : 32 - __clz(sort_n - 1);
        if (sort_n <= RQ_THREADS && sb + cb + rb <= 32) {
            // ---- 8a. 32-bit keys: (score | Java bucket | rank of the first position among the selected entries).  The rank comes
            //          from a bitmap over positions + prefix popcounts, so it orders exactly like the position itself; the id is
            //          parked in sid[rank] and only the key is sorted (one shuffle per compare-exchange). ----
            uint32_t *bm = table_w;                                       // [nw] (the class list is dead)
            const int nw = (n_raw + 31) >> 5;
            uint16_t *pref = reinterpret_cast<uint16_t *>(bm + nw);       // [nw]; both end below the selection list at table + n_raw
            uint32_t *key32 = reinterpret_cast<uint32_t *>(skey);
            for (int i = tid; i < nw; i += RQ_THREADS) bm[i] = 0u;
            __syncthreads();
            const int my_e = tid < m ? (int)sel[tid] : -1;
            if (my_e >= 0) atomicOr(&bm[my_e >> 5], 1u << (my_e & 31));
            __syncthreads();
            for (int w0 = 0; w0 < nw; w0 += RQ_THREADS) {                 // one pass (nw <= 1000)
                const int w = w0 + tid;
                const int c = w < nw ? __popc(bm[w]) : 0;
                int tot;
                const int ex = warp_excl_scan(c, lane, tot);
                if (lane == 31) s_wsum[tid >> 5] = ex + c;
                __syncthreads();
                if (tid < 32) { int t2; const int v = s_wsum[tid]; const int e2 = warp_excl_scan(v, lane, t2); s_wsum[tid] = e2; }
                __syncthreads();
                if (w < nw) pref[w] = (uint16_t)(s_wsum[tid >> 5] + ex);
            }
            __syncthreads();
            uint32_t a = 0xffffffffu;
            if (my_e >= 0) {
                const int32_t idv = ids_s[my_e];
                const uint32_t rank = (uint32_t)pref[my_e >> 5] + (uint32_t)__popc(bm[my_e >> 5] & ((1u << (my_e & 31)) - 1u));
                const uint32_t bucket = java_hash_decimal(idv) & (capF - 1u);
                const uint32_t sc_e = ((inv_bm[my_e >> 5] >> (my_e & 31)) & 1u) ? (uint32_t)(best_s[my_e] & 0x7fu) : (uint32_t)vs_s[my_e >> 6];
                a = (sc_e << (cb + rb)) | (bucket << rb) | rank;
                sid[rank] = idv;                                          // the worklist (which overlaps skey / sid) is dead by now
            }
            const int i = tid;
            if (sort_n == RQ_THREADS) a = bitonic_sort32<RQ_THREADS>(a, i, key32);      // fully unrolled network (B in 513..1024)
            else a = bitonic_sort32<0>(a, i, key32, sort_n);
            __syncthreads();
            if (i < want) { out_id[i] = sid[a & ((1u << rb) - 1u)]; out_sc[i] = (int32_t)(a >> (cb + rb)); }
        } else {
            // one selected entry per thread: full key (score | Java bucket | first position), Java hash evaluated densely here
            for (int i = tid; i < sort_n; i += RQ_THREADS) {
                uint64_t key = ~0ull; int32_t idv = -1;
                if (i < m) {
                    const int e = sel[i];
                    idv = ids_s[e];
                    const uint32_t bucket = java_hash_decimal(idv) & (capF - 1u);
                    const uint32_t sc_e = ((inv_bm[e >> 5] >> (e & 31)) & 1u) ? (uint32_t)(best_s[e] & 0x7fu) : (uint32_t)vs_s[e >> 6];
                    key = ((uint64_t)sc_e << (cb + 16)) | ((uint64_t)bucket << 16) | (uint64_t)e;
                }
                skey[i] = key; sid[i] = idv;                                // the worklist (which overlaps skey / sid) is dead by now
            }
            __syncthreads();
            // ---- 8. bitonic sort: element i lives in thread i (registers), strides < 32 are exchanged with warp shuffles,
            //         larger strides through shared memory ----
            if (sort_n <= RQ_THREADS) {
                const int i = tid;
                uint64_t a = i < sort_n ? skey[i] : ~0ull;
                int32_t av = i < sort_n ? sid[i] : -1;
                for (int k = 2; k <= sort_n; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const bool asc = (i & k) == 0;
                        uint64_t b; int32_t bv;
                        if (j >= 32) {
                            __syncthreads();
                            if (i < sort_n) { skey[i] = a; sid[i] = av; }
                            __syncthreads();
                            b = i < sort_n ? skey[i ^ j] : ~0ull; bv = i < sort_n ? sid[i ^ j] : -1;
                        } else {
                            b = __shfl_xor_sync(0xffffffffu, a, j); bv = __shfl_xor_sync(0xffffffffu, av, j);
                        }
                        const bool lower = (i & j) == 0;                // this thread keeps the smaller of the pair when ascending
                        const bool take_b = (lower == asc) ? (b < a) : (b > a);
                        if (take_b) { a = b; av = bv; }
                    }
                }
                __syncthreads();
                if (i < sort_n) { skey[i] = a; sid[i] = av; }
                __syncthreads();
            } else {
                for (int k = 2; k <= sort_n; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = tid; i < sort_n; i += RQ_THREADS) {
                            const int ixj = i ^ j;
                            if (ixj > i) {
                                const uint64_t a = skey[i], b = skey[ixj];
                                const bool asc = (i & k) == 0;
                                if ((a > b) == asc) { skey[i] = b; skey[ixj] = a; const int32_t t = sid[i]; sid[i] = sid[ixj]; sid[ixj] = t; }
                            }
                        }
                        __syncthreads();
                    }
                }
            }
            for (int i = tid; i < want; i += RQ_THREADS) { out_id[i] = sid[i]; out_sc[i] = (int32_t)(skey[i] >> (cb + 16)); }
        }
        if (tid == 0) { p.n_cand[q] = want; p.unique[q] = n_unique; p.raw_seen[q] = s_raw; }
    }
}

// =====================================================================================================================
// Fast path, second generation (round 2): TWO 512-thread CTAs per SM instead of one 1024-thread CTA, so one query's barriers and
// shared-memory latencies are covered by the other query's work.  The per-query state is cut from ~225 KB to <= 113 KB:
//   * the gathered ids are NOT kept in shared memory: the classify pass and the selection scan re-read them from L2 (the partition
//     rows were just read by the filter pass; 82 KB per query);
//   * everything the exact path needs (id, position, chain link, minimum score) is indexed by WORKLIST slot, not by position, and the
//     worklist has a fixed capacity (~22 % of the positions at C2; a query that exceeds it, a cut class larger than its list or a sort
//     key wider than 32 bits is appended to an overflow list and served by route_fast_kernel in a second launch);
//   * the bitonic sort holds two keys per thread and double-buffers its shared-memory exchanges (one barrier per stage).
// Same results as route_fast_kernel bit for bit (same filters, same exact path, same selection keys).
// =====================================================================================================================
constexpr int R2_THREADS = 512;           // 16 warps per CTA, two CTAs per SM (768 threads = 48 warps per SM at 40 registers was measured slower: 2.00 vs 1.80 ms)
constexpr int R2_U = 5;                   // 16-byte id slices a thread keeps in flight in the gather passes
constexpr int R2_H = 512;                 // the sort works on 2 x R2_H keys held by the first R2_H threads
constexpr int RS_SEGMAX = 16384;          // keys one route_sort_big_kernel CTA orders (16 per thread)
constexpr int RS_MAXSEG = 4;              // segments per query (n_raw <= 32000 needs at most 4 greedy segments of <= RS_SEGMAX)
constexpr int R2_SELBYTES = 7168;          // sel_id int32[1024] + sel_pos uint16[1024] + sel_sc uint8[1024]

struct Route2Layout { int region, cls_cap, wl_cap; size_t smem; };

static bool route2_layout(int TD, int probes, int n_raw, int tbl, Route2Layout &L) {
    const int nvis = TD * probes, nvis16 = (nvis + 15) / 16 * 16;
    const int nbm = ((n_raw + 31) / 32 + 3) / 4 * 4;
    L.region = std::max(2 * tbl, 28672);
    L.cls_cap = std::min(n_raw, ((L.region - R2_SELBYTES) / 4) & ~7);
    const size_t fixed = (size_t)L.region + sizeof(int64_t) * nvis + 2 * (size_t)nvis16 + 3 * sizeof(uint16_t) * (size_t)nvis16 + sizeof(uint32_t) * (size_t)nbm + 64;
    const size_t limit = 114176;                                       // 2 x (dynamic + 1.4 KB static + 1 KB reserve) <= 228 KB per SM
    if (fixed + 9 * 1024 > limit) return false;
    int64_t cap = (int64_t)(limit - fixed) / 9;
    cap = std::min<int64_t>(cap, n_raw);
    int slots_max = 1; while (slots_max * 2 <= L.region / 2) slots_max <<= 1;   // exact table: uint16 slots inside the region, a power of two
    cap = std::min<int64_t>(cap, slots_max / 4 * 3);                   // load factor <= 0.75; slots < 32768 and worklist slots < 32768 fit the 15-bit links
    cap &= ~(int64_t)15;
    if (cap < 16) return false;
    L.wl_cap = (int)cap;
    L.smem = fixed + 9 * (size_t)cap;
    return true;
}

// warp-aggregated slot allocation over a few counters: lanes with in == true and the same `which` share one shared atomic
__device__ __forceinline__ int multi_slot(int32_t *counters, bool in, int which, int lane) {
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (!in) return -1;
    const unsigned peers = __match_any_sync(act, which);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[which], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

// Bitonic sort of TWO 32-bit keys per thread (elements tid and tid + R2_H of threads tid < R2_H; the other threads only keep the barriers
// company); strides < 32 by shuffles, larger ones through the double buffer `buf` (2 x 2*R2_H words; one barrier per stage), stride R2_H
// inside the thread.  N > 0: unrolled network.
template <int N>
__device__ __forceinline__ void bitonic_sort32x2(uint32_t &a0, uint32_t &a1, int tid, uint32_t *buf, int n_rt = 0) {
    const int n = N > 0 ? N : n_rt;
    const bool on = tid < R2_H;
    int phase = 0;
#pragma unroll
    for (int k = 2; k <= n; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == R2_H) {                                         // k == 2 * R2_H: ascending everywhere
                if (a1 < a0) { const uint32_t t = a0; a0 = a1; a1 = t; }
                continue;
            }
            uint32_t b0, b1;
            if (j >= 32) {
                uint32_t *bb = buf + phase * (2 * R2_H);
                if (on) { bb[tid] = a0; bb[tid + R2_H] = a1; }
                __syncthreads();
                b0 = on ? bb[tid ^ j] : a0; b1 = on ? bb[(tid ^ j) + R2_H] : a1;
                phase ^= 1;
            } else {
                b0 = __shfl_xor_sync(0xffffffffu, a0, j); b1 = __shfl_xor_sync(0xffffffffu, a1, j);
            }
            const bool lower = (tid & j) == 0;
            const bool asc0 = (tid & k) == 0, asc1 = ((tid + R2_H) & k) == 0;
            a0 = (lower == asc0) ? min(a0, b0) : max(a0, b0);
            a1 = (lower == asc1) ? min(a1, b1) : max(a1, b1);
        }
    }
}

template <bool HAS_DEL, bool BIG>
__global__ void __launch_bounds__(R2_THREADS, 2) route_fast2_kernel(RoutingView rv, RouteParams p, RouteFastExtra x) {
    extern __shared__ __align__(16) unsigned char rq_smem[];
    const int TD = rv.TD, probes = p.probes, n_raw = x.n_raw, tbl = x.tbl, sort_n = x.sort_n, nvis = TD * probes;
    const int nvis16 = (nvis + 15) / 16 * 16;
    const int region = x.v2_region, cls_cap = x.v2_cls_cap, wl_cap = x.v2_wl_cap;
    uint32_t *F = reinterpret_cast<uint32_t *>(rq_smem);                       // phase 1-3: duplicate filters (2 x 4*tbl two-bit buckets)
    uint32_t *table_w = F;                                                     // phase 4: exact table of worklist slots
    uint16_t *table = reinterpret_cast<uint16_t *>(rq_smem);
    uint16_t *cls_pos = table, *cls_hash = table + cls_cap;                    // phase 6-7: the cut class (position, low 16 bits of the Java hash)
    unsigned char *selb = rq_smem + region - R2_SELBYTES;                      // phase 6-8: the selected entries
    int32_t *sel_id = reinterpret_cast<int32_t *>(selb);
    uint16_t *sel_pos = reinterpret_cast<uint16_t *>(selb + 4096);
    uint8_t *sel_sc = selb + 6144;
    const int nw = (n_raw + 31) >> 5, nwp = (nw + 3) & ~3;
    uint32_t *bm = reinterpret_cast<uint32_t *>(rq_smem);                       // phase 8 (the class list is dead): position bitmap, prefix, sort buffers
    uint16_t *pref = reinterpret_cast<uint16_t *>(bm + nwp);
    uint32_t *key32 = reinterpret_cast<uint32_t *>(rq_smem + (((size_t)6 * nwp + 15) & ~(size_t)15));
    int32_t *sid = reinterpret_cast<int32_t *>(key32 + 4 * R2_H);
    int64_t *vbase_s = reinterpret_cast<int64_t *>(rq_smem + region);
    uint8_t *vs_s = reinterpret_cast<uint8_t *>(vbase_s + nvis);
    uint8_t *vlen_s = vs_s + nvis16;
    uint16_t *lowvis_s = reinterpret_cast<uint16_t *>(vlen_s + nvis16);
    uint16_t *inv_cnt = lowvis_s + nvis16;
    uint16_t *rep_cnt = inv_cnt + nvis16;                                      // per visit: first occurrences of duplicated ids (HARD_CAP accounting)
    uint32_t *inv_bm = reinterpret_cast<uint32_t *>(rep_cnt + nvis16);
    const int nbm = ((n_raw + 31) / 32 + 3) / 4 * 4;
    int32_t *wl_id = reinterpret_cast<int32_t *>(inv_bm + nbm);
    uint16_t *wl_pos = reinterpret_cast<uint16_t *>(wl_id + wl_cap);
    uint16_t *wl_next = wl_pos + wl_cap;
    uint8_t *wl_best = reinterpret_cast<uint8_t *>(wl_next + wl_cap);
    __shared__ int32_t s_hist[256];
    __shared__ int32_t s_raw, s_uniq, s_m, s_ncls, s_nwl, s_nlow, s_pick[3], s_wsum[32], s_cut;
    __shared__ int32_t s_seg_start[RS_MAXSEG + 1], s_seg_cnt[RS_MAXSEG], s_seg_ok;
    __shared__ uint8_t s_seg_of[128];

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t n_ids = rv.n_ids;
    const bool vec_ok = (n_ids & 3) == 0;
    const int n4 = n_raw >> 2;
    const uint8_t *__restrict__ deleted = rv.deleted;

    // ids of the 4 consecutive positions starting at e (-1: beyond the partition row, -2: a deleted id, PIS:739)
    auto gather4 = [&](int e, int32_t (&id4)[4]) {
        const int v = e >> 6, pos = e & 63, len = (int)vlen_s[v];
        const int32_t *src = rv.ids + vbase_s[v] + pos;
        if (vec_ok && pos + 4 <= len) {
            const int4 t = __ldg(reinterpret_cast<const int4 *>(src));
            id4[0] = t.x; id4[1] = t.y; id4[2] = t.z; id4[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) id4[j] = pos + j < len ? __ldg(src + j) : -1;
        }
        if (HAS_DEL) {
#pragma unroll
            for (int j = 0; j < 4; j++) if (id4[j] >= 0 && id4[j] < rv.n_deleted && deleted[id4[j]]) id4[j] = -2;
        }
    };
    // Duplicate filter: tbl/2 32-bit words, each holding EIGHT 2-bit buckets of filter 1 (low half: "taken" bits 0-7, "taken twice" bits 8-15)
    // and eight of filter 2 (high half, same split).  An id
    // picks the word and its filter-1 bucket from one multiplicative hash and its filter-2 bucket from another, so ONE shared atomic sets
    // both "taken" bits and returns both old states, one predicated reduction sets the "taken twice" bits, and the classify pass tests
    // both filters with one load.  (Two separate arrays -- route_fast_kernel -- need twice the atomics and loads for a slightly lower
    // false-positive rate: 2.1 % vs 3.5 % of the positions at C2.)
    const uint32_t F_sa = (uint32_t)__cvta_generic_to_shared(F);
    const int wshift = 32 - (31 - __clz(tbl >> 1));                             // tbl/2 words (a power of two)
    auto filter_bits = [&](int32_t id, uint32_t &word) -> uint32_t {
        const uint32_t h1 = (uint32_t)id * 0x9E3779B1u, h2 = (uint32_t)id * 0x85EBCA6Bu;
        word = h1 >> wshift;
        return (1u << ((h1 >> (wshift - 3)) & 7u)) | (0x10000u << (h2 >> 29));
    };
    auto filter_set = [&](int32_t id) {
        uint32_t word;
        const uint32_t bits = filter_bits(id, word), addr = F_sa + word * 4u;
        uint32_t old;
        asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(bits) : "memory");
        // (predicated: the C++ `if` around an atomic compiles to a divergent branch with convergence barriers)
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p red.shared.or.b32 [%0], %1; }" ::"r"(addr), "r"((old & bits) << 8) : "memory");
    };
    auto filter_twice = [&](int32_t id) -> bool {                        // taken twice in BOTH filters: not provably a single occurrence
        uint32_t word;
        const uint32_t bits = filter_bits(id, word);
        return ((F[word] >> 8) & bits) == bits;
    };
    __shared__ int32_t s_q;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(x.ovf_n + 1, 1);                      // dynamic query assignment: the cost per query varies (worklist size, cut class)
        __syncthreads();
        const int64_t q = s_q;
        if (q >= p.Q) break;
        // ---- 1. stage visits, reset the filters ----
        for (int v = tid; v < nvis; v += R2_THREADS) {
            const int g = v / probes, j = v - g * probes;
            const int64_t vi = (q * TD + g) * probes + j;
            const bool valid = j < (int)x.vis_n[q * TD + g];
            const int64_t row = valid ? (int64_t)x.vis_part[vi] * kBlock : 0;
            vbase_s[v] = (int64_t)g * n_ids + row;
            vlen_s[v] = valid ? (uint8_t)min((int64_t)kBlock, n_ids - row) : 0;
            vs_s[v] = valid ? x.vis_score[vi] : 255;
        }
        for (int i = tid; i < tbl / 8; i += R2_THREADS) reinterpret_cast<uint4 *>(F)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < nbm; i += R2_THREADS) inv_bm[i] = 0u;
        for (int i = tid; i < nvis16; i += R2_THREADS) reinterpret_cast<uint32_t *>(inv_cnt)[i] = 0u;       // inv_cnt and rep_cnt
        if (tid < 256) s_hist[tid] = 0;
        if (tid == 0) { s_raw = 0; s_uniq = 0; s_m = 0; s_ncls = 0; s_nwl = 0; s_nlow = 0; }
        __syncthreads();
        // ---- 2. gather ids (16-byte slices of the 256-byte partition rows, R2_U in flight) and set the two 2-bit filters ----
        for (int q0 = tid; q0 < n4; q0 += R2_U * R2_THREADS) {
            int32_t idv[R2_U][4];
#pragma unroll
            for (int u = 0; u < R2_U; u++) {
                const int e = (q0 + u * R2_THREADS) << 2;
                if (e < n_raw) gather4(e, idv[u]);
                else { idv[u][0] = idv[u][1] = idv[u][2] = idv[u][3] = -1; }
            }
#pragma unroll
            for (int u = 0; u < R2_U; u++) {
                if ((idv[u][0] | idv[u][1] | idv[u][2] | idv[u][3]) >= 0) {      // the common case: four live ids, no per-id test
#pragma unroll
                    for (int j = 0; j < 4; j++) filter_set(idv[u][j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (idv[u][j] >= 0) filter_set(idv[u][j]);
                }
            }
        }
        __syncthreads();
        // ---- 3. classify (ids re-read from L2): single, or involved -> bitmap + per-visit counter + worklist (slot: id, position) ----
        int my_raw = 0, my_uniq = 0;
        bool fits = true;
        for (int q0 = tid; q0 < n4; q0 += R2_U * R2_THREADS) {
            int32_t idv[R2_U][4];
#pragma unroll
            for (int u = 0; u < R2_U; u++) {
                const int e = (q0 + u * R2_THREADS) << 2;
                if (e < n_raw) gather4(e, idv[u]);
                else { idv[u][0] = idv[u][1] = idv[u][2] = idv[u][3] = -1; }
            }
#pragma unroll
            for (int u = 0; u < R2_U; u++) {
                const int e = (q0 + u * R2_THREADS) << 2;
                uint32_t invmask = 0; int dead = 0;
                if ((idv[u][0] | idv[u][1] | idv[u][2] | idv[u][3]) >= 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++) invmask |= filter_twice(idv[u][j]) ? 1u << j : 0u;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (idv[u][j] >= 0) invmask |= filter_twice(idv[u][j]) ? 1u << j : 0u;
                        else if (HAS_DEL && idv[u][j] == -2) dead++;
                    }
                }
                // (warp-collective bookkeeping -- REDUX + shuffles + ballots instead of these three shared atomics -- was measured slower: 1.95 vs 1.86 ms)
                if (invmask | (uint32_t)dead) {
                    const int c = __popc(invmask);
                    atomicAdd(reinterpret_cast<unsigned int *>(inv_cnt) + (e >> 7), (unsigned)(c + dead) << (((e >> 6) & 1) * 16));
                    if (c) {
                        int base = atomicAdd(&s_nwl, c);
                        atomicOr(&inv_bm[e >> 5], invmask << (e & 31));
                        if (base + c <= wl_cap) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if ((invmask >> j) & 1u) { wl_id[base] = idv[u][j]; wl_pos[base] = (uint16_t)(e + j); base++; }
                            }
                        } else fits = false;
                    }
                }
            }
        }
        fits = __syncthreads_and(fits ? 1 : 0) != 0;
        const int n_inv = s_nwl;
        // ---- 4. exact path over the worklist (see route_fast_kernel; links and table entries are worklist slots) ----
        if (fits && n_inv > 0) {
            int tslots = 2 * n_inv <= 1024 ? 1024 : 1 << (32 - __clz(2 * n_inv - 1));
            if (tslots * 2 > region) tslots = region / 2;                  // region is a power of two or 28672 (-> 8192 slots would not fit: clamp to 2^k below)
            tslots = 1 << (31 - __clz(tslots));
            const uint32_t smask = (uint32_t)tslots - 1u;
            const int sshift = 32 - (31 - __clz(tslots));
            for (int i = tid; i < tslots / 2; i += R2_THREADS) table_w[i] = 0xffffffffu;
            __syncthreads();
            for (int i = tid; i < n_inv; i += R2_THREADS) {
                const int32_t id = wl_id[i];
                const uint32_t h = (uint32_t)id * 0x9E3779B1u;
                uint32_t slot = h >> sshift;
                const uint32_t step = ((h >> 7) | 1u) & smask;
                for (;;) {
                    uint32_t cur = *reinterpret_cast<volatile uint16_t *>(&table[slot]);
                    if (cur == RQ_EMPTY) {
                        wl_next[i] = (uint16_t)(0x8000u | slot);
                        cur = cas16(table_w, slot, RQ_EMPTY, (uint32_t)i);
                        if (cur == RQ_EMPTY) break;
                    }
                    if (wl_id[cur] == id) {
                        for (;;) {
                            wl_next[i] = (uint16_t)cur;
                            const uint32_t prev = cas16(table_w, slot, cur, (uint32_t)i);
                            if (prev == cur) break;
                            cur = prev;
                        }
                        break;
                    }
                    slot = (slot + step) & smask;
                }
            }
            __syncthreads();
        }
        // ---- 4b. per-id statistics (+ HARD_CAP).  Pass 0 looks at every visit; it also counts, per visit, the ids that occur there FIRST
        //      (singles + first occurrences of duplicated ids) = by how much that poll grows bestScore.  The reference polls a partition only
        //      while bestScore.size() < HARD_CAP (PIS:657-659; the same test guards every table and division, PIS:624-628) and then adds it
        //      whole, so the visits from the first one that finds the map full are never made: when there is such a visit, pass 1 repeats
        //      the statistics over the visits before it (first occurrences and improvement counts do not depend on later visits; the
        //      minimum score of an id does). ----
        int cut_v = nvis;
        for (int pass = 0; fits && pass < 2; pass++) {
            const uint32_t cutoff = (uint32_t)cut_v * kBlock;
            my_raw = 0; my_uniq = 0;
            if (tid < cut_v) {                                             // the singles of visit `tid` share its score
                const int singles = (int)vlen_s[tid] - (int)inv_cnt[tid];
                if (singles > 0) { atomicAdd(&s_hist[vs_s[tid]], singles); my_raw += singles; my_uniq += singles; }
            }
            for (int i = tid; i < n_inv; i += R2_THREADS) {
                const uint32_t e = wl_pos[i];
                if (e >= cutoff) { wl_best[i] = 255; continue; }
                const uint32_t sc = vs_s[e >> 6];
                uint32_t lk = wl_next[i];
                while (!(lk & 0x8000u)) lk = wl_next[lk];
                const uint32_t head = table[lk & 0x7fffu];
                uint32_t first = 0xffffu, best = 255; bool low = true;
                for (uint32_t y = head;;) {
                    const uint32_t ey = wl_pos[y], sy = vs_s[ey >> 6];
                    first = min(first, ey);
                    if (ey < cutoff) best = min(best, sy);
                    if (ey < e && sy <= sc) low = false;
                    const uint32_t ny = wl_next[y];
                    if (ny & 0x8000u) break;
                    y = ny;
                }
                my_raw += low;
                const bool is_rep = first == e;
                if (is_rep) {
                    atomicAdd(&s_hist[best], 1); my_uniq++;
                    if (pass == 0) atomicAdd(reinterpret_cast<unsigned int *>(rep_cnt) + (e >> 7), 1u << (((e >> 6) & 1) * 16));
                }
                wl_best[i] = (uint8_t)(is_rep ? (RQ_DUP | best) : 255u);
            }
            if (pass == 1 || p.hard_cap > (int64_t)n_raw - kBlock) break;      // the re-count is done / the cap cannot bind (block-uniform)
            __syncthreads();
            const int nu = tid < nvis ? (int)vlen_s[tid] - (int)inv_cnt[tid] + (int)rep_cnt[tid] : 0;
            int tot;
            const int ex = warp_excl_scan(nu, lane, tot);
            if (lane == 31) s_wsum[tid >> 5] = ex + nu;
            if (tid == 0) s_cut = nvis;
            __syncthreads();
            int before = ex;                                               // bestScore.size() when visit `tid` is about to be polled
            for (int w = 0; w < (tid >> 5); w++) before += s_wsum[w];
            if (tid < nvis && (int64_t)before >= p.hard_cap) atomicMin(&s_cut, tid);
            __syncthreads();
            cut_v = s_cut;
            if (cut_v >= nvis) break;
            if (tid < 256) s_hist[tid] = 0;
            if (tid >= cut_v && tid < nvis) vlen_s[tid] = 0;                // the selection scan below skips the visits that were never made
            __syncthreads();
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { my_raw += __shfl_xor_sync(0xffffffffu, my_raw, o); my_uniq += __shfl_xor_sync(0xffffffffu, my_uniq, o); }
        if (lane == 0) { atomicAdd(&s_raw, my_raw); atomicAdd(&s_uniq, my_uniq); }
        __syncthreads();
        const int n_unique = s_uniq;
        const int want = min(p.B, n_unique);
        uint32_t capF = (uint32_t)p.cap0;
        while ((double)n_unique > 0.75 * (double)capF && capF < (1u << 30)) capF <<= 1;
        const int cb = capF <= 1u ? 0 : 32 - __clz(capF - 1u);
        // ---- 5. score histogram -> cut class ----
        if (tid < 32) warp_pick_digit(s_hist, 256, 0, want, lane, s_pick);
        __syncthreads();
        const uint32_t sstar = (uint32_t)s_pick[0];
        int cum = s_pick[1];
        const int cls_n_expected = s_pick[2];
        const bool need_levels = cum + cls_n_expected != want;
        const int sb = 32 - __clz(sstar);
        const int rb = sort_n <= 1 ? 0 : 32 - __clz(sort_n - 1);
        // what this kernel does not hold: more involved positions than worklist slots, a cut class beyond its list, keys wider than 32 bits
        if (!fits || (need_levels && cls_n_expected > cls_cap) || (!BIG && sb + cb + rb > 32) || capF > 65536u) {
            if (tid == 0) { x.ovf_list[atomicAdd(x.ovf_n, 1)] = (int32_t)q; p.n_cand[q] = 0; }
            continue;
        }
        // BIG (refinementLimit > 1024): the selected entries leave as 64-bit keys (score | Java bucket | position) and route_sort_big_kernel orders
        // them.  The score is the leading key field, so the list is cut at score-class boundaries into <= RS_MAXSEG segments of <= RS_SEGMAX keys
        // (greedy over the class histogram); every segment is sorted on its own and they concatenate in order.
        unsigned long long *big = BIG ? x.big_keys + (size_t)q * p.B : nullptr;
        if (BIG) {
            if (tid == 0) {
                int seg = 0, fill = 0, ok = 1;
                s_seg_start[0] = 0;
                for (uint32_t c = 0; c <= sstar && c < 128u; c++) {
                    const int nc = c < sstar ? s_hist[c] : want - cum;
                    if (nc > RS_SEGMAX) ok = 0;
                    if (fill + nc > RS_SEGMAX && fill > 0) {
                        if (seg + 1 >= RS_MAXSEG) { ok = 0; } else { seg++; s_seg_start[seg] = s_seg_start[seg - 1] + fill; fill = 0; }
                    }
                    s_seg_of[c] = (uint8_t)seg;
                    fill += nc;
                }
                for (int g2 = seg + 1; g2 <= RS_MAXSEG; g2++) s_seg_start[g2] = s_seg_start[seg] + fill;
                for (int g2 = 0; g2 < RS_MAXSEG; g2++) s_seg_cnt[g2] = 0;
                s_seg_ok = ok && sstar < 128u;
            }
            __syncthreads();
            if (!s_seg_ok) {                                            // a single score class beyond one segment: the general kernel takes it
                if (tid == 0) { x.ovf_list[atomicAdd(x.ovf_n, 1)] = (int32_t)q; p.n_cand[q] = 0; }
                continue;
            }
        }
        // ---- 6. scan the positions that can hold a score <= s*: the visits with score <= s* (singles; ids re-read) and the worklist ----
        {
            bool low = false;
            if (tid < nvis) low = vlen_s[tid] != 0 && (uint32_t)vs_s[tid] <= sstar;
            const int slot = list_slot(&s_nlow, low, lane);
            if (low) lowvis_s[slot] = (uint16_t)tid;
        }
        __syncthreads();
        {
            const int n_low = s_nlow * kBlock;
            for (int i0 = 0; i0 < n_low + n_inv; i0 += R2_THREADS) {
                const int i = i0 + tid;
                int e = 0; uint32_t sc = 255; int32_t id = -1;
                if (i < n_low) {
                    const int v = (int)lowvis_s[i >> 6], pos = i & 63;
                    e = v * kBlock + pos;
                    if (pos < (int)vlen_s[v] && !((inv_bm[e >> 5] >> (e & 31)) & 1u)) {
                        id = __ldg(rv.ids + vbase_s[v] + pos);
                        if (!(HAS_DEL && id >= 0 && id < rv.n_deleted && deleted[id])) sc = (uint32_t)vs_s[v];
                    }
                } else if (i < n_low + n_inv) {
                    const int w = i - n_low;
                    const uint32_t b = wl_best[w];
                    if (b >= RQ_DUP && b < 255u) { sc = b & 0x7fu; e = (int)wl_pos[w]; id = wl_id[w]; }
                }
                const bool below = sc < sstar || (sc == sstar && !need_levels && sc != 255u);
                const bool incls = need_levels && sc == sstar && sc != 255u;
                const int a = list_slot(&s_m, below, lane);
                if (BIG) {
                    const int sg = below ? (int)s_seg_of[sc] : 0;
                    const int off = multi_slot(s_seg_cnt, below, sg, lane);
                    if (below) {
                        const int slot = s_seg_start[sg] + off;
                        if (slot < p.B) big[slot] = ((unsigned long long)sc << 40) | ((unsigned long long)(java_hash_decimal(id) & (capF - 1u)) << 16) | (unsigned long long)e;
                    }
                } else if (below && a < sort_n) { sel_id[a] = id; sel_pos[a] = (uint16_t)e; sel_sc[a] = (uint8_t)sc; }
                const int c = list_slot(&s_ncls, incls, lane);
                if (incls) { cls_pos[c] = (uint16_t)e; cls_hash[c] = (uint16_t)(java_hash_decimal(id) & 0xffffu); }
            }
        }
        __syncthreads();
        // ---- 7. exact select inside the cut class on key2 = (Java bucket << 16) | first position, 8 bits per level ----
        if (need_levels) {
            const int bits2 = cb + 16;
            uint32_t prefix = 0; int used = 0;
            const int ncls = s_ncls;
            for (;;) {
                const int take = min(8, bits2 - used);
                const int shift = bits2 - used - take;
                if (tid < 256) s_hist[tid] = 0;
                __syncthreads();
                for (int i0 = 0; i0 < ncls; i0 += R2_THREADS) {
                    const int i = i0 + tid;
                    bool in = false; int bin = 0;
                    if (i < ncls) {
                        const uint32_t key2 = (((uint32_t)cls_hash[i] & (capF - 1u)) << 16) | (uint32_t)cls_pos[i];
                        in = used == 0 || (key2 >> (shift + take)) == prefix;
                        bin = (int)((key2 >> shift) & ((1u << take) - 1u));
                    }
                    hist_add(s_hist, in, bin, lane);
                }
                __syncthreads();
                if (tid < 32) warp_pick_digit(s_hist, 1 << take, cum, want, lane, s_pick);
                __syncthreads();
                prefix = (prefix << take) | (uint32_t)s_pick[0];
                used += take;
                cum = s_pick[1];
                const int bin = s_pick[2];
                __syncthreads();
                if (cum + bin == want || used >= bits2) break;
            }
            const int sel_shift = bits2 - used;
            for (int i0 = 0; i0 < ncls; i0 += R2_THREADS) {
                const int i = i0 + tid;
                bool take_it = false; uint32_t e = 0;
                if (i < ncls) {
                    e = cls_pos[i];
                    const uint32_t key2 = (((uint32_t)cls_hash[i] & (capF - 1u)) << 16) | e;
                    take_it = (key2 >> sel_shift) <= prefix;
                }
                const int a = list_slot(&s_m, take_it, lane);
                if (BIG) {
                    const int sg = s_seg_of[sstar];
                    const int off = multi_slot(s_seg_cnt, take_it, sg, lane);
                    if (take_it) {
                        const int slot = s_seg_start[sg] + off;
                        if (slot < p.B) big[slot] = ((unsigned long long)sstar << 40) | ((unsigned long long)((uint32_t)cls_hash[i] & (capF - 1u)) << 16) | (unsigned long long)e;
                    }
                } else if (take_it && a < sort_n) {
                    sel_id[a] = __ldg(rv.ids + vbase_s[e >> 6] + (e & 63));    // the id at a position (single or first occurrence alike)
                    sel_pos[a] = (uint16_t)e; sel_sc[a] = (uint8_t)sstar;
                }
            }
            __syncthreads();
        }
        if (BIG) {
            if (tid <= RS_MAXSEG) x.big_seg[q * (RS_MAXSEG + 1) + tid] = s_seg_start[tid];
            if (tid == 0) { p.n_cand[q] = want; p.unique[q] = n_unique; p.raw_seen[q] = s_raw; }
            continue;
        }
        // ---- 8. 32-bit keys (score | Java bucket | rank of the position among the selected), two per thread, key-only bitonic sort ----
        const int m = min(s_m, sort_n);
        int32_t *out_id = p.cand_ids + (size_t)q * p.B, *out_sc = p.cand_scores + (size_t)q * p.B;
        int my_e[2]; int32_t my_id[2]; uint32_t my_sc[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = tid < R2_H ? tid + u * R2_H : m;
            my_e[u] = i < m ? (int)sel_pos[i] : -1;
            my_id[u] = i < m ? sel_id[i] : -1;
            my_sc[u] = i < m ? (uint32_t)sel_sc[i] : 0u;
        }
        __syncthreads();                                                   // the class list is dead; sel_* are in registers
        for (int i = tid; i < nwp; i += R2_THREADS) bm[i] = 0u;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; u++) if (my_e[u] >= 0) atomicOr(&bm[my_e[u] >> 5], 1u << (my_e[u] & 31));
        __syncthreads();
        for (int w0 = 0; w0 < nwp; w0 += 2 * R2_THREADS) {                 // one pass (nw <= 1000)
            const int w = w0 + 2 * tid;
            const int c0 = w < nwp ? __popc(bm[w]) : 0, c1 = w + 1 < nwp ? __popc(bm[w + 1]) : 0;
            int tot;
            const int ex = warp_excl_scan(c0 + c1, lane, tot);
            if (lane == 31) s_wsum[tid >> 5] = ex + c0 + c1;
            __syncthreads();
            if (tid < 32) { int t2; const int v = tid < R2_THREADS / 32 ? s_wsum[tid] : 0; const int e2 = warp_excl_scan(v, lane, t2); s_wsum[tid] = e2; }
            __syncthreads();
            if (w < nwp) { pref[w] = (uint16_t)(s_wsum[tid >> 5] + ex); pref[w + 1] = (uint16_t)(s_wsum[tid >> 5] + ex + c0); }
        }
        __syncthreads();
        uint32_t a[2] = {0xffffffffu, 0xffffffffu};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (my_e[u] >= 0) {
                const uint32_t rank = (uint32_t)pref[my_e[u] >> 5] + (uint32_t)__popc(bm[my_e[u] >> 5] & ((1u << (my_e[u] & 31)) - 1u));
                const uint32_t bucket = java_hash_decimal(my_id[u]) & (capF - 1u);
                a[u] = (my_sc[u] << (cb + rb)) | (bucket << rb) | rank;
                sid[rank] = my_id[u];
            }
        }
        if (sort_n == 2 * R2_H) bitonic_sort32x2<2 * R2_H>(a[0], a[1], tid, key32);
        else bitonic_sort32x2<0>(a[0], a[1], tid, key32, sort_n);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = tid < R2_H ? tid + u * R2_H : want;
            if (i < want) { out_id[i] = sid[a[u] & ((1u << rb) - 1u)]; out_sc[i] = (int32_t)(a[u] >> (cb + rb)); }
        }
        if (tid == 0) { p.n_cand[q] = want; p.unique[q] = n_unique; p.raw_seen[q] = s_raw; }
    }
}

// Orders the candidates route_fast2_kernel<.., BIG> selected: one CTA per segment, E 64-bit keys (score | Java bucket | position) per thread,
// thread tid OWNS the contiguous elements tid*E .. tid*E+E-1.  Bitonic network: strides below E pair two registers of one thread (50 of
// the 105 stages at E = 16), strides E .. 16E are warp shuffles (thread distance 1 .. 16), only strides >= 32E go through shared memory
// (15 stages; stored transposed, [u][tid], so consecutive lanes touch consecutive words).  Then the id of every position is read back
// from the partition rows.
constexpr int RS_THREADS = 1024;

template <int E>
__global__ void __launch_bounds__(RS_THREADS, 1) route_sort_big_kernel(RoutingView rv, RouteParams p, RouteFastExtra x) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(rs_smem);
    const int tid = threadIdx.x, TD = rv.TD, probes = p.probes;
    constexpr int n = E * RS_THREADS;
    for (int64_t w = blockIdx.x; w < p.Q * RS_MAXSEG; w += gridDim.x) {
        const int sg = (int)(w / p.Q);                                   // segment-major: every CTA gets its share of the (usually only non-empty) first segments
        const int64_t q = w - (int64_t)sg * p.Q;
        if (p.n_cand[q] <= 0) continue;                                  // empty, or handed to the fallback kernel (block-uniform)
        const int s0 = x.big_seg[q * (RS_MAXSEG + 1) + sg], m = x.big_seg[q * (RS_MAXSEG + 1) + sg + 1] - s0;
        if (m <= 0) continue;
        const unsigned long long *src = x.big_keys + (size_t)q * p.B + s0;
        unsigned long long a[E];
#pragma unroll
        for (int u = 0; u < E; u++) { const int i = tid * E + u; a[u] = i < m ? src[i] : ~0ull; }
#pragma unroll 1
        for (int k = 2; k <= n; k <<= 1) {
            const bool asc_t = ((tid * E) & k) == 0;                     // direction of this thread's block once k >= E
#pragma unroll 1
            for (int j = k >> 1; j >= E; j >>= 1) {                      // partner in another thread, td threads away
                const int td = j / E;
                const bool want_min = ((tid & td) == 0) == asc_t;        // this thread keeps the smaller key of the pair
                if (td >= 32) {
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < E; u++) sk[u * RS_THREADS + tid] = a[u];
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < E; u++) {
                        const unsigned long long b = sk[u * RS_THREADS + (tid ^ td)];
                        a[u] = ((b < a[u]) == want_min) ? b : a[u];       // one 64-bit compare + select (min and max separately cost twice that)
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < E; u++) {
                        const unsigned long long b = __shfl_xor_sync(0xffffffffu, a[u], td);
                        a[u] = ((b < a[u]) == want_min) ? b : a[u];
                    }
                }
            }
#pragma unroll
            for (int jl = E / 2; jl > 0; jl >>= 1) {                     // strides inside the thread's block: registers only
                if (jl < k) {
#pragma unroll
                    for (int u = 0; u < E; u++) {
                        if ((u & jl) == 0) {
                            const bool asc = k >= E ? asc_t : ((u & k) == 0);
                            const bool sw = (a[u | jl] < a[u]) == asc;      // out of order for this direction: swap
                            const unsigned long long t0 = a[u], t1 = a[u | jl];
                            a[u] = sw ? t1 : t0; a[u | jl] = sw ? t0 : t1;
                        }
                    }
                }
            }
        }
        int32_t *out_id = p.cand_ids + (size_t)q * p.B + s0, *out_sc = p.cand_scores + (size_t)q * p.B + s0;
#pragma unroll
        for (int u = 0; u < E; u++) {
            const int i = tid * E + u;
            if (i < m) {
                const uint32_t e = (uint32_t)(a[u] & 0xffffull), v = e >> 6;
                const int g = (int)v / probes, jv = (int)v - g * probes;
                const int64_t part = x.vis_part[(q * TD + g) * probes + jv];
                out_id[i] = __ldg(rv.ids + (int64_t)g * rv.n_ids + part * kBlock + (e & 63u));
                out_sc[i] = (int32_t)(a[u] >> 40);
            }
        }
        __syncthreads();
    }
}

bool route_fast_eligible(const RoutingView &rv, int probes, int64_t hard_cap, int B, RouteFastExtra &x, size_t &smem) {
    x.v1_ok = 0; x.v2_smem = 0; x.v2_big = 0;
    if (probes < 1) return false;
    const int64_t n_raw = (int64_t)rv.TD * probes * kBlock;
    if (n_raw > 32000) return false;                         // positions and slots carry a 1-bit tag in 16 bits
    if (rv.m * rv.lambda > 126) return false;               // scores are packed into 7 bits (see best_s encoding)
    if ((int64_t)rv.TD * probes > RQ_THREADS) return false;
    const int64_t b_eff = std::min<int64_t>(B, n_raw);        // a query never has more than n_raw candidates (glove100: B = 40 000 over 12 800 positions)
    int sort_n = 64; while (sort_n < b_eff) sort_n <<= 1;
    int tbl = 1024; while (tbl < n_raw + n_raw / 4) tbl <<= 1;
    const bool tbl_ok = tbl <= 32768;                        // the one-CTA kernel's table holds positions: it needs the full size
    if (!tbl_ok) tbl = 32768;                                // the two-CTA kernel only sizes its filters with it
    x.n_raw = (int)n_raw; x.tbl = tbl; x.sort_n = sort_n; x.wl_extra = 0;
    smem = 0;
    // one-CTA kernel: needs a HARD_CAP that cannot bind (n_raw - 64 < cap: every poll of PIS:657-659 proceeds) and its whole state in shared memory
    if (tbl_ok && n_raw - kBlock < hard_cap) {
        int tbl1 = tbl;
        if (tbl1 < n_raw + sort_n) tbl1 <<= 1;              // room for the class list and the selection list
        const size_t limit = 227 * 1024 - 2048;
        if (tbl1 <= 32768 && route_fast_smem(rv.TD, probes, (int)n_raw, tbl1, sort_n, 0) <= limit) {
            const size_t base = route_fast_smem(rv.TD, probes, (int)n_raw, tbl1, sort_n, 0);
            // worklist of involved positions: wl_extra dedicated entries + the 6*sort_n that overlay skey / sid
            int64_t wl_extra = std::max<int64_t>(0, n_raw - 6 * (int64_t)sort_n);
            wl_extra = std::min<int64_t>(wl_extra, (int64_t)(limit - base) / 2);
            wl_extra &= ~(int64_t)7;                         // keeps skey 16-byte aligned
            x.v1_ok = 1; x.tbl1 = tbl1; x.wl_extra = (int)wl_extra;
            smem = route_fast_smem(rv.TD, probes, (int)n_raw, tbl1, sort_n, (int)wl_extra);
        }
    }
    // two-CTA kernel: any HARD_CAP; B <= 1024 sorted in the kernel, B <= 32768 by route_sort_big_kernel (segments of <= 16384 keys); a worklist of at least 1/8 of the positions
    Route2Layout L{};
    if (sort_n <= 2 * RS_SEGMAX && (int64_t)rv.TD * probes <= R2_THREADS && route2_layout(rv.TD, probes, (int)n_raw, tbl, L) &&
        (L.wl_cap >= n_raw / 8 || L.wl_cap >= n_raw)) {
        x.v2_region = L.region; x.v2_cls_cap = L.cls_cap; x.v2_wl_cap = L.wl_cap; x.v2_smem = L.smem;
        x.v2_big = sort_n > 2 * R2_H;
    }
    return x.v1_ok || x.v2_smem;
}

int configure_route_kernels() {   // per-device opt-in, see configure_tokengen_kernels
    return opt_in_smem(route_fast_kernel) || opt_in_smem(route_fast2_kernel<false, false>) || opt_in_smem(route_fast2_kernel<true, false>) ||
           opt_in_smem(route_fast2_kernel<false, true>) || opt_in_smem(route_fast2_kernel<true, true>) || opt_in_smem(route_sort_big_kernel<1>) ||
           opt_in_smem(route_sort_big_kernel<2>) || opt_in_smem(route_sort_big_kernel<4>) || opt_in_smem(route_sort_big_kernel<8>) ||
           opt_in_smem(route_sort_big_kernel<16>) || opt_in_smem(route_kernel) ? -1 : 0;
}

// Fast path of Route.  x.v2_smem != 0: route_fast2_kernel (+ route_sort_big_kernel for B > 1024); the queries it hands back (x.ovf_list) go to
// route_fast_kernel when that kernel is eligible (x.v1_ok), otherwise to the general kernel (`pg` = its parameters, scratch included).
int launch_route_fast(cudaStream_t s, const RoutingView &rv, const RouteParams &p, RouteFastExtra x, size_t smem, int sm_count,
                      int32_t *vis_part, uint8_t *vis_score, uint8_t *vis_n, const RouteParams *pg, int grid_g) {
    if (p.Q <= 0) return 0;
    const int64_t nthreads = p.Q * rv.TD;
    route_probe_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(rv, p.Q, p.codes, p.probes, vis_part, vis_score, vis_n);
    x.vis_part = vis_part; x.vis_score = vis_score; x.vis_n = vis_n;
    const int grid = (int)std::min<int64_t>(p.Q, sm_count);
    int launches = 1;
    if (x.v2_smem && x.ovf_n) {
        if (cudaMemsetAsync(x.ovf_n, 0, 2 * sizeof(int32_t), s) != cudaSuccess) return -1;     // [0] overflow count, [1] next query
        const int grid2 = (int)std::min<int64_t>(p.Q, 2 * (int64_t)sm_count);
        if (x.v2_big) {
            if (rv.deleted) route_fast2_kernel<true, true><<<grid2, R2_THREADS, x.v2_smem, s>>>(rv, p, x);
            else route_fast2_kernel<false, true><<<grid2, R2_THREADS, x.v2_smem, s>>>(rv, p, x);
            const int seg_n = std::min(x.sort_n, RS_SEGMAX);
            const size_t ssm = sizeof(unsigned long long) * (size_t)seg_n;
            switch (seg_n / RS_THREADS) {
                case 2: route_sort_big_kernel<2><<<grid, RS_THREADS, ssm, s>>>(rv, p, x); break;
                case 4: route_sort_big_kernel<4><<<grid, RS_THREADS, ssm, s>>>(rv, p, x); break;
                case 8: route_sort_big_kernel<8><<<grid, RS_THREADS, ssm, s>>>(rv, p, x); break;
                case 16: route_sort_big_kernel<16><<<grid, RS_THREADS, ssm, s>>>(rv, p, x); break;
                default: return -1;
            }
            launches++;
        } else {
            if (rv.deleted) route_fast2_kernel<true, false><<<grid2, R2_THREADS, x.v2_smem, s>>>(rv, p, x);
            else route_fast2_kernel<false, false><<<grid2, R2_THREADS, x.v2_smem, s>>>(rv, p, x);
        }
        launches += 2;
        if (x.v1_ok) {
            x.qlist = x.ovf_list; x.qlist_n = x.ovf_n; x.tbl = x.tbl1;
            route_fast_kernel<<<grid, RQ_THREADS, smem, s>>>(rv, p, x);
        } else {
            if (!pg) return -1;
            RouteParams g = *pg;
            g.qlist = x.ovf_list; g.qlist_n = x.ovf_n;
            const size_t gsm = sizeof(int32_t) * ((size_t)2 * rv.TD * p.probes + rv.TD);
            route_kernel<<<grid_g, RT_THREADS, gsm, s>>>(rv, g);
        }
        return cudaGetLastError() == cudaSuccess ? launches : -1;
    }
    if (!x.v1_ok) return -1;
    x.qlist = nullptr; x.qlist_n = nullptr; x.tbl = x.tbl1;
    route_fast_kernel<<<grid, RQ_THREADS, smem, s>>>(rv, p, x);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

}  // namespace fsp
