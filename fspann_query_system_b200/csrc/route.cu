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

    for (int64_t q = blockIdx.x; q < p.Q; q += gridDim.x) {
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
                        if (len >= 8) s_overflow = 1;    // 9th entry in a bin: Java would treeify
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

}  // namespace fsp
