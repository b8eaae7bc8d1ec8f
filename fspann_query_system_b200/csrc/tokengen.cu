// tokengen.cu -- stage 1: routing codes C(v) for every (table, division) of every query.
//
// Replaces QueryTokenFactory.create's coding loop (query/.../QueryTokenFactory.java:98-131) ->
// Coding.C / Coding.H / Coding.dot (index/.../paper/Coding.java:285-301, 250-258, 349-353).
//
// Bit-exactness: h_j = (int)floor((sum_i v_i*alpha_ji + r_j) / omega_j) must match Java's strict FP64, so the
// dot product is accumulated in index order with separate round-to-nearest multiply and add (no FMA
// contraction: __dmul_rn / __dadd_rn), then IEEE division and a saturating floor-to-int32 (Java (int) cast).
// The FP64 pipe does this at a few hundred microseconds per 10k-query batch (3.9 GFLOP), so the exact path
// IS the fast path at query time; there is nothing to re-check.
//
// Layout: one CTA = 64 queries (rows staged once in shared memory, padded to an odd stride so the per-query
// walks are bank-conflict free) x a contiguous chunk of (t,d) groups whose alpha tiles stream through shared
// memory.  A warp holds 32 queries of one projection slice, so alpha reads are broadcasts.
#include "fspann_internal.cuh"

namespace fsp {

constexpr int TG_QT = 64;        // queries per CTA
constexpr int TG_THREADS = 256;  // 4 projection slices x 64 queries
constexpr int TG_JB = 8;         // accumulators per register block

__global__ void __launch_bounds__(TG_THREADS) tokengen_kernel(RoutingView rv, int64_t Q, const double *__restrict__ queries,
                                                              uint64_t *__restrict__ codes, int groups_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dim = rv.dim, m = rv.m, W = rv.W, lambda = rv.lambda;
    const int qstride = dim | 1;                                   // odd stride in doubles
    double *qs = reinterpret_cast<double *>(smem_raw);             // [TG_QT][qstride]
    double *as = qs + (size_t)TG_QT * qstride;                     // [m][dim] alpha tile of the current group
    double *rs = as + (size_t)m * dim;                             // [m] r
    double *os = rs + m;                                           // [m] omega
    uint32_t *cs = reinterpret_cast<uint32_t *>(os + m);           // [TG_QT][2*W] code words being assembled

    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * TG_QT;
    const int nq = (int)min((int64_t)TG_QT, Q - q0);

    // stage the query tile (coalesced over the row-major [Q][dim] input)
    for (int idx = tid; idx < nq * dim; idx += TG_THREADS) {
        const int qi = idx / dim, i = idx - qi * dim;
        qs[(size_t)qi * qstride + i] = queries[(q0 + qi) * dim + i];
    }

    const int qi = tid & (TG_QT - 1);
    const int slice = tid / TG_QT;                 // 0..3
    const int jper = (m + 3) / 4;
    const int j_lo = slice * jper, j_hi = min(m, j_lo + jper);
    const int g_lo = blockIdx.y * groups_per_cta, g_hi = min(rv.TD, g_lo + groups_per_cta);

    for (int g = g_lo; g < g_hi; g++) {
        __syncthreads();  // previous group's tile / code words fully consumed; query tile staged
        const double *ag = rv.alpha + (size_t)g * m * dim;
        for (int idx = tid; idx < m * dim; idx += TG_THREADS) as[idx] = ag[idx];
        for (int idx = tid; idx < m; idx += TG_THREADS) { rs[idx] = rv.r[(size_t)g * m + idx]; os[idx] = rv.omega[(size_t)g * m + idx]; }
        for (int idx = tid; idx < TG_QT * 2 * W; idx += TG_THREADS) cs[idx] = 0u;
        __syncthreads();

        if (qi < nq) {
            const double *qrow = qs + (size_t)qi * qstride;
            for (int jb = j_lo; jb < j_hi; jb += TG_JB) {
                double acc[TG_JB];
#pragma unroll
                for (int u = 0; u < TG_JB; u++) acc[u] = 0.0;
                const int nj = min(TG_JB, j_hi - jb);
                if (nj == TG_JB) {
                    for (int i = 0; i < dim; i++) {
                        const double v = qrow[i];
#pragma unroll
                        for (int u = 0; u < TG_JB; u++) acc[u] = __dadd_rn(acc[u], __dmul_rn(v, as[(size_t)(jb + u) * dim + i]));
                    }
                } else {
                    for (int i = 0; i < dim; i++) {
                        const double v = qrow[i];
#pragma unroll
                        for (int u = 0; u < TG_JB; u++)
                            if (u < nj) acc[u] = __dadd_rn(acc[u], __dmul_rn(v, as[(size_t)(jb + u) * dim + i]));
                    }
                }
#pragma unroll
                for (int u = 0; u < TG_JB; u++) {
                    if (u < nj) {
                        const int j = jb + u;
                        const double y = __dadd_rn(acc[u], rs[j]);                 // dot(v, alpha_j) + r_j   (Coding:254)
                        const double f = floor(__ddiv_rn(y, os[j]));               // Math.floor(y / omega_j) (Coding:255)
                        const int32_t h = __double2int_rz(f);                      // Java (int): saturating, NaN -> 0
                        const uint32_t hj = (uint32_t)h ^ 0x80000000u;             // Coding:293
                        for (int ib = 0; ib < lambda; ib++) {
                            const int bit = ib < 32 ? (int)((hj >> ib) & 1u) : 0;
                            if (bit) {
                                const int pos = (lambda - 1 - ib) * m + j;        // MSB-first bit planes (Coding:291-299)
                                atomicOr(&cs[(size_t)qi * 2 * W + (pos >> 5)], 1u << (pos & 31));
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // write the group's codes: [Q][TD][W] uint64, little-endian word pairs
        for (int idx = tid; idx < nq * W; idx += TG_THREADS) {
            const int qq = idx / W, w = idx - qq * W;
            const uint64_t lo = cs[(size_t)qq * 2 * W + 2 * w], hi = cs[(size_t)qq * 2 * W + 2 * w + 1];
            codes[((q0 + qq) * rv.TD + g) * W + w] = lo | (hi << 32);
        }
    }
}

int launch_tokengen(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, uint64_t *codes) {
    if (Q <= 0) return 0;
    const int qstride = rv.dim | 1;
    const size_t smem = sizeof(double) * ((size_t)TG_QT * qstride + (size_t)rv.m * rv.dim + 2 * (size_t)rv.m) + sizeof(uint32_t) * (size_t)TG_QT * 2 * rv.W;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(tokengen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    const int64_t tiles = (Q + TG_QT - 1) / TG_QT;
    int gsplit = (int)((296 + tiles - 1) / tiles);
    if (gsplit < 1) gsplit = 1;
    if (gsplit > rv.TD) gsplit = rv.TD;
    const int groups_per_cta = (rv.TD + gsplit - 1) / gsplit;
    gsplit = (rv.TD + groups_per_cta - 1) / groups_per_cta;
    dim3 grid((unsigned)tiles, (unsigned)gsplit);
    tokengen_kernel<<<grid, TG_THREADS, smem, s>>>(rv, Q, queries, codes, groups_per_cta);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace fsp
