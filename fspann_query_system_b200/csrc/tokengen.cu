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
// Layout: one CTA = 64 queries (rows staged once in shared memory) x a contiguous chunk of (t,d) groups whose alpha tiles
// stream through shared memory.  A thread owns TWO queries (lane, lane+32) and up to 8 projections of one of four
// projection slices, and walks the dimension two values at a time: per step 2 conflict-free 128-bit loads of its query rows
// (row stride = an odd number of 16-byte granules) and one broadcast 128-bit load per projection feed 4 multiply-adds per
// projection, so the loop is bound by the FP64 pipe (2 x 2*m*T*D*d flop per query), not by shared-memory wavefronts.
// Two CTAs (91 KB at d=128, m=24) share an SM so one's tile load overlaps the other's arithmetic.
#include "fspann_internal.cuh"

namespace fsp {

constexpr int TG_QT = 64;        // queries per CTA
constexpr int TG_THREADS = 128;  // 4 projection slices (warps) x 32 query pairs
constexpr int TG_JB = 8;         // projections per register block

// Accumulates NJ projections (rows jb.. of the alpha tile) for the two query rows qa / qb; strictly sequential per accumulator.
template <int NJ>
__device__ __forceinline__ void tg_dot_block(const double *qa, const double *qb, const double *as, int astride, int dim, int jb, double accA[TG_JB],
                                             double accB[TG_JB]) {
#pragma unroll
    for (int u = 0; u < NJ; u++) { accA[u] = 0.0; accB[u] = 0.0; }
    const double2 *qa2 = reinterpret_cast<const double2 *>(qa), *qb2 = reinterpret_cast<const double2 *>(qb);
    const double2 *a2 = reinterpret_cast<const double2 *>(as + (size_t)jb * astride);
    const int n2 = dim >> 1, as2 = astride >> 1;
#pragma unroll 2
    for (int i2 = 0; i2 < n2; i2++) {
        const double2 va = qa2[i2], vb = qb2[i2];
#pragma unroll
        for (int u = 0; u < NJ; u++) {
            const double2 al = a2[u * as2 + i2];
            accA[u] = __dadd_rn(accA[u], __dmul_rn(va.x, al.x));
            accB[u] = __dadd_rn(accB[u], __dmul_rn(vb.x, al.x));
            accA[u] = __dadd_rn(accA[u], __dmul_rn(va.y, al.y));
            accB[u] = __dadd_rn(accB[u], __dmul_rn(vb.y, al.y));
        }
    }
    if (dim & 1) {
        const double va = qa[dim - 1], vb = qb[dim - 1];
#pragma unroll
        for (int u = 0; u < NJ; u++) {
            const double al = as[(size_t)(jb + u) * astride + dim - 1];
            accA[u] = __dadd_rn(accA[u], __dmul_rn(va, al));
            accB[u] = __dadd_rn(accB[u], __dmul_rn(vb, al));
        }
    }
}

__global__ void __launch_bounds__(TG_THREADS, 2) tokengen_kernel(RoutingView rv, int64_t Q, const double *__restrict__ queries,
                                                                 uint64_t *__restrict__ codes, int groups_per_cta, int qstride, int astride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dim = rv.dim, m = rv.m, W = rv.W, lambda = rv.lambda;
    double *qs = reinterpret_cast<double *>(smem_raw);             // [TG_QT][qstride]
    double *as = qs + (size_t)TG_QT * qstride;                     // [m][astride] alpha tile of the current group
    double *rs = as + (size_t)m * astride;                         // [m] r
    double *os = rs + m;                                           // [m] omega
    uint32_t *cs = reinterpret_cast<uint32_t *>(os + m);           // [TG_QT][2*W] code words being assembled

    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * TG_QT;
    const int nq = (int)min((int64_t)TG_QT, Q - q0);

    // stage the query tile (coalesced over the row-major [Q][dim] input); rows past the batch are zero
    for (int idx = tid; idx < TG_QT * dim; idx += TG_THREADS) {
        const int qi = idx / dim, i = idx - qi * dim;
        qs[(size_t)qi * qstride + i] = qi < nq ? queries[(q0 + qi) * dim + i] : 0.0;
    }

    const int lane = tid & 31, slice = tid >> 5;   // slice 0..3
    const int jper = (m + 3) / 4;
    const int j_lo = slice * jper, j_hi = min(m, j_lo + jper);
    const int g_lo = blockIdx.y * groups_per_cta, g_hi = min(rv.TD, g_lo + groups_per_cta);
    const double *qa = qs + (size_t)lane * qstride, *qb = qs + (size_t)(lane + 32) * qstride;

    for (int g = g_lo; g < g_hi; g++) {
        __syncthreads();  // previous group's tile / code words fully consumed; query tile staged
        const double *ag = rv.alpha + (size_t)g * m * dim;
        for (int idx = tid; idx < m * dim; idx += TG_THREADS) { const int j = idx / dim, i = idx - j * dim; as[(size_t)j * astride + i] = ag[idx]; }
        for (int idx = tid; idx < m; idx += TG_THREADS) { rs[idx] = rv.r[(size_t)g * m + idx]; os[idx] = rv.omega[(size_t)g * m + idx]; }
        for (int idx = tid; idx < TG_QT * 2 * W; idx += TG_THREADS) cs[idx] = 0u;
        __syncthreads();

        for (int jb = j_lo; jb < j_hi; jb += TG_JB) {
            double accA[TG_JB], accB[TG_JB];
            const int nj = min(TG_JB, j_hi - jb);
            switch (nj) {
                case 8: tg_dot_block<8>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 7: tg_dot_block<7>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 6: tg_dot_block<6>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 5: tg_dot_block<5>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 4: tg_dot_block<4>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 3: tg_dot_block<3>(qa, qb, as, astride, dim, jb, accA, accB); break;
                case 2: tg_dot_block<2>(qa, qb, as, astride, dim, jb, accA, accB); break;
                default: tg_dot_block<1>(qa, qb, as, astride, dim, jb, accA, accB); break;
            }
#pragma unroll
            for (int u = 0; u < TG_JB; u++) {
                if (u < nj) {
                    const int j = jb + u;
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int qi = lane + 32 * half;
                        const double y = __dadd_rn(half ? accB[u] : accA[u], rs[j]);   // dot(v, alpha_j) + r_j   (Coding:254)
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
    const int dim_even = (rv.dim + 1) & ~1;
    int qstride = dim_even + 2;                      // doubles; 16-byte aligned rows whose granule count is odd => conflict-free 128-bit loads
    if (((qstride >> 1) & 1) == 0) qstride += 2;
    const int astride = dim_even;
    const size_t smem = sizeof(double) * ((size_t)TG_QT * qstride + (size_t)rv.m * astride + 2 * (size_t)rv.m) + sizeof(uint32_t) * (size_t)TG_QT * 2 * rv.W;
    if (smem > 227 * 1024) return -1;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(tokengen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    // enough CTAs for >= 8 waves of the 2-per-SM slots when the batch allows it (tail below ~10 %)
    const int64_t tiles = (Q + TG_QT - 1) / TG_QT;
    int gsplit = (int)((296 * 8 + tiles - 1) / tiles);
    if (gsplit < 1) gsplit = 1;
    if (gsplit > rv.TD) gsplit = rv.TD;
    const int groups_per_cta = (rv.TD + gsplit - 1) / gsplit;
    gsplit = (rv.TD + groups_per_cta - 1) / groups_per_cta;
    dim3 grid((unsigned)tiles, (unsigned)gsplit);
    tokengen_kernel<<<grid, TG_THREADS, smem, s>>>(rv, Q, queries, codes, groups_per_cta, qstride, astride);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace fsp
