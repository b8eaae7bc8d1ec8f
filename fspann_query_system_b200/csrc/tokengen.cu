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
                                                                 uint64_t *__restrict__ codes, int groups_per_cta, int qstride, int astride,
                                                                 const int32_t *__restrict__ run_if) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (run_if && *run_if == 0) return;                           // fallback launch behind the FP32 pre-filter: only after a worklist overflow
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

// =====================================================================================================================
// FP32 pre-filter + exact re-check.  h_j = floor((y + r_j) / omega_j) only needs y = v . alpha_j to within the distance of
// (y + r_j) / omega_j from the nearest integer, and omega_j is a sizeable fraction of the data's projection range
// (Coding:224-237), so a single-precision dot product with a PROVEN error bound decides almost every h_j; the few that fall
// inside the bound of a quantisation boundary are listed and recomputed with the exact sequential FP64 arithmetic above.
//   |fl32(sum) - y_java| <= E(v, j) = C(dim) * ||v||_2 * ||alpha_j||_2,   C = 4 * ((dim + 2) * 2^-24 * 1.01 + 2^-22)
//   (inputs rounded to FP32: 2 * 2^-24 per product; dim sequential FMAs: gamma_dim; Java's own FP64 rounding is < 2^-45; x4 safety)
// decided  <=>  t = ((double)sum + r_j) / omega_j is farther than E / omega_j + 1e-9 * (1 + |t|) from both neighbouring integers
//              and |t| < 2^30 (no saturation of Java's (int) cast).
// Codes are bit-identical to the exact kernel by construction; tests/ compare the two paths over the full 1M-vector base set.
// =====================================================================================================================
constexpr int TF_QT = 128;       // queries per CTA: 2 halves of 64, a thread owns 2 queries (lane, lane + 32) of one half
constexpr int TF_THREADS = 128;  // 4 warps = 2 query halves x 2 projection slices
constexpr int TF_JB = 12;        // projections per thread (slice = ceil(m / 2) <= 12, i.e. m <= 24; larger m runs the exact kernel)

__global__ void alpha_prepare_kernel(const double *__restrict__ alpha, int64_t rows, int dim, float *__restrict__ alpha_f, float *__restrict__ norm) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double s = 0.0;
    for (int i = 0; i < dim; i++) { const double a = alpha[r * dim + i]; alpha_f[r * dim + i] = (float)a; s += a * a; }
    norm[r] = __double2float_ru(sqrt(s) * 1.000001);
}
int launch_alpha_prepare(cudaStream_t s, const double *alpha, int64_t rows, int dim, float *alpha_f, float *norm) {
    if (rows <= 0) return 0;
    alpha_prepare_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, s>>>(alpha, rows, dim, alpha_f, norm);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// FP32 copy of the query rows (zero-padded to dim4 = a multiple of 4 floats) and an upper bound of every row's 2-norm.
__global__ void tokengen_prep_kernel(const double *__restrict__ queries, int64_t Q, int dim, int dim4, float *__restrict__ qf, float *__restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= Q) return;
    float ss = 0.f;
    for (int i = lane; i < dim4; i += 32) {
        const float v = i < dim ? (float)queries[row * dim + i] : 0.f;
        qf[row * dim4 + i] = v;
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) norm[row] = __fmul_ru(__fsqrt_ru(ss), 1.0001f);
}

template <int NJ>
__device__ __forceinline__ void tf_dot_block(const float *qa, const float *qb, const float *as, int astride, int n4, int jb, float accA[TF_JB], float accB[TF_JB]) {
#pragma unroll
    for (int u = 0; u < NJ; u++) { accA[u] = 0.f; accB[u] = 0.f; }
    const float4 *qa4 = reinterpret_cast<const float4 *>(qa), *qb4 = reinterpret_cast<const float4 *>(qb);
    const float4 *a4 = reinterpret_cast<const float4 *>(as + (size_t)jb * astride);
    const int as4 = astride >> 2;
#pragma unroll 2
    for (int i4 = 0; i4 < n4; i4++) {
        const float4 va = qa4[i4], vb = qb4[i4];
#pragma unroll
        for (int u = 0; u < NJ; u++) {
            const float4 al = a4[u * as4 + i4];
            accA[u] = fmaf(va.x, al.x, accA[u]); accB[u] = fmaf(vb.x, al.x, accB[u]);
            accA[u] = fmaf(va.y, al.y, accA[u]); accB[u] = fmaf(vb.y, al.y, accB[u]);
            accA[u] = fmaf(va.z, al.z, accA[u]); accB[u] = fmaf(vb.z, al.z, accB[u]);
            accA[u] = fmaf(va.w, al.w, accA[u]); accB[u] = fmaf(vb.w, al.w, accB[u]);
        }
    }
}

// work[0] = number of listed (q, g, j); work[1] = overflow flag; list entries = q << 24 | g << 8 | j  (g < 65536, j < 256)
__global__ void __launch_bounds__(TF_THREADS, 2) tokengen_fast_kernel(RoutingView rv, int64_t Q, const float *__restrict__ qf, const float *__restrict__ qnorm,
                                                                      uint64_t *__restrict__ codes, int groups_per_cta, int qstride, int astride,
                                                                      float cbound, int32_t *__restrict__ work, unsigned long long *__restrict__ list,
                                                                      int64_t list_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dim = rv.dim, m = rv.m, W = rv.W, lambda = rv.lambda;
    float *qs = reinterpret_cast<float *>(smem_raw);               // [TF_QT][qstride] (rows zero-padded to a multiple of 4)
    float *as = qs + (size_t)TF_QT * qstride;                      // [m][astride]
    double *rs = reinterpret_cast<double *>(as + (size_t)m * astride);   // [m] r
    double *os = rs + m;                                           // [m] omega
    float *na = reinterpret_cast<float *>(os + m);                 // [m] ||alpha_j||
    float *rf = na + ((m + 3) & ~3);                               // [m] r_j in FP32
    float *iof = rf + ((m + 3) & ~3);                              // [m] 1 / omega_j in FP32
    float *nv = iof + ((m + 3) & ~3);                              // [TF_QT] ||v||
    uint32_t *cs = reinterpret_cast<uint32_t *>(nv + TF_QT);       // [TF_QT][2*W] code words being assembled

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * TF_QT;
    const int nq = (int)min((int64_t)TF_QT, Q - q0);
    const int n4 = (dim + 3) >> 2;

    // stage the FP32 query tile (128-bit loads from the prepared copy, several in flight per thread) and the row norms
    for (int idx = tid; idx < TF_QT * n4; idx += TF_THREADS) {
        const int r = idx / n4, c = idx - r * n4;
        const float4 v = r < nq ? __ldg(reinterpret_cast<const float4 *>(qf + (q0 + r) * (int64_t)(4 * n4)) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(qs + (size_t)r * qstride + 4 * c) = v;
    }
    for (int r = tid; r < TF_QT; r += TF_THREADS) nv[r] = r < nq ? qnorm[q0 + r] : 0.f;

    const int half = warp & 1, slice = warp >> 1;
    const int jper = (m + 1) / 2;
    const int j_lo = slice * jper, j_hi = min(m, j_lo + jper);
    const int g_lo = blockIdx.y * groups_per_cta, g_hi = min(rv.TD, g_lo + groups_per_cta);
    const int ra = half * 64 + lane, rb = ra + 32;
    const float *qa = qs + (size_t)ra * qstride, *qb = qs + (size_t)rb * qstride;

    for (int g = g_lo; g < g_hi; g++) {
        __syncthreads();  // previous group's tile / code words fully consumed; query tile staged
        const float *ag = rv.alpha_f32 + (size_t)g * m * dim;
        if ((dim & 3) == 0) {
            for (int idx = tid; idx < m * n4; idx += TF_THREADS) reinterpret_cast<float4 *>(as)[idx] = __ldg(reinterpret_cast<const float4 *>(ag) + idx);
        } else {
            for (int idx = tid; idx < m * 4 * n4; idx += TF_THREADS) { const int j = idx / (4 * n4), i = idx - j * 4 * n4; as[(size_t)j * astride + i] = i < dim ? ag[(size_t)j * dim + i] : 0.f; }
        }
        for (int idx = tid; idx < m; idx += TF_THREADS) { rs[idx] = rv.r[(size_t)g * m + idx]; os[idx] = rv.omega[(size_t)g * m + idx]; na[idx] = rv.alpha_norm[(size_t)g * m + idx];
                                                           rf[idx] = (float)rs[idx]; iof[idx] = (float)(1.0 / os[idx]); }
        for (int idx = tid; idx < TF_QT * 2 * W; idx += TF_THREADS) cs[idx] = 0u;
        __syncthreads();

        float accA[TF_JB], accB[TF_JB];
        const int nj = j_hi - j_lo;
        switch (nj) {
            case 12: tf_dot_block<12>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 11: tf_dot_block<11>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 10: tf_dot_block<10>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 9: tf_dot_block<9>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 8: tf_dot_block<8>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 7: tf_dot_block<7>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 6: tf_dot_block<6>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 5: tf_dot_block<5>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 4: tf_dot_block<4>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 3: tf_dot_block<3>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 2: tf_dot_block<2>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            case 1: tf_dot_block<1>(qa, qb, as, astride, n4, j_lo, accA, accB); break;
            default: break;
        }
        // ---- decide h_j in single precision; undecided projections go to the exact re-check list and contribute no bits here ----
        int32_t hA[TF_JB], hB[TF_JB];
        uint32_t decA = 0, decB = 0;                                   // bit u: projection j_lo + u was decided here
#pragma unroll
        for (int u = 0; u < TF_JB; u++) {
            hA[u] = 0; hB[u] = 0;
            if (u < nj) {
                const int j = j_lo + u;
                const float rj = rf[j], io = iof[j], bj = cbound * na[j] * io;
#pragma unroll
                for (int hb = 0; hb < 2; hb++) {
                    const int qi = hb ? rb : ra;
                    const float t = ((hb ? accB[u] : accA[u]) + rj) * io;
                    const float f = floorf(t);
                    // |t - t_java| <= E / omega (dot product) + a few 2^-24 relative roundings of the FP32 epilogue (r, 1/omega, +, *)
                    const float slack = bj * nv[qi] + 2e-6f * (1.0f + fabsf(t));
                    const bool decided = (t - f > slack) && (f + 1.0f - t > slack) && fabsf(t) < 4194304.0f;      // false for NaN / Inf too
                    if (decided) { if (hb) { hB[u] = (int32_t)f; decB |= 1u << u; } else { hA[u] = (int32_t)f; decA |= 1u << u; } }
                    else if (qi < nq) {
                        const int at = atomicAdd(&work[0], 1);
                        if (at < list_cap) list[at] = ((unsigned long long)(q0 + qi) << 24) | ((unsigned long long)g << 8) | (unsigned long long)j;
                        else work[1] = 1;
                    }
                }
            }
        }
        // bit planes (Coding:291-299): plane ib holds bit ib of h_j at position (lambda-1-ib)*m + j; this thread's projections are
        // consecutive, so a plane's bits form one <= 12-bit field spanning at most two 32-bit words
        for (int ib = 0; ib < lambda && ib < 32; ib++) {
            uint32_t fa = 0, fb = 0;
#pragma unroll
            for (int u = 0; u < TF_JB; u++) {
                if (u < nj) {                                           // (h ^ 0x80000000) only flips bit 31 (Coding:293)
                    fa |= ((((uint32_t)hA[u] ^ 0x80000000u) >> ib) & 1u) << u;
                    fb |= ((((uint32_t)hB[u] ^ 0x80000000u) >> ib) & 1u) << u;
                }
            }
            fa &= decA; fb &= decB;                                     // undecided entries get their bits from the exact re-check
            const int pos0 = (lambda - 1 - ib) * m + j_lo, w0 = pos0 >> 5, sh = pos0 & 31;
            if (ra < nq) {
                if (fa << sh) atomicOr(&cs[(size_t)ra * 2 * W + w0], fa << sh);
                if (sh && (fa >> (32 - sh))) atomicOr(&cs[(size_t)ra * 2 * W + w0 + 1], fa >> (32 - sh));
            }
            if (rb < nq) {
                if (fb << sh) atomicOr(&cs[(size_t)rb * 2 * W + w0], fb << sh);
                if (sh && (fb >> (32 - sh))) atomicOr(&cs[(size_t)rb * 2 * W + w0 + 1], fb >> (32 - sh));
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nq * W; idx += TF_THREADS) {
            const int qq = idx / W, w = idx - qq * W;
            const uint64_t lo = cs[(size_t)qq * 2 * W + 2 * w], hi = cs[(size_t)qq * 2 * W + 2 * w + 1];
            codes[((q0 + qq) * rv.TD + g) * W + w] = lo | (hi << 32);
        }
    }
}

// Exact sequential FP64 value of every listed projection (Coding:250-258, 349-353), OR-ed into the code words.
__global__ void tokengen_recheck_kernel(RoutingView rv, const double *__restrict__ queries, uint64_t *__restrict__ codes, const int32_t *__restrict__ work,
                                        const unsigned long long *__restrict__ list, int64_t list_cap) {
    if (work[1]) return;                                            // overflow: the exact kernel recomputes the whole batch
    const int64_t n = min((int64_t)work[0], list_cap);
    const int dim = rv.dim, m = rv.m, W = rv.W, lambda = rv.lambda;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < n; it += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long e = list[it];
        const int64_t q = (int64_t)(e >> 24);
        const int g = (int)((e >> 8) & 0xffffu), j = (int)(e & 0xffu);
        const double *v = queries + q * dim, *al = rv.alpha + ((size_t)g * m + j) * dim;
        double acc = 0.0;
        int i = 0;
        // strictly sequential adds (Coding:349-353), loads software-pipelined: the 16 loads of the next 8 terms are issued before the adds of the
        // current 8, so a thread waits for one memory latency per 8 terms only while the pipeline fills (few threads have work: latency-bound)
        if (dim >= 8) {
            double a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { a[u] = __ldg(v + u); b[u] = __ldg(al + u); }
            for (; i + 16 <= dim; i += 8) {
                double a2[8], b2[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { a2[u] = __ldg(v + i + 8 + u); b2[u] = __ldg(al + i + 8 + u); }
#pragma unroll
                for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, __dmul_rn(a[u], b[u]));
#pragma unroll
                for (int u = 0; u < 8; u++) { a[u] = a2[u]; b[u] = b2[u]; }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, __dmul_rn(a[u], b[u]));
            i += 8;
        }
        for (; i < dim; i++) acc = __dadd_rn(acc, __dmul_rn(v[i], al[i]));
        const double y = __dadd_rn(acc, rv.r[(size_t)g * m + j]);
        const double f = floor(__ddiv_rn(y, rv.omega[(size_t)g * m + j]));
        const uint32_t hj = (uint32_t)__double2int_rz(f) ^ 0x80000000u;
        for (int ib = 0; ib < lambda; ib++) {
            if (ib < 32 && ((hj >> ib) & 1u)) {
                const int pos = (lambda - 1 - ib) * m + j;
                atomicOr(reinterpret_cast<unsigned long long *>(codes + (q * rv.TD + g) * W + (pos >> 6)), 1ull << (pos & 63));
            }
        }
    }
}

// The opt-in to > 48 KB of dynamic shared memory is a per-DEVICE attribute of each kernel: fspann_ctx_create calls this after
// cudaSetDevice, so every context's device is configured no matter how many GPUs one process drives.
int configure_tokengen_kernels() {
    return opt_in_smem(tokengen_kernel) || opt_in_smem(tokengen_fast_kernel) || configure_tokengen_tc_kernels() ? -1 : 0;
}

static int launch_tokengen_exact(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, uint64_t *codes, const int32_t *run_if, int sm_count) {
    const int dim_even = (rv.dim + 1) & ~1;
    int qstride = dim_even + 2;                      // doubles; 16-byte aligned rows whose granule count is odd => conflict-free 128-bit loads
    if (((qstride >> 1) & 1) == 0) qstride += 2;
    const int astride = dim_even;
    const size_t smem = sizeof(double) * ((size_t)TG_QT * qstride + (size_t)rv.m * astride + 2 * (size_t)rv.m) + sizeof(uint32_t) * (size_t)TG_QT * 2 * rv.W;
    if (smem > (size_t)kMaxDynSmem) return -1;
    // enough CTAs for >= 8 waves of the 2-per-SM slots when the batch allows it (tail below ~10 %)
    const int64_t tiles = (Q + TG_QT - 1) / TG_QT;
    int gsplit = (int)((2LL * sm_count * 8 + tiles - 1) / tiles);
    if (gsplit < 1) gsplit = 1;
    if (gsplit > rv.TD) gsplit = rv.TD;
    const int groups_per_cta = (rv.TD + gsplit - 1) / gsplit;
    gsplit = (rv.TD + groups_per_cta - 1) / groups_per_cta;
    dim3 grid((unsigned)tiles, (unsigned)gsplit);
    tokengen_kernel<<<grid, TG_THREADS, smem, s>>>(rv, Q, queries, codes, groups_per_cta, qstride, astride, run_if);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int64_t tokengen_list_capacity(const RoutingView &rv, int64_t Q) {
    const int64_t total = Q * (int64_t)rv.TD * rv.m;
    return std::max<int64_t>(65536, total / 8);
}

// work: int32[2] + list: uint64[tokengen_list_capacity] scratch.  mode: 0 = automatic (tensor-core pre-filter when the shape allows it, else
// the FP32 pre-filter, else exact), 1 = the exact FP64 kernel alone, 2 = FP32 pre-filter (no tensor cores).  *path_out: 1 exact, 2 FP32, 3 tensor.
int launch_tokengen(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, uint64_t *codes, int32_t *work,
                    unsigned long long *list, int64_t list_cap, float *qf, float *qnorm, int mode, int sm_count, int *path_out) {
    if (Q <= 0) return 0;
    const bool exact_only = mode == 1;
    if (path_out) *path_out = 1;
    if (mode == 0 && work && list && rv.alpha_tc && rv.alpha_norm && Q * (int64_t)rv.TD * rv.m < (1LL << 31)) {
        if (cudaMemsetAsync(work, 0, 2 * sizeof(int32_t), s) != cudaSuccess) return -1;
        const int n_tc = launch_tokengen_tc(s, rv, Q, queries, rv.alpha_tc, codes, work, list, list_cap, sm_count);
        if (n_tc < 0) return -1;
        if (n_tc > 0) {
            tokengen_recheck_kernel<<<sm_count * 2, 256, 0, s>>>(rv, queries, codes, work, list, list_cap);
            if (cudaGetLastError() != cudaSuccess) return -1;
            const int n = launch_tokengen_exact(s, rv, Q, queries, codes, work + 1, sm_count);      // runs only after a worklist overflow
            if (path_out) *path_out = 3;
            return n < 0 ? -1 : n_tc + 1 + n;
        }
    }
    const bool fast_ok = !exact_only && work && list && qf && qnorm && rv.alpha_f32 && rv.alpha_norm && (rv.m + 1) / 2 <= TF_JB && rv.TD < 65536 && rv.m < 256 &&
                         Q * (int64_t)rv.TD * rv.m < (1LL << 31);        // the re-check counter is an int32: it can never wrap
    if (!fast_ok) return launch_tokengen_exact(s, rv, Q, queries, codes, nullptr, sm_count);
    const int n4 = (rv.dim + 3) >> 2;
    int qg = n4 + 1;                                 // row stride in 16-byte granules, odd => conflict-free 128-bit loads
    if ((qg & 1) == 0) qg++;
    const int qstride = 4 * qg, astride = 4 * n4;
    const size_t smem = sizeof(float) * ((size_t)TF_QT * qstride + (size_t)rv.m * astride) + sizeof(double) * 2 * (size_t)rv.m +
                        sizeof(float) * (3 * (((size_t)rv.m + 3) / 4 * 4) + TF_QT) + sizeof(uint32_t) * (size_t)TF_QT * 2 * rv.W;
    if (smem > 113 * 1024) return launch_tokengen_exact(s, rv, Q, queries, codes, nullptr, sm_count);
    if (cudaMemsetAsync(work, 0, 2 * sizeof(int32_t), s) != cudaSuccess) return -1;
    const int64_t tiles = (Q + TF_QT - 1) / TF_QT;
    int gsplit = (int)((2LL * sm_count * 4 + tiles - 1) / tiles);
    if (gsplit < 1) gsplit = 1;
    if (gsplit > rv.TD) gsplit = rv.TD;
    int groups_per_cta = (rv.TD + gsplit - 1) / gsplit;
    if (groups_per_cta < 4) groups_per_cta = std::min(4, (int)rv.TD);   // amortise the tile staging over >= 4 groups
    gsplit = (rv.TD + groups_per_cta - 1) / groups_per_cta;
    const float cbound = 4.0f * ((float)(rv.dim + 2) * 5.9604645e-8f * 1.01f + 2.3841858e-7f);
    dim3 grid((unsigned)tiles, (unsigned)gsplit);
    tokengen_prep_kernel<<<(unsigned)((Q * 32 + 255) / 256), 256, 0, s>>>(queries, Q, rv.dim, 4 * n4, qf, qnorm);
    tokengen_fast_kernel<<<grid, TF_THREADS, smem, s>>>(rv, Q, qf, qnorm, codes, groups_per_cta, qstride, astride, cbound, work, list, list_cap);
    tokengen_recheck_kernel<<<sm_count * 2, 256, 0, s>>>(rv, queries, codes, work, list, list_cap);
    if (cudaGetLastError() != cudaSuccess) return -1;
    const int n = launch_tokengen_exact(s, rv, Q, queries, codes, work + 1, sm_count);      // runs only after a worklist overflow
    if (path_out) *path_out = 2;
    return n < 0 ? -1 : 3 + n;
}

}  // namespace fsp
