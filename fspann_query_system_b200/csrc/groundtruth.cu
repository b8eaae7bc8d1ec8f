// groundtruth.cu -- exact ground truth and recall@K on the device, so "queries/sec at recall@10" can be
// evaluated on the box for synthetic data.
//
// Replaces GroundtruthPrecompute.run (api/src/main/java/com/fspann/api/GroundtruthPrecompute.java:218-276): for every query the
// K nearest base vectors by squared L2 with the reference's arithmetic (VecReader.l2sq, :144-163): per dimension
//   double d = q[i] - b[i]   -- a FLOAT subtraction (both operands are float / float-converted byte), widened to double --
//   sum += d * d             -- sequential FP64, no sqrt,
// ordered by (sum, id) ascending (BY_D_THEN_ID, :168-189); and the recall of ForwardSecureANNSystem.computeMetricsAtK
// (api/.../ForwardSecureANNSystem.java:785-794): |gt[0..K) ∩ result[0..min(K,|result|))| / K.
//
// Two kernels per chunk of queries: (1) a register-tiled distance kernel (4 queries x 4 base vectors per thread, tiles staged
// in shared memory, sums strictly sequential per pair) writing the FP64 distance matrix chunk; (2) one CTA per query: exact
// K-th smallest distance by repeated linear 2048-bin histograms over the shrinking key range, then collection of everything
// below it plus the smallest ids among the ties, and a small bitonic sort.
#include "fspann_internal.cuh"

namespace fsp {

constexpr int GT_TQ = 64, GT_TB = 64, GT_DK = 32, GT_THREADS = 256;

__global__ void __launch_bounds__(GT_THREADS) gt_dist_kernel(const float *__restrict__ base, int64_t N, int dim, const float *__restrict__ queries,
                                                             int Qc, double *__restrict__ dist /* [Qc][N] */) {
    __shared__ __align__(16) float qs[GT_DK][GT_TQ];
    __shared__ __align__(16) float bs[GT_DK][GT_TB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t b0 = (int64_t)blockIdx.x * GT_TB;
    const int q0 = blockIdx.y * GT_TQ;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    for (int d0 = 0; d0 < dim; d0 += GT_DK) {
        __syncthreads();
        // consecutive lanes take consecutive ROWS, so the transposed shared-memory stores are conflict free; each lane reads 16
        // bytes of its row (the neighbouring 16 bytes of the sector are picked up by the next column group from L1)
        if ((dim & 3) == 0) {
            for (int idx = tid; idx < GT_TQ * (GT_DK / 4); idx += GT_THREADS) {
                const int r = idx & (GT_TQ - 1), c = idx / GT_TQ;
                const bool in = d0 + 4 * c < dim;
                float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), bv = qv;
                if (in && q0 + r < Qc) qv = __ldg(reinterpret_cast<const float4 *>(queries + (size_t)(q0 + r) * dim + d0 + 4 * c));
                if (in && b0 + r < N) bv = __ldg(reinterpret_cast<const float4 *>(base + (size_t)(b0 + r) * dim + d0 + 4 * c));
                qs[4 * c + 0][r] = qv.x; qs[4 * c + 1][r] = qv.y; qs[4 * c + 2][r] = qv.z; qs[4 * c + 3][r] = qv.w;
                bs[4 * c + 0][r] = bv.x; bs[4 * c + 1][r] = bv.y; bs[4 * c + 2][r] = bv.z; bs[4 * c + 3][r] = bv.w;
            }
        } else {
            for (int idx = tid; idx < GT_TQ * GT_DK; idx += GT_THREADS) {
                const int r = idx & (GT_TQ - 1), dd = idx / GT_TQ;
                const bool in = d0 + dd < dim;
                qs[dd][r] = (in && q0 + r < Qc) ? queries[(size_t)(q0 + r) * dim + d0 + dd] : 0.f;
                bs[dd][r] = (in && b0 + r < N) ? base[(size_t)(b0 + r) * dim + d0 + dd] : 0.f;
            }
        }
        __syncthreads();
        const int nd = min(GT_DK, dim - d0);
        for (int dd = 0; dd < nd; dd++) {
            const float4 qv = *reinterpret_cast<const float4 *>(&qs[dd][4 * ty]);
            const float4 bv = *reinterpret_cast<const float4 *>(&bs[dd][4 * tx]);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double d = (double)__fsub_rn(qa[a], ba[b]);                  // float subtraction, then widened
                    acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(d, d));
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int q = q0 + 4 * ty + a;
        if (q >= Qc) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t bi = b0 + 4 * tx + b;
            if (bi < N) dist[(size_t)q * N + bi] = acc[a][b];
        }
    }
}

constexpr int GS_THREADS = 1024, GS_BINS = 2048, GS_TIES = 4096, GS_KMAX = 1024;

__device__ __forceinline__ void gs_hist_add(int32_t *hist, bool active, int bin, int lane) {
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned peers = __match_any_sync(act, bin);
    if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

// One CTA per query row of the distance chunk.  Keys = IEEE bits of the (non-negative) FP64 sums: ordered like the values.
__global__ void __launch_bounds__(GS_THREADS) gt_select_kernel(const double *__restrict__ dist, int64_t N, int Qc, int K, int32_t *__restrict__ out_ids,
                                                               double *__restrict__ out_d2, int32_t *__restrict__ tie_overflow) {
    __shared__ int32_t hist[GS_BINS];
    __shared__ unsigned long long s_lo, s_hi, s_red[2][GS_THREADS / 32];
    __shared__ int32_t s_less, s_bin, s_n_out, s_n_tie, s_scan[GS_THREADS / 32];
    __shared__ unsigned long long okey[GS_KMAX];
    __shared__ int32_t oid[GS_KMAX], tie[GS_TIES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = blockIdx.x; q < Qc; q += gridDim.x) {
        const unsigned long long *row = reinterpret_cast<const unsigned long long *>(dist + (size_t)q * N);
        __syncthreads();
        // ---- key range ----
        unsigned long long mn = ~0ull, mx = 0ull;
        for (int64_t i = tid; i < N; i += GS_THREADS) { const unsigned long long k = row[i]; mn = min(mn, k); mx = max(mx, k); }
#pragma unroll
        for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long a = ~0ull, b = 0ull;
            for (int w = 0; w < GS_THREADS / 32; w++) { a = min(a, s_red[0][w]); b = max(b, s_red[1][w]); }
            s_lo = a; s_hi = b; s_less = 0;
        }
        __syncthreads();
        // ---- exact K-th smallest key T: shrink [lo, hi] with linear histograms until it is a single value ----
        for (;;) {
            const unsigned long long lo = s_lo, hi = s_hi;
            if (lo == hi) break;
            int sh = 0; while (((hi - lo) >> sh) >= (unsigned long long)GS_BINS) sh++;
            for (int i = tid; i < GS_BINS; i += GS_THREADS) hist[i] = 0;
            __syncthreads();
            for (int64_t i0 = 0; i0 < N; i0 += GS_THREADS) {
                const int64_t i = i0 + tid;
                bool in = false; int bin = 0;
                if (i < N) { const unsigned long long k = row[i]; in = k >= lo && k <= hi; bin = in ? (int)((k - lo) >> sh) : 0; }
                gs_hist_add(hist, in, bin, lane);
            }
            __syncthreads();
            if (tid == 0) {
                int c = s_less, b = 0;
                for (; b < GS_BINS - 1; b++) { if (c + hist[b] >= K) break; c += hist[b]; }
                s_less = c; s_bin = b;
            }
            __syncthreads();
            const int b = s_bin;
            __syncthreads();
            if (tid == 0) {
                const unsigned long long nlo = lo + ((unsigned long long)b << sh);
                unsigned long long nhi = nlo + ((1ull << sh) - 1ull);
                if (nhi > hi) nhi = hi;
                s_lo = nlo; s_hi = nhi;
            }
            __syncthreads();
        }
        const unsigned long long T = s_lo;
        const int c_less = s_less, need = K - c_less;          // c_less keys < T (< K), `need` of the keys == T with the smallest ids
        if (tid == 0) { s_n_out = 0; s_n_tie = 0; }
        __syncthreads();
        // ---- collect: everything below T, and the ids of the ties ----
        for (int64_t i = tid; i < N; i += GS_THREADS) {
            const unsigned long long k = row[i];
            if (k < T) { const int a = atomicAdd(&s_n_out, 1); okey[a] = k; oid[a] = (int32_t)i; }
            else if (k == T) { const int a = atomicAdd(&s_n_tie, 1); if (a < GS_TIES) tie[a] = (int32_t)i; }
        }
        __syncthreads();
        const int n_tie = s_n_tie;
        if (n_tie > GS_TIES) {
            // more exact ties than the list holds: ordered pass, ids ascending, first `need` of them
            if (tid == 0) { *tie_overflow = 1; s_n_tie = 0; }
            __syncthreads();
            for (int64_t i0 = 0; i0 < N && s_n_tie < need; i0 += GS_THREADS) {
                const int64_t i = i0 + tid;
                const bool f = i < N && row[i] == T;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if (lane == 0) s_scan[warp] = __popc(bal);
                __syncthreads();
                int basew = s_n_tie;
                for (int w = 0; w < warp; w++) basew += s_scan[w];
                int tot = 0;
                for (int w = 0; w < GS_THREADS / 32; w++) tot += s_scan[w];
                const int r = basew + __popc(bal & ((1u << lane) - 1u));
                if (f && r < need) tie[r] = (int32_t)i;
                __syncthreads();
                if (tid == 0) s_n_tie += tot;
                __syncthreads();
            }
        } else {
            // sort the tie ids ascending (bitonic over the next power of two)
            int n2 = 1; while (n2 < n_tie) n2 <<= 1;
            for (int i = n_tie + tid; i < n2; i += GS_THREADS) tie[i] = 0x7fffffff;
            __syncthreads();
            for (int k2 = 2; k2 <= n2; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < n2; i += GS_THREADS) {
                        const int x = i ^ j;
                        if (x > i) { const int a = tie[i], b = tie[x]; if ((a > b) == ((i & k2) == 0)) { tie[i] = b; tie[x] = a; } }
                    }
                    __syncthreads();
                }
        }
        for (int i = tid; i < need; i += GS_THREADS) { okey[c_less + i] = T; oid[c_less + i] = tie[i]; }
        __syncthreads();
        // ---- final order by (key, id): bitonic over K entries (the ties are already last and in id order) ----
        int n2 = 1; while (n2 < K) n2 <<= 1;
        for (int i = K + tid; i < n2; i += GS_THREADS) { okey[i] = ~0ull; oid[i] = 0x7fffffff; }
        __syncthreads();
        for (int k2 = 2; k2 <= n2; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n2; i += GS_THREADS) {
                    const int x = i ^ j;
                    if (x > i) {
                        const unsigned long long a = okey[i], b = okey[x]; const int ia = oid[i], ib = oid[x];
                        const bool gt = a > b || (a == b && ia > ib);
                        if (gt == ((i & k2) == 0)) { okey[i] = b; okey[x] = a; oid[i] = ib; oid[x] = ia; }
                    }
                }
                __syncthreads();
            }
        for (int i = tid; i < K; i += GS_THREADS) {
            out_ids[(size_t)q * K + i] = oid[i];
            if (out_d2) out_d2[(size_t)q * K + i] = __longlong_as_double((long long)okey[i]);
        }
    }
}

// recall@K per query (FSA:785-794): hits among the first min(K, n_ret) results that appear in gt[0..K), divided by K.
__global__ void recall_kernel(int Q, int K, const int32_t *__restrict__ gt, int gt_stride, const int32_t *__restrict__ res, int res_stride,
                              const int32_t *__restrict__ n_ret, double *__restrict__ recall) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int n = min(K, n_ret ? n_ret[q] : K);
    int hits = 0;
    for (int i = 0; i < n; i++) {
        const int32_t id = res[(size_t)q * res_stride + i];
        bool in = false;
        for (int j = 0; j < K; j++) in |= gt[(size_t)q * gt_stride + j] == id;
        hits += in;
    }
    recall[q] = (double)hits / (double)K;
}

int launch_gt_chunk(cudaStream_t s, const float *base, int64_t N, int dim, const float *queries, int Qc, int K, double *dist, int32_t *out_ids,
                    double *out_d2, int32_t *tie_overflow, int sm_count) {
    if (Qc <= 0) return 0;
    dim3 grid((unsigned)((N + GT_TB - 1) / GT_TB), (unsigned)((Qc + GT_TQ - 1) / GT_TQ));
    gt_dist_kernel<<<grid, GT_THREADS, 0, s>>>(base, N, dim, queries, Qc, dist);
    gt_select_kernel<<<std::min(Qc, sm_count), GS_THREADS, 0, s>>>(dist, N, Qc, K, out_ids, out_d2, tie_overflow);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}
int launch_recall(cudaStream_t s, int Q, int K, const int32_t *gt, int gt_stride, const int32_t *res, int res_stride, const int32_t *n_ret,
                  double *recall) {
    if (Q <= 0) return 0;
    recall_kernel<<<(Q + 127) / 128, 128, 0, s>>>(Q, K, gt, gt_stride, res, res_stride, n_ret, recall);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int gt_max_k() { return GS_KMAX; }

}  // namespace fsp
