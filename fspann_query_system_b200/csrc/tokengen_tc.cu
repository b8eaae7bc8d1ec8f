// tokengen_tc.cu -- stage 1 on the 5th-generation tensor cores: the projection v . alpha_j of TokenGen's pre-filter as a tcgen05.mma
// contraction with the accumulator in TMEM (sm_100a).
//
// Same contract as tokengen_fast_kernel (tokengen.cu): a reduced-precision dot product with an error bound decides almost every
// h_j = floor((v . alpha_j + r_j) / omega_j) (Coding.H, index/.../paper/Coding.java:250-258); the projections that fall within the bound
// of a quantisation boundary go to the re-check list and tokengen_recheck_kernel recomputes them with the reference's exact sequential
// FP64 arithmetic, so the codes are bit-identical to the exact kernel.  Only the contraction moves: FP32 FMA pipe -> tensor pipe.
//
// Arithmetic.  BF16 x BF16 products are exact in FP32 (8 x 8 significand bits), so each FP32 operand is split into three BF16 pieces
// x = x1 + x2 + x3 (exact: 24 = 8 + 8 + 8 bits) and the products of piece pairs (a, b) with a + b <= 2 are accumulated in FP32 in TMEM:
//   v has one non-zero piece (integer-valued SIFT / .bvecs data are exact in BF16): 3 pair products, nothing dropped;
//   otherwise: 6 pair products, the dropped ones are < 2^-23 |v_i alpha_ji| each.
// Error model of the bound (per projection, S = ||v||_2 ||alpha_j||_2 >= sum |v_i alpha_ji|): inputs rounded to FP32 2^-23 S, dropped
// pieces 2^-23 S, tensor-core FP32 accumulation <= (#MMA x 17) x 2^-24 S assuming every MMA of K = 16 aligns its 17 addends to the
// largest and truncates below 2^-24 of it (24 MMAs for one-piece v: 2.4e-5 S; 48 MMAs: 4.9e-5 S).  The kernel uses C_tc = 6e-5 / 1.2e-4
// (2.4x that sum; published measurements of NVIDIA tensor cores find MORE than 24 bits kept, so the model itself is pessimistic).  The
// hardware's internal accumulation is not documented, so the bound is an ASSUMPTION that tests/ verify empirically: codes over the full
// 1M-vector base set (1.5 G projections, millions of them within 1e-3 of a boundary) equal the exact kernel's.
//
// Structure (one CTA per SM, persistent over tiles of 128 vectors; 21 warps):
//   all warps    stage the A tile (FP64 -> FP32 -> 3 BF16 pieces, canonical K-major no-swizzle UMMA layout);
//   warps 0-19   epilogue: warp w reads TMEM lanes 32 * (w % 4) .. +31 (= 32 vectors) with tcgen05.ld and takes the (t,d) groups
//                w / 4, w / 4 + 5, ... of the tile: decision, bit planes, code word store, undecided projections -> re-check list;
//   warp 20      lane 0 streams the pre-split alpha tiles from HBM/L2 with cp.async.bulk (TMA) into shared memory, issues the
//                tcgen05.mma sequence (M = 128, N <= 128, K = 16 per instruction) into one of TWO accumulator buffers in TMEM and
//                commits it to an mbarrier: tile t+1's TMA and MMAs run under tile t's epilogue.
// The kernel is bound by the epilogue's integer / FP32 work (one decision per projection), not by the tensor pipe.
#include <cuda_bf16.h>

#include "fspann_internal.cuh"

namespace fsp {

constexpr int TC_M = 128;          // vectors per tile = TMEM lanes
constexpr int TC_N = 128;          // accumulator columns (projections) per tile
constexpr int TC_EPI_SLOTS = 5;                        // epilogue warps per TMEM lane quarter: warp w reads lanes 32 * (w % 4), takes groups w / 4, + 5, ...
constexpr int TC_EPI_WARPS = 4 * TC_EPI_SLOTS;
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 1);    // + the TMA / MMA-issue warp
constexpr int TC_ISSUER = 32 * TC_EPI_WARPS;           // its lane 0
constexpr int TC_CHUNK_BYTES = 16;                    // one K chunk of a core matrix: 8 BF16
constexpr int TC_SBO = 8 * TC_CHUNK_BYTES;            // 8-row group stride (bytes): a core matrix is 8 rows x 16 bytes, contiguous
constexpr int TC_LBO = TC_M * TC_CHUNK_BYTES;         // K-chunk stride (bytes): all 128 rows of one 16-byte K chunk are contiguous

struct TcPlan {
    int kpad;              // dim rounded up to 16 (BF16 MMA K)
    int groups_per_tile;   // (t,d) groups per alpha tile = floor(TC_N / m)
    int n_tiles;           // ceil(TD / groups_per_tile)
    int piece_bytes;       // one BF16 piece of one tile (A or B): TC_M * kpad * 2
    size_t smem;
};

bool tokengen_tc_plan(const RoutingView &rv, TcPlan &pl) {
    if (rv.W != 1 || rv.m > 32 || rv.m < 1 || rv.lambda > 4 || rv.dim > 128 || rv.TD >= 65536) return false;
    pl.kpad = (rv.dim + 15) / 16 * 16;
    pl.groups_per_tile = TC_N / rv.m;
    pl.n_tiles = (rv.TD + pl.groups_per_tile - 1) / pl.groups_per_tile;
    pl.piece_bytes = TC_M * pl.kpad * 2;
    const size_t consts = sizeof(float4) * (size_t)rv.TD * rv.m;
    pl.smem = 6 * (size_t)pl.piece_bytes + consts + sizeof(float) * TC_M + 256;
    return pl.smem <= (size_t)kMaxDynSmem - 1024;
}
size_t tokengen_tc_alpha_bytes(const RoutingView &rv) {
    TcPlan pl;
    return tokengen_tc_plan(rv, pl) ? (size_t)pl.n_tiles * 3 * pl.piece_bytes : 0;
}

// byte offset of element (row, k) inside one canonical K-major, no-swizzle piece (UMMA SmemDescriptor with LBO = TC_LBO, SBO = TC_SBO)
__host__ __device__ inline int tc_offset(int row, int k) { return (k >> 3) * TC_LBO + (row >> 3) * TC_SBO + (row & 7) * TC_CHUNK_BYTES + (k & 7) * 2; }

__device__ __forceinline__ void bf16_split3(float x, uint16_t &b1, uint16_t &b2, uint16_t &b3) {
    const __nv_bfloat16 h1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h1);                 // exact
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(h2);                // exact, <= 8 significant bits left
    const __nv_bfloat16 h3 = __float2bfloat16_rn(r2);
    b1 = __bfloat16_as_ushort(h1); b2 = __bfloat16_as_ushort(h2); b3 = __bfloat16_as_ushort(h3);
}

// alpha FP64 [TD*m][dim] -> out [n_tiles][3][canonical piece]: tile t holds the projections of groups t*G .. t*G+G-1 as its rows
__global__ void alpha_tc_prepare_kernel(const double *__restrict__ alpha, int TD, int m, int dim, int kpad, int G, int n_tiles,
                                        uint16_t *__restrict__ out) {
    const int64_t total = (int64_t)n_tiles * TC_N * kpad;
    const int piece_elems = TC_M * kpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / ((int64_t)TC_N * kpad));
        const int rem = (int)(i - (int64_t)t * TC_N * kpad);
        const int row = rem / kpad, k = rem - row * kpad;
        const int64_t proj = (int64_t)t * G * m + row;            // global projection index (group-major)
        float x = 0.f;
        if (row < G * m && proj < (int64_t)TD * m && k < dim) x = (float)alpha[proj * dim + k];
        uint16_t b1, b2, b3;
        bf16_split3(x, b1, b2, b3);
        const int off = tc_offset(row, k) >> 1;
        uint16_t *base = out + (size_t)t * 3 * piece_elems;
        base[off] = b1; base[piece_elems + off] = b2; base[2 * piece_elems + off] = b3;
    }
}
int launch_alpha_tc_prepare(cudaStream_t s, const RoutingView &rv, uint16_t *out) {
    TcPlan pl;
    if (!tokengen_tc_plan(rv, pl)) return 0;
    const int64_t total = (int64_t)pl.n_tiles * TC_N * pl.kpad;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)cur_sm_count() * 8);
    alpha_tc_prepare_kernel<<<grid, 256, 0, s>>>(rv.alpha, rv.TD, rv.m, rv.dim, pl.kpad, pl.groups_per_tile, pl.n_tiles, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TCWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TCDONE_%=;\n\t"
        "bra TCWAIT_%=;\n\t"
        "TCDONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// work[0] = number of listed (q, g, j); work[1] = overflow flag; list entries = q << 24 | g << 8 | j (as tokengen_fast_kernel)
__global__ void __launch_bounds__(TC_THREADS, 1) tokengen_tc_kernel(RoutingView rv, TcPlan pl, int64_t Q, const double *__restrict__ queries,
                                                                    const uint16_t *__restrict__ alpha_tc, uint64_t *__restrict__ codes, float cbound,
                                                                    int32_t *__restrict__ work, unsigned long long *__restrict__ list, int64_t list_cap, int split) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar_tma, s_bar_mma;
    __shared__ uint32_t s_tmem;
    __shared__ int s_multi;
    const int dim = rv.dim, m = rv.m, lambda = rv.lambda, TD = rv.TD, kpad = pl.kpad, G = pl.groups_per_tile;
    const int n_proj = TD * m;
    unsigned char *sA = tc_smem;                                      // 3 pieces
    unsigned char *sB = sA + 3 * (size_t)pl.piece_bytes;              // 3 pieces
    float4 *cst = reinterpret_cast<float4 *>(sB + 3 * (size_t)pl.piece_bytes);   // [n_proj] (r_j, 1 / omega_j, C_tc * ||alpha_j|| / omega_j, -)
    float *nv2 = reinterpret_cast<float *>(cst + n_proj);             // [TC_M] sum of squares of the tile's rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == TC_ISSUER) { tc_mbar_init(&s_bar_tma, 1); tc_mbar_init(&s_bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {                                                  // two accumulator buffers of TC_N columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(2 * TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < n_proj; i += TC_THREADS) {
        const float io = (float)(1.0 / rv.omega[i]);
        cst[i] = make_float4((float)rv.r[i], io, rv.alpha_norm[i] * io, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_tmem;
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    const int k_steps = kpad / 16;
    const int64_t a_tiles = (Q + TC_M - 1) / TC_M;
    uint32_t tma_phase = 0, mma_phase = 0;
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_M >> 4) << 24);

    // TMA of alpha tile `t` into the (single) B buffer; MMA sequence of tile t into accumulator buffer `buf` (issuer thread only)
    auto load_b = [&](int t) {
        tc_mbar_expect_tx(&s_bar_tma, 3u * (uint32_t)pl.piece_bytes);
        for (int pc = 0; pc < 3; pc++)
            tc_tma_load(sB + (size_t)pc * pl.piece_bytes, alpha_tc + ((size_t)t * 3 + pc) * (pl.piece_bytes / 2), (uint32_t)pl.piece_bytes, &s_bar_tma);
    };
    auto issue_mma = [&](int t, int buf, int n_a) {
        tc_mbar_wait(&s_bar_tma, tma_phase);                           // the alpha tile has landed
        tma_phase ^= 1u;
        tc_fence_after();
        const int ncols = min(G, TD - t * G) * m;
        const uint32_t idesc = idesc_base | ((uint32_t)(((ncols + 15) / 16 * 16) >> 3) << 17);
        uint32_t acc = 0;
        for (int a = 0; a < n_a; a++)                                  // piece pairs (a, b) with a + b <= 2; a single non-zero piece of v: a = 0 only
            for (int b = 0; a + b <= 2; b++)
                for (int ks = 0; ks < k_steps; ks++) {
                    const uint64_t ad = tc_smem_desc(sA_u + (uint32_t)a * pl.piece_bytes + (uint32_t)ks * 2u * TC_LBO);
                    const uint64_t bd = tc_smem_desc(sB_u + (uint32_t)b * pl.piece_bytes + (uint32_t)ks * 2u * TC_LBO);
                    tc_mma_bf16(taddr + (uint32_t)buf * TC_N, ad, bd, idesc, acc);
                    acc = 1;
                }
        tc_commit(&s_bar_mma);                                         // arrives when every MMA above has completed
    };

    // Few A tiles (a small batch): `split` CTAs share one A tile and take the alpha tiles t0, t0 + split, ... so the batch still spreads over the SMs
    const int t0 = (int)(blockIdx.x % (unsigned)split), a_lane = (int)(blockIdx.x / (unsigned)split), a_step = (int)(gridDim.x / (unsigned)split);
    if (tid == TC_ISSUER && a_lane < a_tiles && t0 < pl.n_tiles) load_b(t0);
    int buf = 0;
    for (int64_t at = a_lane; at < a_tiles && t0 < pl.n_tiles; at += a_step) {
        const int64_t q0 = at * TC_M;
        // ---- stage the A tile: FP64 -> FP32 -> three BF16 pieces in the canonical layout; row norms; "more than one piece" flag ----
        if (tid < TC_M) nv2[tid] = 0.f;
        if (tid == 0) s_multi = 0;
        __syncthreads();
        {
            const int chunks = kpad >> 3;
            int multi = 0;
            for (int task = tid; task < TC_M * chunks; task += TC_THREADS) {
                const int kc = task / TC_M, row = task - kc * TC_M;
                const int64_t q = q0 + row;
                uint16_t p1[8], p2[8], p3[8];
                float ss = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int k = kc * 8 + e;
                    const float x = (q < Q && k < dim) ? (float)queries[q * dim + k] : 0.f;
                    bf16_split3(x, p1[e], p2[e], p3[e]);
                    multi |= (p2[e] | p3[e]) & 0x7fff;               // a piece that is +-0 does not count
                    ss = fmaf(x, x, ss);
                }
                const int off = kc * TC_LBO + (row >> 3) * TC_SBO + (row & 7) * TC_CHUNK_BYTES;
                *reinterpret_cast<uint4 *>(sA + off) = make_uint4(p1[0] | (p1[1] << 16), p1[2] | (p1[3] << 16), p1[4] | (p1[5] << 16), p1[6] | (p1[7] << 16));
                *reinterpret_cast<uint4 *>(sA + pl.piece_bytes + off) = make_uint4(p2[0] | (p2[1] << 16), p2[2] | (p2[3] << 16), p2[4] | (p2[5] << 16), p2[6] | (p2[7] << 16));
                *reinterpret_cast<uint4 *>(sA + 2 * pl.piece_bytes + off) = make_uint4(p3[0] | (p3[1] << 16), p3[2] | (p3[3] << 16), p3[4] | (p3[5] << 16), p3[6] | (p3[7] << 16));
                atomicAdd(&nv2[row], ss);
            }
            if (multi) s_multi = 1;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A -> visible to the tensor core's async proxy
        __syncthreads();
        const int n_a = s_multi ? 3 : 1;
        const int quarter = warp & 3, slot = warp >> 2;
        const int row = quarter * 32 + lane;                          // TMEM lane = vector of the tile handled by this thread (epilogue warps)
        // C_tc * (upper bound of ||v||_2; float atomics: any order).  C_tc scales with the number of MMAs accumulated per projection.
        const float nv = warp < TC_EPI_WARPS ? __fmul_ru(__fmul_ru(__fsqrt_ru(nv2[row]), 1.0002f), n_a == 1 ? cbound : 2.0f * cbound) : 0.f;
        if (tid == TC_ISSUER) issue_mma(t0, buf, n_a);

        for (int t = t0; t < pl.n_tiles; t += split) {
            const int g_lo = t * G, g_n = min(G, TD - g_lo);
            tc_mbar_wait(&s_bar_mma, mma_phase);                       // accumulator `buf` holds tile t; the B buffer is free again
            mma_phase ^= 1u;
            tc_fence_after();
            if (tid == TC_ISSUER) {
                // next alpha tile -> B buffer, and (same A tile) its MMAs into the OTHER accumulator buffer, all under this tile's epilogue
                if (t + split < pl.n_tiles) { load_b(t + split); issue_mma(t + split, buf ^ 1, n_a); }
                else if (at + a_step < a_tiles) load_b(t0);            // first tile of the next A tile: only the TMA (A is not staged yet)
            }
            // ---- epilogue: warp = (TMEM lane quarter, group slot); thread = vector; one (t,d) group of m projections at a time ----
            if (warp < TC_EPI_WARPS) {
                const int64_t q = q0 + row;
                for (int gl = slot; gl < g_n; gl += TC_EPI_SLOTS) {
                    const int g = g_lo + gl;
                    const float4 *cg = cst + (size_t)g * m;
                    uint32_t plane[4] = {0u, 0u, 0u, 0u}, und = 0u;
                    for (int c8 = 0; c8 < m; c8 += 8) {
                        uint32_t v[8];
                        tc_ld8(taddr + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_N + gl * m + c8), v);
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int j = c8 + u;
                            if (j < m) {
                                const float4 c = cg[j];
                                const float tt = (__uint_as_float(v[u]) + c.x) * c.y;
                                const float f = floorf(tt);
                                const float slack = fmaf(c.z, nv, 2e-6f * (1.0f + fabsf(tt)));
                                const bool decided = (tt - f > slack) && (f + 1.0f - tt > slack) && fabsf(tt) < 4194304.0f;   // false for NaN / Inf too
                                const uint32_t hj = decided ? (uint32_t)(int32_t)f : 0u;           // (h ^ 0x80000000 only flips bit 31, Coding:293)
#pragma unroll
                                for (int ib = 0; ib < 4; ib++) plane[ib] |= ((hj >> ib) & 1u) << j;
                                und |= (decided ? 0u : 1u) << j;
                            }
                        }
                    }
                    unsigned long long code = 0ull;                   // MSB-first bit planes (Coding:291-299): plane ib at bit (lambda-1-ib)*m
#pragma unroll
                    for (int ib = 0; ib < 4; ib++) if (ib < lambda) code |= (unsigned long long)plane[ib] << ((lambda - 1 - ib) * m);
                    if (q >= Q) und = 0u;
                    else codes[q * TD + g] = code;
                    // undecided projections of the warp's 32 vectors -> re-check list: one global atomic per warp and group
                    const int cnt = __popc(und);
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {                                      // warp-uniform
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&work[0], total);
                        base = __shfl_sync(0xffffffffu, base, 0) + incl - cnt;
                        while (und) {
                            const int j = __ffs(und) - 1;
                            und &= und - 1u;
                            if (base < list_cap) list[base] = ((unsigned long long)q << 24) | ((unsigned long long)g << 8) | (unsigned long long)j;
                            else work[1] = 1;
                            base++;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();                                           // accumulator `buf` drained; after the last tile also: A free for the next staging
            tc_fence_after();
            buf ^= 1;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(2 * TC_N) : "memory");
}

int configure_tokengen_tc_kernels() { return opt_in_smem(tokengen_tc_kernel); }

// returns kernels launched (1) or -1; 0 when the shape is not supported (caller falls back to the FP32 pre-filter)
int launch_tokengen_tc(cudaStream_t s, const RoutingView &rv, int64_t Q, const double *queries, const uint16_t *alpha_tc, uint64_t *codes, int32_t *work,
                       unsigned long long *list, int64_t list_cap, int sm_count) {
    TcPlan pl;
    if (!alpha_tc || !tokengen_tc_plan(rv, pl)) return 0;
    const int64_t a_tiles = (Q + TC_M - 1) / TC_M;
    int split = 1;                                    // CTAs per A tile
    if (a_tiles < sm_count) split = (int)std::min<int64_t>(pl.n_tiles, sm_count / a_tiles);
    const int grid = (int)std::min<int64_t>(a_tiles * split, (int64_t)(sm_count / split) * split);
    const float cbound = 6e-5f;                       // one-piece v (24 MMAs per projection); doubled in the kernel for three-piece v (48 MMAs)
    tokengen_tc_kernel<<<grid, TC_THREADS, pl.smem, s>>>(rv, pl, Q, queries, alpha_tc, codes, cbound, work, list, list_cap, split);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace fsp
