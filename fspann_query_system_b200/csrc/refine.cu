// refine.cu -- stage 3: bounded trusted refinement.
//
// Replaces QueryServiceImpl.search stage B + C (query/.../service/QueryServiceImpl.java:238-322):
//   loadPointIfActive (PIS:717-724) -> keyService.getVersion(ep.keyVersion) (KRS:82-88) ->
//   AesGcmCryptoService.decryptFromPoint (crypto/.../AesGcmCryptoService.java:126-166, AAD EP:80-83) ->
//   isValid (QSI:407-413) -> l2 (QSI:364-372) -> stable sort by distance, first K (QSI:298-316).
//
// B200 design: the batch is processed RECORD-major.  A 10k-query batch at B=1024 names ~10M (query, candidate)
// pairs but at most N distinct records, and AES-256-GCM on the SM (no AES/CLMUL instructions) is ALU/LDS bound,
// far below the HBM rate.  So pairs are grouped by record id (count -> scan -> fill), each distinct record is
// authenticated and decrypted ONCE, its plaintext lives only in shared memory, and every pair that selected it is
// scored from there (exact sequential FP64, bit-identical to the Java loop).  Per-query top-k runs afterwards on
// the scalar distances.  Plaintext never reaches global memory.
#include <algorithm>

#include "fspann_internal.cuh"

namespace fsp {

// ------------------------------------------------------------------------------------------------------------------
// grouping: pairs (q, rank) -> per-record lists
// ------------------------------------------------------------------------------------------------------------------
// record bytes are read-only inside every kernel that loads them through this helper (Migrate rewrites a record only after
// its warp has loaded everything it needs from it), so the non-coherent path is safe
template <class T>
__device__ __forceinline__ T ld_rec(const T *p) { return __ldg(p); }

__device__ __forceinline__ bool is_deleted(const StoreView &sv, int32_t id) {
    return sv.deleted && id < sv.n_deleted && sv.deleted[id];
}

// q = pair / stride for pair < 2^31 with the host-precomputed multiplier (division by an invariant: one IMAD.HI, one add, one shift)
__device__ __forceinline__ uint32_t pair_query(const RefineParams &p, uint32_t pair) { return (__umulhi(pair, p.div_magic) + pair) >> p.div_shift; }

__global__ void refine_count_kernel(StoreView sv, RefineParams p) {
    const uint32_t total = (uint32_t)(p.Q * (int64_t)p.stride);          // < 2^31 (checked by the caller)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t q = pair_query(p, i);
        const int r = (int)(i - q * (uint32_t)p.stride);
        uint8_t v = 0xff;                               // 0xff = slot not part of the candidate list
        if (r < p.n_cand[q] && (!p.qfinite || p.qfinite[q])) {   // QSI:137: a query with NaN/Inf returns empty and touches nothing
            const int32_t id = p.cand_ids[i];
            if (id < 0 || id >= sv.n_global || is_deleted(sv, id)) v = FSPANN_V_NOT_FOUND;     // PIS:717-724
            else if (id < sv.id_base || id >= sv.id_base + sv.N) v = 0xfd;                      // 0xfd = lives in another shard
            else { atomicAdd(&p.cnt[id - sv.id_base], 1); v = 0xfe; }                           // 0xfe = pending
        }
        p.verdict[i] = v;
    }
}

// Two-level exclusive scan of (cnt, cnt>0) over N+1 entries.  SCAN_ITEMS per block.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_ITEMS = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ int2 block_excl_scan2(int2 v, int2 *total) {
    __shared__ int2 s_w[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int2 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, x.x, o), yb = __shfl_up_sync(0xffffffffu, x.y, o);
        if (lane >= o) { x.x += ya; x.y += yb; }
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    int2 base = make_int2(0, 0), tot = make_int2(0, 0);
    for (int w = 0; w < SCAN_THREADS / 32; w++) { if (w < warp) { base.x += s_w[w].x; base.y += s_w[w].y; } tot.x += s_w[w].x; tot.y += s_w[w].y; }
    __syncthreads();
    *total = tot;
    return make_int2(base.x + x.x - v.x, base.y + x.y - v.y);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_partial_kernel(const int32_t *__restrict__ cnt, int64_t n, int32_t *__restrict__ block_sums) {
    const int64_t b0 = (int64_t)blockIdx.x * SCAN_ITEMS;
    int2 acc = make_int2(0, 0);
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const int64_t i = b0 + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) { const int c = cnt[i]; acc.x += c; acc.y += c > 0; }
    }
    int2 tot;
    block_excl_scan2(acc, &tot);
    if (threadIdx.x == 0) { block_sums[2 * blockIdx.x] = tot.x; block_sums[2 * blockIdx.x + 1] = tot.y; }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int32_t *block_sums, int nblocks, int32_t *totals, int32_t *uoff) {
    __shared__ int2 carry;
    if (threadIdx.x == 0) carry = make_int2(0, 0);
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
        const int i = b0 + threadIdx.x;
        int2 v = i < nblocks ? make_int2(block_sums[2 * i], block_sums[2 * i + 1]) : make_int2(0, 0);
        int2 tot;
        int2 ex = block_excl_scan2(v, &tot);
        const int2 c = carry;
        if (i < nblocks) { block_sums[2 * i] = ex.x + c.x; block_sums[2 * i + 1] = ex.y + c.y; }
        __syncthreads();
        if (threadIdx.x == 0) { carry.x = c.x + tot.x; carry.y = c.y + tot.y; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = carry.x; totals[1] = carry.y; totals[2] = 0; totals[3] = 0; uoff[carry.y] = carry.x; }   // sentinel: end of the last record's pairs
}
// Rewrites cnt[] in place to exclusive offsets, writes flag prefix, the unique-id list and zeroes fill[].
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int32_t *__restrict__ cnt, int64_t n, const int32_t *__restrict__ block_sums,
                                                                  int32_t *__restrict__ uniq, int32_t *__restrict__ uoff, int32_t *__restrict__ fill) {
    const int64_t b0 = (int64_t)blockIdx.x * SCAN_ITEMS;
    // thread owns SCAN_PER_THREAD consecutive items so the in-block order is the id order
    int c[SCAN_PER_THREAD];
    int2 acc = make_int2(0, 0);
    const int64_t t0 = b0 + (int64_t)threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) { const int64_t i = t0 + k; c[k] = i < n ? cnt[i] : 0; acc.x += c[k]; acc.y += c[k] > 0; }
    int2 tot;
    int2 ex = block_excl_scan2(acc, &tot);
    ex.x += block_sums[2 * blockIdx.x]; ex.y += block_sums[2 * blockIdx.x + 1];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const int64_t i = t0 + k;
        if (i < n) {
            cnt[i] = ex.x;
            fill[i] = 0;
            if (c[k] > 0) { uniq[ex.y] = (int32_t)i; uoff[ex.y] = ex.x; }   // distinct record #ex.y and where its pairs start
            ex.x += c[k]; ex.y += c[k] > 0;
        }
    }
}
__global__ void refine_fill_kernel(StoreView sv, RefineParams p) {
    const int64_t total = p.Q * (int64_t)p.stride;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    // four slots per thread and iteration: the returning atomics (an L2 round trip each) of the four are in flight together
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * step) {
        int32_t id[4]; int base[4], off[4]; bool live[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int64_t i = i0 + u * step; live[u] = i < total && p.verdict[i] == 0xfe; }
#pragma unroll
        for (int u = 0; u < 4; u++) id[u] = live[u] ? (int32_t)(p.cand_ids[i0 + u * step] - sv.id_base) : 0;
#pragma unroll
        for (int u = 0; u < 4; u++) { base[u] = live[u] ? p.cnt[id[u]] : 0; off[u] = live[u] ? atomicAdd(&p.fill[id[u]], 1) : 0; }
#pragma unroll
        for (int u = 0; u < 4; u++) if (live[u]) p.pairs[base[u] + off[u]] = (uint32_t)(i0 + u * step);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Small batches (Q * stride <= GS_MAX pairs: a handful of queries): ONE CTA classifies the candidate slots, sorts the pending
// (record, pair) keys in shared memory and emits the same grouping (uniq / uoff / pairs / totals) -- instead of a memset, a count,
// three scans over the N-entry arrays and a fill (six launches touching ~16 MB at N = 1 M).
// ------------------------------------------------------------------------------------------------------------------
constexpr int GS_THREADS = 1024;
constexpr int GS_MAX = 8192;                     // pairs; 8 per thread; 64 KB of 64-bit keys
__global__ void __launch_bounds__(GS_THREADS) refine_group_small_kernel(StoreView sv, RefineParams p) {
    extern __shared__ __align__(16) unsigned char gs_smem[];
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(gs_smem);   // [n_sort] (local record index << 13) | pair index; ~0 = not pending
    __shared__ int s_wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = (int)(p.Q * (int64_t)p.stride);
    int n_sort = 1024; while (n_sort < total) n_sort <<= 1;
    for (int i = tid; i < n_sort; i += GS_THREADS) {
        unsigned long long key = ~0ull;
        if (i < total) {
            const uint32_t q = pair_query(p, (uint32_t)i);
            const int r = i - (int)q * p.stride;
            uint8_t v = 0xff;
            if (r < p.n_cand[q] && (!p.qfinite || p.qfinite[q])) {
                const int32_t id = p.cand_ids[i];
                if (id < 0 || id >= sv.n_global || is_deleted(sv, id)) v = FSPANN_V_NOT_FOUND;
                else if (id < sv.id_base || id >= sv.id_base + sv.N) v = 0xfd;
                else { v = 0xfe; key = ((unsigned long long)(uint32_t)(id - sv.id_base) << 13) | (unsigned long long)i; }
            }
            p.verdict[i] = v;
        }
        s_key[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= n_sort; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_sort; i += GS_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s_key[i], b = s_key[ixj];
                    if ((a > b) == ((i & k) == 0)) { s_key[i] = b; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // pending pairs are now sorted by (record, pair); thread t owns the contiguous run [t*per, (t+1)*per)
    const int per = n_sort / GS_THREADS, lo = tid * per;
    int heads = 0, pend = 0;
    for (int i = lo; i < lo + per; i++) {
        const unsigned long long key = s_key[i];
        if (key == ~0ull) continue;
        pend++;
        heads += (i == 0 || (s_key[i - 1] >> 13) != (key >> 13)) ? 1 : 0;
    }
    // block-wide exclusive prefix of (heads, pend)
    int ih = heads, ip = pend;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, ih, o), b = __shfl_up_sync(0xffffffffu, ip, o);
        if (lane >= o) { ih += a; ip += b; }
    }
    if (lane == 31) s_wsum[warp] = ih;
    __syncthreads();
    int bh = 0, th = 0;
    for (int w = 0; w < 32; w++) { const int c = s_wsum[w]; if (w < warp) bh += c; th += c; }
    __syncthreads();
    if (lane == 31) s_wsum[warp] = ip;
    __syncthreads();
    int bp = 0, tp = 0;
    for (int w = 0; w < 32; w++) { const int c = s_wsum[w]; if (w < warp) bp += c; tp += c; }
    int u = bh + ih - heads, pos = bp + ip - pend;
    for (int i = lo; i < lo + per; i++) {
        const unsigned long long key = s_key[i];
        if (key == ~0ull) continue;
        if (i == 0 || (s_key[i - 1] >> 13) != (key >> 13)) { p.uniq[u] = (int32_t)(key >> 13); p.uoff[u] = pos; u++; }
        p.pairs[pos++] = (uint32_t)(key & 0x1fffu);
    }
    if (tid == 0) { p.totals[0] = tp; p.totals[1] = th; p.totals[2] = 0; p.totals[3] = 0; p.uoff[th] = tp; }
}

int launch_refine_group(cudaStream_t s, const StoreView &sv, const RefineParams &p) {
    const int64_t total = p.Q * (int64_t)p.stride;
    if (total <= GS_MAX) {
        int n_sort = 1024; while (n_sort < total) n_sort <<= 1;
        refine_group_small_kernel<<<1, GS_THREADS, sizeof(unsigned long long) * (size_t)n_sort, s>>>(sv, p);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    const int64_t n = sv.N + 1;  // one extra slot so cnt[N] = total after the scan
    if (cudaMemsetAsync(p.cnt, 0, sizeof(int32_t) * (size_t)n, s) != cudaSuccess) return -1;
    int grid = (int)((total + 255) / 256); { const int cap = cur_sm_count() * 16; if (grid > cap) grid = cap; } if (grid < 1) grid = 1;
    refine_count_kernel<<<grid, 256, 0, s>>>(sv, p);
    const int nblocks = (int)((n + SCAN_ITEMS - 1) / SCAN_ITEMS);
    scan_partial_kernel<<<nblocks, SCAN_THREADS, 0, s>>>(p.cnt, n, p.block_sums);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, s>>>(p.block_sums, nblocks, p.totals, p.uoff);
    scan_apply_kernel<<<nblocks, SCAN_THREADS, 0, s>>>(p.cnt, n, p.block_sums, p.uniq, p.uoff, p.fill);
    refine_fill_kernel<<<grid, 256, 0, s>>>(sv, p);
    return cudaGetLastError() == cudaSuccess ? 5 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// decrypt + distance.  One warp owns up to 32 consecutive distinct records at a time; per authenticated record the lanes
// split the ciphertext blocks (coalesced 128-bit loads), each lane runs AES-CTR on its blocks, plaintext goes to the
// warp's shared-memory row; then lanes take the (query, rank) pairs of the record and accumulate the exact FP64 distance
// from shared memory.  Authentication (GHASH + tag) is done beforehand by refine_verify_kernel.
// AES uses one T-table (Te0) replicated for the 32 banks so every lookup is conflict free; Te1..3 are rotations.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RF_THREADS = 768;    // one CTA per SM (the 64 KB-aligned AES table costs up to 128 KB of shared memory); 24 warps measured best of 16..32
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int DBG_THREADS = 256;
[[maybe_unused]] constexpr int DBG_WARPS = DBG_THREADS / 32;

struct TeSmem {
    const uint32_t *t;  // te_s + lane
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const { return t[x << 5]; }
};
struct RkSmem {
    const uint32_t *r;
    __device__ __forceinline__ uint32_t operator()(int i) const { return r[i]; }
};

__device__ __forceinline__ void aes256_encrypt_fast(uint32_t K0, const uint32_t *rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                                                    uint32_t out[4]);
struct AesFast {      // production decrypt kernel
    uint32_t K0;
    __device__ __forceinline__ void operator()(const uint32_t *rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t out[4]) const {
        aes256_encrypt_fast(K0, rk, s0, s1, s2, s3, out);
    }
};
struct AesPlain {     // same arithmetic through the generic FSP_HD routine (debug tap kernel)
    TeSmem te;
    __device__ __forceinline__ void operator()(const uint32_t *rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t out[4]) const {
        aes256_encrypt(te, RkSmem{rk}, s0, s1, s2, s3, out);
    }
};

__device__ __forceinline__ int find_key_slot(const int32_t *s_ver, int nkeys, int32_t version) {
    int slot = -1;
    for (int i = 0; i < nkeys; i++) if (s_ver[i] == version) slot = i;
    return slot;
}

// ---- AES-256 with single-instruction table addressing --------------------------------------------------------------
// Table layout in shared memory (64 KB, placed on a 64 KB boundary of the shared window): entry x occupies 256 bytes =
// 64 four-byte columns; column `lane` holds Te0[x], column 32+lane holds Te2[x] = rot16(Te0[x]).  A lookup address is
// tbase + x*256 + lane*4 (+128): x lands exactly in byte 1 of the address, so ONE PRMT builds it from the state word and
// the per-lane constant K = tbase + lane*4 (whose byte 1 is zero), and every lane stays in its own bank (conflict free).
// Te1 / Te3 are byte rotations of the looked-up Te0 / Te2 words.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));     // read-only table: free to schedule
    return v;
}
__device__ __forceinline__ uint32_t ror8(uint32_t x) { return __funnelshift_r(x, x, 8); }
#define FSP_T0(s, k) lds32(__byte_perm((s), K0, 0x7604 | ((k) << 4)))
#define FSP_T2(s, k) lds32(__byte_perm((s), K2, 0x7604 | ((k) << 4)))
// second 64 KB block, same layout: column `lane` holds Te1[x] = ror8(Te0[x]), column 32+lane holds Te3[x] = ror8(Te2[x]): the rotate and one
// XOR per state word and round leave the ALU pipe (which the decrypt kernel fills to 81 %) for twice the table space
#define FSP_T1(s, k) lds32(__byte_perm((s), K1, 0x7604 | ((k) << 4)))
#define FSP_T3(s, k) lds32(__byte_perm((s), K3, 0x7604 | ((k) << 4)))
__device__ __forceinline__ void aes256_encrypt_fast(uint32_t K0, const uint32_t *rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3,
                                                    uint32_t out[4]) {
    const uint32_t K2 = K0 + 128u, K1 = K0 + 0x10000u, K3 = K1 + 128u;
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
#pragma unroll
    for (int r = 1; r < 14; r++) {
        // t_i = Te0[s_i.b3] ^ Te1[s_{i+1}.b2] ^ Te2[s_{i+2}.b1] ^ Te3[s_{i+3}.b0] ^ rk
        const uint32_t a0 = FSP_T0(s0, 3), b0 = FSP_T1(s1, 2), c0 = FSP_T2(s2, 1), d0 = FSP_T3(s3, 0);
        const uint32_t a1 = FSP_T0(s1, 3), b1 = FSP_T1(s2, 2), c1 = FSP_T2(s3, 1), d1 = FSP_T3(s0, 0);
        const uint32_t a2 = FSP_T0(s2, 3), b2 = FSP_T1(s3, 2), c2 = FSP_T2(s0, 1), d2 = FSP_T3(s1, 0);
        const uint32_t a3 = FSP_T0(s3, 3), b3 = FSP_T1(s0, 2), c3 = FSP_T2(s1, 1), d3 = FSP_T3(s2, 0);
        s0 = a0 ^ b0 ^ c0 ^ d0 ^ rk[4 * r + 0];
        s1 = a1 ^ b1 ^ c1 ^ d1 ^ rk[4 * r + 1];
        s2 = a2 ^ b2 ^ c2 ^ d2 ^ rk[4 * r + 2];
        s3 = a3 ^ b3 ^ c3 ^ d3 ^ rk[4 * r + 3];
    }
    // last round: S[x] is byte 3 and byte 0 of Te2[x] = (s, 3s, 2s, s) and bytes 2, 1 of Te0[x] = (2s, s, s, 3s)
#define FSP_LAST(a, b, c, d) ((FSP_T2(a, 3) & 0xff000000u) | (FSP_T0(b, 2) & 0x00ff0000u) | (FSP_T0(c, 1) & 0x0000ff00u) | (FSP_T2(d, 0) & 0x000000ffu))
    out[0] = FSP_LAST(s0, s1, s2, s3) ^ rk[56];
    out[1] = FSP_LAST(s1, s2, s3, s0) ^ rk[57];
    out[2] = FSP_LAST(s2, s3, s0, s1) ^ rk[58];
    out[3] = FSP_LAST(s3, s0, s1, s2) ^ rk[59];
#undef FSP_LAST
}

// ---- CTR-mode specialisation: rounds 1 and 2 hoisted out of the per-block work ------------------------------------
// Every counter block of one record is IV || ctr with the same IV, and ctr = 2 .. c+1 differs only in its low byte while
// c <= 254 (dim <= 508).  After AddRoundKey only byte 0 of column 3 varies, so after round 1 only column 0 varies
// (through Te3[s3.b0]) and in round 2 only the four look-ups fed by that column vary.  The 27 look-ups that depend on
// (IV, key) alone are evaluated ONCE per record (CtrPre, one lane per record of a 32-record chunk); a block then costs
// 1 + 4 look-ups for rounds 1-2 instead of 32 (197 per block instead of 224).
struct CtrPre { uint32_t c0, d0, d1, d2, d3; };

__device__ __forceinline__ CtrPre aes256_ctr_precompute(uint32_t K0, const uint32_t *rk, uint32_t iv0, uint32_t iv1, uint32_t iv2) {
    const uint32_t K2 = K0 + 128u, K1 = K0 + 0x10000u, K3 = K1 + 128u;
    const uint32_t s0 = iv0 ^ rk[0], s1 = iv1 ^ rk[1], s2 = iv2 ^ rk[2], s3 = rk[3];     // ctr bytes 1..3 are zero
    CtrPre o;
    o.c0 = FSP_T0(s0, 3) ^ FSP_T2(s2, 1) ^ rk[4] ^ FSP_T1(s1, 2);                     // + Te3[s3.b0] per block
    const uint32_t c1 = FSP_T0(s1, 3) ^ FSP_T2(s3, 1) ^ rk[5] ^ FSP_T1(s2, 2) ^ FSP_T3(s0, 0);
    const uint32_t c2 = FSP_T0(s2, 3) ^ FSP_T2(s0, 1) ^ rk[6] ^ FSP_T1(s3, 2) ^ FSP_T3(s1, 0);
    const uint32_t c3 = FSP_T0(s3, 3) ^ FSP_T2(s1, 1) ^ rk[7] ^ FSP_T1(s0, 2) ^ FSP_T3(s2, 0);
    o.d0 = FSP_T2(c2, 1) ^ rk[8] ^ FSP_T1(c1, 2) ^ FSP_T3(c3, 0);                     // + Te0[t0.b3]
    o.d1 = FSP_T0(c1, 3) ^ FSP_T2(c3, 1) ^ rk[9] ^ FSP_T1(c2, 2);                     // + Te3[t0.b0]
    o.d2 = FSP_T0(c2, 3) ^ rk[10] ^ FSP_T1(c3, 2) ^ FSP_T3(c1, 0);                    // + Te2[t0.b1]
    o.d3 = FSP_T0(c3, 3) ^ FSP_T2(c1, 1) ^ rk[11] ^ FSP_T3(c2, 0);                    // + Te1[t0.b2]
    return o;
}

// E_K(IV || ctr) for ctr < 256 given the record's CtrPre.
__device__ __forceinline__ void aes256_ctr_block(uint32_t K0, const uint32_t *rk, const CtrPre &pre, uint32_t ctr, uint32_t out[4]) {
    const uint32_t K2 = K0 + 128u, K1 = K0 + 0x10000u, K3 = K1 + 128u;
    const uint32_t t0 = pre.c0 ^ FSP_T3(ctr ^ rk[3], 0);
    uint32_t s0 = pre.d0 ^ FSP_T0(t0, 3);
    uint32_t s1 = pre.d1 ^ FSP_T3(t0, 0);
    uint32_t s2 = pre.d2 ^ FSP_T2(t0, 1);
    uint32_t s3 = pre.d3 ^ FSP_T1(t0, 2);
#pragma unroll
    for (int r = 3; r < 14; r++) {
        const uint32_t a0 = FSP_T0(s0, 3), b0 = FSP_T1(s1, 2), c0 = FSP_T2(s2, 1), d0 = FSP_T3(s3, 0);
        const uint32_t a1 = FSP_T0(s1, 3), b1 = FSP_T1(s2, 2), c1 = FSP_T2(s3, 1), d1 = FSP_T3(s0, 0);
        const uint32_t a2 = FSP_T0(s2, 3), b2 = FSP_T1(s3, 2), c2 = FSP_T2(s0, 1), d2 = FSP_T3(s1, 0);
        const uint32_t a3 = FSP_T0(s3, 3), b3 = FSP_T1(s0, 2), c3 = FSP_T2(s1, 1), d3 = FSP_T3(s2, 0);
        s0 = a0 ^ b0 ^ c0 ^ d0 ^ rk[4 * r + 0];
        s1 = a1 ^ b1 ^ c1 ^ d1 ^ rk[4 * r + 1];
        s2 = a2 ^ b2 ^ c2 ^ d2 ^ rk[4 * r + 2];
        s3 = a3 ^ b3 ^ c3 ^ d3 ^ rk[4 * r + 3];
    }
#define FSP_LAST(a, b, c, d) ((FSP_T2(a, 3) & 0xff000000u) | (FSP_T0(b, 2) & 0x00ff0000u) | (FSP_T0(c, 1) & 0x0000ff00u) | (FSP_T2(d, 0) & 0x000000ffu))
    out[0] = FSP_LAST(s0, s1, s2, s3) ^ rk[56];
    out[1] = FSP_LAST(s1, s2, s3, s0) ^ rk[57];
    out[2] = FSP_LAST(s2, s3, s0, s1) ^ rk[58];
    out[3] = FSP_LAST(s3, s0, s1, s2) ^ rk[59];
#undef FSP_LAST
}
#undef FSP_T0
#undef FSP_T2
#undef FSP_T1
#undef FSP_T3

// AES-CTR decryption of record `id` by the whole warp (authentication already done by refine_verify_kernel).
// Plaintext doubles -> pt_row[0..dim).  Returns true when every value is finite (warp-uniform).
template <class AES>
__device__ __forceinline__ bool warp_decrypt_record(const StoreView &sv, const uint8_t *rec, uint4 hdr, int slot, const AES &aes,
                                                    const uint32_t *s_rk, double *pt_row, int lane) {
    const int nbytes = 8 * sv.dim;
    const int c = (nbytes + 15) >> 4;                     // ciphertext blocks (the last may hold only 8 bytes)
    const uint32_t iv0 = bswap32(hdr.x), iv1 = bswap32(hdr.y), iv2 = bswap32(hdr.z);
    const uint32_t *rk = s_rk + slot * 60;
    const uint4 *ctv = reinterpret_cast<const uint4 *>(rec + 16);
    bool finite = true;
    for (int blk = lane; blk < c; blk += 32) {
        const uint4 w = __ldg(ctv + blk);
        uint32_t ks[4];
        aes(rk, iv0, iv1, iv2, (uint32_t)(blk + 2), ks);
        const uint32_t p0 = bswap32(w.x) ^ ks[0], p1 = bswap32(w.y) ^ ks[1];
        // big-endian FP64 (AGC:261-277): first word is the high half
        pt_row[2 * blk] = __hiloint2double((int)p0, (int)p1);
        finite &= ((p0 >> 20) & 0x7ffu) != 0x7ffu;
        if (!((blk == c - 1) && (nbytes & 15))) {          // odd dim: the last block holds one double, the rest is tag
            const uint32_t p2 = bswap32(w.z) ^ ks[2], p3 = bswap32(w.w) ^ ks[3];
            pt_row[2 * blk + 1] = __hiloint2double((int)p2, (int)p3);
            finite &= ((p2 >> 20) & 0x7ffu) != 0x7ffu;
        }
    }
    return __all_sync(0xffffffffu, finite);
}

// Integer test of one plaintext double given as its IEEE words: true when the value is an integer in [0, 255] (+0.0 included, -0.0 not);
// *byte receives it.  SIFT / .bvecs descriptors (loader/.../BvecsLoader.java) are such integers: for them the squared differences and
// every partial sum of QSI:364-372 are exact integers < 2^53, so the distance may be accumulated in integer arithmetic bit for bit.
__device__ __forceinline__ bool small_int_of_double(uint32_t hi, uint32_t lo, uint32_t &byte) {
    const int e = (int)(hi >> 20) - 1023;                      // a set sign bit makes e > 7
    const uint32_t frac = hi & 0xfffffu;
    const bool zero = (hi | lo) == 0u;
    const int sh = 20 - (e & 31);
    const bool ok = lo == 0u && e >= 0 && e <= 7 && (frac & ((1u << sh) - 1u)) == 0u;
    byte = zero ? 0u : ((0x100000u | frac) >> sh) & 0xffu;
    return zero || ok;
}

// Same with the record's hoisted round-1/2 constants (production kernel, c <= 254).  u8_row != nullptr: also writes the plaintext as bytes
// and reports (warp-uniform) in *all_small_int whether EVERY value of the record is an integer in [0, 255].
__device__ __forceinline__ bool warp_decrypt_record_ctr(const StoreView &sv, const uint8_t *rec, uint32_t K0, const uint32_t *rk, const CtrPre &pre,
                                                        double *pt_row, int lane, uint8_t *u8_row, bool *all_small_int) {
    const int nbytes = 8 * sv.dim;
    const int c = (nbytes + 15) >> 4;
    const uint4 *ctv = reinterpret_cast<const uint4 *>(rec + 16);
    bool finite = true, small = true;
    for (int blk = lane; blk < c; blk += 32) {
        const uint4 w = __ldg(ctv + blk);
        uint32_t ks[4];
        aes256_ctr_block(K0, rk, pre, (uint32_t)(blk + 2), ks);
        const uint32_t p0 = bswap32(w.x) ^ ks[0], p1 = bswap32(w.y) ^ ks[1];
        finite &= ((p0 >> 20) & 0x7ffu) != 0x7ffu;
        if (!((blk == c - 1) && (nbytes & 15))) {
            const uint32_t p2 = bswap32(w.z) ^ ks[2], p3 = bswap32(w.w) ^ ks[3];
            finite &= ((p2 >> 20) & 0x7ffu) != 0x7ffu;
            // big-endian FP64 (AGC:261-277): first word is the high half; one 128-bit store per block
            *reinterpret_cast<double2 *>(pt_row + 2 * blk) = make_double2(__hiloint2double((int)p0, (int)p1), __hiloint2double((int)p2, (int)p3));
            if (u8_row) {
                uint32_t b0, b1;
                small &= small_int_of_double(p0, p1, b0);
                small &= small_int_of_double(p2, p3, b1);
                *reinterpret_cast<uint16_t *>(u8_row + 2 * blk) = (uint16_t)(b0 | (b1 << 8));
            }
        } else {
            pt_row[2 * blk] = __hiloint2double((int)p0, (int)p1);   // odd dim: the last block holds one double, the rest is tag
            small = false;                                          // the byte path needs dim % 16 == 0 anyway
        }
    }
    if (all_small_int) *all_small_int = u8_row != nullptr && __all_sync(0xffffffffu, small);
    return __all_sync(0xffffffffu, finite);
}

// ---- TMA staging of records (cp.async.bulk, 1-D) ------------------------------------------------------------------------------------
// A warp's elected lane arms the warp's mbarrier with the byte count of the group's records and issues ONE bulk copy per record
// (global -> that record's shared-memory row; the record is 16-byte aligned and a multiple of 16 bytes long).  The copy engine moves the
// bytes; the lanes wait on the barrier's phase, then AES-CTR runs IN PLACE on the staged row (ciphertext words out, plaintext doubles in),
// so the ciphertext is read from HBM exactly once per batch and never travels through the LSU's global path.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// AES-CTR decryption IN PLACE of a record staged in shared memory by tma_load_1d: row = [iv 12 | version 4 | ciphertext | tag], the
// plaintext doubles replace the ciphertext at row + 16.  Same outputs as warp_decrypt_record_ctr.
__device__ __forceinline__ bool warp_decrypt_staged_ctr(int dim, unsigned char *row, uint32_t K0, const uint32_t *rk, const CtrPre &pre, int lane,
                                                        uint8_t *u8_row, bool *all_small_int) {
    const int nbytes = 8 * dim;
    const int c = (nbytes + 15) >> 4;
    uint4 *ctv = reinterpret_cast<uint4 *>(row + 16);
    bool finite = true, small = true;
    for (int blk = lane; blk < c; blk += 32) {
        const uint4 w = ctv[blk];
        uint32_t ks[4];
        aes256_ctr_block(K0, rk, pre, (uint32_t)(blk + 2), ks);
        const uint32_t p0 = bswap32(w.x) ^ ks[0], p1 = bswap32(w.y) ^ ks[1];
        finite &= ((p0 >> 20) & 0x7ffu) != 0x7ffu;
        if (!((blk == c - 1) && (nbytes & 15))) {
            const uint32_t p2 = bswap32(w.z) ^ ks[2], p3 = bswap32(w.w) ^ ks[3];
            finite &= ((p2 >> 20) & 0x7ffu) != 0x7ffu;
            // big-endian FP64 (AGC:261-277): first word is the high half; little-endian double = (lo, hi)
            ctv[blk] = make_uint4(p1, p0, p3, p2);
            if (u8_row) {
                uint32_t b0, b1;
                small &= small_int_of_double(p0, p1, b0);
                small &= small_int_of_double(p2, p3, b1);
                *reinterpret_cast<uint16_t *>(u8_row + 2 * blk) = (uint16_t)(b0 | (b1 << 8));
            }
        } else {
            *reinterpret_cast<uint2 *>(row + 16 + 16 * blk) = make_uint2(p1, p0);   // odd dim: the last block holds one double, the rest is tag
            small = false;
        }
    }
    if (all_small_int) *all_small_int = u8_row != nullptr && __all_sync(0xffffffffu, small);
    return __all_sync(0xffffffffu, finite);
}

// ------------------------------------------------------------------------------------------------------------------
// Authentication: one LANE per distinct record.  GHASH runs as Horner's rule Y <- (Y ^ X_i) * H with the multiply done
// by 16 look-ups into the key version's 8-bit Shoup table (64 KB in shared memory, one live version at a time), over the
// AAD blocks ("id:<id>|v:<ver>|d:<dim>", EP:80-83), the ciphertext blocks and the length block; tag = GHASH ^ E_K(J0)
// (NIST SP 800-38D).  Writes one verdict per distinct record: OK / TAG_FAIL (NO_KEY is the preset default).
// ------------------------------------------------------------------------------------------------------------------
constexpr int VF_THREADS = 512;            // one CTA per SM: 64 KB Shoup table + 32 KB Te + 16 warps x 5 KB of ciphertext staging
constexpr int VF_SLICE = 4;                // ciphertext blocks staged per record and step (64 bytes)
constexpr int VF_ROW = VF_SLICE * 16 + 16; // staged row pitch: 80 bytes -> the 128-bit reads of a quarter warp fall into 8 distinct bank groups
constexpr int VF_STAGE = 32 * VF_ROW;      // one buffer of one warp

struct ShoupSmem {
    const uint4 *t;   // [256][16] byte-major
    int rot;          // lane & 15: this lane visits byte positions rot, rot+1, ... (mod 16)
    // y <- y * H.  The 16 table look-ups are XOR-accumulated in a lane-rotated order: at step s lane l reads byte position
    // (s + l) mod 16, whose entries sit in bank group (s + l) mod 8 -- distinct for the 8 lanes of a quarter warp.
    __device__ __forceinline__ void mul(uint32_t &y0, uint32_t &y1, uint32_t &y2, uint32_t &y3) const {
        // rotate the 16-byte vector left by `rot` bytes: whole words first (two conditional stages), then bytes
        uint32_t a0 = y0, a1 = y1, a2 = y2, a3 = y3;
        if (rot & 4) { const uint32_t t0 = a0; a0 = a1; a1 = a2; a2 = a3; a3 = t0; }
        if (rot & 8) { const uint32_t t0 = a0, t1 = a1; a0 = a2; a1 = a3; a2 = t0; a3 = t1; }
        const int sh = (rot & 3) * 8;
        const uint32_t w0 = __funnelshift_l(a1, a0, sh), w1 = __funnelshift_l(a2, a1, sh), w2 = __funnelshift_l(a3, a2, sh),
                       w3 = __funnelshift_l(a0, a3, sh);
        const char *base = reinterpret_cast<const char *>(t);
        const uint32_t r16 = (uint32_t)rot * 16u;
        uint4 z = make_uint4(0, 0, 0, 0), u;
        // byte s of the rotated vector is byte (s + rot) & 15 of y; entry (b, j) lives at byte offset b*256 + j*16
#define FSP_ACC(st, wd, kb) u = *reinterpret_cast<const uint4 *>(base + __byte_perm((wd), 0, 0x4404 | ((kb) << 4)) + ((r16 + (st) * 16u) & 0xf0u)); \
                         z.x ^= u.x; z.y ^= u.y; z.z ^= u.z; z.w ^= u.w;
        FSP_ACC(0, w0, 3) FSP_ACC(1, w0, 2) FSP_ACC(2, w0, 1) FSP_ACC(3, w0, 0)
        FSP_ACC(4, w1, 3) FSP_ACC(5, w1, 2) FSP_ACC(6, w1, 1) FSP_ACC(7, w1, 0)
        FSP_ACC(8, w2, 3) FSP_ACC(9, w2, 2) FSP_ACC(10, w2, 1) FSP_ACC(11, w2, 0)
        FSP_ACC(12, w3, 3) FSP_ACC(13, w3, 2) FSP_ACC(14, w3, 1) FSP_ACC(15, w3, 0)
#undef FSP_ACC
        y0 = z.x; y1 = z.y; y2 = z.z; y3 = z.w;
    }
};

// GCM tag of one record by ONE lane: tag = GHASH_H(AAD || ciphertext || lengths) ^ E_K(J0).  `sh` = Shoup table of the key
// version named in the record header, rk = its round keys, id = the GLOBAL id the AAD binds (EP:80-83).
template <class SH>
__device__ __forceinline__ void lane_gcm_tag(const StoreView &sv, const uint8_t *rec, int32_t id, uint4 hdr, const SH &sh, const TeSmem &te,
                                             const RkSmem &rk, uint32_t tag[4]) {
    const int dim = sv.dim, nbytes = 8 * dim, c = (nbytes + 15) >> 4;
    const int32_t version = (int32_t)hdr.w;
    // ---- GHASH over AAD || ciphertext || lengths ----
    uint32_t y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    uint8_t aad[FSP_AAD_MAX];
    const int alen = build_aad((int64_t)id, version, dim, aad);
    for (int j = 0; j < (alen + 15) >> 4; j++) {
        const uint64_t hi = load_be64(aad + 16 * j), lo = load_be64(aad + 16 * j + 8);
        y0 ^= (uint32_t)(hi >> 32); y1 ^= (uint32_t)hi; y2 ^= (uint32_t)(lo >> 32); y3 ^= (uint32_t)lo;
        sh.mul(y0, y1, y2, y3);
    }
    const uint4 *ctv = reinterpret_cast<const uint4 *>(rec + 16);
    for (int b0 = 0; b0 < c; b0 += 8) {
        uint4 w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = (b0 + k < c) ? ld_rec(ctv + b0 + k) : make_uint4(0, 0, 0, 0);   // 128 contiguous bytes of this record
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (b0 + k < c) {
                uint32_t x0 = bswap32(w[k].x), x1 = bswap32(w[k].y), x2 = bswap32(w[k].z), x3 = bswap32(w[k].w);
                if ((b0 + k == c - 1) && (nbytes & 15)) { x2 = 0; x3 = 0; }     // odd dim: zero-pad the half block
                y0 ^= x0; y1 ^= x1; y2 ^= x2; y3 ^= x3;
                sh.mul(y0, y1, y2, y3);
            }
        }
    }
    y1 ^= (uint32_t)alen * 8u;                       // [len(A)]64 || [len(C)]64 in bits
    y3 ^= (uint32_t)nbytes * 8u;
    sh.mul(y0, y1, y2, y3);
    // ---- tag = GHASH ^ E_K(J0), J0 = IV || 0x00000001 ----
    uint32_t ej0[4];
    aes256_encrypt(te, rk, bswap32(hdr.x), bswap32(hdr.y), bswap32(hdr.z), 1u, ej0);
    tag[0] = y0 ^ ej0[0]; tag[1] = y1 ^ ej0[1]; tag[2] = y2 ^ ej0[2]; tag[3] = y3 ^ ej0[3];
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The same tag for the 32 records of a WARP (lane l owns record l; rec == nullptr: the lane idles), with the ciphertext staged through
// shared memory: a lane-per-record read of 16 bytes touches 32 different 128-byte lines per warp instruction and costs the LSU as much as
// half of the Shoup look-ups.  Here the warp copies VF_SLICE blocks of all its records per step with cp.async (4 lanes per record: 64
// contiguous bytes, whole sectors; global -> shared without registers), double-buffered, and every lane then reads its own row with
// conflict-free 128-bit shared loads.  `stage` = this warp's 2 x VF_STAGE bytes.
template <class SH>
__device__ __forceinline__ void warp_gcm_tag_staged(const StoreView &sv, const uint8_t *rec, int32_t id, uint4 hdr, const SH &sh, const TeSmem &te,
                                                    const RkSmem &rk, unsigned char *stage, int lane, uint32_t tag[4]) {
    const int dim = sv.dim, nbytes = 8 * dim, c = (nbytes + 15) >> 4;
    const int n_slices = (c + VF_SLICE - 1) / VF_SLICE;
    const int32_t version = (int32_t)hdr.w;
    const unsigned long long ctp = rec ? (unsigned long long)(uintptr_t)(rec + 16) : 0ull;
    auto issue = [&](int sl, int b) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int src = j * 8 + (lane >> 2);
            const unsigned long long pr = __shfl_sync(0xffffffffu, ctp, src);
            const int blk = sl * VF_SLICE + (lane & 3);
            if (pr && blk < c) cp_async16(stage + b * VF_STAGE + src * VF_ROW + (lane & 3) * 16, reinterpret_cast<const void *>((uintptr_t)(pr + (size_t)blk * 16)));
        }
        cp_async_commit();
    };
    issue(0, 0);
    // ---- GHASH over AAD || ciphertext || lengths ----
    uint32_t y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    int alen = 0;
    if (rec) {
        uint8_t aad[FSP_AAD_MAX];
        alen = build_aad((int64_t)id, version, dim, aad);
        for (int j = 0; j < (alen + 15) >> 4; j++) {
            const uint64_t hi = load_be64(aad + 16 * j), lo = load_be64(aad + 16 * j + 8);
            y0 ^= (uint32_t)(hi >> 32); y1 ^= (uint32_t)hi; y2 ^= (uint32_t)(lo >> 32); y3 ^= (uint32_t)lo;
            sh.mul(y0, y1, y2, y3);
        }
    }
    for (int sl = 0; sl < n_slices; sl++) {
        if (sl + 1 < n_slices) { issue(sl + 1, (sl + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        if (rec) {
            const uint4 *row = reinterpret_cast<const uint4 *>(stage + (sl & 1) * VF_STAGE + lane * VF_ROW);
#pragma unroll
            for (int k = 0; k < VF_SLICE; k++) {
                const int blk = sl * VF_SLICE + k;
                if (blk < c) {
                    const uint4 w = row[k];
                    uint32_t x0 = bswap32(w.x), x1 = bswap32(w.y), x2 = bswap32(w.z), x3 = bswap32(w.w);
                    if ((blk == c - 1) && (nbytes & 15)) { x2 = 0; x3 = 0; }     // odd dim: zero-pad the half block
                    y0 ^= x0; y1 ^= x1; y2 ^= x2; y3 ^= x3;
                    sh.mul(y0, y1, y2, y3);
                }
            }
        }
        __syncwarp();                                                           // the buffer is refilled two steps later
    }
    if (!rec) return;
    y1 ^= (uint32_t)alen * 8u;                       // [len(A)]64 || [len(C)]64 in bits
    y3 ^= (uint32_t)nbytes * 8u;
    sh.mul(y0, y1, y2, y3);
    // ---- tag = GHASH ^ E_K(J0), J0 = IV || 0x00000001 ----
    uint32_t ej0[4];
    aes256_encrypt(te, rk, bswap32(hdr.x), bswap32(hdr.y), bswap32(hdr.z), 1u, ej0);
    tag[0] = y0 ^ ej0[0]; tag[1] = y1 ^ ej0[1]; tag[2] = y2 ^ ej0[2]; tag[3] = y3 ^ ej0[3];
}

// GHASH + tag check of one record by ONE lane (AGC:145-158: doFinal throws AEADBadTagException on a mismatch).
template <class SH>
__device__ __forceinline__ bool lane_verify_record(const StoreView &sv, const uint8_t *rec, int32_t id, uint4 hdr, const SH &sh,
                                                   const TeSmem &te, const RkSmem &rk) {
    uint32_t tag[4];
    lane_gcm_tag(sv, rec, id, hdr, sh, te, rk, tag);
    const uint2 *tagp = reinterpret_cast<const uint2 *>(rec + 16 + 8 * sv.dim);
    const uint2 t0 = ld_rec(tagp), t1 = ld_rec(tagp + 1);
    return tag[0] == bswap32(t0.x) && tag[1] == bswap32(t0.y) && tag[2] == bswap32(t1.x) && tag[3] == bswap32(t1.y);
}

// Shared body of the authentication kernels: one LANE per listed record, one key version (= one Shoup table in shared memory)
// at a time.  list[u] = record index inside this store view; gid[u] = global id for the AAD (or nullptr: index + id_base).
// write_flag == nullptr: verify, verdict[u] = OK / TAG_FAIL (records of unknown versions keep their preset verdict).
// write_flag != nullptr: (re-)compute and STORE the tag of every record with write_flag[u] != 0 (encrypt / Migrate).
// vorder / voff (optional): list positions bucketed by key-version slot (voff[slot] .. voff[slot+1]) so that with several live
// versions every pass walks a DENSE list of its own records instead of all records with most lanes masked off.
__device__ __forceinline__ void gcm_tag_body(const StoreView &sv, const int32_t *__restrict__ list, const int32_t *__restrict__ gid, int n_list,
                                             uint8_t *__restrict__ verdict, const uint8_t *__restrict__ write_flag, unsigned char *vf_smem,
                                             const int32_t *__restrict__ vorder = nullptr, const int32_t *__restrict__ voff = nullptr) {
    uint4 *shoup_s = reinterpret_cast<uint4 *>(vf_smem);                      // [16][256]
    uint32_t *te_s = reinterpret_cast<uint32_t *>(shoup_s + 4096);            // [256][32]
    uint32_t *s_rk = te_s + 256 * 32;                                         // [kMaxKeys][60]
    int32_t *s_ver = reinterpret_cast<int32_t *>(s_rk + kMaxKeys * 60);
    unsigned char *stage_s = reinterpret_cast<unsigned char *>(s_ver + kMaxKeys);   // [warps][2][VF_STAGE]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256 * 32; i += VF_THREADS) te_s[i] = sv.te0[i >> 5];
    const int nkeys = sv.keys->n;
    for (int i = tid; i < nkeys * 60; i += VF_THREADS) s_rk[i] = sv.keys->rk[i / 60][i % 60];
    for (int i = tid; i < kMaxKeys; i += VF_THREADS) s_ver[i] = i < nkeys ? sv.keys->version[i] : INT32_MIN;
    const TeSmem te{te_s + lane};
    const ShoupSmem sh{shoup_s, lane & 15};
    const int warps_total = gridDim.x * (VF_THREADS / 32);

    for (int vi = 0; vi < nkeys; vi++) {
        const int v_lo = vorder ? voff[vi] : 0, v_n = vorder ? voff[vi + 1] - v_lo : n_list;
        if (v_n <= 0) continue;                                             // uniform across the grid
        __syncthreads();
        for (int i = tid; i < 4096; i += VF_THREADS) shoup_s[i] = sv.shoup[(size_t)vi * 4096 + i];
        __syncthreads();
        const int32_t version = s_ver[vi];
        const RkSmem rk{s_rk + vi * 60};
        const int n_chunks = (v_n + 31) >> 5;
        for (int chunk = blockIdx.x * (VF_THREADS / 32) + warp; chunk < n_chunks; chunk += warps_total) {   // whole warps stay in the loop
            const int idx = (chunk << 5) + lane;
            const uint8_t *rec = nullptr;
            int32_t id = 0, u = -1;
            uint4 hdr = make_uint4(0, 0, 0, 0);
            if (idx < v_n) {
                u = vorder ? vorder[v_lo + idx] : idx;
                if (!(write_flag && !write_flag[u])) {
                    const int32_t li = list[u];                                 // index inside this shard
                    id = gid ? gid[u] : (int32_t)(li + sv.id_base);             // global id: what the AAD binds (EP:80-83)
                    const uint8_t *r = sv.rec + (size_t)li * sv.rec_stride;
                    hdr = ld_rec(reinterpret_cast<const uint4 *>(r));
                    if ((int32_t)hdr.w == version) rec = r;
                }
            }
            uint32_t tag[4] = {0, 0, 0, 0};
            warp_gcm_tag_staged(sv, rec, id, hdr, sh, te, rk, stage_s + (size_t)warp * 2 * VF_STAGE, lane, tag);
            if (!rec) continue;
            uint2 *tagp = reinterpret_cast<uint2 *>(const_cast<uint8_t *>(rec) + 16 + 8 * sv.dim);
            if (!write_flag) {
                const uint2 t0 = ld_rec(tagp), t1 = ld_rec(tagp + 1);         // AGC:145-158: doFinal throws AEADBadTagException on a mismatch
                const bool ok = tag[0] == bswap32(t0.x) && tag[1] == bswap32(t0.y) && tag[2] == bswap32(t1.x) && tag[3] == bswap32(t1.y);
                verdict[u] = ok ? FSPANN_V_OK : FSPANN_V_TAG_FAIL;          // AGC:159-165
            } else {
                tagp[0] = make_uint2(bswap32(tag[0]), bswap32(tag[1]));
                tagp[1] = make_uint2(bswap32(tag[2]), bswap32(tag[3]));
            }
        }
    }
}

__global__ void __launch_bounds__(VF_THREADS, 1) refine_verify_kernel(StoreView sv, RefineParams p) {
    extern __shared__ __align__(16) unsigned char vf_smem[];
    gcm_tag_body(sv, p.uniq, nullptr, p.totals[1], p.rec_verdict, nullptr, vf_smem, p.vorder, p.voff);
}

// Buckets the distinct records of the batch by key-version slot (only launched when several versions are live).
//   pass 0: vcnt[slot]++            pass 1 (after the host-free prefix in version_offsets_kernel): vorder[voff[slot] + cursor[slot]++] = u
__global__ void version_bucket_kernel(StoreView sv, RefineParams p, int pass) {
    const int n = p.totals[1], nkeys = sv.keys->n, lane = threadIdx.x & 31;
    const int n_round = (n + 31) & ~31;                                     // whole warps stay in the loop (warp-aggregated atomics)
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_round; u += gridDim.x * blockDim.x) {
        int slot = -1;
        if (u < n) {
            const uint4 hdr = ld_rec(reinterpret_cast<const uint4 *>(sv.rec + (size_t)p.uniq[u] * sv.rec_stride));
            for (int i = 0; i < nkeys; i++) if (sv.keys->version[i] == (int32_t)hdr.w) slot = i;
        }
        // unknown / retired version (slot -1): verdict stays NO_KEY.  One atomic per (warp, slot).
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
        if (slot < 0) continue;
        const int leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (lane == leader) base = atomicAdd(&p.vcnt[(pass ? kMaxKeys : 0) + slot], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        if (pass) p.vorder[p.voff[slot] + base + rank] = u;
    }
}
__global__ void version_offsets_kernel(RefineParams p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kMaxKeys; i++) { p.voff[i] = run; run += p.vcnt[i]; }
        p.voff[kMaxKeys] = run;
    }
}
int launch_version_bucket(cudaStream_t s, const StoreView &sv, const RefineParams &p, int64_t n_upper) {
    if (cudaMemsetAsync(p.vcnt, 0, sizeof(int32_t) * 2 * kMaxKeys, s) != cudaSuccess) return -1;
    int grid = (int)std::min<int64_t>((n_upper + 255) / 256, (int64_t)cur_sm_count() * 8); if (grid < 1) grid = 1;
    version_bucket_kernel<<<grid, 256, 0, s>>>(sv, p, 0);
    version_offsets_kernel<<<1, 32, 0, s>>>(p);
    version_bucket_kernel<<<grid, 256, 0, s>>>(sv, p, 1);
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}

// The same over an explicit host-sized list (Migrate / bulk encryption).
__global__ void __launch_bounds__(VF_THREADS, 1) gcm_tag_kernel(StoreView sv, const int32_t *list, const int32_t *gid, int n_list, uint8_t *verdict,
                                                             const uint8_t *write_flag) {
    extern __shared__ __align__(16) unsigned char vf_smem[];
    gcm_tag_body(sv, list, gid, n_list, verdict, write_flag, vf_smem);
}

static size_t verify_smem_bytes() {
    return sizeof(uint4) * 4096 + sizeof(uint32_t) * (256 * 32 + kMaxKeys * 60) + sizeof(int32_t) * kMaxKeys + (size_t)(VF_THREADS / 32) * 2 * VF_STAGE;
}

int launch_gcm_tag(cudaStream_t s, const StoreView &sv, const int32_t *list, const int32_t *gid, int n_list, uint8_t *verdict,
                   const uint8_t *write_flag, int sm_count) {
    if (n_list <= 0) return 0;
    const size_t smem = verify_smem_bytes();
    int grid = (n_list + VF_THREADS - 1) / VF_THREADS; if (grid > sm_count) grid = sm_count;
    gcm_tag_kernel<<<grid, VF_THREADS, smem, s>>>(sv, list, gid, n_list, verdict, write_flag);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_refine_verify(cudaStream_t s, const StoreView &sv, const RefineParams &p, int sm_count) {
    refine_verify_kernel<<<sm_count, VF_THREADS, verify_smem_bytes(), s>>>(sv, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Exact sequential FP64 squared distance between one query and the plaintext row `row` (shared memory), QSI:364-372: strict
// FP64, index order, no FMA contraction.  The three query encodings are value-identical, so the result is bit-identical.
// Both the query and the record hold integers in [0, 255]: sum of squared byte differences with |a-b| per byte (VABSDIFF4) and a 4-way
// integer dot product (DP4A).  Every partial sum of the reference's FP64 loop is then an exact integer, so (double)acc IS its result.
__device__ __forceinline__ double pair_dist2_bytes(const RefineParams &p, int dim, const uint8_t *row_u8, uint32_t q) {
    const uint4 *q16 = reinterpret_cast<const uint4 *>(p.queries_u8 + (size_t)q * dim);
    const uint4 *v16 = reinterpret_cast<const uint4 *>(row_u8);
    const int n16 = dim >> 4;
    uint32_t acc = 0;
    for (int i0 = 0; i0 < n16; i0 += 8) {
        uint4 qq[8];
#pragma unroll
        for (int u = 0; u < 8; u++) qq[u] = (i0 + u < n16) ? __ldg(q16 + i0 + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (i0 + u < n16) {
                const uint4 vv = v16[i0 + u];
                uint32_t d;
                d = __vabsdiffu4(qq[u].x, vv.x); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(qq[u].y, vv.y); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(qq[u].z, vv.z); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(qq[u].w, vv.w); acc = __dp4a(d, d, acc);
            }
        }
    }
    return (double)acc;
}

__device__ __forceinline__ double pair_dist2(const RefineParams &p, bool use_u8, bool use_f32, int dim, const double *row, uint32_t q) {
    double s = 0.0;
    const double2 *v2 = reinterpret_cast<const double2 *>(row);
    if (use_u8) {
        // every query value is an integer in [0, 255] (SIFT / .bvecs descriptors, loader/BvecsLoader): one 128-byte
        // line holds 128 dimensions, an eighth of the FP64 row.  Bytes are widened exactly (2^52 magic constant).
        const uint4 *q16 = reinterpret_cast<const uint4 *>(p.queries_u8 + (size_t)q * dim);
        const int n16 = dim >> 4;
        for (int i0 = 0; i0 < n16; i0 += 8) {
            uint4 qq[8];
#pragma unroll
            for (int u = 0; u < 8; u++) qq[u] = (i0 + u < n16) ? __ldg(q16 + i0 + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (i0 + u < n16) {
                    const uint32_t wds[4] = {qq[u].x, qq[u].y, qq[u].z, qq[u].w};
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        const double2 va = v2[8 * (i0 + u) + 2 * c4], vb = v2[8 * (i0 + u) + 2 * c4 + 1];
                        const double q0 = __hiloint2double(0x43300000, (int)(wds[c4] & 0xffu)) - 4503599627370496.0;
                        const double q1 = __hiloint2double(0x43300000, (int)((wds[c4] >> 8) & 0xffu)) - 4503599627370496.0;
                        const double q2 = __hiloint2double(0x43300000, (int)((wds[c4] >> 16) & 0xffu)) - 4503599627370496.0;
                        const double q3 = __hiloint2double(0x43300000, (int)(wds[c4] >> 24)) - 4503599627370496.0;
                        const double d0 = __dsub_rn(q0, va.x), d1 = __dsub_rn(q1, va.y), d2 = __dsub_rn(q2, vb.x), d3 = __dsub_rn(q3, vb.y);
                        s = __dadd_rn(s, __dmul_rn(d0, d0));
                        s = __dadd_rn(s, __dmul_rn(d1, d1));
                        s = __dadd_rn(s, __dmul_rn(d2, d2));
                        s = __dadd_rn(s, __dmul_rn(d3, d3));
                    }
                }
            }
        }
    } else if (use_f32) {
        // every query value is exactly representable in FP32 (the reference's loaders read float32 and widen,
        // FvecsLoader.java:27-30): read the compact FP32 copy -- half the bytes through L1 -- and widen back
        const float4 *q4 = reinterpret_cast<const float4 *>(p.queries_f32 + (size_t)q * dim);
        const int n4 = dim >> 2;
        int i = 0;
        for (; i + 8 <= n4; i += 8) {
            float4 qq[8];
#pragma unroll
            for (int u = 0; u < 8; u++) qq[u] = __ldg(q4 + i + u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double2 va = v2[2 * (i + u)], vb = v2[2 * (i + u) + 1];
                const double d0 = __dsub_rn((double)qq[u].x, va.x), d1 = __dsub_rn((double)qq[u].y, va.y);
                const double d2 = __dsub_rn((double)qq[u].z, vb.x), d3 = __dsub_rn((double)qq[u].w, vb.y);
                s = __dadd_rn(s, __dmul_rn(d0, d0));
                s = __dadd_rn(s, __dmul_rn(d1, d1));
                s = __dadd_rn(s, __dmul_rn(d2, d2));
                s = __dadd_rn(s, __dmul_rn(d3, d3));
            }
        }
        for (; i < n4; i++) {
            const float4 q1 = __ldg(q4 + i);
            const double2 va = v2[2 * i], vb = v2[2 * i + 1];
            const double d0 = __dsub_rn((double)q1.x, va.x), d1 = __dsub_rn((double)q1.y, va.y);
            const double d2 = __dsub_rn((double)q1.z, vb.x), d3 = __dsub_rn((double)q1.w, vb.y);
            s = __dadd_rn(s, __dmul_rn(d0, d0));
            s = __dadd_rn(s, __dmul_rn(d1, d1));
            s = __dadd_rn(s, __dmul_rn(d2, d2));
            s = __dadd_rn(s, __dmul_rn(d3, d3));
        }
    } else if ((dim & 1) == 0) {
        // 128-bit loads, 16 values in flight per lane; the adds stay strictly sequential
        const double2 *q2 = reinterpret_cast<const double2 *>(p.queries + (size_t)q * dim);
        const int n2 = dim >> 1;
        int i = 0;
        for (; i + 8 <= n2; i += 8) {
            double2 qq[8];
#pragma unroll
            for (int u = 0; u < 8; u++) qq[u] = __ldg(q2 + i + u);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double2 vv = v2[i + u];
                const double d0 = __dsub_rn(qq[u].x, vv.x), d1 = __dsub_rn(qq[u].y, vv.y);
                s = __dadd_rn(s, __dmul_rn(d0, d0));
                s = __dadd_rn(s, __dmul_rn(d1, d1));
            }
        }
        for (; i < n2; i++) {
            const double2 q1 = __ldg(q2 + i), vv = v2[i];
            const double d0 = __dsub_rn(q1.x, vv.x), d1 = __dsub_rn(q1.y, vv.y);
            s = __dadd_rn(s, __dmul_rn(d0, d0));
            s = __dadd_rn(s, __dmul_rn(d1, d1));
        }
    } else {
        const double *qv = p.queries + (size_t)q * dim;
        for (int i = 0; i < dim; i++) {
            const double d = __dsub_rn(__ldg(qv + i), row[i]);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
    }
    return s;
}

// Shared-memory map of refine_decrypt_kernel (dynamic, up to 227 KB):
//   [region A: up to the next 64 KB boundary of the shared window][AES tables, 64 KB][region B]
// Every warp owns `rows` plaintext rows of row_bytes = 8*dim_pad + 16 (the 16-byte skew puts the same column of different
// rows into different bank groups); warps [0, warps_a) keep theirs in region A, the others in region B.
struct DecryptLayout { int rows, warps_a, row_bytes, u8_row_bytes, rec_bytes; size_t smem; };   // u8_row_bytes > 0: byte copies of the rows follow region B;
                                                                                             // rec_bytes > 0: rows are TMA-staged records (plaintext at +16)

__global__ void __launch_bounds__(RF_THREADS, 1) refine_decrypt_kernel(const __grid_constant__ DevKeyRing ring, StoreView sv, RefineParams p, DecryptLayout lay) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    const uint32_t base_sa = (uint32_t)__cvta_generic_to_shared(rf_smem);
    const uint32_t pad = (0x10000u - (base_sa & 0xffffu)) & 0xffffu;
    uint32_t *tab = reinterpret_cast<uint32_t *>(rf_smem + pad);                         // [256][64]
    const int dim = sv.dim;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = lay.rows;
    if ((uint32_t)(lay.warps_a * R * lay.row_bytes) > pad) __trap();                      // launcher assumed a larger region A
    __shared__ __align__(8) uint64_t s_mbar[RF_WARPS];                                    // one TMA barrier per warp
    if (lane == 0) { mbar_init(&s_mbar[warp], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    uint32_t phase = 0;
    const int pt_off = lay.rec_bytes ? 16 : 0;                                            // staged rows: [iv | version | plaintext-in-place | tag]

    for (int i = tid; i < 256 * 64; i += RF_THREADS) {
        const uint32_t t0 = sv.te0[i >> 6];
        const uint32_t v = (i & 32) ? __funnelshift_r(t0, t0, 16) : t0;    // Te0 | Te2
        tab[i] = v;
        tab[i + 256 * 64] = __funnelshift_r(v, v, 8);                      // Te1 | Te3 in the next 64 KB
    }
    const int nkeys = ring.n;
    __syncthreads();

    const AesFast te{base_sa + pad + (uint32_t)lane * 4u};
    unsigned char *rows_base = warp < lay.warps_a ? rf_smem + (size_t)warp * R * lay.row_bytes
                                                  : rf_smem + pad + 0x20000u + (size_t)(warp - lay.warps_a) * R * lay.row_bytes;
    unsigned char *u8_base = rf_smem + pad + 0x20000u + (size_t)(RF_WARPS - lay.warps_a) * R * lay.row_bytes + (size_t)warp * R * lay.u8_row_bytes;
    const bool use_u8 = p.queries_u8 != nullptr && (dim & 15) == 0 && p.f32_exact[1] != 0;
    const bool use_f32 = p.queries_f32 != nullptr && (dim & 3) == 0 && p.f32_exact[0] != 0;
    const int n_uniq = p.totals[1];
    // records per chunk: 32 for a full batch; fewer when the batch names few distinct records (a small batch, or one shard of a sharded
    // store), so that every warp of the grid gets work instead of 32 records queueing behind one warp
    const int cs = max(1, min(32, n_uniq / (int)(gridDim.x * RF_WARPS)));
    const int n_chunks = (n_uniq + cs - 1) / cs;
    const bool ctr_fast = ((8 * dim + 15) >> 4) <= 254;       // every counter 2..c+1 fits in one byte

    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(&p.totals[2], 1);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const int base = chunk * cs;
        const int nrec = min(cs, n_uniq - base);

        // lane l fetches the header, the authentication verdict and the pair range of record l.  uniq[] is sorted by id and
        // pairs[] is grouped in id order, so the pairs of consecutive distinct records are contiguous.
        int32_t my_id = -1; int my_slot = -1; uint4 my_hdr = make_uint4(0, 0, 0, 0); int my_verdict = FSPANN_V_NO_KEY;
        int my_end = 0, my_off = 0;
        bool my_small = false;                               // record `lane` of the chunk holds only integers in [0, 255] (byte row valid)
        if (lane < nrec) {
            my_id = p.uniq[base + lane];
            my_hdr = __ldg(reinterpret_cast<const uint4 *>(sv.rec + (size_t)my_id * sv.rec_stride));
            for (int i = 0; i < nkeys; i++) if (ring.version[i] == (int32_t)my_hdr.w) my_slot = i;
            my_verdict = my_slot < 0 ? FSPANN_V_NO_KEY : (int)p.rec_verdict[base + lane];        // KRS:82-88 -> QSI:265-270
            my_off = p.uoff[base + lane]; my_end = p.uoff[base + lane + 1];
        }
        // lane l: the (IV, key)-only part of AES rounds 1-2 of record l (see CtrPre)
        CtrPre my_pre{0, 0, 0, 0, 0};
        if (ctr_fast) {
            for (int vi = 0; vi < nkeys; vi++)          // warp-uniform vi: round keys stay on the uniform datapath
                if (vi == my_slot && my_verdict == FSPANN_V_OK)
                    my_pre = aes256_ctr_precompute(te.K0, ring.rk[vi], bswap32(my_hdr.x), bswap32(my_hdr.y), bswap32(my_hdr.z));
        }

        for (int g0 = 0; g0 < nrec; g0 += R) {
            const int gn = min(R, nrec - g0);
            // ---- stage the group's authenticated records: one bulk copy (TMA) per record into its row, completion on the warp's mbarrier ----
            if (lay.rec_bytes) {
                const unsigned okmask = __ballot_sync(0xffffffffu, lane >= g0 && lane < g0 + gn && my_verdict == FSPANN_V_OK);
                if (okmask) {                                           // warp-uniform
                    if (lane == 0) {
                        fence_proxy_async();                            // the rows were last touched through the generic proxy
                        mbar_expect_tx(&s_mbar[warp], (uint32_t)(__popc(okmask) * lay.rec_bytes));
                    }
                    for (int i = 0; i < gn; i++) {
                        const int32_t id = __shfl_sync(0xffffffffu, my_id, g0 + i);
                        if (lane == 0 && ((okmask >> (g0 + i)) & 1u))
                            tma_load_1d(rows_base + (size_t)i * lay.row_bytes, sv.rec + (size_t)id * sv.rec_stride, (uint32_t)lay.rec_bytes, &s_mbar[warp]);
                    }
                    mbar_wait(&s_mbar[warp], phase);
                    phase ^= 1u;
                }
            }
            // ---- decrypt the group's records in their rows (plaintext never leaves shared memory) ----
            for (int i = 0; i < gn; i++) {
                const int r = g0 + i;
                const int32_t id = __shfl_sync(0xffffffffu, my_id, r);
                const int slot = __shfl_sync(0xffffffffu, my_slot, r);
                const int verdict = __shfl_sync(0xffffffffu, my_verdict, r);
                unsigned char *row = rows_base + (size_t)i * lay.row_bytes;
                double *pt_row = reinterpret_cast<double *>(row + pt_off);
                if (verdict == FSPANN_V_OK) {      // plaintext is produced only for authenticated records
                    // the loop counter is warp-uniform, so the round keys are read from the kernel-parameter constant bank
                    // through the uniform datapath instead of 60 shared-memory loads per block
                    bool finite = true;
                    if (ctr_fast) {
                        CtrPre pre;
                        pre.c0 = __shfl_sync(0xffffffffu, my_pre.c0, r); pre.d0 = __shfl_sync(0xffffffffu, my_pre.d0, r);
                        pre.d1 = __shfl_sync(0xffffffffu, my_pre.d1, r); pre.d2 = __shfl_sync(0xffffffffu, my_pre.d2, r);
                        pre.d3 = __shfl_sync(0xffffffffu, my_pre.d3, r);
                        bool small = false;
                        uint8_t *u8_row = (use_u8 && lay.u8_row_bytes) ? u8_base + (size_t)i * lay.u8_row_bytes : nullptr;
                        for (int vi = 0; vi < nkeys; vi++) {
                            if (vi != slot) continue;
                            if (lay.rec_bytes) finite = warp_decrypt_staged_ctr(dim, row, te.K0, ring.rk[vi], pre, lane, u8_row, &small);
                            else finite = warp_decrypt_record_ctr(sv, sv.rec + (size_t)id * sv.rec_stride, te.K0, ring.rk[vi], pre, pt_row, lane, u8_row, &small);
                        }
                        if (lane == r) my_small = small;
                    } else {
                        uint4 hdr;
                        hdr.x = __shfl_sync(0xffffffffu, my_hdr.x, r); hdr.y = __shfl_sync(0xffffffffu, my_hdr.y, r);
                        hdr.z = __shfl_sync(0xffffffffu, my_hdr.z, r); hdr.w = __shfl_sync(0xffffffffu, my_hdr.w, r);
                        for (int vi = 0; vi < nkeys; vi++)
                            if (vi == slot) finite = warp_decrypt_record(sv, sv.rec + (size_t)id * sv.rec_stride, hdr, 0, te, ring.rk[vi], pt_row, lane);
                    }
                    if (!finite && lane == r) my_verdict = FSPANN_V_NON_FINITE;                  // QSI:253
                }
            }
            if (lane >= g0 && lane < g0 + gn && my_verdict == FSPANN_V_OK) atomicOr(&p.touched[my_id >> 5], 1u << (my_id & 31));   // QSI:262
            __syncwarp();

            // ---- score the group's (query, rank) pairs: lanes = pairs, 32 at a time across record boundaries ----
            const int off0 = __shfl_sync(0xffffffffu, my_off, g0);
            const int offE = __shfl_sync(0xffffffffu, my_end, g0 + gn - 1);
            for (int j0 = off0; j0 < offE; j0 += 32) {
                const int j = j0 + lane;
                int i = 0;                                           // which record of the group owns pair j
                for (int t = 0; t + 1 < gn; t++) i += (j >= __shfl_sync(0xffffffffu, my_end, g0 + t)) ? 1 : 0;
                const int verdict = __shfl_sync(0xffffffffu, my_verdict, g0 + i);
                const bool small = __shfl_sync(0xffffffffu, my_small ? 1 : 0, g0 + i) != 0;
                const bool live = j < offE;
                const uint32_t pair = live ? p.pairs[j] : 0u;
                const bool ok = live && verdict == FSPANN_V_OK;
                const uint32_t q = ok ? pair_query(p, pair) : 0u;
                // Byte path, cooperative form (dim <= 128): EIGHT lanes read one pair's 128-byte query row and record row 16 bytes each -- one
                // coalesced line per pair instead of eight scattered 16-byte reads per lane (the kernel sits on the LSU; with two AES tables the
                // extra shuffles made this slower, with four tables the ALU pipe has the room) -- four pairs per step, partial sums reduced over
                // the 8 lanes, the pair's own lane picks its sum up.
                if (dim <= 128 && __all_sync(0xffffffffu, !ok || small)) {
                    uint32_t mine = 0;
                    const int sub = lane & 7;
#pragma unroll
                    for (int st = 0; st < 8; st++) {
                        const int src = st * 4 + (lane >> 3);
                        const uint32_t q_s = __shfl_sync(0xffffffffu, q, src);
                        const int i_s = __shfl_sync(0xffffffffu, i, src);
                        const bool ok_s = __shfl_sync(0xffffffffu, ok ? 1 : 0, src) != 0;
                        uint32_t part = 0;
                        if (ok_s && sub * 16 < dim) {
                            const uint4 qq = __ldg(reinterpret_cast<const uint4 *>(p.queries_u8 + (size_t)q_s * dim) + sub);
                            const uint4 vv = *(reinterpret_cast<const uint4 *>(u8_base + (size_t)i_s * lay.u8_row_bytes) + sub);
                            uint32_t d;
                            d = __vabsdiffu4(qq.x, vv.x); part = __dp4a(d, d, part);
                            d = __vabsdiffu4(qq.y, vv.y); part = __dp4a(d, d, part);
                            d = __vabsdiffu4(qq.z, vv.z); part = __dp4a(d, d, part);
                            d = __vabsdiffu4(qq.w, vv.w); part = __dp4a(d, d, part);
                        }
                        part += __shfl_xor_sync(0xffffffffu, part, 1);
                        part += __shfl_xor_sync(0xffffffffu, part, 2);
                        part += __shfl_xor_sync(0xffffffffu, part, 4);
                        const uint32_t r = __shfl_sync(0xffffffffu, part, (lane & 3) * 8);
                        if ((lane >> 2) == st) mine = r;
                    }
                    if (ok) p.dist[pair] = __dsqrt_rn((double)mine);
                } else if (ok) {
                    double d2;
                    if (small) d2 = pair_dist2_bytes(p, dim, u8_base + (size_t)i * lay.u8_row_bytes, q);
                    else d2 = pair_dist2(p, use_u8, use_f32, dim, reinterpret_cast<const double *>(rows_base + (size_t)i * lay.row_bytes + pt_off), q);
                    p.dist[pair] = __dsqrt_rn(d2);
                }
                if (live) p.verdict[pair] = (uint8_t)verdict;
            }
            __syncwarp();
        }
    }
}

// Picks the number of plaintext rows per warp that fits (see DecryptLayout).
static bool decrypt_layout(int dim, int64_t rec_stride, DecryptLayout &lay) {
    const int dim_pad = (dim + 1) & ~1;
    const bool ctr_fast = ((8 * dim + 15) >> 4) <= 254;
    lay.rec_bytes = ctr_fast ? (int)rec_stride : 0;          // TMA staging needs the in-place CTR path
    lay.row_bytes = lay.rec_bytes ? lay.rec_bytes + 16 : 8 * dim_pad + 16;   // + 16: the same column of different rows lands in different bank groups
    const int a_avail = 0x10000 - 2048;                       // region A: the window starts with <= 2 KB of reserved / static memory
    const int b_avail = 227 * 1024 - 3 * 0x10000 - 64;
    lay.u8_row_bytes = (dim & 15) == 0 ? dim + 16 : 0;       // byte rows (integer-valued data), 16-byte skew like the FP64 rows
    for (int rows = 8; rows >= 1; rows--) {
        const int per_warp = rows * lay.row_bytes;
        const int wa = std::min(RF_WARPS, a_avail / per_warp), wb = RF_WARPS - wa;
        const int u8_total = RF_WARPS * rows * lay.u8_row_bytes;
        if (wb * per_warp + u8_total <= b_avail) {
            lay.rows = rows; lay.warps_a = wa;
            lay.smem = (size_t)3 * 0x10000 + (size_t)wb * per_warp + (size_t)u8_total + 16;
            return true;
        }
    }
    return false;
}

int configure_refine_kernels() {   // per-device opt-in to > 48 KB of dynamic shared memory (see opt_in_smem)
    if (opt_in_smem(gcm_tag_kernel) || opt_in_smem(refine_verify_kernel) || opt_in_smem(refine_decrypt_kernel) || opt_in_smem(refine_group_small_kernel)) return -1;
    return 0;
}

int launch_refine_decrypt(cudaStream_t s, const StoreView &sv, const RefineParams &p, int sm_count) {
    DecryptLayout lay;
    if (!decrypt_layout(sv.dim, sv.rec_stride, lay)) return -1;
    refine_decrypt_kernel<<<sm_count, RF_THREADS, lay.smem, s>>>(*sv.keys_host, sv, p, lay);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Compact copies of the query batch (FP32 and uint8) + flags telling whether every value survived the round trip exactly:
// exact[0] for FP32, exact[1] for uint8 (integers 0..255).
// One warp per query row; qfinite[q] = 1 when every value of the row is finite (isValid, QSI:407-413).
__global__ void queries_compact_kernel(const double *__restrict__ q, float *__restrict__ out32, uint8_t *__restrict__ out8, int64_t Q, int dim,
                                       int32_t *exact, uint8_t *__restrict__ qfinite) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool ok32 = true, ok8 = true, fin_all = true;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < Q; row += warps) {
        bool fin = true;
        for (int i = lane; i < dim; i += 32) {
            const int64_t at = row * dim + i;
            const double v = q[at];
            const float f = (float)v;
            out32[at] = f;
            ok32 &= (double)f == v;
            const bool in8 = v >= 0.0 && v <= 255.0;
            const int b = in8 ? (int)v : 0;
            out8[at] = (uint8_t)b;
            ok8 &= in8 && (double)b == v;
            fin &= (((unsigned long long)__double_as_longlong(v) >> 52) & 0x7ffull) != 0x7ffull;
        }
        fin = __all_sync(0xffffffffu, fin);
        if (lane == 0) qfinite[row] = fin ? 1 : 0;
        fin_all &= fin;
    }
    if (!__all_sync(0xffffffffu, ok32) && lane == 0) atomicAnd(&exact[0], 0);
    if (!__all_sync(0xffffffffu, ok8) && lane == 0) atomicAnd(&exact[1], 0);
    if (!__all_sync(0xffffffffu, fin_all) && lane == 0) atomicAnd(&exact[2], 0);
}
// exact[0] / exact[1]: the FP32 / uint8 copy is value-identical; exact[2]: every value is finite.
int launch_queries_to_f32(cudaStream_t s, const double *q, float *out, uint8_t *out8, int64_t Q, int dim, int32_t *exact, uint8_t *qfinite) {
    if (Q <= 0) return 0;
    if (cudaMemsetAsync(exact, 0xff, 4 * sizeof(int32_t), s) != cudaSuccess) return -1;
    int grid = (int)((Q * 32 + 255) / 256); { const int cap = cur_sm_count() * 8; if (grid > cap) grid = cap; }
    queries_compact_kernel<<<grid, 256, 0, s>>>(q, out, out8, Q, dim, exact, qfinite);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Adaptive retry decision on the device (QSI:327-337 + 444-447): rows[] = the queries, IN INCREASING ORDER, whose first pass decrypted
// something but returned < k results or decrypted < 10*k candidates; out[0] = their number, out[1] = 1 when some query of the batch
// holds NaN/Inf.  One 1024-thread block, ordered compaction (the sharded search needs the same row order on every rank).
__global__ void __launch_bounds__(1024) retry_select_kernel(int64_t Q, int k, const int32_t *__restrict__ n_ret, const int32_t *__restrict__ n_dec,
                                                            const int32_t *__restrict__ exact, int32_t *__restrict__ rows, int32_t *__restrict__ out) {
    __shared__ int s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (Q + 1023) / 1024, lo = min(Q, tid * per), hi = min(Q, lo + per);    // thread t owns rows [lo, hi): order = thread order
    int mine = 0;
    for (int64_t q = lo; q < hi; q++) { const int nd = n_dec[q]; mine += nd > 0 && (n_ret[q] < k || nd < 10 * k); }   // QSI:293: nothing decrypted -> no retry
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int before = incl - mine, total = 0;
    for (int w = 0; w < 32; w++) { const int c = s_w[w]; if (w < warp) before += c; total += c; }
    if (mine) for (int64_t q = lo; q < hi; q++) { const int nd = n_dec[q]; if (nd > 0 && (n_ret[q] < k || nd < 10 * k)) rows[before++] = (int32_t)q; }
    if (tid == 0) { out[0] = total; out[1] = exact[2] == 0 ? 1 : 0; }
}
int launch_retry_select(cudaStream_t s, int64_t Q, int k, const int32_t *n_ret, const int32_t *n_dec, const int32_t *exact, int32_t *rows, int32_t *out) {
    if (Q <= 0) return 0;
    retry_select_kernel<<<1, 1024, 0, s>>>(Q, k, n_ret, n_dec, exact, rows, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// per-query stable top-k (QSI:298-316): order by (Double.compare(dist), candidate order); distances are >= +0.0 so
// the IEEE bit pattern orders like the value.  One CTA per query, k rounds of block-wide arg-min.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 128;

__global__ void __launch_bounds__(TK_THREADS) refine_topk_kernel(RefineParams p) {
    __shared__ unsigned long long s_key[TK_THREADS / 32];
    __shared__ int s_rank[TK_THREADS / 32];
    __shared__ unsigned long long s_last_key;
    __shared__ int s_last_rank, s_ndec;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n_cand[q];
    const double *dist = p.dist + (size_t)q * p.stride;
    const uint8_t *ver = p.verdict + (size_t)q * p.stride;

    int ndec = 0;
    for (int r = tid; r < n; r += TK_THREADS) ndec += ver[r] == FSPANN_V_OK;
#pragma unroll
    for (int o = 16; o; o >>= 1) ndec += __shfl_xor_sync(0xffffffffu, ndec, o);
    if (tid == 0) { s_ndec = 0; s_last_rank = -1; s_last_key = 0ull; }
    __syncthreads();
    if (lane == 0) atomicAdd(&s_ndec, ndec);
    __syncthreads();
    const int total_ok = s_ndec;
    const int eff = min(p.k, total_ok);

    for (int round = 0; round < eff; round++) {
        const unsigned long long lk = s_last_key; const int lr = s_last_rank;
        unsigned long long best = ~0ull; int best_r = 0x7fffffff;
        for (int r = tid; r < n; r += TK_THREADS) {
            if (ver[r] != FSPANN_V_OK) continue;
            const unsigned long long key = (unsigned long long)__double_as_longlong(dist[r]);
            // strictly after the previously selected (key, rank)
            const bool after = round == 0 || key > lk || (key == lk && r > lr);
            if (after && (key < best || (key == best && r < best_r))) { best = key; best_r = r; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
            const int orr = __shfl_xor_sync(0xffffffffu, best_r, o);
            if (ok < best || (ok == best && orr < best_r)) { best = ok; best_r = orr; }
        }
        if (lane == 0) { s_key[warp] = best; s_rank[warp] = best_r; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = s_key[0]; int br = s_rank[0];
            for (int w = 1; w < TK_THREADS / 32; w++) if (s_key[w] < b || (s_key[w] == b && s_rank[w] < br)) { b = s_key[w]; br = s_rank[w]; }
            s_last_key = b; s_last_rank = br;
            p.topk_ids[(size_t)q * p.k + round] = p.cand_ids[(size_t)q * p.stride + br];
            p.topk_dist[(size_t)q * p.k + round] = __longlong_as_double((long long)b);
            if (p.topk_rank) p.topk_rank[(size_t)q * p.k + round] = p.rank_map ? p.rank_map[(size_t)q * p.stride + br] : br;
        }
        __syncthreads();
    }
    for (int i = eff + tid; i < p.k; i += TK_THREADS) {
        p.topk_ids[(size_t)q * p.k + i] = -1;
        p.topk_dist[(size_t)q * p.k + i] = __longlong_as_double(0x7ff8000000000000ll);
        if (p.topk_rank) p.topk_rank[(size_t)q * p.k + i] = 0x7fffffff;
    }
    if (tid == 0) { p.n_ret[q] = eff; if (p.n_dec) p.n_dec[q] = total_ok; }
}

// Register-resident form, ONE WARP per query, for candidate lists of <= 32 * E entries (B <= 1024 with E = 32): every lane loads its E
// (verdict, distance) entries ONCE (rank r = lane + 32 * j, coalesced; all verdict bytes first, then the distances of the decrypted
// ones), then k rounds of: lane-local minimum -> warp shuffle minimum -> the owner lane retires the selected entry.  No shared memory,
// no block barrier, no re-reads.  Ties on the distance keep the lower candidate rank (the reference's stable sort, QSI:298).
constexpr int TKW_WARPS = 4;
template <int E>
__global__ void __launch_bounds__(TKW_WARPS * 32) refine_topk_warp_kernel(RefineParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * TKW_WARPS + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int n = p.n_cand[q];
    const double *dist = p.dist + (size_t)q * p.stride;
    const uint8_t *ver = p.verdict + (size_t)q * p.stride;
    const int ne = (n + 31) >> 5;                                // entries per lane that can be live (warp-uniform): short lists (one shard of a
                                                                 // sharded store, ragged lists) skip the rest of the unrolled loops
    uint8_t vv[E];
#pragma unroll
    for (int j = 0; j < E; j++) { const int r = lane + 32 * j; vv[j] = (j < ne && r < n) ? ver[r] : (uint8_t)0xff; }
    unsigned long long key[E];
    int ndec = 0;
#pragma unroll
    for (int j = 0; j < E; j++) {
        const bool ok = j < ne && vv[j] == FSPANN_V_OK;
        key[j] = ok ? (unsigned long long)__double_as_longlong(dist[lane + 32 * j]) : ~0ull;     // dist >= +0.0: the bits order like the values
        ndec += ok;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ndec += __shfl_xor_sync(0xffffffffu, ndec, o);
    const int eff = min(p.k, ndec);
    // The selected (distance, position) of round t is parked in lane t & 31 and written out 32 results at a time: an id look-up + three
    // stores inside every round would put a dependent global load on the critical path of each of the k rounds.
    unsigned long long my_b = 0ull; int my_br = 0;
    auto flush = [&](int base, int cnt) {                             // results base .. base + cnt - 1 sit in lanes 0 .. cnt - 1
        if (lane < cnt) {
            const size_t o = (size_t)q * p.k + base + lane;
            p.topk_ids[o] = p.cand_ids[(size_t)q * p.stride + my_br];    // QSI:298-316
            p.topk_dist[o] = __longlong_as_double((long long)my_b);
            if (p.topk_rank) p.topk_rank[o] = p.rank_map ? p.rank_map[(size_t)q * p.stride + my_br] : my_br;
        }
    };
    // Small k: prune first.  Let T be the eff-th smallest of the 32 lane minima: at least eff entries are <= T, so every entry of the true
    // top-eff is <= T as well.  Typically ~eff..2*eff of the up to 1024 entries survive; they are compacted (ordered ballots) into <= 64
    // slots = two per lane, and the eff selection rounds then cost two compares + one warp reduction each instead of a scan of E entries.
    bool pruned = false;
    if (eff > 0 && eff <= 16) {
        __shared__ unsigned long long s_pk[TKW_WARPS][64];
        __shared__ int s_pr[TKW_WARPS][64];
        const int w = threadIdx.x >> 5;
        unsigned long long lmin = ~0ull;
#pragma unroll
        for (int j = 0; j < E; j++) if (j < ne) lmin = min(lmin, key[j]);
        unsigned long long x = lmin;                              // bitonic sort of the lane minima across the warp
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, jj);
                const bool lower = (lane & jj) == 0, asc = (lane & kk) == 0;
                x = (lower == asc) ? min(x, y) : max(x, y);
            }
        }
        const unsigned long long T = __shfl_sync(0xffffffffu, x, eff - 1);
        int run = 0;
#pragma unroll
        for (int j = 0; j < E; j++) {
            if (j < ne) {
                const bool keep = key[j] <= T && key[j] != ~0ull;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) { const int o = run + __popc(bal & ((1u << lane) - 1u)); if (o < 64) { s_pk[w][o] = key[j]; s_pr[w][o] = lane + 32 * j; } }
                run += __popc(bal);
            }
        }
        if (run <= 64) {                                          // warp-uniform
            __syncwarp();
            unsigned long long k0 = lane < run ? s_pk[w][lane] : ~0ull, k1 = lane + 32 < run ? s_pk[w][lane + 32] : ~0ull;
            const int r0 = lane < run ? s_pr[w][lane] : 0x7fffffff, r1 = lane + 32 < run ? s_pr[w][lane + 32] : 0x7fffffff;
            for (int round = 0; round < eff; round++) {
                const bool first = k0 < k1 || (k0 == k1 && r0 < r1);
                unsigned long long b = first ? k0 : k1; int br = first ? r0 : r1;
                const unsigned long long mine = b; const int mine_r = br;
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, b, o);
                    const int orr = __shfl_xor_sync(0xffffffffu, br, o);
                    if (ok < b || (ok == b && orr < br)) { b = ok; br = orr; }
                }
                if (br == mine_r && b == mine) { if (first) k0 = ~0ull; else k1 = ~0ull; }
                if (lane == (round & 31)) { my_b = b; my_br = br; }
            }
            flush(0, eff);
            pruned = true;
        }
    }
    for (int round = 0; !pruned && round < eff; round++) {
        unsigned long long best = ~0ull; int bj = 0;
#pragma unroll
        for (int j = 0; j < E; j++) if (j < ne && key[j] < best) { best = key[j]; bj = j; }   // ties: the lower rank (smaller j) stays
        int best_r = lane + 32 * bj;
        unsigned long long b = best; int br = best_r;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, b, o);
            const int orr = __shfl_xor_sync(0xffffffffu, br, o);
            if (ok < b || (ok == b && orr < br)) { b = ok; br = orr; }
        }
        if (br == best_r && b == best) {                              // the owner lane retires the entry
#pragma unroll
            for (int j = 0; j < E; j++) if (j < ne) key[j] = (j == bj) ? ~0ull : key[j];
        }
        if (lane == (round & 31)) { my_b = b; my_br = br; }
        if ((round & 31) == 31) flush(round - 31, 32);
    }
    if (!pruned && (eff & 31)) flush(eff & ~31, eff & 31);
    for (int i = eff + lane; i < p.k; i += 32) {
        p.topk_ids[(size_t)q * p.k + i] = -1;
        p.topk_dist[(size_t)q * p.k + i] = __longlong_as_double(0x7ff8000000000000ll);
        if (p.topk_rank) p.topk_rank[(size_t)q * p.k + i] = 0x7fffffff;
    }
    if (lane == 0) { p.n_ret[q] = eff; if (p.n_dec) p.n_dec[q] = ndec; }
}

// Candidate lists longer than 1024 (e.g. the reference's SIFT_P6 operating point, B = 16000, k = 100): exact radix SELECT of the k-th
// smallest (distance bits, rank) instead of k block-wide arg-min rounds.  One 256-thread CTA per query:
//   1. eight 8-bit levels (most significant first) over the IEEE bit patterns of the decrypted candidates' distances -> the longest
//      prefix P of the k-th smallest key and how many keys are smaller (warp-aggregated shared histograms);
//   2. one ORDERED pass over the list: keys below P are taken, keys equal to P in candidate order until k are taken (the reference's
//      stable sort keeps the earlier candidate on a tie, QSI:298) -> <= k (key, rank) pairs, in candidate order;
//   3. bitonic sort of those on (key, rank), write-out.
constexpr int TS_THREADS = 256;
constexpr int TS_MAXK = 2048;

__global__ void __launch_bounds__(TS_THREADS) refine_topk_select_kernel(RefineParams p) {
    __shared__ int32_t s_hist[256];
    __shared__ int32_t s_w[TS_THREADS / 32 + 1], s_w2[TS_THREADS / 32 + 1];
    __shared__ int32_t s_pick[3], s_cnt;
    __shared__ unsigned long long s_key[TS_MAXK];
    __shared__ int32_t s_rank[TS_MAXK];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n_cand[q];
    const double *dist = p.dist + (size_t)q * p.stride;
    const uint8_t *ver = p.verdict + (size_t)q * p.stride;

    int ndec = 0;
    for (int r = tid; r < n; r += TS_THREADS) ndec += ver[r] == FSPANN_V_OK;
#pragma unroll
    for (int o = 16; o; o >>= 1) ndec += __shfl_xor_sync(0xffffffffu, ndec, o);
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (lane == 0) atomicAdd(&s_cnt, ndec);
    __syncthreads();
    const int total_ok = s_cnt;
    const int eff = min(p.k, total_ok);
    __syncthreads();

    // ---- 0. prune (k <= 256): with T = the eff-th smallest of the 256 per-thread minima at least eff keys are <= T, so the true top-eff is
    //         among the keys <= T -- typically a few hundred of the thousands of candidates.  When they fit the sort buffer the radix
    //         levels (eight passes over the whole list) are skipped: one ordered pass collects them and the sort below finishes. ----
    bool pruned = false;
    unsigned long long T = 0ull;
    int n_sorted = eff;
    if (eff > 0 && eff <= TS_THREADS) {
        unsigned long long x = ~0ull;
        for (int r = tid; r < n; r += TS_THREADS)
            if (ver[r] == FSPANN_V_OK) x = min(x, (unsigned long long)__double_as_longlong(dist[r]));
#pragma unroll 1
        for (int kk = 2; kk <= TS_THREADS; kk <<= 1) {            // bitonic sort of the minima, one per thread
#pragma unroll 1
            for (int j = kk >> 1; j > 0; j >>= 1) {
                unsigned long long y;
                if (j >= 32) { __syncthreads(); s_key[tid] = x; __syncthreads(); y = s_key[tid ^ j]; }
                else y = __shfl_xor_sync(0xffffffffu, x, j);
                const bool lower = (tid & j) == 0, asc = (tid & kk) == 0;
                x = (lower == asc) ? min(x, y) : max(x, y);
            }
        }
        __syncthreads();
        s_key[tid] = x;
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        T = s_key[eff - 1];
        __syncthreads();
        if (T != ~0ull) {
            int c = 0;
            for (int r = tid; r < n; r += TS_THREADS)
                c += ver[r] == FSPANN_V_OK && (unsigned long long)__double_as_longlong(dist[r]) <= T;
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) atomicAdd(&s_cnt, c);
            __syncthreads();
            const int S = s_cnt;
            __syncthreads();
            if (S <= TS_MAXK) { pruned = true; n_sorted = S; }
        }
    }
    if (eff > 0) {
        // ---- 1. radix select: prefix of the eff-th smallest key ----
        unsigned long long prefix = pruned ? T : 0ull;
        int cum = 0, used = pruned ? 64 : 0;
        for (; used < 64; used += 8) {
            const int shift = 56 - used;
            s_hist[tid] = 0;
            __syncthreads();
            for (int r0 = 0; r0 < n; r0 += TS_THREADS) {
                const int r = r0 + tid;
                bool in = false; int bin = 0;
                if (r < n && ver[r] == FSPANN_V_OK) {
                    const unsigned long long key = (unsigned long long)__double_as_longlong(dist[r]);
                    in = used == 0 || (key >> (shift + 8)) == prefix;
                    bin = (int)((key >> shift) & 0xffull);
                }
                const unsigned act = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const unsigned peers = __match_any_sync(act, bin);
                    if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[bin], __popc(peers));
                }
            }
            __syncthreads();
            if (warp == 0) {                                   // first bin d with cum + sum(hist[0..d]) >= eff
                int loc[8], sum = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) { loc[u] = s_hist[lane * 8 + u]; sum += loc[u]; }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                const int excl = cum + incl - sum;
                const unsigned hit = __ballot_sync(0xffffffffu, excl + sum >= eff);
                const int owner = __ffs(hit) - 1;              // total_ok >= eff: some lane hits
                if (lane == owner) {
                    int c = excl, u = 0;
                    for (; u < 7; u++) { if (c + loc[u] >= eff) break; c += loc[u]; }
                    s_pick[0] = lane * 8 + u; s_pick[1] = c; s_pick[2] = loc[u];
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned long long)s_pick[0];
            cum = s_pick[1];
            const int bin_n = s_pick[2];
            __syncthreads();
            if (cum + bin_n == eff) { used += 8; break; }      // the whole bin is selected: no finer split needed
        }
        const int pshift = 64 - used;                          // keys compare through their top `used` bits
        const int need_tie = pruned ? 0x7fffffff : eff - cum;  // pruned: every key <= T is collected, the sort picks the first eff
        // ---- 2. ordered pass: below the prefix -> taken; equal -> taken in candidate order until need_tie ----
        int run_sel = 0, run_tie = 0;                          // block-uniform running counts
        for (int r0 = 0; r0 < n; r0 += TS_THREADS) {
            const int r = r0 + tid;
            bool less = false, tie = false; unsigned long long key = 0ull;
            if (r < n && ver[r] == FSPANN_V_OK) {
                key = (unsigned long long)__double_as_longlong(dist[r]);
                const unsigned long long kp = pshift >= 64 ? 0ull : key >> pshift;
                less = kp < prefix; tie = kp == prefix;
            }
            const unsigned bt = __ballot_sync(0xffffffffu, tie), bl = __ballot_sync(0xffffffffu, less);
            if (lane == 0) { s_w[warp] = __popc(bt); s_w2[warp] = __popc(bl); }
            __syncthreads();
            int tie_before = run_tie, tie_total = 0, less_before = 0, less_total = 0;
            for (int w = 0; w < TS_THREADS / 32; w++) {
                const int ct = s_w[w], cl = s_w2[w];
                if (w < warp) { tie_before += ct; less_before += cl; }
                tie_total += ct; less_total += cl;
            }
            const unsigned lt = (1u << lane) - 1u;
            const int my_tie_ord = tie_before + __popc(bt & lt);
            const bool take = less || (tie && my_tie_ord < need_tie);
            // slot: ordered over (less | taken ties) -- ties taken before this thread in the chunk: min(ordinal, need_tie) bookkeeping
            const unsigned btake = __ballot_sync(0xffffffffu, take);
            __syncthreads();
            if (lane == 0) s_w[warp] = __popc(btake);
            __syncthreads();
            int before = run_sel, tot = 0;
            for (int w = 0; w < TS_THREADS / 32; w++) { const int c = s_w[w]; if (w < warp) before += c; tot += c; }
            if (take) { const int slot = before + __popc(btake & lt); s_key[slot] = key; s_rank[slot] = r; }
            run_sel += tot; run_tie += tie_total;
            (void)less_before; (void)less_total;
            __syncthreads();
        }
        // ---- 3. bitonic sort of the collected pairs on (key, rank) ----
        int m = 1; while (m < n_sorted) m <<= 1;
        for (int i = n_sorted + tid; i < m; i += TS_THREADS) { s_key[i] = ~0ull; s_rank[i] = 0x7fffffff; }
        __syncthreads();
        for (int kk = 2; kk <= m; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < m; i += TS_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = s_key[i], b = s_key[ixj];
                        const int ra = s_rank[i], rb = s_rank[ixj];
                        const bool gt = a > b || (a == b && ra > rb);
                        if (gt == ((i & kk) == 0)) { s_key[i] = b; s_key[ixj] = a; s_rank[i] = rb; s_rank[ixj] = ra; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < eff; i += TS_THREADS) {
            const int br = s_rank[i];
            p.topk_ids[(size_t)q * p.k + i] = p.cand_ids[(size_t)q * p.stride + br];
            p.topk_dist[(size_t)q * p.k + i] = __longlong_as_double((long long)s_key[i]);
            if (p.topk_rank) p.topk_rank[(size_t)q * p.k + i] = p.rank_map ? p.rank_map[(size_t)q * p.stride + br] : br;
        }
    }
    for (int i = eff + tid; i < p.k; i += TS_THREADS) {
        p.topk_ids[(size_t)q * p.k + i] = -1;
        p.topk_dist[(size_t)q * p.k + i] = __longlong_as_double(0x7ff8000000000000ll);
        if (p.topk_rank) p.topk_rank[(size_t)q * p.k + i] = 0x7fffffff;
    }
    if (tid == 0) { p.n_ret[q] = eff; if (p.n_dec) p.n_dec[q] = total_ok; }
}

// Database-sharded search: keeps, per query, only the candidates whose id lies in this shard's range [id_lo, id_hi), in their original
// order, with their original positions (rank_out) -- the grouping passes and the per-shard top-k then look at ~B/W slots per query instead
// of scanning all B for the ones they own.  Ids outside every shard (negative, >= n_global) only ever produce "not found" verdicts, which
// no output of the sharded search reports, so they are dropped too.  One warp per query, ordered (ballot) compaction.
__global__ void shard_compact_kernel(int64_t Q, int stride, const int32_t *__restrict__ cand, const int32_t *__restrict__ n_cand, int64_t id_lo,
                                     int64_t id_hi, int32_t *__restrict__ out_ids, int32_t *__restrict__ out_rank, int32_t *__restrict__ out_n) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int n = n_cand[q];
    const int32_t *src = cand + (size_t)q * stride;
    int32_t *dst = out_ids + (size_t)q * stride, *dstr = out_rank + (size_t)q * stride;
    int run = 0;
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int r = r0 + lane;
        const int32_t id = r < n ? src[r] : -1;
        const bool own = r < n && (int64_t)id >= id_lo && (int64_t)id < id_hi;
        const unsigned b = __ballot_sync(0xffffffffu, own);
        if (own) { const int o = run + __popc(b & ((1u << lane) - 1u)); dst[o] = id; dstr[o] = r; }
        run += __popc(b);
    }
    if (lane == 0) out_n[q] = run;
}
int launch_shard_compact(cudaStream_t s, int64_t Q, int stride, const int32_t *cand, const int32_t *n_cand, int64_t id_lo, int64_t id_hi,
                         int32_t *out_ids, int32_t *out_rank, int32_t *out_n) {
    if (Q <= 0) return 0;
    shard_compact_kernel<<<(unsigned)((Q * 32 + 255) / 256), 256, 0, s>>>(Q, stride, cand, n_cand, id_lo, id_hi, out_ids, out_rank, out_n);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_refine_topk(cudaStream_t s, const RefineParams &p) {
    if (p.Q <= 0) return 0;
    const unsigned grid = (unsigned)((p.Q + TKW_WARPS - 1) / TKW_WARPS);
    if (p.stride <= 32 * 8) refine_topk_warp_kernel<8><<<grid, TKW_WARPS * 32, 0, s>>>(p);
    else if (p.stride <= 32 * 16) refine_topk_warp_kernel<16><<<grid, TKW_WARPS * 32, 0, s>>>(p);
    else if (p.stride <= 32 * 32) refine_topk_warp_kernel<32><<<grid, TKW_WARPS * 32, 0, s>>>(p);
    else if (p.k <= TS_MAXK) refine_topk_select_kernel<<<(unsigned)p.Q, TS_THREADS, 0, s>>>(p);
    else refine_topk_kernel<<<(unsigned)p.Q, TK_THREADS, 0, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// cross-shard merge (database-sharded store, BASELINE config 4): S per-shard top-k lists [S][Q][k] of (distance, candidate
// rank, id), padded with id = -1 -> the global top-k ordered by (distance, rank).  The rank is the candidate's position in the
// (identical on every shard) ordered candidate list, so this reproduces the reference's stable sort over all B candidates
// (QSI:298-316) bit for bit.  One warp per query, k rounds of warp-wide arg-min.
// ------------------------------------------------------------------------------------------------------------------
__global__ void merge_topk_kernel(int S, int64_t Q, int k, const double *__restrict__ dist, const int32_t *__restrict__ rank, const int32_t *__restrict__ ids,
                                  int32_t *__restrict__ out_ids, double *__restrict__ out_dist, int32_t *__restrict__ out_nret) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int n = S * k;
    unsigned long long lk = 0ull; int lr = -1;                       // last selected (key, rank)
    int got = 0;
    for (int round = 0; round < k; round++) {
        unsigned long long best = ~0ull; int best_r = 0x7fffffff, best_id = -1;
        for (int j = lane; j < n; j += 32) {
            const int s = j / k, i = j - s * k;
            const size_t at = ((size_t)s * Q + q) * k + i;
            const int32_t id = ids[at];
            if (id < 0) continue;
            const unsigned long long key = (unsigned long long)__double_as_longlong(dist[at]);
            const int r = rank[at];
            const bool after = round == 0 || key > lk || (key == lk && r > lr);
            if (after && (key < best || (key == best && r < best_r))) { best = key; best_r = r; best_id = id; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
            const int orr = __shfl_xor_sync(0xffffffffu, best_r, o), oi = __shfl_xor_sync(0xffffffffu, best_id, o);
            if (ok < best || (ok == best && orr < best_r)) { best = ok; best_r = orr; best_id = oi; }
        }
        if (best_id < 0) break;                                      // fewer than k results over all shards
        lk = best; lr = best_r; got++;
        if (lane == 0) { out_ids[q * k + round] = best_id; out_dist[q * k + round] = __longlong_as_double((long long)best); }
    }
    for (int i = got + lane; i < k; i += 32) { out_ids[q * k + i] = -1; out_dist[q * k + i] = __longlong_as_double(0x7ff8000000000000ll); }
    if (lane == 0) out_nret[q] = got;
}
int launch_merge_topk(cudaStream_t s, int S, int64_t Q, int k, const double *dist, const int32_t *rank, const int32_t *ids, int32_t *out_ids,
                      double *out_dist, int32_t *out_nret) {
    if (Q <= 0) return 0;
    merge_topk_kernel<<<(unsigned)((Q * 32 + 255) / 256), 256, 0, s>>>(S, Q, k, dist, rank, ids, out_ids, out_dist, out_nret);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
__global__ void counters_kernel(int64_t Q, const int32_t *raw, const int32_t *uniq, const int32_t *n_dec, const int32_t *n_ret,
                                const int32_t *n_cand, int32_t retried, int64_t *counters, const uint8_t *qfinite) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int64_t *c = counters + q * FSPANN_COUNTERS;
    if (qfinite && !qfinite[q]) { for (int i = 0; i < FSPANN_COUNTERS; i++) c[i] = 0; return; }   // QSI:137 returns before any lookup
    c[0] = raw ? raw[q] : 0; c[1] = uniq ? uniq[q] : 0; c[2] = n_dec[q]; c[3] = n_ret[q]; c[4] = retried; c[5] = n_cand[q];
}
int launch_counters(cudaStream_t s, int64_t Q, const int32_t *raw, const int32_t *uniq, const int32_t *n_dec, const int32_t *n_ret,
                    const int32_t *n_cand, int32_t retried, int64_t *counters, const uint8_t *qfinite) {
    if (Q <= 0) return 0;
    counters_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, s>>>(Q, raw, uniq, n_dec, n_ret, n_cand, retried, counters, qfinite);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// dst[i] = src[rows[i]] (gather) or dst[rows[i]] = src[i] (scatter); rows of row_bytes (multiple of 4)
__global__ void gather_rows_kernel(const uint32_t *src, uint32_t *dst, const int32_t *rows, int64_t n_rows, int64_t row_words, int scatter) {
    const int64_t total = n_rows * row_words;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / row_words, w = i - r * row_words;
        if (scatter) dst[(int64_t)rows[r] * row_words + w] = src[i];
        else dst[i] = src[(int64_t)rows[r] * row_words + w];
    }
}
int launch_gather_rows(cudaStream_t s, const void *src, void *dst, const int32_t *rows, int64_t n_rows, int64_t row_bytes, bool scatter) {
    if (n_rows <= 0) return 0;
    const int64_t words = row_bytes / 4, total = n_rows * words;
    int grid = (int)((total + 255) / 256); { const int cap = cur_sm_count() * 8; if (grid > cap) grid = cap; }
    gather_rows_kernel<<<grid, 256, 0, s>>>((const uint32_t *)src, (uint32_t *)dst, rows, n_rows, words, scatter ? 1 : 0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Packs host-layout arrays (iv[n][12], ct[n][8*dim+16], ver[n]) staged on the device into the record layout.
__global__ void store_pack_kernel(uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *ids, const uint8_t *iv,
                                  const uint8_t *ct, const int32_t *ver) {
    const int64_t ct_bytes = 8LL * dim + 16, words = (16 + ct_bytes) / 4;   // 8*dim+32 is a multiple of 8
    const int64_t total = n * words;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words, w = i - r * words;
        const int64_t dst_row = ids ? ids[r] : r;
        uint32_t v;
        if (w < 3) v = reinterpret_cast<const uint32_t *>(iv)[r * 3 + w];
        else if (w == 3) v = (uint32_t)ver[r];
        else v = reinterpret_cast<const uint32_t *>(ct)[r * (ct_bytes / 4) + (w - 4)];
        reinterpret_cast<uint32_t *>(rec + dst_row * rec_stride)[w] = v;
    }
}
int launch_store_pack(cudaStream_t s, uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *ids, const uint8_t *iv,
                      const uint8_t *ct, const int32_t *ver) {
    if (n <= 0) return 0;
    const int64_t total = n * ((32 + 8LL * dim) / 4);
    int grid = (int)((total + 255) / 256); { const int cap = cur_sm_count() * 16; if (grid > cap) grid = cap; }
    store_pack_kernel<<<grid, 256, 0, s>>>(rec, rec_stride, dim, n, ids, iv, ct, ver);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// Record encryption on the device:
//   Migrate  = KeyRotationServiceImpl.reencryptTouched (keymanagement/.../KeyRotationServiceImpl.java:215-289): decrypt under the
//              stored version, re-encrypt under the target version with a fresh IV, in place in the HBM store;
//   Encrypt  = AesGcmCryptoService.encryptToPoint (crypto/.../AesGcmCryptoService.java:55-112) for a batch of vectors.
// Both are AES-CTR passes (one warp per record, a lane per 16-byte block) followed by gcm_tag_kernel in write mode.  In Migrate
// the plaintext exists only as keystream XORs in registers: new_ct = old_ct ^ E_old(ctr) ^ E_new(ctr').
// ------------------------------------------------------------------------------------------------------------------
constexpr int XC_THREADS = 256;
constexpr int XC_WARPS = XC_THREADS / 32;

struct XcryptSmem {
    uint32_t *te_s, *s_rk; int32_t *s_ver; int nkeys;
    __device__ __forceinline__ void init(unsigned char *smem, const StoreView &sv) {
        te_s = reinterpret_cast<uint32_t *>(smem);
        s_rk = te_s + 256 * 32;
        s_ver = reinterpret_cast<int32_t *>(s_rk + kMaxKeys * 60);
        nkeys = sv.keys->n;
        for (int i = threadIdx.x; i < 256 * 32; i += XC_THREADS) te_s[i] = sv.te0[i >> 5];
        for (int i = threadIdx.x; i < nkeys * 60; i += XC_THREADS) s_rk[i] = sv.keys->rk[i / 60][i % 60];
        for (int i = threadIdx.x; i < kMaxKeys; i += XC_THREADS) s_ver[i] = i < nkeys ? sv.keys->version[i] : INT32_MIN;
        __syncthreads();
    }
};
static size_t xcrypt_smem_bytes() { return sizeof(uint32_t) * (256 * 32 + kMaxKeys * 60) + sizeof(int32_t) * kMaxKeys; }

// verdict[j]: authentication verdict of list entry j under its STORED version (gcm_tag_kernel, verify mode; NO_KEY preset).
// flag[j] <- 1 when the record was re-encrypted: stored version < target, key known, tag valid (KRS:243-279).
__global__ void __launch_bounds__(XC_THREADS) migrate_xcrypt_kernel(StoreView sv, int n, const int32_t *__restrict__ list,
                                                                    const uint8_t *__restrict__ fresh_iv, int32_t target_version,
                                                                    const uint8_t *__restrict__ verdict, uint8_t *__restrict__ flag) {
    extern __shared__ __align__(16) unsigned char xc_smem[];
    XcryptSmem sm; sm.init(xc_smem, sv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TeSmem te{sm.te_s + lane};
    const int slot_new = find_key_slot(sm.s_ver, sm.nkeys, target_version);
    const int nbytes = 8 * sv.dim, c = (nbytes + 15) >> 4;
    for (int j = blockIdx.x * XC_WARPS + warp; j < n; j += gridDim.x * XC_WARPS) {
        uint8_t *rec = const_cast<uint8_t *>(sv.rec) + (size_t)list[j] * sv.rec_stride;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(rec);
        const int slot_old = find_key_slot(sm.s_ver, sm.nkeys, (int32_t)hdr.w);
        const bool go = slot_new >= 0 && slot_old >= 0 && (int32_t)hdr.w < target_version && verdict[j] == FSPANN_V_OK;
        if (lane == 0) flag[j] = go ? 1 : 0;
        if (!go) continue;                                                   // warp-uniform
        const uint32_t *iv = reinterpret_cast<const uint32_t *>(fresh_iv + (size_t)j * 12);
        const uint32_t n0 = iv[0], n1 = iv[1], n2 = iv[2];
        const RkSmem rk_old{sm.s_rk + slot_old * 60}, rk_new{sm.s_rk + slot_new * 60};
        uint4 *ctv = reinterpret_cast<uint4 *>(rec + 16);
        for (int blk = lane; blk < c; blk += 32) {
            uint4 w = ctv[blk];
            uint32_t ka[4], kb[4];
            aes256_encrypt(te, rk_old, bswap32(hdr.x), bswap32(hdr.y), bswap32(hdr.z), (uint32_t)(blk + 2), ka);
            aes256_encrypt(te, rk_new, bswap32(n0), bswap32(n1), bswap32(n2), (uint32_t)(blk + 2), kb);
            w.x ^= bswap32(ka[0] ^ kb[0]); w.y ^= bswap32(ka[1] ^ kb[1]);
            if (!((blk == c - 1) && (nbytes & 15))) { w.z ^= bswap32(ka[2] ^ kb[2]); w.w ^= bswap32(ka[3] ^ kb[3]); }   // odd dim: the rest is tag
            ctv[blk] = w;
        }
        __syncwarp();
        if (lane == 0) *reinterpret_cast<uint4 *>(rec) = make_uint4(n0, n1, n2, (uint32_t)target_version);
    }
}

// Builds n records [iv | version | AES-CTR(big-endian FP64 vector) | (tag left for gcm_tag_kernel)] in `out` (record layout of
// the store, stride sv.rec_stride).  sv.rec must point at `out`; sv.dim is the vector length.
__global__ void __launch_bounds__(XC_THREADS) encrypt_xcrypt_kernel(StoreView sv, int n, const double *__restrict__ vectors,
                                                                    const uint8_t *__restrict__ ivs, int32_t version, uint8_t *__restrict__ flag) {
    extern __shared__ __align__(16) unsigned char xc_smem[];
    XcryptSmem sm; sm.init(xc_smem, sv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TeSmem te{sm.te_s + lane};
    const int slot = find_key_slot(sm.s_ver, sm.nkeys, version);
    const int dim = sv.dim, nbytes = 8 * dim, c = (nbytes + 15) >> 4;
    for (int j = blockIdx.x * XC_WARPS + warp; j < n; j += gridDim.x * XC_WARPS) {
        uint8_t *rec = const_cast<uint8_t *>(sv.rec) + (size_t)j * sv.rec_stride;
        if (lane == 0) flag[j] = slot >= 0 ? 1 : 0;
        if (slot < 0) continue;
        const uint32_t *iv = reinterpret_cast<const uint32_t *>(ivs + (size_t)j * 12);
        const uint32_t n0 = iv[0], n1 = iv[1], n2 = iv[2];
        const RkSmem rk{sm.s_rk + slot * 60};
        const double *v = vectors + (size_t)j * dim;
        uint4 *ctv = reinterpret_cast<uint4 *>(rec + 16);
        for (int blk = lane; blk < c; blk += 32) {
            uint32_t ks[4];
            aes256_encrypt(te, rk, bswap32(n0), bswap32(n1), bswap32(n2), (uint32_t)(blk + 2), ks);
            const double a = v[2 * blk];                                    // serializeVector: big-endian FP64 (AGC:240-259)
            uint4 w;
            w.x = bswap32((uint32_t)__double2hiint(a) ^ ks[0]); w.y = bswap32((uint32_t)__double2loint(a) ^ ks[1]);
            if (2 * blk + 1 < dim) {
                const double b = v[2 * blk + 1];
                w.z = bswap32((uint32_t)__double2hiint(b) ^ ks[2]); w.w = bswap32((uint32_t)__double2loint(b) ^ ks[3]);
            } else { w.z = 0; w.w = 0; }
            ctv[blk] = w;
        }
        if (lane == 0) *reinterpret_cast<uint4 *>(rec) = make_uint4(n0, n1, n2, (uint32_t)version);
    }
}

int launch_migrate_xcrypt(cudaStream_t s, const StoreView &sv, int n, const int32_t *list, const uint8_t *fresh_iv, int32_t target_version,
                          const uint8_t *verdict, uint8_t *flag, int sm_count) {
    if (n <= 0) return 0;
    const size_t smem = xcrypt_smem_bytes();
    int grid = (n + XC_WARPS - 1) / XC_WARPS; if (grid > sm_count * 4) grid = sm_count * 4;
    migrate_xcrypt_kernel<<<grid, XC_THREADS, smem, s>>>(sv, n, list, fresh_iv, target_version, verdict, flag);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_encrypt_xcrypt(cudaStream_t s, const StoreView &sv, int n, const double *vectors, const uint8_t *ivs, int32_t version, uint8_t *flag,
                          int sm_count) {
    if (n <= 0) return 0;
    const size_t smem = xcrypt_smem_bytes();
    int grid = (n + XC_WARPS - 1) / XC_WARPS; if (grid > sm_count * 4) grid = sm_count * 4;
    encrypt_xcrypt_kernel<<<grid, XC_THREADS, smem, s>>>(sv, n, vectors, ivs, version, flag);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Inverse of store_pack_kernel for listed records: rec[rows[r]] (or rec[r]) -> iv[r][12], ct[r][8*dim+16], ver[r].
__global__ void store_unpack_kernel(const uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *rows, uint8_t *iv, uint8_t *ct,
                                    int32_t *ver) {
    const int64_t ct_bytes = 8LL * dim + 16, words = (16 + ct_bytes) / 4;
    const int64_t total = n * words;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words, w = i - r * words;
        const int64_t src_row = rows ? rows[r] : r;
        const uint32_t v = reinterpret_cast<const uint32_t *>(rec + src_row * rec_stride)[w];
        if (w < 3) reinterpret_cast<uint32_t *>(iv)[r * 3 + w] = v;
        else if (w == 3) ver[r] = (int32_t)v;
        else reinterpret_cast<uint32_t *>(ct)[r * (ct_bytes / 4) + (w - 4)] = v;
    }
}
int launch_store_unpack(cudaStream_t s, const uint8_t *rec, int64_t rec_stride, int32_t dim, int64_t n, const int32_t *rows, uint8_t *iv,
                        uint8_t *ct, int32_t *ver) {
    if (n <= 0) return 0;
    const int64_t total = n * ((32 + 8LL * dim) / 4);
    int grid = (int)((total + 255) / 256); { const int cap = cur_sm_count() * 16; if (grid > cap) grid = cap; }
    store_unpack_kernel<<<grid, 256, 0, s>>>(rec, rec_stride, dim, n, rows, iv, ct, ver);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------------------------
// debug tap (only with -DFSPANN_DEBUG_TAP): plaintext of listed records to global memory, for the parity test that
// "decrypted plaintexts are bit-exact".  Not compiled into the production library.
// ------------------------------------------------------------------------------------------------------------------
#ifdef FSPANN_DEBUG_TAP
__global__ void __launch_bounds__(DBG_THREADS) debug_decrypt_kernel(StoreView sv, int64_t n, const int32_t *ids, double *pt, uint8_t *verdict) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    uint32_t *te_s = reinterpret_cast<uint32_t *>(rf_smem);
    uint32_t *s_rk = te_s + 256 * 32;
    int32_t *s_ver = reinterpret_cast<int32_t *>(s_rk + kMaxKeys * 60);
    double *pt_all = reinterpret_cast<double *>(s_ver + kMaxKeys);
    const int dim = sv.dim, dim_pad = (dim + 1) & ~1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256 * 32; i += DBG_THREADS) te_s[i] = sv.te0[i >> 5];
    const int nkeys = sv.keys->n;
    for (int i = tid; i < nkeys * 60; i += DBG_THREADS) s_rk[i] = sv.keys->rk[i / 60][i % 60];
    for (int i = tid; i < kMaxKeys; i += DBG_THREADS) s_ver[i] = i < nkeys ? sv.keys->version[i] : INT32_MIN;
    __syncthreads();
    const TeSmem te{te_s + lane};
    double *pt_row = pt_all + (size_t)warp * dim_pad;
    for (int64_t j = (int64_t)blockIdx.x * DBG_WARPS + warp; j < n; j += (int64_t)gridDim.x * DBG_WARPS) {
        const int32_t id = ids[j];
        int v;
        if (id < sv.id_base || id >= sv.id_base + sv.N || is_deleted(sv, id)) v = FSPANN_V_NOT_FOUND;
        else {
            const uint8_t *rec = sv.rec + (size_t)(id - sv.id_base) * sv.rec_stride;
            const uint4 hdr = __ldg(reinterpret_cast<const uint4 *>(rec));
            const int slot = find_key_slot(s_ver, nkeys, (int32_t)hdr.w);
            if (slot < 0) v = FSPANN_V_NO_KEY;
            else {
                // lane 0 authenticates with the Shoup table read from global memory (debug path, speed irrelevant)
                bool ok = false;
                if (lane == 0) ok = lane_verify_record(sv, rec, id, hdr, ShoupSmem{sv.shoup + (size_t)slot * 4096, 0}, te, RkSmem{s_rk + slot * 60});
                ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                if (!ok) v = FSPANN_V_TAG_FAIL;
                else v = warp_decrypt_record(sv, rec, hdr, slot, AesPlain{te}, s_rk, pt_row, lane) ? FSPANN_V_OK : FSPANN_V_NON_FINITE;
            }
        }
        __syncwarp();
        for (int i = lane; i < dim; i += 32) pt[j * dim + i] = (v == FSPANN_V_OK || v == FSPANN_V_NON_FINITE) ? pt_row[i] : __longlong_as_double(0x7ff8000000000000ll);
        if (lane == 0) verdict[j] = (uint8_t)v;
        __syncwarp();
    }
}
int launch_debug_decrypt(cudaStream_t s, const StoreView &sv, int64_t n, const int32_t *ids, double *pt, uint8_t *verdict) {
    const int dim_pad = (sv.dim + 1) & ~1;
    const size_t smem = sizeof(uint32_t) * (256 * 32 + kMaxKeys * 60) + sizeof(int32_t) * kMaxKeys + sizeof(double) * (size_t)DBG_WARPS * dim_pad;
    if (cudaFuncSetAttribute(debug_decrypt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int grid = (int)((n + DBG_WARPS - 1) / DBG_WARPS); { const int cap = cur_sm_count() * 2; if (grid > cap) grid = cap; } if (grid < 1) grid = 1;
    debug_decrypt_kernel<<<grid, DBG_THREADS, smem, s>>>(sv, n, ids, pt, verdict);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
#else
int launch_debug_decrypt(cudaStream_t, const StoreView &, int64_t, const int32_t *, double *, uint8_t *) { return -2; }
#endif

}  // namespace fsp
