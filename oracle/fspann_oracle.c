/*
 * fspann_oracle.c -- CPU restatement of the FSPANN query hot path (TokenGen -> Route -> Refine).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libfspann_gpu.so) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests for Route/Refine outputs -- the
 * reference holds no golden vectors, known-answer tests or fixtures for this path and
 * cannot be executed here (no JVM).  What IS pinned: AES-256-GCM and HMAC-SHA256 (JDK SunJCE, not in the
 * reference tree; JDK 21) are NIST SP 800-38D / RFC 2104 standards -- this file uses OpenSSL's
 * implementation and tests/ check it against NIST GCM known-answer vectors, RFC 4231 HMAC vectors and
 * Python `cryptography`.  Everything else follows the Java sources line by line, including the
 * java.util.HashMap iteration order, java.util.PriorityQueue sift rules and List.sort stability the
 * reference's results depend on.
 *
 * All file:line citations are relative to /root/reference/fsp-anns-parent/ :
 *   Coding   = index/src/main/java/com/fspann/index/paper/Coding.java
 *   GFR      = index/src/main/java/com/fspann/index/paper/GFunctionRegistry.java
 *   GP       = index/src/main/java/com/fspann/index/paper/GreedyPartitioner.java
 *   PIS      = index/src/main/java/com/fspann/index/paper/PartitionedIndexService.java
 *   QSI      = query/src/main/java/com/fspann/query/service/QueryServiceImpl.java
 *   QTF      = query/src/main/java/com/fspann/query/core/QueryTokenFactory.java
 *   AGC      = crypto/src/main/java/com/fspann/crypto/AesGcmCryptoService.java
 *   EP       = common/src/main/java/com/fspann/common/EncryptedPoint.java
 *   KM       = keymanagement/src/main/java/com/fspann/key/KeyManager.java
 *   KRS      = keymanagement/src/main/java/com/fspann/key/KeyRotationServiceImpl.java
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <openssl/evp.h>
#include <openssl/hmac.h>

#define ORC_BLOCK 64 /* DEFAULT_GREEDY_BLOCK_SIZE, PIS:92 */

/* ------------------------------------------------------------------------------------------------
 * java.util.SplittableRandom (JDK): SplitMix64 with the default gamma.  Coding:136,189,342-347.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t seed, gamma; } orc_sr_t;
#define ORC_GOLDEN_GAMMA 0x9e3779b97f4a7c15ULL
static uint64_t orc_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void orc_sr_init(orc_sr_t *s, int64_t seed) { s->seed = (uint64_t)seed; s->gamma = ORC_GOLDEN_GAMMA; }
static uint64_t orc_sr_next_long(orc_sr_t *s) { s->seed += s->gamma; return orc_mix64(s->seed); }
static double orc_sr_next_double(orc_sr_t *s) { return (double)(orc_sr_next_long(s) >> 11) * 0x1.0p-53; }

/* Coding:342-347 nextGaussian (Box-Muller; libm log/cos stand in for Java Math.log/cos -- the
 * resulting arrays are treated as DATA by everything downstream, so last-ulp differences from a real
 * JVM do not affect parity of the hot path). */
static double orc_next_gaussian(orc_sr_t *s) {
    double u1 = orc_sr_next_double(s);
    if (u1 < 4.9e-324) u1 = 4.9e-324; /* Math.max(Double.MIN_VALUE, u1) */
    double u2 = orc_sr_next_double(s);
    double mag = sqrt(-2.0 * log(u1));
    return mag * cos(2.0 * M_PI * u2);
}

/* Coding:349-353 dot(): strictly sequential FP64, no contraction (Java forbids FMA fusion). */
static double orc_dot(const double *a, const double *b, int n) {
    double acc = 0.0; /* built with -ffp-contract=off, no -ffast-math: mul then add, index order */
    for (int i = 0; i < n; i++) acc += a[i] * b[i];
    return acc;
}

/* Coding:184-241 buildFromSample.  sample is n x d row-major. */
void orc_build_from_sample(const double *sample, int n, int d, int m, int lambda, int64_t seed,
                           double *alpha /* m*d */, double *r /* m */, double *omega /* m */) {
    (void)lambda;
    orc_sr_t rnd; orc_sr_init(&rnd, seed);
    for (int j = 0; j < m; j++) {
        double norm = 0.0;
        for (int i = 0; i < d; i++) { double v = orc_next_gaussian(&rnd); alpha[(size_t)j * d + i] = v; norm += v * v; }
        norm = sqrt(norm > 1e-12 ? norm : 1e-12);
        for (int i = 0; i < d; i++) alpha[(size_t)j * d + i] /= norm;
    }
    double *mn = (double *)malloc(sizeof(double) * m), *mx = (double *)malloc(sizeof(double) * m);
    for (int j = 0; j < m; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    for (int s = 0; s < n; s++)
        for (int j = 0; j < m; j++) {
            double y = orc_dot(sample + (size_t)s * d, alpha + (size_t)j * d, d);
            if (y < mn[j]) mn[j] = y;
            if (y > mx[j]) mx[j] = y;
        }
    for (int j = 0; j < m; j++) {
        double range = mx[j] - mn[j]; if (!(range > 1e-6)) range = 1e-6; /* Math.max(1e-6, range) */
        double om = range / 2.5;                                          /* OMEGA_DIVISOR, Coding:224 */
        if (!(om > 0)) om = 1e-3;
        omega[j] = om;
        r[j] = orc_sr_next_double(&rnd) * om;
    }
    free(mn); free(mx);
}

/* GFR:63-147 initialize + GFR:291-293 computeSeed: one GFunction per (table, division). */
void orc_registry_init(const double *sample, int n, int d, int m, int lambda, int64_t base_seed, int T, int D,
                       double *alpha /* T*D*m*d */, double *r /* T*D*m */, double *omega /* T*D*m */) {
    for (int t = 0; t < T; t++)
        for (int dv = 0; dv < D; dv++) {
            int64_t seed = base_seed + ((int64_t)t * 1000003LL) + dv;
            size_t g = (size_t)t * D + dv;
            orc_build_from_sample(sample, n, d, m, lambda, seed, alpha + g * m * d, r + g * m, omega + g * m);
        }
}

/* Java (int) cast of a double: NaN -> 0, saturating. */
static int32_t orc_d2i(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return INT32_MAX;
    if (x <= -2147483648.0) return INT32_MIN;
    return (int32_t)x;
}

/* Coding:250-258 H(v). */
void orc_H(const double *v, int d, int m, const double *alpha, const double *r, const double *omega, int32_t *out) {
    for (int j = 0; j < m; j++) {
        double y = orc_dot(v, alpha + (size_t)j * d, d) + r[j];
        double q = y / omega[j];
        out[j] = orc_d2i(floor(q));
    }
}

/* Coding:285-301 C(v): bit position pos = (lambda-1-i)*m + j holds bit i of (H[j]^0x80000000).
 * Output layout = java.util.BitSet.toLongArray(): bit p in word p>>6, bit p&63; W = ceil(m*lambda/64). */
void orc_C(const double *v, int d, int m, int lambda, const double *alpha, const double *r, const double *omega,
           uint64_t *code, int W) {
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * m);
    orc_H(v, d, m, alpha, r, omega, H);
    for (int w = 0; w < W; w++) code[w] = 0;
    int pos = 0;
    for (int i = lambda - 1; i >= 0; i--)
        for (int j = 0; j < m; j++) {
            uint32_t hj = (uint32_t)H[j] ^ 0x80000000u;
            if ((hj >> i) & 1u) code[pos >> 6] |= 1ULL << (pos & 63);
            pos++;
        }
    free(H);
}

/* QTF:98-131 / GFR:272-280: codes for every (t, d).  codes is [T*D][W]. */
void orc_code_all(const double *v, int d, int m, int lambda, int T, int D, const double *alpha, const double *r,
                  const double *omega, uint64_t *codes, int W) {
    for (int g = 0; g < T * D; g++)
        orc_C(v, d, m, lambda, alpha + (size_t)g * m * d, r + (size_t)g * m, omega + (size_t)g * m, codes + (size_t)g * W, W);
}

/* GP:87-96 computeKey: code bit i (i < 63) -> key bit 62-i. */
int64_t orc_compute_key(const uint64_t *code, int W) {
    int64_t v = 0;
    for (int i = 0; i < 63 && i < W * 64; i++)
        if ((code[i >> 6] >> (i & 63)) & 1ULL) v |= (int64_t)(1ULL << (62 - i));
    return v;
}

/* GP:78-82 hamming. */
int64_t orc_hamming(const uint64_t *a, const uint64_t *b, int W) {
    int64_t c = 0;
    for (int w = 0; w < W; w++) c += __builtin_popcountll(a[w] ^ b[w]);
    return c;
}

/* ------------------------------------------------------------------------------------------------
 * java.util.HashMap<String, ?> iteration order for decimal-string keys.
 *   String.hashCode: h = 31*h + c; HashMap.hash: h ^ (h >>> 16); index = hash & (cap-1);
 *   HashMap(int n): first table = tableSizeFor(n); resize (x2) when ++size > 0.75*cap; split keeps
 *   relative order; iteration = buckets ascending, insertion order inside a bucket.
 * Valid while no bin is ever treeified (chain length 9); max_chain_out reports the longest chain seen
 * at any table size so callers can assert that.
 * ---------------------------------------------------------------------------------------------- */
static uint32_t orc_java_string_hash_decimal(int64_t id) {
    char buf[24];
    int n = snprintf(buf, sizeof buf, "%lld", (long long)id);
    uint32_t h = 0;
    for (int i = 0; i < n; i++) h = 31u * h + (uint32_t)(unsigned char)buf[i];
    return h ^ (h >> 16);
}
uint32_t orc_java_hash_decimal(int64_t id) { return orc_java_string_hash_decimal(id); }

static uint32_t orc_table_size_for(int64_t cap) { /* HashMap.tableSizeFor */
    if (cap <= 1) return 1;
    uint32_t n = 1;
    while ((int64_t)n < cap && n < (1u << 30)) n <<= 1;
    return n;
}
uint32_t orc_hashmap_final_cap(int64_t initial_capacity, int64_t size) {
    uint32_t cap = orc_table_size_for(initial_capacity);
    if (initial_capacity == 0) cap = 1;
    while ((double)size > 0.75 * (double)cap && cap < (1u << 30)) cap <<= 1;
    return cap;
}

typedef struct { uint32_t bucket; uint32_t seq; } orc_bs_t;
static int orc_bs_cmp(const void *a, const void *b) {
    const orc_bs_t *x = (const orc_bs_t *)a, *y = (const orc_bs_t *)b;
    if (x->bucket != y->bucket) return x->bucket < y->bucket ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
/* keys[0..n) in insertion order (all distinct).  order_out[i] = index (into keys) of the i-th entry the
 * HashMap iterator yields.  Returns the longest chain that existed at any table size. */
int orc_hashmap_order(const int32_t *keys, int64_t n, int64_t initial_capacity, int32_t *order_out) {
    uint32_t cap = orc_hashmap_final_cap(initial_capacity, n);
    orc_bs_t *e = (orc_bs_t *)malloc(sizeof(orc_bs_t) * (size_t)(n ? n : 1));
    for (int64_t i = 0; i < n; i++) { e[i].bucket = orc_java_string_hash_decimal(keys[i]) & (cap - 1); e[i].seq = (uint32_t)i; }
    qsort(e, (size_t)n, sizeof(orc_bs_t), orc_bs_cmp);
    for (int64_t i = 0; i < n; i++) order_out[i] = (int32_t)e[i].seq;
    free(e);
    /* chain-length audit at every intermediate table size */
    int max_chain = 0;
    uint32_t c = orc_table_size_for(initial_capacity);
    if (initial_capacity == 0) c = 1;
    for (;;) {
        int64_t upto = n;
        int64_t thr = (int64_t)(0.75 * (double)c);
        if (c < cap && thr + 1 < upto) upto = thr + 1; /* entries present just before this table was outgrown */
        uint16_t *cnt = (uint16_t *)calloc(c, sizeof(uint16_t));
        for (int64_t i = 0; i < upto; i++) {
            uint32_t b = orc_java_string_hash_decimal(keys[i]) & (c - 1);
            if (++cnt[b] > max_chain) max_chain = cnt[b];
        }
        free(cnt);
        if (c >= cap) break;
        c <<= 1;
    }
    return max_chain;
}

/* ------------------------------------------------------------------------------------------------
 * GP:37-76 build, for one (table, division).
 *   ids[0..n) / codes[i*W..] in STAGED order (PIS:412-420 puts them into new HashMap<>(n) in that order),
 *   entries are then taken in HashMap iteration order (GP:45-48), stably sorted by key (GP:51), cut into
 *   blocks of ORC_BLOCK (GP:54-73).  Outputs: P = ceil(n/64) partitions.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t key; int32_t pos; int32_t idx; } orc_kp_t;
static int orc_kp_cmp(const void *a, const void *b) {
    const orc_kp_t *x = (const orc_kp_t *)a, *y = (const orc_kp_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos); /* stability */
}
int orc_partition_build(const int32_t *ids, const uint64_t *codes, int64_t n, int W, int64_t code_stride_words,
                        int64_t *min_key, int64_t *max_key, uint64_t *rep_code /* P*W */, int32_t *part_ids /* n */) {
    if (n <= 0) return 0;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int max_chain = orc_hashmap_order(ids, n, n, order);
    orc_kp_t *kp = (orc_kp_t *)malloc(sizeof(orc_kp_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) {
        int32_t src = order[i];
        kp[i].key = orc_compute_key(codes + (size_t)src * code_stride_words, W);
        kp[i].pos = (int32_t)i;
        kp[i].idx = src;
    }
    qsort(kp, (size_t)n, sizeof(orc_kp_t), orc_kp_cmp);
    int64_t p = 0;
    for (int64_t i = 0; i < n; i += ORC_BLOCK, p++) {
        int64_t end = i + ORC_BLOCK < n ? i + ORC_BLOCK : n;
        min_key[p] = kp[i].key;
        max_key[p] = kp[end - 1].key;
        int64_t mid = i + ((end - i - 1) >> 1);
        memcpy(rep_code + (size_t)p * W, codes + (size_t)kp[mid].idx * code_stride_words, sizeof(uint64_t) * W);
        for (int64_t j = i; j < end; j++) part_ids[j] = ids[kp[j].idx];
    }
    free(order); free(kp);
    return max_chain;
}

/* GP:101-130 findNearestPartition. */
int64_t orc_find_nearest(const int64_t *min_key, const int64_t *max_key, int64_t P, int64_t q) {
    if (P <= 0) return 0;
    int64_t lo = 0, hi = P - 1;
    while (lo <= hi) {
        int64_t mid = (int64_t)(((uint64_t)lo + (uint64_t)hi) >> 1);
        if (q < min_key[mid]) hi = mid - 1;
        else if (q > max_key[mid]) lo = mid + 1;
        else return mid;
    }
    if (lo <= 0) return 0;
    if (lo >= P) return P - 1;
    int64_t l = lo - 1, rr = lo;
    int64_t dl = q < min_key[l] ? min_key[l] - q : (q > max_key[l] ? q - max_key[l] : 0);
    int64_t dr = q < min_key[rr] ? min_key[rr] - q : (q > max_key[rr] ? q - max_key[rr] : 0);
    return dl <= dr ? l : rr;
}

/* ------------------------------------------------------------------------------------------------
 * Flat routing index (the same layout the C ABI uploads; see include/fspann_gpu.h).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t dim, T, D, m, lambda, W;
    int64_t N;              /* ids per (t,d) */
    int64_t P;              /* partitions per (t,d) = ceil(N/64) */
    const int64_t *min_key; /* [T*D][P] */
    const int64_t *max_key; /* [T*D][P] */
    const uint64_t *rep;    /* [T*D][P][W] */
    const int32_t *ids;     /* [T*D][N] */
    const uint8_t *deleted; /* [max_id+1] or NULL  (RDB:203-224 isDeleted) */
    int64_t n_deleted_flags;
} orc_index_t;

/* java.util.PriorityQueue<long[]> with Comparator.comparingLong(a -> a[1]) (PIS:643-644). */
typedef struct { int64_t idx, dist; } orc_pq_e;
typedef struct { orc_pq_e q[8]; int size; } orc_pq_t;
static void orc_pq_add(orc_pq_t *pq, orc_pq_e x) {
    int k = pq->size++;
    while (k > 0) { int parent = (k - 1) >> 1; if (x.dist >= pq->q[parent].dist) break; pq->q[k] = pq->q[parent]; k = parent; }
    pq->q[k] = x;
}
static orc_pq_e orc_pq_poll(orc_pq_t *pq) {
    orc_pq_e res = pq->q[0];
    int n = --pq->size;
    orc_pq_e x = pq->q[n];
    if (n > 0) {
        int k = 0, half = n >> 1;
        while (k < half) {
            int child = 2 * k + 1, right = child + 1;
            orc_pq_e c = pq->q[child];
            if (right < n && c.dist > pq->q[right].dist) c = pq->q[child = right];
            if (x.dist <= c.dist) break;
            pq->q[k] = c; k = child;
        }
        pq->q[k] = x;
    }
    return res;
}

/* HashMap<String,Long> bestScore emulation (PIS:619): chained table keyed by id, remembers the first
 * insertion sequence so the iteration order can be reproduced afterwards. */
typedef struct { int32_t id; int32_t score; int32_t seq; int32_t next; } orc_node_t;
typedef struct { int32_t *head; uint32_t mask; orc_node_t *nodes; int32_t n, cap_nodes; } orc_map_t;
static void orc_map_init(orc_map_t *m, int64_t max_nodes) {
    uint32_t sz = 1024; while (sz < 2 * (uint64_t)max_nodes) sz <<= 1;
    m->head = (int32_t *)malloc(sizeof(int32_t) * sz); memset(m->head, 0xff, sizeof(int32_t) * sz);
    m->mask = sz - 1; m->nodes = (orc_node_t *)malloc(sizeof(orc_node_t) * (size_t)(max_nodes + 1)); m->n = 0; m->cap_nodes = (int32_t)max_nodes;
}
static void orc_map_free(orc_map_t *m) { free(m->head); free(m->nodes); }
static orc_node_t *orc_map_get(orc_map_t *m, int32_t id) {
    uint32_t b = ((uint32_t)id * 2654435761u) & m->mask;
    for (int32_t i = m->head[b]; i >= 0; i = m->nodes[i].next) if (m->nodes[i].id == id) return &m->nodes[i];
    return NULL;
}
static void orc_map_put_new(orc_map_t *m, int32_t id, int32_t score) {
    uint32_t b = ((uint32_t)id * 2654435761u) & m->mask;
    orc_node_t *nd = &m->nodes[m->n];
    nd->id = id; nd->score = score; nd->seq = m->n; nd->next = m->head[b]; m->head[b] = m->n++;
}

typedef struct { int32_t score; uint32_t bucket; int32_t seq; int32_t id; } orc_ent_t;
static int orc_ent_cmp(const void *a, const void *b) { /* stable sort by score over HashMap iteration order */
    const orc_ent_t *x = (const orc_ent_t *)a, *y = (const orc_ent_t *)b;
    if (x->score != y->score) return x->score < y->score ? -1 : 1;
    if (x->bucket != y->bucket) return x->bucket < y->bucket ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

/* PIS:592-715 lookupCandidatesWithScores (+ PIS:726-753 collectPartitionOrdered).
 *   codes: [T*D][W] query codes; probes = effectiveMaxProbes() (PIS:880-888); hard_cap = HARD_CAP (PIS:612-615).
 *   out_ids/out_scores need room for max_out entries; returns number of unique candidates (all are written
 *   up to max_out).  raw_seen_out = lastRawVisited (PIS:703).  max_chain_out: longest bucket chain in the
 *   emulated HashMap (>= 9 would mean Java treeified and the order is not reproduced). */
int64_t orc_route(const orc_index_t *ix, const uint64_t *codes, int probes, int64_t hard_cap, int32_t *out_ids,
                  int32_t *out_scores, int64_t max_out, int64_t *raw_seen_out, int *max_chain_out) {
    const int TD = ix->T * ix->D, W = ix->W;
    const int64_t P = ix->P, N = ix->N;
    int64_t max_nodes = (int64_t)TD * (probes > 0 ? probes : 0) * ORC_BLOCK + ORC_BLOCK;
    if (max_nodes > hard_cap + ORC_BLOCK) max_nodes = hard_cap + ORC_BLOCK;
    if (max_nodes < ORC_BLOCK) max_nodes = ORC_BLOCK;
    orc_map_t best; orc_map_init(&best, max_nodes);
    int64_t raw_seen = 0;
    uint8_t *visited = (uint8_t *)malloc((size_t)(P > 0 ? P : 1));

    for (int g = 0; g < TD && best.n < hard_cap; g++) {   /* PIS:624,628 (t-major, d-minor == g order) */
        if (P <= 0) continue;
        const int64_t *mn = ix->min_key + (size_t)g * P, *mx = ix->max_key + (size_t)g * P;
        const uint64_t *rep = ix->rep + (size_t)g * P * W;
        const int32_t *ids = ix->ids + (size_t)g * N;
        const uint64_t *qb = codes + (size_t)g * W;
        int64_t qkey = orc_compute_key(qb, W);                 /* PIS:640 */
        int64_t center = orc_find_nearest(mn, mx, P, qkey);    /* PIS:641 */
        memset(visited, 0, (size_t)P);
        orc_pq_t pq; pq.size = 0;
        orc_pq_e ce = { center, orc_hamming(qb, rep + (size_t)center * W, W) };
        orc_pq_add(&pq, ce); visited[center] = 1;
        int used = 0;
        while (pq.size > 0 && used < probes && best.n < hard_cap) {   /* PIS:657-659 */
            orc_pq_e cur = orc_pq_poll(&pq);
            int64_t idx = cur.idx; used++;
            /* collectPartitionOrdered PIS:726-753 */
            int32_t pd = (int32_t)orc_hamming(qb, rep + (size_t)idx * W, W);
            int64_t b0 = idx * ORC_BLOCK, b1 = b0 + ORC_BLOCK < N ? b0 + ORC_BLOCK : N;
            for (int64_t j = b0; j < b1; j++) {
                int32_t id = ids[j];
                if (ix->deleted && id >= 0 && id < ix->n_deleted_flags && ix->deleted[id]) continue;
                orc_node_t *nd = orc_map_get(&best, id);
                if (!nd) { orc_map_put_new(&best, id, pd); raw_seen++; }
                else if (pd < nd->score) { nd->score = pd; raw_seen++; }
            }
            int64_t left = idx - 1, right = idx + 1;
            if (left >= 0 && !visited[left]) { visited[left] = 1; orc_pq_e e = { left, orc_hamming(qb, rep + (size_t)left * W, W) }; orc_pq_add(&pq, e); }
            if (right < P && !visited[right]) { visited[right] = 1; orc_pq_e e = { right, orc_hamming(qb, rep + (size_t)right * W, W) }; orc_pq_add(&pq, e); }
        }
    }
    free(visited);

    /* PIS:690-696: entries in HashMap iteration order, stable sort by score. */
    int64_t n = best.n;
    int64_t init_cap = hard_cap < 65536 ? hard_cap : 65536;      /* PIS:619 */
    uint32_t cap = orc_hashmap_final_cap(init_cap, n);
    orc_ent_t *ent = (orc_ent_t *)malloc(sizeof(orc_ent_t) * (size_t)(n ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        ent[i].score = best.nodes[i].score; ent[i].seq = best.nodes[i].seq; ent[i].id = best.nodes[i].id;
        ent[i].bucket = orc_java_string_hash_decimal(best.nodes[i].id) & (cap - 1);
    }
    if (max_chain_out) {
        int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1)), *ord = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
        for (int64_t i = 0; i < n; i++) keys[i] = best.nodes[i].id;
        *max_chain_out = orc_hashmap_order(keys, n, init_cap, ord);
        free(keys); free(ord);
    }
    qsort(ent, (size_t)n, sizeof(orc_ent_t), orc_ent_cmp);
    for (int64_t i = 0; i < n && i < max_out; i++) { out_ids[i] = ent[i].id; out_scores[i] = ent[i].score; }
    free(ent); orc_map_free(&best);
    if (raw_seen_out) *raw_seen_out = raw_seen;
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Crypto.  KM:221-237 (KDF), AGC:55-112 (encryptToPoint), AGC:126-166 (decryptFromPoint), EP:80-83 (AAD),
 * AGC:169-204 (query encrypt/decrypt, no AAD), AGC:240-277 (big-endian FP64 (de)serialisation).
 * ---------------------------------------------------------------------------------------------- */
void orc_kdf(const uint8_t master[32], int32_t version, uint8_t key_out[32]) {
    uint8_t salt[4] = { (uint8_t)(version >> 24), (uint8_t)(version >> 16), (uint8_t)(version >> 8), (uint8_t)version };
    unsigned int len = 32; uint8_t out[EVP_MAX_MD_SIZE];
    HMAC(EVP_sha256(), master, 32, salt, 4, out, &len);
    memcpy(key_out, out, 32);
}
void orc_hmac_sha256(const uint8_t *key, int klen, const uint8_t *msg, int mlen, uint8_t out32[32]) {
    unsigned int len = 32; uint8_t out[EVP_MAX_MD_SIZE];
    HMAC(EVP_sha256(), key, klen, msg, (size_t)mlen, out, &len);
    memcpy(out32, out, 32);
}

static void orc_serialize(const double *v, int d, uint8_t *out) {
    for (int i = 0; i < d; i++) { uint64_t b; memcpy(&b, &v[i], 8); for (int k = 0; k < 8; k++) out[i * 8 + k] = (uint8_t)(b >> (56 - 8 * k)); }
}
static void orc_deserialize(const uint8_t *in, int nbytes, double *v) {
    for (int i = 0; i < nbytes / 8; i++) { uint64_t b = 0; for (int k = 0; k < 8; k++) b |= (uint64_t)in[i * 8 + k] << (56 - 8 * k); memcpy(&v[i], &b, 8); }
}
int orc_aad(int64_t id, int32_t key_version, int32_t dim, char *out /* >= 64 */) {
    return snprintf(out, 64, "id:%lld|v:%d|d:%d", (long long)id, key_version, dim);
}

/* Generic AES-256-GCM (96-bit IV, 128-bit tag appended, Java doFinal layout).  Returns 0 on success. */
int orc_gcm_encrypt(const uint8_t key[32], const uint8_t iv[12], const uint8_t *aad, int aad_len, const uint8_t *pt,
                    int pt_len, uint8_t *ct_and_tag) {
    EVP_CIPHER_CTX *c = EVP_CIPHER_CTX_new(); int len = 0, ok = 1;
    ok &= EVP_EncryptInit_ex(c, EVP_aes_256_gcm(), NULL, NULL, NULL);
    ok &= EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, 12, NULL);
    ok &= EVP_EncryptInit_ex(c, NULL, NULL, key, iv);
    if (aad_len > 0) ok &= EVP_EncryptUpdate(c, NULL, &len, aad, aad_len);
    if (pt_len > 0) ok &= EVP_EncryptUpdate(c, ct_and_tag, &len, pt, pt_len);
    ok &= EVP_EncryptFinal_ex(c, ct_and_tag + pt_len, &len);
    ok &= EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_GET_TAG, 16, ct_and_tag + pt_len);
    EVP_CIPHER_CTX_free(c);
    return ok ? 0 : -1;
}
/* Returns 0 = ok, 3 = tag mismatch / any cipher failure (AGC:159-165 wraps everything in RuntimeException). */
int orc_gcm_decrypt(const uint8_t key[32], const uint8_t iv[12], const uint8_t *aad, int aad_len, const uint8_t *ct_and_tag,
                    int ct_len_with_tag, uint8_t *pt) {
    if (ct_len_with_tag < 16) return 3;
    int ct_len = ct_len_with_tag - 16, len = 0, ok = 1;
    EVP_CIPHER_CTX *c = EVP_CIPHER_CTX_new();
    ok &= EVP_DecryptInit_ex(c, EVP_aes_256_gcm(), NULL, NULL, NULL);
    ok &= EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, 12, NULL);
    ok &= EVP_DecryptInit_ex(c, NULL, NULL, key, iv);
    if (aad_len > 0) ok &= EVP_DecryptUpdate(c, NULL, &len, aad, aad_len);
    if (ct_len > 0) ok &= EVP_DecryptUpdate(c, pt, &len, ct_and_tag, ct_len);
    ok &= EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_TAG, 16, (void *)(ct_and_tag + ct_len));
    int fin = EVP_DecryptFinal_ex(c, pt + ct_len, &len);
    EVP_CIPHER_CTX_free(c);
    return (ok && fin > 0) ? 0 : 3;
}

/* AGC:55-112 encryptToPoint with a caller-supplied IV (the reference draws it from SecureRandom). */
int orc_encrypt_point(int64_t id, int32_t key_version, const double *vec, int d, const uint8_t key[32], const uint8_t iv[12],
                      uint8_t *ct_out /* 8d+16 */) {
    char aad[64]; int al = orc_aad(id, key_version, d, aad);
    uint8_t *pt = (uint8_t *)malloc((size_t)d * 8 + 8);
    orc_serialize(vec, d, pt);
    int rc = orc_gcm_encrypt(key, iv, (const uint8_t *)aad, al, pt, d * 8, ct_out);
    free(pt);
    return rc;
}
/* AGC:126-166 decryptFromPoint. */
int orc_decrypt_point(int64_t id, int32_t key_version, int d, const uint8_t key[32], const uint8_t iv[12],
                      const uint8_t *ct /* 8d+16 */, double *vec_out) {
    char aad[64]; int al = orc_aad(id, key_version, d, aad);
    uint8_t *pt = (uint8_t *)malloc((size_t)d * 8 + 32);
    int rc = orc_gcm_decrypt(key, iv, (const uint8_t *)aad, al, ct, d * 8 + 16, pt);
    if (rc == 0) orc_deserialize(pt, d * 8, vec_out);
    free(pt);
    return rc;
}
/* AGC:169-186 / AGC:189-204: query encryption, no AAD. */
int orc_encrypt_query(const double *vec, int d, const uint8_t key[32], const uint8_t iv[12], uint8_t *ct_out) {
    uint8_t *pt = (uint8_t *)malloc((size_t)d * 8 + 8);
    orc_serialize(vec, d, pt);
    int rc = orc_gcm_encrypt(key, iv, NULL, 0, pt, d * 8, ct_out);
    free(pt); return rc;
}
int orc_decrypt_query(const uint8_t *ct, int ct_len_with_tag, const uint8_t key[32], const uint8_t iv[12], double *vec_out) {
    uint8_t *pt = (uint8_t *)malloc((size_t)ct_len_with_tag + 16);
    int rc = orc_gcm_decrypt(key, iv, NULL, 0, ct, ct_len_with_tag, pt);
    if (rc == 0) orc_deserialize(pt, ct_len_with_tag - 16, vec_out);
    free(pt); return rc;
}

/* Bulk helper for test/bench set-up: encrypt n records (ids[i], vecs[i]) under one key version. */
int orc_encrypt_store(const int32_t *ids, int64_t n, const double *vecs, int d, int32_t key_version, const uint8_t key[32],
                      const uint8_t *ivs /* n*12 */, uint8_t *cts /* n*(8d+16) */) {
    for (int64_t i = 0; i < n; i++)
        if (orc_encrypt_point(ids[i], key_version, vecs + (size_t)i * d, d, key, ivs + (size_t)i * 12, cts + (size_t)i * (8 * d + 16))) return -1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Record store + key ring (stand-ins for RDB:530-544 loadEncryptedPoint and KRS:82-88 getVersion).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t N; int32_t dim;
    const uint8_t *iv;           /* [N][12] */
    const uint8_t *ct;           /* [N][8*dim+16] */
    const int32_t *key_version;  /* [N] */
    const uint8_t *deleted;      /* [N] or NULL */
    const uint8_t *present;      /* [N] or NULL (NULL = all present) */
    int32_t n_keys;
    const int32_t *key_versions; /* [n_keys] */
    const uint8_t *keys;         /* [n_keys][32] */
} orc_store_t;

static const uint8_t *orc_find_key(const orc_store_t *st, int32_t version) {
    for (int i = 0; i < st->n_keys; i++) if (st->key_versions[i] == version) return st->keys + (size_t)i * 32;
    return NULL;
}

/* verdict codes shared with the C ABI */
enum { ORC_OK = 0, ORC_NOT_FOUND = 1, ORC_NO_KEY = 2, ORC_TAG_FAIL = 3, ORC_NON_FINITE = 4 };

/* QSI:364-372 l2 */
static double orc_l2(const double *a, const double *b, int len) {
    double s = 0.0;
    for (int i = 0; i < len; i++) { double dd = a[i] - b[i]; s += dd * dd; }
    return sqrt(s);
}
static int orc_all_finite(const double *v, int n) { for (int i = 0; i < n; i++) if (!isfinite(v[i])) return 0; return 1; }

typedef struct { double dist; int32_t pos; int32_t id; } orc_sc_t;
static int orc_sc_cmp(const void *a, const void *b) { /* Double.compare + stability (QSI:298) */
    const orc_sc_t *x = (const orc_sc_t *)a, *y = (const orc_sc_t *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/* QSI:238-322 stage B + C for one query: candidates in order -> verdicts, (id, dist) top-k.
 * Optional plaintext tap (pt_out: [n_cand][dim] doubles) for the debug parity check. */
int orc_refine(const orc_store_t *st, const double *q, const int32_t *cand, int n_cand, int k, int32_t *top_ids,
               double *top_dist, uint8_t *verdict /* n_cand */, int32_t *n_decrypted_out, double *cand_dist_out /* n_cand or NULL */,
               double *pt_out /* or NULL */) {
    int d = st->dim; size_t rec = (size_t)8 * d + 16;
    orc_sc_t *sc = (orc_sc_t *)malloc(sizeof(orc_sc_t) * (size_t)(n_cand ? n_cand : 1));
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    int ns = 0;
    for (int i = 0; i < n_cand; i++) {
        int32_t id = cand[i];
        if (cand_dist_out) cand_dist_out[i] = NAN;
        if (id < 0 || id >= st->N || (st->deleted && st->deleted[id]) || (st->present && !st->present[id])) { verdict[i] = ORC_NOT_FOUND; continue; } /* PIS:717-724 */
        const uint8_t *key = orc_find_key(st, st->key_version[id]);
        if (!key) { verdict[i] = ORC_NO_KEY; continue; }                                     /* KRS:82-88 -> QSI:265 */
        if (orc_decrypt_point(id, st->key_version[id], d, key, st->iv + (size_t)id * 12, st->ct + (size_t)id * rec, v)) { verdict[i] = ORC_TAG_FAIL; continue; }
        if (pt_out) memcpy(pt_out + (size_t)i * d, v, sizeof(double) * (size_t)d);
        if (!orc_all_finite(v, d)) { verdict[i] = ORC_NON_FINITE; continue; }                 /* QSI:253 */
        verdict[i] = ORC_OK;
        sc[ns].dist = orc_l2(q, v, d); sc[ns].pos = ns; sc[ns].id = id;
        if (cand_dist_out) cand_dist_out[i] = sc[ns].dist;
        ns++;
    }
    if (n_decrypted_out) *n_decrypted_out = ns;
    qsort(sc, (size_t)ns, sizeof(orc_sc_t), orc_sc_cmp);
    int eff = ns < k ? ns : k;
    for (int i = 0; i < eff; i++) { top_ids[i] = sc[i].id; top_dist[i] = sc[i].dist; }
    free(sc); free(v);
    return eff;
}

/* QSI:100-352 search(token) for one query.
 *   q_plain: the query vector (the token's AES-GCM wrapping a4/a5 is exercised by orc_encrypt_query /
 *   orc_decrypt_query; callers pass the decrypted vector here -- QSI:131-140 returns empty if it is not finite).
 *   refinement_limit = getEffectiveRefinementLimit (QSI:170-171), ham_threshold = hammingPrefilterThreshold.
 * Outputs describe the LAST pass (the one whose result is returned): cand_ids/cand_scores (<= B), verdicts,
 * counters[0..5] = {candTotal(raw), candKept(unique), candDecrypted, returned, retried, n_cand_refined}.
 * touched (bitmap over ids, or NULL) accumulates over both passes like touchedThisSession (QSI:117,262). */
int orc_search(const orc_index_t *ix, const orc_store_t *st, const double *q_plain, const uint64_t *codes, int k,
               int probes_default, int64_t hard_cap, int refinement_limit, int ham_threshold, int32_t *top_ids,
               double *top_dist, int32_t *cand_ids /* B */, int32_t *cand_scores /* B */, uint8_t *verdicts /* B */,
               int64_t *counters /* 6 */, uint8_t *touched /* N or NULL */) {
    int d = ix->dim;
    for (int i = 0; i < 6; i++) counters[i] = 0;
    if (!orc_all_finite(q_plain, d)) return 0;
    int probes = probes_default, retried = 0;
    int64_t max_u = (int64_t)ix->T * ix->D * 16 * ORC_BLOCK + hard_cap + 2 * ORC_BLOCK; /* room for probes <= 16 */
    int32_t *all_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_u), *all_sc = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_u);
    int ret = 0;
    for (;;) {
        int64_t raw = 0;
        int64_t uniq = orc_route(ix, codes, probes, hard_cap, all_ids, all_sc, max_u, &raw, NULL);
        counters[0] = raw; counters[1] = uniq;
        if (uniq == 0) { ret = 0; break; }
        /* QSI:161-214 stage A.5.  Input is sorted by score, so both branches yield the first <= B entries;
         * restated literally to keep the threshold edge cases. */
        int B = refinement_limit, nc = 0;
        if (ham_threshold > 0) {
            for (int64_t i = 0; i < uniq; i++) if (all_sc[i] <= ham_threshold) { cand_ids[nc] = all_ids[i]; cand_scores[nc] = all_sc[i]; nc++; if (nc >= B) break; }
            if (nc < B) for (int64_t i = 0; i < uniq; i++) if (all_sc[i] > ham_threshold) { cand_ids[nc] = all_ids[i]; cand_scores[nc] = all_sc[i]; nc++; if (nc >= B) break; }
        } else {
            for (int64_t i = 0; i < uniq; i++) { cand_ids[nc] = all_ids[i]; cand_scores[nc] = all_sc[i]; nc++; if (nc >= B) break; }
        }
        int refine_limit = nc < B ? nc : B;                      /* QSI:218 */
        if (refine_limit < 0) refine_limit = 0;
        int32_t ndec = 0;
        int eff = orc_refine(st, q_plain, cand_ids, refine_limit, k, top_ids, top_dist, verdicts, &ndec, NULL, NULL);
        if (touched) for (int i = 0; i < refine_limit; i++) if (verdicts[i] == ORC_OK) touched[cand_ids[i]] = 1;
        counters[2] = ndec; counters[5] = refine_limit;
        if (ndec == 0) { counters[3] = 0; ret = 0; break; }      /* QSI:293 */
        counters[3] = eff; ret = eff;
        if (!retried && (eff < k || ndec < 10 * k)) { retried = 1; counters[4] = 1; probes = 10; continue; } /* QSI:327-337,444-447 */
        break;
    }
    free(all_ids); free(all_sc);
    return ret;
}

/* KRS:215-289 reencryptTouched (Migrate), host side: decrypt each listed record under its stored version,
 * re-encrypt under target_version with the caller's fresh IVs; skips records already >= target, records whose
 * key is gone or whose tag fails.  Writes in place into mutable copies of the store arrays.  Returns #re-encrypted. */
int64_t orc_migrate(int64_t N, int d, uint8_t *iv, uint8_t *ct, int32_t *key_version, const int32_t *ids, int64_t n_ids,
                    const uint8_t *fresh_ivs /* n_ids*12 */, int32_t target_version, int32_t n_keys, const int32_t *key_versions,
                    const uint8_t *keys) {
    size_t rec = (size_t)8 * d + 16; int64_t done = 0;
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    const uint8_t *tkey = NULL;
    for (int i = 0; i < n_keys; i++) if (key_versions[i] == target_version) tkey = keys + (size_t)i * 32;
    for (int64_t j = 0; j < n_ids && tkey; j++) {
        int32_t id = ids[j];
        if (id < 0 || id >= N) continue;
        int32_t old = key_version[id];
        if (old >= target_version) continue;
        const uint8_t *okey = NULL;
        for (int i = 0; i < n_keys; i++) if (key_versions[i] == old) okey = keys + (size_t)i * 32;
        if (!okey) continue;
        if (orc_decrypt_point(id, old, d, okey, iv + (size_t)id * 12, ct + (size_t)id * rec, v)) continue;
        memcpy(iv + (size_t)id * 12, fresh_ivs + (size_t)j * 12, 12);
        orc_encrypt_point(id, target_version, v, d, tkey, iv + (size_t)id * 12, ct + (size_t)id * rec);
        key_version[id] = target_version; done++;
    }
    free(v);
    return done;
}

/* ------------------------------------------------------------------------------------------------
 * Batch drivers used by tests and by the CPU baseline (one thread per call; callers fan out threads).
 * ---------------------------------------------------------------------------------------------- */
void orc_tokengen_batch(const double *queries, int64_t Q, int d, int m, int lambda, int T, int D, const double *alpha,
                        const double *r, const double *omega, uint64_t *codes, int W) {
    for (int64_t q = 0; q < Q; q++) orc_code_all(queries + (size_t)q * d, d, m, lambda, T, D, alpha, r, omega, codes + (size_t)q * T * D * W, W);
}

/* Full TokenGen -> Route -> Refine for queries [q0, q1): the analogue of FSA:636-747's sequential loop. */
void orc_search_batch(const orc_index_t *ix, const orc_store_t *st, const double *queries, int64_t q0, int64_t q1,
                      const double *alpha, const double *r, const double *omega, int k, int probes, int64_t hard_cap, int B,
                      int ham_threshold, int32_t *top_ids /* [Q][k] */, double *top_dist, int32_t *n_ret /* [Q] */,
                      int64_t *counters /* [Q][6] or NULL */) {
    int TD = ix->T * ix->D, W = ix->W, d = ix->dim;
    uint64_t *codes = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)TD * W);
    int32_t *cid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B + 1)), *csc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B + 1));
    uint8_t *ver = (uint8_t *)malloc((size_t)B + 1);
    int64_t cnt[6];
    for (int64_t q = q0; q < q1; q++) {
        const double *qv = queries + (size_t)q * d;
        orc_code_all(qv, d, ix->m, ix->lambda, ix->T, ix->D, alpha, r, omega, codes, W);
        for (int i = 0; i < k; i++) { top_ids[(size_t)q * k + i] = -1; top_dist[(size_t)q * k + i] = NAN; }
        n_ret[q] = orc_search(ix, st, qv, codes, k, probes, hard_cap, B, ham_threshold, top_ids + (size_t)q * k, top_dist + (size_t)q * k, cid, csc, ver, cnt, NULL);
        if (counters) memcpy(counters + (size_t)q * 6, cnt, sizeof cnt);
    }
    free(codes); free(cid); free(csc); free(ver);
}

/* ------------------------------------------------------------------------------------------------
 * Evaluation: exact ground truth and recall@K.
 *   GTP = api/src/main/java/com/fspann/api/GroundtruthPrecompute.java,  FSA = api/.../ForwardSecureANNSystem.java
 * ---------------------------------------------------------------------------------------------- */
/* VecReader.l2sq (GTP:144-163): `double d = q[i] - v` is a FLOAT subtraction widened to double; sum += d*d in FP64. */
static double orc_l2sq_f32(const float *q, const float *b, int d) {
    double sum = 0.0;
    for (int i = 0; i < d; i++) { float df = q[i] - b[i]; double dd = (double)df; sum += dd * dd; }
    return sum;
}
/* BY_D_THEN_ID (GTP:168-171) */
static int orc_gt_less(double da, int32_t ia, double db, int32_t ib) { return da < db || (da == db && ia < ib); }

/* GTP:218-276 run(): per query a bounded max-heap over all base vectors (HeapK, GTP:173-189), ids ascending by (d, id).
 * K must already be clamped to [1, N] (kFinal, GTP:238).  ids [Q][K], d2 [Q][K] (may be NULL). */
void orc_groundtruth(const float *base, int64_t N, int d, const float *queries, int64_t Q, int K, int32_t *ids, double *d2) {
    double *hd = (double *)malloc(sizeof(double) * (size_t)K);
    int32_t *hi = (int32_t *)malloc(sizeof(int32_t) * (size_t)K);
    for (int64_t q = 0; q < Q; q++) {
        int n = 0;
        for (int64_t b = 0; b < N; b++) {
            double dist = orc_l2sq_f32(queries + (size_t)q * d, base + (size_t)b * d, d);
            if (n == K && !orc_gt_less(dist, (int32_t)b, hd[0], hi[0])) continue;      /* GTP:180 */
            int i;
            if (n < K) { i = n++; }
            else {  /* replace the root (the current worst) and sift down */
                i = 0;
                for (;;) {
                    int l = 2 * i + 1, r = l + 1, big = -1;
                    if (l < n) big = l;
                    if (r < n && orc_gt_less(hd[l], hi[l], hd[r], hi[r])) big = r;
                    if (big < 0 || !orc_gt_less(dist, (int32_t)b, hd[big], hi[big])) break;
                    hd[i] = hd[big]; hi[i] = hi[big]; i = big;
                }
                hd[i] = dist; hi[i] = (int32_t)b;
                continue;
            }
            /* sift up (max-heap on (d,id)) */
            while (i > 0) {
                int p = (i - 1) / 2;
                if (!orc_gt_less(hd[p], hi[p], dist, (int32_t)b)) break;
                hd[i] = hd[p]; hi[i] = hi[p]; i = p;
            }
            hd[i] = dist; hi[i] = (int32_t)b;
        }
        /* idsAscending (GTP:182-188): insertion sort of the K survivors by (d, id) */
        for (int a = 1; a < n; a++) {
            double dv = hd[a]; int32_t iv = hi[a]; int j = a - 1;
            while (j >= 0 && orc_gt_less(dv, iv, hd[j], hi[j])) { hd[j + 1] = hd[j]; hi[j + 1] = hi[j]; j--; }
            hd[j + 1] = dv; hi[j + 1] = iv;
        }
        for (int a = 0; a < K; a++) { ids[(size_t)q * K + a] = a < n ? hi[a] : -1; if (d2) d2[(size_t)q * K + a] = a < n ? hd[a] : NAN; }
    }
    free(hd); free(hi);
}

/* FSA:785-794: hits = results[0..min(K, n_ret)) found in gt[0..K); recall = hits / K. */
double orc_recall_at_k(const int32_t *gt, const int32_t *res, int n_ret, int K) {
    int hits = 0, n = n_ret < K ? n_ret : K;
    for (int i = 0; i < n; i++) { int in = 0; for (int j = 0; j < K; j++) in |= gt[j] == res[i]; hits += in; }
    return (double)hits / (double)K;
}
