// host_check.cpp -- compiles the FSP_HD primitives of aes_gcm.cuh for the CPU so tests can check the exact
// arithmetic the kernels use (AES-256 T-table rounds, hole-multiply GHASH, AAD builder, Java hash) against the
// oracle / OpenSSL without a GPU.  TEST-ONLY: not linked into libfspann_gpu.so, never a fallback.
#include <cstring>
#include <vector>
#include "aes_gcm.cuh"
using namespace fsp;

namespace {
struct TeArr { const uint32_t *t; uint32_t operator()(uint32_t x) const { return t[x]; } };
struct RkArr { const uint32_t *r; uint32_t operator()(int i) const { return r[i]; } };
}

extern "C" {

void fsp_hc_te0(uint32_t *out) { aes_make_te0(out); }
void fsp_hc_expand(const uint8_t *key, uint32_t *rk) { aes256_expand_key(key, rk); }

void fsp_hc_aes_block(const uint8_t *key, const uint8_t *in, uint8_t *out) {
    uint32_t te[256], rk[60], o[4];
    aes_make_te0(te); aes256_expand_key(key, rk);
    uint32_t s[4];
    for (int i = 0; i < 4; i++) s[i] = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
    aes256_encrypt(TeArr{te}, RkArr{rk}, s[0], s[1], s[2], s[3], o);
    for (int i = 0; i < 4; i++) { out[4 * i] = o[i] >> 24; out[4 * i + 1] = o[i] >> 16; out[4 * i + 2] = o[i] >> 8; out[4 * i + 3] = o[i]; }
}

void fsp_hc_gfmul(const uint64_t *x, const uint64_t *y, uint64_t *z_fast, uint64_t *z_ref) {
    u128 a{x[0], x[1]}, b{y[0], y[1]};
    u128 f = gf128_mul(a, b), r = gf128_mul_ref(a, b);
    z_fast[0] = f.hi; z_fast[1] = f.lo; z_ref[0] = r.hi; z_ref[1] = r.lo;
}

// Shoup 8-bit table multiply (the verify kernel's GHASH step) against the bit-serial reference.  Returns #mismatches.
int fsp_hc_shoup_check(const uint64_t *h, const uint64_t *xs, int n) {
    u128 H{h[0], h[1]};
    std::vector<uint32_t> tab(4096 * 4);
    ghash_make_shoup8(H, tab.data());
    int bad = 0;
    for (int i = 0; i < n; i++) {
        u128 x{xs[2 * i], xs[2 * i + 1]};
        const u128 ref = gf128_mul_ref(x, H);
        uint32_t y[4] = {(uint32_t)(x.hi >> 32), (uint32_t)x.hi, (uint32_t)(x.lo >> 32), (uint32_t)x.lo}, z[4] = {0, 0, 0, 0};
        for (int j = 0; j < 16; j++) {
            const uint32_t b = (y[j / 4] >> (24 - 8 * (j % 4))) & 0xff;
            for (int k = 0; k < 4; k++) z[k] ^= tab[((size_t)b * 16 + j) * 4 + k];
        }
        if (z[0] != (uint32_t)(ref.hi >> 32) || z[1] != (uint32_t)ref.hi || z[2] != (uint32_t)(ref.lo >> 32) || z[3] != (uint32_t)ref.lo) bad++;
    }
    return bad;
}

int fsp_hc_aad(int64_t id, int32_t ver, int32_t dim, uint8_t *buf) { return build_aad(id, ver, dim, buf); }
uint32_t fsp_hc_java_hash(int32_t id) { return java_hash_decimal(id); }

// Full record decrypt the way the refine kernel does it: CTR keystream per block, GHASH as
// sum_i X_i * H^(p_i) over a power table, tag = GHASH ^ E_K(J0).  Returns 1 if the tag verifies.
int fsp_hc_decrypt_record(const uint8_t *key, const uint8_t *iv, int64_t id, int32_t ver, int32_t dim,
                          const uint8_t *ct /* 8*dim+16 */, uint8_t *plain /* 8*dim */) {
    uint32_t te[256], rk[60];
    aes_make_te0(te); aes256_expand_key(key, rk);
    TeArr T{te}; RkArr R{rk};
    const int nbytes = 8 * dim, c = (nbytes + 15) / 16;
    uint32_t ivw[3];
    for (int i = 0; i < 3; i++) ivw[i] = ((uint32_t)iv[4 * i] << 24) | ((uint32_t)iv[4 * i + 1] << 16) | ((uint32_t)iv[4 * i + 2] << 8) | iv[4 * i + 3];
    uint32_t h[4], ej0[4];
    aes256_encrypt(T, R, 0, 0, 0, 0, h);
    aes256_encrypt(T, R, ivw[0], ivw[1], ivw[2], 1, ej0);
    u128 H{((uint64_t)h[0] << 32) | h[1], ((uint64_t)h[2] << 32) | h[3]};
    uint8_t aad[FSP_AAD_MAX];
    const int alen = build_aad(id, ver, dim, aad), a = (alen + 15) / 16;
    const int npow = c + 1 + 3;
    std::vector<u128> hp(npow + 1);
    hp[1] = H;
    for (int p = 2; p <= npow; p++) hp[p] = gf128_mul_ref(hp[p - 1], H);
    u128 acc{0, 0};
    auto add = [&](u128 x, int p) { u128 z = gf128_mul(x, hp[p]); acc.hi ^= z.hi; acc.lo ^= z.lo; };
    for (int j = 0; j < a; j++) add(u128{load_be64(aad + 16 * j), load_be64(aad + 16 * j + 8)}, c + 1 + a - j);
    for (int i = 0; i < c; i++) {
        uint8_t blk[16] = {0};
        int nb = nbytes - 16 * i < 16 ? nbytes - 16 * i : 16;
        memcpy(blk, ct + 16 * i, nb);
        add(u128{load_be64(blk), load_be64(blk + 8)}, c + 1 - i);
        uint32_t ks[4];
        aes256_encrypt(T, R, ivw[0], ivw[1], ivw[2], (uint32_t)(i + 2), ks);
        for (int b = 0; b < nb; b++) plain[16 * i + b] = blk[b] ^ (uint8_t)(ks[b / 4] >> (24 - 8 * (b % 4)));
    }
    add(u128{(uint64_t)alen * 8, (uint64_t)nbytes * 8}, 1);
    uint64_t thi = acc.hi ^ (((uint64_t)ej0[0] << 32) | ej0[1]), tlo = acc.lo ^ (((uint64_t)ej0[2] << 32) | ej0[3]);
    const uint8_t *tag = ct + nbytes;
    return thi == load_be64(tag) && tlo == load_be64(tag + 8);
}
}
