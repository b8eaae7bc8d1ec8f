// aes_gcm.cuh -- AES-256 / GHASH / AAD primitives shared by the device kernels and the host-side key set-up.
//
// Everything here is FSP_HD (__host__ __device__) so the exact arithmetic the kernels run can also be
// compiled with g++ and checked on the CPU (tests/test_primitives_host.py) without a GPU.
//
// Replaces, on the device, what the reference gets from JDK SunJCE through
//   crypto/src/main/java/com/fspann/crypto/AesGcmCryptoService.java:126-166 (decryptFromPoint)
//   common/src/main/java/com/fspann/common/EncryptedPoint.java:80-83        (AAD "id:%s|v:%d|d:%d")
// AES-256-GCM per NIST SP 800-38D: 96-bit IV => J0 = IV || 0x00000001, data counters start at 2, 128-bit tag
// appended to the ciphertext (Java doFinal layout).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FSP_HD __host__ __device__ __forceinline__
#define FSP_D __device__ __forceinline__
#else
#define FSP_HD static inline
#endif

namespace fsp {

// ------------------------------------------------------------------------------------------------
// AES tables.  Te0[x] = (2*S[x], S[x], S[x], 3*S[x]) as a big-endian word; Te1..3 are byte rotations of it.
// Generated arithmetically (GF(2^8) inverse + affine map) rather than typed in.
// ------------------------------------------------------------------------------------------------
static inline uint8_t gf256_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) p ^= a;
        uint8_t hi = a & 0x80;
        a = (uint8_t)(a << 1);
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}
static inline void aes_make_sbox(uint8_t sbox[256]) {
    uint8_t inv[256];
    inv[0] = 0;
    for (int x = 1; x < 256; x++) {
        // x^254 = x^-1 in GF(2^8)
        uint8_t r = 1, b = (uint8_t)x;
        int e = 254;
        while (e) { if (e & 1) r = gf256_mul(r, b); b = gf256_mul(b, b); e >>= 1; }
        inv[x] = r;
    }
    for (int x = 0; x < 256; x++) {
        uint8_t v = inv[x], s = v;
        for (int i = 1; i <= 4; i++) s ^= (uint8_t)((v << i) | (v >> (8 - i)));
        sbox[x] = (uint8_t)(s ^ 0x63);
    }
}
static inline void aes_make_te0(uint32_t te0[256]) {
    uint8_t sbox[256];
    aes_make_sbox(sbox);
    for (int x = 0; x < 256; x++) {
        uint8_t s = sbox[x], s2 = gf256_mul(s, 2), s3 = (uint8_t)(s2 ^ s);
        te0[x] = ((uint32_t)s2 << 24) | ((uint32_t)s << 16) | ((uint32_t)s << 8) | (uint32_t)s3;
    }
}
// AES-256 key schedule: 60 big-endian round-key words (host side, once per key version).
static inline void aes256_expand_key(const uint8_t key[32], uint32_t rk[60]) {
    uint8_t sbox[256];
    aes_make_sbox(sbox);
    for (int i = 0; i < 8; i++)
        rk[i] = ((uint32_t)key[4 * i] << 24) | ((uint32_t)key[4 * i + 1] << 16) | ((uint32_t)key[4 * i + 2] << 8) | key[4 * i + 3];
    uint32_t rcon = 1;
    for (int i = 8; i < 60; i++) {
        uint32_t t = rk[i - 1];
        if (i % 8 == 0) {
            t = (t << 8) | (t >> 24);
            t = ((uint32_t)sbox[t >> 24] << 24) | ((uint32_t)sbox[(t >> 16) & 0xff] << 16) | ((uint32_t)sbox[(t >> 8) & 0xff] << 8) | sbox[t & 0xff];
            t ^= rcon << 24;
            rcon = (uint32_t)gf256_mul((uint8_t)rcon, 2);
        } else if (i % 8 == 4) {
            t = ((uint32_t)sbox[t >> 24] << 24) | ((uint32_t)sbox[(t >> 16) & 0xff] << 16) | ((uint32_t)sbox[(t >> 8) & 0xff] << 8) | sbox[t & 0xff];
        }
        rk[i] = rk[i - 8] ^ t;
    }
}

FSP_HD uint32_t ror32(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}
FSP_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// One AES-256 block encryption (14 rounds).  TE is a functor x -> Te0[x] (x in 0..255) so the same code runs
// against the lane-replicated shared-memory table on the device and a plain array on the host.
// State and round keys are big-endian column words.
template <class TE, class RK>
FSP_HD void aes256_encrypt(const TE &te, const RK &rk, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t out[4]) {
    s0 ^= rk(0); s1 ^= rk(1); s2 ^= rk(2); s3 ^= rk(3);
#pragma unroll
    for (int r = 1; r < 14; r++) {
        uint32_t t0 = te(s0 >> 24) ^ ror32(te((s1 >> 16) & 0xff), 8) ^ ror32(te((s2 >> 8) & 0xff), 16) ^ ror32(te(s3 & 0xff), 24) ^ rk(4 * r + 0);
        uint32_t t1 = te(s1 >> 24) ^ ror32(te((s2 >> 16) & 0xff), 8) ^ ror32(te((s3 >> 8) & 0xff), 16) ^ ror32(te(s0 & 0xff), 24) ^ rk(4 * r + 1);
        uint32_t t2 = te(s2 >> 24) ^ ror32(te((s3 >> 16) & 0xff), 8) ^ ror32(te((s0 >> 8) & 0xff), 16) ^ ror32(te(s1 & 0xff), 24) ^ rk(4 * r + 2);
        uint32_t t3 = te(s3 >> 24) ^ ror32(te((s0 >> 16) & 0xff), 8) ^ ror32(te((s1 >> 8) & 0xff), 16) ^ ror32(te(s2 & 0xff), 24) ^ rk(4 * r + 3);
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    // last round: SubBytes + ShiftRows + AddRoundKey; S[x] is byte 2 (and 1) of Te0[x]
#define FSP_SB(x) ((te(x) >> 8) & 0xffu)
    out[0] = (FSP_SB(s0 >> 24) << 24 | FSP_SB((s1 >> 16) & 0xff) << 16 | FSP_SB((s2 >> 8) & 0xff) << 8 | FSP_SB(s3 & 0xff)) ^ rk(56);
    out[1] = (FSP_SB(s1 >> 24) << 24 | FSP_SB((s2 >> 16) & 0xff) << 16 | FSP_SB((s3 >> 8) & 0xff) << 8 | FSP_SB(s0 & 0xff)) ^ rk(57);
    out[2] = (FSP_SB(s2 >> 24) << 24 | FSP_SB((s3 >> 16) & 0xff) << 16 | FSP_SB((s0 >> 8) & 0xff) << 8 | FSP_SB(s1 & 0xff)) ^ rk(58);
    out[3] = (FSP_SB(s3 >> 24) << 24 | FSP_SB((s0 >> 16) & 0xff) << 16 | FSP_SB((s1 >> 8) & 0xff) << 8 | FSP_SB(s2 & 0xff)) ^ rk(59);
#undef FSP_SB
}

// ------------------------------------------------------------------------------------------------
// GF(2^128) for GHASH.  A block b[0..15] is held as hi = be64(b[0..7]), lo = be64(b[8..15]); in that integer
// bit (127-i) is the coefficient of x^i (GCM's reflected convention).  The SM has no carry-less multiply, so
// the 32x32->64 products are emulated with integer multiplies on operands whose bits are spread into
// "holes" (one payload bit per nibble: at most 8 partial products land on a nibble, so no carry crosses it).
// ------------------------------------------------------------------------------------------------
struct u128 { uint64_t hi, lo; };

FSP_HD uint64_t mulwide(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)a * (uint64_t)b;  // IMAD.WIDE.U32
#else
    return (uint64_t)a * (uint64_t)b;
#endif
}
// carry-less 32 x 32 -> 64
FSP_HD uint64_t clmul32(uint32_t x, uint32_t y) {
    const uint32_t x0 = x & 0x11111111u, x1 = x & 0x22222222u, x2 = x & 0x44444444u, x3 = x & 0x88888888u;
    const uint32_t y0 = y & 0x11111111u, y1 = y & 0x22222222u, y2 = y & 0x44444444u, y3 = y & 0x88888888u;
    uint64_t z0 = mulwide(x0, y0) ^ mulwide(x1, y3) ^ mulwide(x2, y2) ^ mulwide(x3, y1);
    uint64_t z1 = mulwide(x0, y1) ^ mulwide(x1, y0) ^ mulwide(x2, y3) ^ mulwide(x3, y2);
    uint64_t z2 = mulwide(x0, y2) ^ mulwide(x1, y1) ^ mulwide(x2, y0) ^ mulwide(x3, y3);
    uint64_t z3 = mulwide(x0, y3) ^ mulwide(x1, y2) ^ mulwide(x2, y1) ^ mulwide(x3, y0);
    z0 &= 0x1111111111111111ull; z1 &= 0x2222222222222222ull; z2 &= 0x4444444444444444ull; z3 &= 0x8888888888888888ull;
    return z0 | z1 | z2 | z3;
}
// carry-less 64 x 64 -> 128 (Karatsuba over 32-bit halves)
FSP_HD void clmul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p0 = clmul32(a0, b0), p2 = clmul32(a1, b1), pm = clmul32(a0 ^ a1, b0 ^ b1) ^ p0 ^ p2;
    lo = p0 ^ (pm << 32);
    hi = p2 ^ (pm >> 32);
}
// Z = X * Y in GF(2^128) with GCM's polynomial x^128 + x^7 + x^2 + x + 1 (reflected representation above).
FSP_HD u128 gf128_mul(const u128 &x, const u128 &y) {
    uint64_t z0h, z0l, z1h, z1l, z2h, z2l;
    clmul64(x.lo, y.lo, z0h, z0l);
    clmul64(x.hi, y.hi, z1h, z1l);
    clmul64(x.lo ^ x.hi, y.lo ^ y.hi, z2h, z2l);
    z2h ^= z0h ^ z1h; z2l ^= z0l ^ z1l;
    // 255-bit product (v3:v2:v1:v0), integer bit j <-> coefficient 254-j
    uint64_t v0 = z0l, v1 = z0h ^ z2l, v2 = z1l ^ z2h, v3 = z1h;
    // align to 256 bits (the reflected product is short by one bit)
    v3 = (v3 << 1) | (v2 >> 63); v2 = (v2 << 1) | (v1 >> 63); v1 = (v1 << 1) | (v0 >> 63); v0 = v0 << 1;
    // fold the low half (degrees >= 128) back in
    v2 ^= v0 ^ (v0 >> 1) ^ (v0 >> 2) ^ (v0 >> 7);
    v1 ^= (v0 << 63) ^ (v0 << 62) ^ (v0 << 57);
    v3 ^= v1 ^ (v1 >> 1) ^ (v1 >> 2) ^ (v1 >> 7);
    v2 ^= (v1 << 63) ^ (v1 << 62) ^ (v1 << 57);
    u128 r; r.hi = v3; r.lo = v2;
    return r;
}
// Bit-serial reference multiply (NIST SP 800-38D algorithm 1); host-side check of gf128_mul and power tables.
static inline u128 gf128_mul_ref(const u128 &x, const u128 &y) {
    u128 z = {0, 0}, v = y;
    for (int i = 0; i < 128; i++) {
        uint64_t bit = i < 64 ? (x.hi >> (63 - i)) & 1 : (x.lo >> (127 - i)) & 1;
        if (bit) { z.hi ^= v.hi; z.lo ^= v.lo; }
        uint64_t lsb = v.lo & 1;
        v.lo = (v.lo >> 1) | (v.hi << 63);
        v.hi >>= 1;
        if (lsb) v.hi ^= 0xe100000000000000ull;
    }
    return z;
}

// Shoup 8-bit tables for multiplication by a fixed H: T[j][b] = (byte b at byte position j of a block) * H, so
// V * H = XOR_j T[j][byte_j(V)].  4096 entries of 4 big-endian words (64 KB per key version), stored BYTE-major:
// entry (b, j) at index b*16 + j.  Its 16-byte bank group is then j mod 8, so a warp whose lanes walk the 16 byte
// positions in lane-rotated order reads shared memory without bank conflicts.  Host side.
static inline void ghash_make_shoup8(const u128 &H, uint32_t *table /* [256][16][4] */) {
    for (int j = 0; j < 16; j++)
        for (int b = 0; b < 256; b++) {
            u128 v{0, 0};
            if (j < 8) v.hi = (uint64_t)b << (56 - 8 * j); else v.lo = (uint64_t)b << (56 - 8 * (j - 8));
            const u128 z = gf128_mul_ref(v, H);
            uint32_t *t = table + ((size_t)b * 16 + j) * 4;
            t[0] = (uint32_t)(z.hi >> 32); t[1] = (uint32_t)z.hi; t[2] = (uint32_t)(z.lo >> 32); t[3] = (uint32_t)z.lo;
        }
}

// ------------------------------------------------------------------------------------------------
// AAD = UTF-8 of String.format("id:%s|v:%d|d:%d", id, keyVersion, dimension) (EP:80-83); ids are the decimal
// strings the reference's facade assigns (api/.../ForwardSecureANNSystem.java:501,515).  Writes at most 48
// bytes (zero padded to FSP_AAD_MAX) and returns the length.
// ------------------------------------------------------------------------------------------------
#define FSP_AAD_MAX 48
FSP_HD int put_dec(uint8_t *buf, int pos, int64_t v) {
    char tmp[20];
    int n = 0;
    bool neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    do { tmp[n++] = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
    if (neg) buf[pos++] = '-';
    while (n) buf[pos++] = (uint8_t)tmp[--n];
    return pos;
}
FSP_HD int build_aad(int64_t id, int32_t key_version, int32_t dim, uint8_t buf[FSP_AAD_MAX]) {
    for (int i = 0; i < FSP_AAD_MAX; i++) buf[i] = 0;
    int p = 0;
    buf[p++] = 'i'; buf[p++] = 'd'; buf[p++] = ':';
    p = put_dec(buf, p, id);
    buf[p++] = '|'; buf[p++] = 'v'; buf[p++] = ':';
    p = put_dec(buf, p, key_version);
    buf[p++] = '|'; buf[p++] = 'd'; buf[p++] = ':';
    p = put_dec(buf, p, dim);
    return p;
}
FSP_HD uint64_t load_be64(const uint8_t *b) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | b[i];
    return v;
}

// java.lang.String.hashCode of the decimal string of a non-negative id, spread like java.util.HashMap.hash
// (h ^ (h >>> 16)).  Drives the reference's candidate ordering (PIS:619,690-696).
FSP_HD uint32_t java_hash_decimal(int32_t id) {
    // h = sum_i c_i * 31^(n-1-i): peel digits from the right, the power of 31 grows with the position
    uint32_t u = (uint32_t)id, h = 0, p = 1;
    do {
        const uint32_t q = u / 10u;
        h += (48u + (u - q * 10u)) * p;
        p *= 31u;
        u = q;
    } while (u);
    return h ^ (h >> 16);
}

}  // namespace fsp
