// gcm_core.cuh -- GF(2^128) arithmetic of GHASH (NIST SP 800-38D) for the GCM STREAM cipher mode
// (reference: lib/src/cipher/gcm.rs, aes-gcm 0.11 `AesGcm<C, U12>` over Aes256 / Camellia256).
//
// An element is four 32-bit words, w[3] the most significant: w[3] = big-endian bytes 0..3 of the block, w[0] = bytes 12..15.
// Multiplication by a FIXED element uses a 16-entry table per element (4 bits at a time, the reduction of the four bits
// shifted out comes from a 16-entry constant table); the tables of H, H^2, H^4, ... are what the kernels combine partial
// hashes with, so no general multiplication is needed anywhere.  PNA_HD: tests/host compiles the same code with g++.
#pragma once
#include "common.cuh"

namespace pna {
namespace gcm {

struct alignas(16) G128 { uint32_t w[4]; };
struct GTab { G128 e[16]; };   // e[n] = n(x) * P for the 4-bit polynomial n (bit 3 = degree 0)

// reduction of the 4 bits shifted out at the low end, already placed at the top 16 bits of w[3]
#define PNA_GCM_LAST4 { 0x00000000u, 0x1c200000u, 0x38400000u, 0x24600000u, 0x70800000u, 0x6ca00000u, 0x48c00000u, 0x54e00000u, \
                        0xe1000000u, 0xfd200000u, 0xd9400000u, 0xc5600000u, 0x91800000u, 0x8da00000u, 0xa9c00000u, 0xb5e00000u }

PNA_HD G128 from_le_words(uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3) {   // block bytes as loaded (little-endian words)
    G128 g;
    g.w[3] = bswap32(d0); g.w[2] = bswap32(d1); g.w[1] = bswap32(d2); g.w[0] = bswap32(d3);
    return g;
}
PNA_HD void to_le_words(const G128& g, uint32_t d[4]) {
    d[0] = bswap32(g.w[3]); d[1] = bswap32(g.w[2]); d[2] = bswap32(g.w[1]); d[3] = bswap32(g.w[0]);
}
PNA_HD void gxor(G128& a, const G128& b) { a.w[0] ^= b.w[0]; a.w[1] ^= b.w[1]; a.w[2] ^= b.w[2]; a.w[3] ^= b.w[3]; }

// table of P: e[8] = P, e[4] = P*x, e[2] = P*x^2, e[1] = P*x^3, the rest by linearity
PNA_HD void make_table(const G128& p, GTab* t) {
    G128 v = p;
    t->e[0] = G128{{0, 0, 0, 0}};
    t->e[8] = v;
    for (int i = 4; i > 0; i >>= 1) {
        const uint32_t carry = v.w[0] & 1u;
        v.w[0] = (v.w[0] >> 1) | (v.w[1] << 31);
        v.w[1] = (v.w[1] >> 1) | (v.w[2] << 31);
        v.w[2] = (v.w[2] >> 1) | (v.w[3] << 31);
        v.w[3] = (v.w[3] >> 1) ^ (carry ? 0xe1000000u : 0u);
        t->e[i] = v;
    }
    for (int i = 2; i <= 8; i <<= 1)
        for (int j = 1; j < i; j++) {
            G128 s = t->e[i];
            gxor(s, t->e[j]);
            t->e[i + j] = s;
        }
}

// x * P with P's table; nibbles of x from the least significant up (Shoup's method).  `tab` and `last4` may live in any
// address space (shared memory in the kernels).
template <class TabPtr, class RedPtr>
PNA_HD G128 mul_table(const G128& x, TabPtr tab, RedPtr last4) {
    G128 z = tab[x.w[0] & 15u];
#pragma unroll
    for (int q = 1; q < 32; q++) {
        const uint32_t nib = (x.w[q >> 3] >> (4 * (q & 7))) & 15u;
        const uint32_t rem = z.w[0] & 15u;
#if defined(__CUDA_ARCH__)
        z.w[0] = __funnelshift_r(z.w[0], z.w[1], 4);
        z.w[1] = __funnelshift_r(z.w[1], z.w[2], 4);
        z.w[2] = __funnelshift_r(z.w[2], z.w[3], 4);
#else
        z.w[0] = (z.w[0] >> 4) | (z.w[1] << 28);
        z.w[1] = (z.w[1] >> 4) | (z.w[2] << 28);
        z.w[2] = (z.w[2] >> 4) | (z.w[3] << 28);
#endif
        z.w[3] = (z.w[3] >> 4) ^ last4[rem];
        const G128 t = tab[nib];
        z.w[0] ^= t.w[0]; z.w[1] ^= t.w[1]; z.w[2] ^= t.w[2]; z.w[3] ^= t.w[3];
    }
    return z;
}

// the powers every key needs: H^(2^k) for k = 0..5 (lane tree + the strided Horner step H^32) and H^1024 (tile to tile)
constexpr int GCM_N_POW = 7;
constexpr int GCM_POW_TILE = 6;          // index of H^1024
constexpr int GCM_POW_STRIDE = 5;        // index of H^32
struct GcmPow { GTab t[GCM_N_POW]; };    // 1792 bytes per key

// h = E_K(0^128) as an element.  Fills out->t[]; `scratch` holds the table of the running power.
PNA_HD void make_powers(const G128& h, GcmPow* out) {
    const uint32_t last4[16] = PNA_GCM_LAST4;
    GTab cur;
    G128 p = h;
    for (int k = 0; k <= 10; k++) {     // p = H^(2^k)
        make_table(p, &cur);
        if (k <= 5) out->t[k] = cur;
        if (k == 10) out->t[GCM_POW_TILE] = cur;
        p = mul_table(p, cur.e, last4);  // square
    }
}

}  // namespace gcm
}  // namespace pna
