// aead_host.hpp -- host-side key schedule of the GCM STREAM cipher mode (lib/src/cipher/aead.rs:152-208): HKDF-SHA-256 of the
// master key over the entry context, and the key confirmation value.  Runs once per entry on the CPU, like the password KDF
// (lib/src/hash.rs): nothing here is data-parallel.  SHA-256 / HMAC / HKDF are FIPS 180-4 / RFC 2104 / RFC 5869; the hkdf 0.12
// crate treats an absent-or-empty salt as HashLen zero bytes.
#pragma once
#include <stdint.h>
#include <string.h>

namespace pna {
namespace aead {

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    Sha256() {
        static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h, iv, sizeof h);
    }
    static uint32_t rotr(uint32_t v, int s) { return (v >> s) | (v << (32 - s)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
            0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
            0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
            0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
            0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
            0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        size_t fill = (size_t)(len & 63);
        len += n;
        if (fill) {
            const size_t k = n < 64 - fill ? n : 64 - fill;
            memcpy(buf + fill, p, k);
            p += k; n -= k; fill += k;
            if (fill < 64) return;
            block(buf);
        }
        for (; n >= 64; p += 64, n -= 64) block(p);
        if (n) memcpy(buf, p, n);
    }
    void finish(uint8_t out[32]) {
        const uint64_t bits = len * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while ((len & 63) != 56) update(&zero, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
    }
};
inline void sha256(const void* a, size_t na, const void* b, size_t nb, uint8_t out[32]) {
    Sha256 s;
    if (na) s.update(a, na);
    if (nb) s.update(b, nb);
    s.finish(out);
}
inline void hmac_sha256(const uint8_t* key, size_t nk, const void* a, size_t na, const void* b, size_t nb, uint8_t out[32]) {
    uint8_t k0[64] = {0}, pad[64], inner[32];
    if (nk > 64) sha256(key, nk, nullptr, 0, k0); else if (nk) memcpy(k0, key, nk);
    Sha256 si;
    for (int i = 0; i < 64; i++) pad[i] = k0[i] ^ 0x36;
    si.update(pad, 64);
    if (na) si.update(a, na);
    if (nb) si.update(b, nb);
    si.finish(inner);
    Sha256 so;
    for (int i = 0; i < 64; i++) pad[i] = k0[i] ^ 0x5c;
    so.update(pad, 64);
    so.update(inner, 32);
    so.finish(out);
}
// aead.rs:152-158 hkdf_sha256: extract with `salt`, expand one block with `info`
inline void hkdf_sha256(const uint8_t* ikm, size_t n_ikm, const uint8_t* salt, size_t n_salt, const uint8_t* info, size_t n_info, uint8_t okm[32]) {
    const uint8_t zeros[32] = {0};
    uint8_t prk[32];
    hmac_sha256(n_salt ? salt : zeros, n_salt ? n_salt : 32, ikm, n_ikm, nullptr, 0, prk);
    const uint8_t one = 1;
    hmac_sha256(prk, 32, info, n_info, &one, 1, okm);
}
constexpr size_t STREAM_HEADER_LEN = 75;
// aead.rs:162-164
inline void key_confirmation(const uint8_t k_master[32], uint8_t out[32]) {
    hkdf_sha256(k_master, 32, nullptr, 0, reinterpret_cast<const uint8_t*>("PNA-KC-v1"), 9, out);
}
// aead.rs:166-208 entry_context + derive_stream_key
inline void derive_stream_key(const uint8_t k_master[32], const uint8_t header[STREAM_HEADER_LEN], const uint8_t header_type[4],
                              const uint8_t* header_data, size_t header_len, const uint8_t* phsf, size_t phsf_len, uint8_t out[32]) {
    uint8_t ctx[88];
    memcpy(ctx, "PNA-STREAM-v1", 13);
    sha256(header_type, 4, header_data, header_len, ctx + 13);
    sha256(phsf, phsf_len, nullptr, 0, ctx + 45);
    memcpy(ctx + 77, header + 32, 7);    // nonce prefix
    memcpy(ctx + 84, header + 39, 4);    // segment size, big endian as on the wire
    hkdf_sha256(k_master, 32, header, 32, ctx, sizeof ctx, out);
}

}  // namespace aead
}  // namespace pna
