/*
 * pna_oracle.c -- CPU ORACLE for the PNA per-entry data-chunk pipeline.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (portable-network-archive_b200/csrc) never links, loads or calls anything here.
 *
 * What it restates (citations relative to /root/reference, PNA v0.37.0):
 *   chunk CRC            lib/src/format/chunk.rs:7-21      (crc32fast 1.5.0 == CRC-32/ISO-HDLC)
 *   CTR decrypt/encrypt  lib/src/cipher/stream/read.rs:39-43, write.rs:50-58, cipher.rs:21-55
 *                        (ctr 0.10.1 Ctr128BE: keystream block i = E_K(IV +128 i), big endian)
 *   CBC decrypt/encrypt  lib/src/cipher/block/read.rs:31-116, write.rs:48-124 (cbc 0.2.1 + PKCS#7)
 *   IV prefix            lib/src/entry/read.rs:79-103, lib/src/entry/write.rs:46-50
 *   GCM STREAM segments  lib/src/cipher/gcm.rs:44-63,206-290, lib/src/cipher/aead.rs:92-150,210-217 (aes-gcm 0.11 == SP 800-38D)
 *   decompress           lib/src/entry/read.rs:171-190 (zstd::Decoder::with_buffer -> libzstd streaming,
 *                        flate2::bufread::ZlibDecoder -> one zlib stream)
 *   compress             lib/src/entry/write.rs:251-265 (ZstdEncoder level, no pledged size; ZlibEncoder)
 *   task-per-entry pool  cli/src/command/extract.rs:987, cli/src/command/core.rs:510 (rayon)
 *
 * Third-party arithmetic the reference pulls from Cargo (NOT under /root/reference) and
 * what stands in for it here:
 *   crc32fast 1.5.0            -> table-driven restatement below (checked against zlib crc32())
 *   aes 0.9.2 / camellia 0.2.1 -> OpenSSL 3.0 EVP_aes_256_ecb / EVP_camellia_256_ecb block primitive;
 *                                 the CTR128-BE counter and the CBC/PKCS#7 chaining are restated here
 *   zstd-sys 2.0.14+zstd.1.5.7 -> system libzstd.so.1 (dlopen; same upstream C code)
 *   flate2 1.1.9/miniz_oxide   -> zlib 1.3 inflate()/deflate() (same RFC 1950/1951 format; encode
 *                                 bytes differ from miniz_oxide, decode is a unique function)
 *
 * Parity is PINNED: tests/test_oracle.py and tests/test_gcm.py check this file against the reference's own
 * golden archives (resources/test/NAME.pna vs resources/test/raw) and the KATs in
 * lib/src/format/chunk.rs:31, lib/src/io.rs:179, lib/src/cipher.rs:256-292.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <openssl/evp.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* status codes mirror include/pna_cuda.h (io::ErrorKind classes of the reference) */
enum {
    ORA_OK = 0,
    ORA_INVALID_DATA = 1,
    ORA_UNEXPECTED_EOF = 2,
    ORA_INVALID_INPUT = 3,
    ORA_UNSUPPORTED = 4,
    ORA_NOSPACE = 5,
    ORA_OOM = 6,
    ORA_INTERNAL = 7,
};

/* ------------------------------------------------------------------ CRC-32 */
/* lib/src/format/chunk.rs:7-12: Hasher::new(); update(type); update(data); finalize() */
static uint32_t crc_tab[8][256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
        crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
}
uint32_t pna_oracle_crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    pthread_once(&crc_once, crc_init);
    uint32_t c = ~crc;
    while (n && ((uintptr_t)p & 7)) { c = (c >> 8) ^ crc_tab[0][(c ^ *p++) & 0xFF]; n--; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = crc_tab[7][w & 0xFF] ^ crc_tab[6][(w >> 8) & 0xFF] ^ crc_tab[5][(w >> 16) & 0xFF] ^
            crc_tab[4][(w >> 24) & 0xFF] ^ crc_tab[3][(w >> 32) & 0xFF] ^ crc_tab[2][(w >> 40) & 0xFF] ^
            crc_tab[1][(w >> 48) & 0xFF] ^ crc_tab[0][(w >> 56) & 0xFF];
        p += 8; n -= 8;
    }
    while (n--) c = (c >> 8) ^ crc_tab[0][(c ^ *p++) & 0xFF];
    return ~c;
}
uint32_t pna_oracle_chunk_crc(const uint8_t type[4], const uint8_t* data, size_t n) {
    uint32_t c = pna_oracle_crc32_update(0, type, 4);
    return pna_oracle_crc32_update(c, data, n);
}
/* the faster system CRC (zlib's braided/PCLMUL one) for the CPU-baseline timing leg */
uint32_t pna_oracle_crc32_zlib(const uint8_t* p, size_t n) {
    uLong c = crc32(0L, Z_NULL, 0);
    while (n) { uInt k = n > (1u << 30) ? (1u << 30) : (uInt)n; c = crc32(c, p, k); p += k; n -= k; }
    return (uint32_t)c;
}

/* ------------------------------------------------------------------ ciphers */
static const EVP_CIPHER* block_cipher(int encryption) {
    if (encryption == 1) return EVP_aes_256_ecb();      /* Encryption::AES       options.rs:483 */
    if (encryption == 2) return EVP_camellia_256_ecb(); /* Encryption::CAMELLIA  options.rs:487 */
    return NULL;
}
typedef struct { EVP_CIPHER_CTX* c; } blk_t;
static int blk_init(blk_t* b, int encryption, const uint8_t key[32], int enc) {
    const EVP_CIPHER* ci = block_cipher(encryption);
    if (!ci) return ORA_UNSUPPORTED;
    b->c = EVP_CIPHER_CTX_new();
    if (!b->c) return ORA_OOM;
    if (EVP_CipherInit_ex(b->c, ci, NULL, key, NULL, enc) != 1) return ORA_INTERNAL;
    EVP_CIPHER_CTX_set_padding(b->c, 0);
    return ORA_OK;
}
static void blk_do(blk_t* b, const uint8_t* in, uint8_t* out, size_t n) {
    while (n) {
        int k = n > (1u << 30) ? (1 << 30) : (int)n, ol = 0;
        EVP_CipherUpdate(b->c, out, &ol, in, k);
        in += k; out += k; n -= (size_t)k;
    }
}
static void blk_free(blk_t* b) { if (b->c) EVP_CIPHER_CTX_free(b->c); b->c = NULL; }

static void ctr_add(uint8_t ctr[16], uint64_t add) { /* 128-bit big-endian add */
    for (int i = 15; i >= 0 && add; i--) {
        uint64_t s = (uint64_t)ctr[i] + (add & 0xFF);
        ctr[i] = (uint8_t)s;
        add = (add >> 8) + (s >> 8);
    }
}
/* CTR128-BE keystream XOR (both directions).  cipher.rs:26,33; stream/read.rs:39-43 */
int pna_oracle_ctr_restated(int encryption, const uint8_t key[32], const uint8_t iv[16], const uint8_t* in, size_t n,
                            uint8_t* out) {
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 1);
    if (rc) return rc;
    enum { B = 4096 };
    uint8_t ctrs[B * 16], ks[B * 16], ctr[16];
    memcpy(ctr, iv, 16);
    size_t done = 0;
    while (done < n) {
        size_t nb = (n - done + 15) / 16;
        if (nb > B) nb = B;
        for (size_t i = 0; i < nb; i++) { memcpy(ctrs + 16 * i, ctr, 16); ctr_add(ctr, 1); }
        blk_do(&b, ctrs, ks, nb * 16);
        size_t k = nb * 16;
        if (k > n - done) k = n - done;
        for (size_t i = 0; i < k; i++) out[done + i] = in[done + i] ^ ks[i];
        done += k;
    }
    blk_free(&b);
    return ORA_OK;
}
/* Same function through OpenSSL's own CTR mode (full-width 128-bit big-endian counter == ctr::Ctr128BE); this is the
 * pipelined AES-NI path and is what the CPU baseline times.  tests/test_oracle.py checks it against the restated
 * counter above. */
int pna_oracle_ctr(int encryption, const uint8_t key[32], const uint8_t iv[16], const uint8_t* in, size_t n,
                   uint8_t* out) {
    const EVP_CIPHER* ci = encryption == 1 ? EVP_aes_256_ctr() : encryption == 2 ? EVP_camellia_256_ctr() : NULL;
    if (!ci) return ORA_UNSUPPORTED;
    EVP_CIPHER_CTX* c = EVP_CIPHER_CTX_new();
    if (!c) return ORA_OOM;
    if (EVP_EncryptInit_ex(c, ci, NULL, key, iv) != 1) { EVP_CIPHER_CTX_free(c); return ORA_INTERNAL; }
    while (n) {
        int k = n > (1u << 30) ? (1 << 30) : (int)n, ol = 0;
        EVP_EncryptUpdate(c, out, &ol, in, k);
        in += k; out += k; n -= (size_t)k;
    }
    EVP_CIPHER_CTX_free(c);
    return ORA_OK;
}
/* CBC decrypt + PKCS#7 unpad.  cipher/block/read.rs:31-116 */
int pna_oracle_cbc_decrypt(int encryption, const uint8_t key[32], const uint8_t iv[16], const uint8_t* in,
                           size_t n, uint8_t* out, size_t* out_len) {
    *out_len = 0;
    if (n < 16) return ORA_UNEXPECTED_EOF;       /* new(): read_exact(first block)  read.rs:36 */
    if (n % 16) return ORA_UNEXPECTED_EOF;       /* partial trailing block           read.rs:90 */
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 0);
    if (rc) return rc;
    blk_do(&b, in, out, n);
    for (size_t i = 0; i < n; i++) out[i] ^= (i < 16) ? iv[i] : in[i - 16];
    blk_free(&b);
    uint8_t pad = out[n - 1];                    /* Pkcs7::unpad                     read.rs:101 */
    if (pad == 0 || pad > 16) return ORA_INVALID_DATA;
    for (size_t i = n - pad; i < n; i++) if (out[i] != pad) return ORA_INVALID_DATA;
    *out_len = n - pad;
    return ORA_OK;
}
/* CBC encrypt + PKCS#7 pad (always appends 1..16).  cipher/block/write.rs:48-124 */
int pna_oracle_cbc_encrypt(int encryption, const uint8_t key[32], const uint8_t iv[16], const uint8_t* in,
                           size_t n, uint8_t* out, size_t* out_len) {
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 1);
    if (rc) return rc;
    size_t total = (n / 16 + 1) * 16;
    uint8_t prev[16], blk[16];
    memcpy(prev, iv, 16);
    for (size_t off = 0; off < total; off += 16) {
        size_t have = off < n ? (n - off >= 16 ? 16 : n - off) : 0;
        uint8_t pad = (uint8_t)(16 - have);
        for (size_t i = 0; i < 16; i++) blk[i] = (i < have ? in[off + i] : pad) ^ prev[i];
        blk_do(&b, blk, prev, 16);
        memcpy(out + off, prev, 16);
    }
    blk_free(&b);
    *out_len = total;
    return ORA_OK;
}
/* single-block ECB primitive, exposed so tests can pin the GPU block ciphers directly */
int pna_oracle_ecb(int encryption, int encrypt, const uint8_t key[32], const uint8_t* in, size_t n, uint8_t* out) {
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, encrypt);
    if (rc) return rc;
    blk_do(&b, in, out, n & ~(size_t)15);
    blk_free(&b);
    return ORA_OK;
}

/* ------------------------------------------------------------------ GCM STREAM (cipher mode 2) */
/* lib/src/cipher/gcm.rs:206-290 (segment reader), :44-63 (segment writer), lib/src/cipher/aead.rs:92-150,210-217
 * (stream header layout, segment nonce).  aes-gcm 0.11 `AesGcm<C, U12>` is standard GCM (NIST SP 800-38D) over a
 * 128-bit block cipher with a 96-bit nonce, no AAD, 16-byte detached tag: restated here over the ECB primitive
 * (so Camellia works too); gcm_ghash_mul is SP 800-38D Algorithm 1, bit by bit.  tests/test_oracle.py checks the
 * restatement against OpenSSL's EVP_aes_256_gcm and against the reference's GCM fixtures. */
enum { GCM_HDR = 75, GCM_TAG = 16, GCM_MAX_SEG = 67108864 };
static void gcm_ghash_mul(uint8_t x[16], const uint8_t h[16]) { /* x <- x * h */
    uint8_t z[16] = {0}, v[16];
    memcpy(v, h, 16);
    for (int i = 0; i < 128; i++) {
        if ((x[i >> 3] >> (7 - (i & 7))) & 1) for (int k = 0; k < 16; k++) z[k] ^= v[k];
        int lsb = v[15] & 1;
        for (int k = 15; k > 0; k--) v[k] = (uint8_t)((v[k] >> 1) | (v[k - 1] << 7));
        v[0] >>= 1;
        if (lsb) v[0] ^= 0xE1;
    }
    memcpy(x, z, 16);
}
/* one segment, both directions: out = in ^ CTR keystream (inc32 from J0+1); tag over the CIPHERTEXT */
static void gcm_segment(blk_t* b, const uint8_t h[16], const uint8_t nonce[12], const uint8_t* in, size_t n, uint8_t* out,
                        int in_is_ciphertext, uint8_t tag[16]) {
    uint8_t j0[16], ek0[16], y[16] = {0}, ctr[16], ks[16];
    memcpy(j0, nonce, 12); j0[12] = j0[13] = j0[14] = 0; j0[15] = 1;
    blk_do(b, j0, ek0, 16);
    memcpy(ctr, j0, 16);
    for (size_t off = 0; off < n; off += 16) {
        size_t k = n - off < 16 ? n - off : 16;
        uint32_t c = ((uint32_t)ctr[12] << 24 | (uint32_t)ctr[13] << 16 | (uint32_t)ctr[14] << 8 | ctr[15]) + 1u; /* inc32 */
        ctr[12] = (uint8_t)(c >> 24); ctr[13] = (uint8_t)(c >> 16); ctr[14] = (uint8_t)(c >> 8); ctr[15] = (uint8_t)c;
        blk_do(b, ctr, ks, 16);
        uint8_t cblk[16] = {0};
        for (size_t i = 0; i < k; i++) {
            uint8_t o = in[off + i] ^ ks[i];
            cblk[i] = in_is_ciphertext ? in[off + i] : o;
            out[off + i] = o;
        }
        for (int i = 0; i < 16; i++) y[i] ^= cblk[i];
        gcm_ghash_mul(y, h);
    }
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 0; i < 8; i++) y[15 - i] ^= (uint8_t)(bits >> (8 * i)); /* len(A)=0 || len(C) */
    gcm_ghash_mul(y, h);
    for (int i = 0; i < 16; i++) tag[i] = y[i] ^ ek0[i];
}
/* AES segments go through OpenSSL's AES-256-GCM (AES-NI + PCLMUL: what the CPU baseline should time, like the aes-gcm crate's
 * hardware back ends); Camellia has no OpenSSL GCM and uses the restatement.  tests/test_gcm.py checks one against the other. */
static int gcm_segment_aes_evp(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* in, size_t n, uint8_t* out, int decrypt,
                               uint8_t tag[16]) {
    EVP_CIPHER_CTX* c = EVP_CIPHER_CTX_new();
    if (!c) return ORA_OOM;
    int ol = 0, rc = ORA_OK;
    if (EVP_CipherInit_ex(c, EVP_aes_256_gcm(), NULL, NULL, NULL, !decrypt) != 1 ||
        EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, 12, NULL) != 1 ||
        EVP_CipherInit_ex(c, NULL, NULL, key, nonce, !decrypt) != 1) rc = ORA_INTERNAL;
    size_t done = 0;
    while (rc == ORA_OK && done < n) {
        int k = n - done > (1u << 30) ? (1 << 30) : (int)(n - done);
        if (EVP_CipherUpdate(c, out + done, &ol, in + done, k) != 1) rc = ORA_INTERNAL;
        done += (size_t)k;
    }
    if (rc == ORA_OK && decrypt) {
        if (EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_TAG, 16, tag) != 1) rc = ORA_INTERNAL;
        else if (EVP_CipherFinal_ex(c, out + n, &ol) != 1) rc = ORA_INVALID_DATA;   /* tag mismatch */
    } else if (rc == ORA_OK) {
        if (EVP_CipherFinal_ex(c, out + n, &ol) != 1 || EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_GET_TAG, 16, tag) != 1) rc = ORA_INTERNAL;
    }
    EVP_CIPHER_CTX_free(c);
    return rc;
}
static void gcm_nonce(const uint8_t prefix[7], uint32_t counter, int is_final, uint8_t nonce[12]) { /* aead.rs:210 */
    memcpy(nonce, prefix, 7);
    nonce[7] = (uint8_t)(counter >> 24); nonce[8] = (uint8_t)(counter >> 16); nonce[9] = (uint8_t)(counter >> 8);
    nonce[10] = (uint8_t)counter; nonce[11] = is_final ? 1 : 0;
}
/* stream = header(75) || { ciphertext(<= segment_size) || tag(16) }...; key = the derived STREAM key.
 * Only tag-verified plaintext counts: on any error *out_len = bytes of the verified segments before it. */
int pna_oracle_gcm_decrypt_stream(int encryption, const uint8_t key[32], const uint8_t* stream, size_t n, uint8_t* out,
                                  size_t* out_len) {
    *out_len = 0;
    if (n < GCM_HDR) return ORA_INVALID_DATA;                /* "datastream shorter than the stream header" read.rs:108 */
    const uint32_t seg = (uint32_t)stream[39] << 24 | (uint32_t)stream[40] << 16 | (uint32_t)stream[41] << 8 | stream[42];
    if (seg == 0 || seg > GCM_MAX_SEG) return ORA_INVALID_DATA; /* "segment size out of range" aead.rs:141 */
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 1);
    if (rc) return rc;
    uint8_t h[16], zero[16] = {0};
    blk_do(&b, zero, h, 16);
    const uint8_t* p = stream + GCM_HDR;
    size_t rest = n - GCM_HDR, opos = 0;
    for (uint32_t i = 0;; i++) {
        size_t take = rest < (size_t)seg + GCM_TAG ? rest : (size_t)seg + GCM_TAG;
        int is_final = take == rest;                          /* look-ahead byte absent  gcm.rs:233-235 */
        if (take < GCM_TAG) { rc = ORA_INVALID_DATA; break; } /* malformed (i == 0) or truncation  gcm.rs:249-258 */
        uint8_t nonce[12], tag[16];
        gcm_nonce(stream + 32, i, is_final, nonce);
        if (encryption == 1) {
            memcpy(tag, p + take - GCM_TAG, 16);
            rc = gcm_segment_aes_evp(key, nonce, p, take - GCM_TAG, out + opos, 1, tag);
            if (rc) break;
        } else {
            gcm_segment(&b, h, nonce, p, take - GCM_TAG, out + opos, 1, tag);
            uint8_t diff = 0;
            for (int k = 0; k < 16; k++) diff |= (uint8_t)(tag[k] ^ p[take - GCM_TAG + k]);
            if (diff) { rc = ORA_INVALID_DATA; break; }       /* AuthenticationFailure  gcm.rs:283 */
        }
        opos += take - GCM_TAG; p += take; rest -= take;
        if (is_final) break;
        if (i == 0xFFFFFFFFu) { rc = ORA_INVALID_DATA; break; }
    }
    blk_free(&b);
    *out_len = opos;
    return rc;
}
/* GcmEncryptWriter: full segments are flushed when more data follows, finish() always emits a final one. */
size_t pna_oracle_gcm_encrypt_bound(size_t n, uint32_t seg) { return GCM_HDR + n + (n / seg + 1) * (size_t)GCM_TAG; }
int pna_oracle_gcm_encrypt_stream(int encryption, const uint8_t key[32], const uint8_t header[75], const uint8_t* plain,
                                  size_t n, uint8_t* out, size_t* out_len) {
    *out_len = 0;
    const uint32_t seg = (uint32_t)header[39] << 24 | (uint32_t)header[40] << 16 | (uint32_t)header[41] << 8 | header[42];
    if (seg == 0 || seg > GCM_MAX_SEG) return ORA_INVALID_INPUT;
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 1);
    if (rc) return rc;
    uint8_t h[16], zero[16] = {0};
    blk_do(&b, zero, h, 16);
    memcpy(out, header, GCM_HDR);
    size_t opos = GCM_HDR, ipos = 0;
    for (uint32_t i = 0;; i++) {
        size_t take = n - ipos < seg ? n - ipos : seg;
        int is_final = ipos + take == n;
        uint8_t nonce[12];
        gcm_nonce(header + 32, i, is_final, nonce);
        if (encryption == 1) {
            uint8_t tg[16];
            int r2 = gcm_segment_aes_evp(key, nonce, plain + ipos, take, out + opos, 0, tg);
            if (r2) { blk_free(&b); return r2; }
            memcpy(out + opos + take, tg, 16);
        } else gcm_segment(&b, h, nonce, plain + ipos, take, out + opos, 0, out + opos + take);
        opos += take + GCM_TAG; ipos += take;
        if (is_final) break;
    }
    blk_free(&b);
    *out_len = opos;
    return ORA_OK;
}
/* OpenSSL's own AES-256-GCM of one message (check of the restatement above; AES only) */
int pna_oracle_gcm_openssl(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* in, size_t n, uint8_t* out,
                           uint8_t tag[16]) {
    EVP_CIPHER_CTX* c = EVP_CIPHER_CTX_new();
    if (!c) return ORA_OOM;
    int ol = 0, rc = ORA_OK;
    if (EVP_EncryptInit_ex(c, EVP_aes_256_gcm(), NULL, NULL, NULL) != 1 ||
        EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_SET_IVLEN, 12, NULL) != 1 ||
        EVP_EncryptInit_ex(c, NULL, NULL, key, nonce) != 1) rc = ORA_INTERNAL;
    if (rc == ORA_OK && n && EVP_EncryptUpdate(c, out, &ol, in, (int)n) != 1) rc = ORA_INTERNAL;
    if (rc == ORA_OK && EVP_EncryptFinal_ex(c, out + ol, &ol) != 1) rc = ORA_INTERNAL;
    if (rc == ORA_OK && EVP_CIPHER_CTX_ctrl(c, EVP_CTRL_GCM_GET_TAG, 16, tag) != 1) rc = ORA_INTERNAL;
    EVP_CIPHER_CTX_free(c);
    return rc;
}
/* one restated segment, exposed for that check */
int pna_oracle_gcm_segment(int encryption, const uint8_t key[32], const uint8_t nonce[12], const uint8_t* in, size_t n,
                           uint8_t* out, uint8_t tag[16]) {
    blk_t b = {0};
    int rc = blk_init(&b, encryption, key, 1);
    if (rc) return rc;
    uint8_t h[16], zero[16] = {0};
    blk_do(&b, zero, h, 16);
    gcm_segment(&b, h, nonce, in, n, out, 0, tag);
    blk_free(&b);
    return ORA_OK;
}

/* ------------------------------------------------------------------ libzstd via dlopen */
typedef struct { const void* src; size_t size; size_t pos; } ZIn;
typedef struct { void* dst; size_t size; size_t pos; } ZOut;
static struct {
    void* h;
    void* (*createDStream)(void);
    size_t (*initDStream)(void*);
    size_t (*decompressStream)(void*, ZOut*, ZIn*);
    size_t (*freeDStream)(void*);
    unsigned (*isError)(size_t);
    void* (*createCCtx)(void);
    size_t (*freeCCtx)(void*);
    size_t (*setParameter)(void*, int, int);
    size_t (*compressStream2)(void*, ZOut*, ZIn*, int);
    unsigned (*versionNumber)(void);
} Z;
static pthread_once_t z_once = PTHREAD_ONCE_INIT;
static void z_load(void) {
    const char* env = getenv("PNA_ORACLE_LIBZSTD");
    const char* names[] = {env, "libzstd.so.1", "libzstd.so", NULL};
    for (int i = 0; i < 4 && !Z.h; i++) if (names[i]) Z.h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!Z.h) return;
    Z.createDStream = dlsym(Z.h, "ZSTD_createDStream");
    Z.initDStream = dlsym(Z.h, "ZSTD_initDStream");
    Z.decompressStream = dlsym(Z.h, "ZSTD_decompressStream");
    Z.freeDStream = dlsym(Z.h, "ZSTD_freeDStream");
    Z.isError = dlsym(Z.h, "ZSTD_isError");
    Z.createCCtx = dlsym(Z.h, "ZSTD_createCCtx");
    Z.freeCCtx = dlsym(Z.h, "ZSTD_freeCCtx");
    Z.setParameter = dlsym(Z.h, "ZSTD_CCtx_setParameter");
    Z.compressStream2 = dlsym(Z.h, "ZSTD_compressStream2");
    Z.versionNumber = dlsym(Z.h, "ZSTD_versionNumber");
}
unsigned pna_oracle_zstd_version(void) {
    pthread_once(&z_once, z_load);
    return Z.h && Z.versionNumber ? Z.versionNumber() : 0;
}

/* ------------------------------------------------------------------ decompress */
/* entry/read.rs:171-190.  out==NULL/cap==0 is allowed: the function then only sizes.
 * On ORA_NOSPACE *out_len holds the full decoded size (two-pass sizing contract). */
static int zstd_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    pthread_once(&z_once, z_load);
    if (!Z.h) return ORA_INTERNAL;
    void* ds = Z.createDStream();
    if (!ds) return ORA_OOM;
    Z.initDStream(ds);
    ZIn zi = {in, n, 0};
    size_t total = 0, ret = 0;
    int rc = ORA_OK, overflow = 0;
    uint8_t scratch[1 << 17];
    /* zstd-rs Decoder (zio::Reader) keeps decoding concatenated frames until EOF */
    while (zi.pos < zi.size) {
        ZOut zo;
        int to_scratch = overflow || total >= cap;
        if (!to_scratch) { zo.dst = out + total; zo.size = cap - total; }
        else { zo.dst = scratch; zo.size = sizeof scratch; }
        zo.pos = 0;
        ret = Z.decompressStream(ds, &zo, &zi);
        if (Z.isError(ret)) { rc = ORA_INVALID_DATA; break; }
        total += zo.pos;
        if (to_scratch && zo.pos) overflow = 1;
        if (zo.pos == 0 && zi.pos >= zi.size) break;
    }
    /* flush what is still buffered inside the decoder once input is exhausted */
    while (rc == ORA_OK && ret != 0) {
        ZOut zo;
        int to_scratch = overflow || total >= cap;
        if (!to_scratch) { zo.dst = out + total; zo.size = cap - total; }
        else { zo.dst = scratch; zo.size = sizeof scratch; }
        zo.pos = 0;
        ret = Z.decompressStream(ds, &zo, &zi);
        if (Z.isError(ret)) { rc = ORA_INVALID_DATA; break; }
        total += zo.pos;
        if (to_scratch && zo.pos) overflow = 1;
        if (zo.pos == 0) break;
    }
    Z.freeDStream(ds);
    *out_len = total;
    if (rc) return rc;
    if (ret != 0) return ORA_UNEXPECTED_EOF; /* "incomplete frame" */
    return overflow ? ORA_NOSPACE : ORA_OK;
}
static int zlib_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) return ORA_OOM;
    uint8_t scratch[1 << 16];
    size_t total = 0, ipos = 0;
    int rc = ORA_OK, overflow = 0, zr = Z_OK;
    for (;;) {
        size_t ik = n - ipos > (1u << 30) ? (1u << 30) : n - ipos;
        zs.next_in = (Bytef*)(in + ipos);
        zs.avail_in = (uInt)ik;
        size_t ok;
        int to_scratch = overflow || total >= cap;
        if (!to_scratch) { zs.next_out = out + total; ok = cap - total; if (ok > (1u << 30)) ok = 1u << 30; }
        else { zs.next_out = scratch; ok = sizeof scratch; }
        zs.avail_out = (uInt)ok;
        zr = inflate(&zs, Z_NO_FLUSH);
        ipos += ik - zs.avail_in;
        total += ok - zs.avail_out;
        if (to_scratch && ok != zs.avail_out) overflow = 1;
        if (zr == Z_STREAM_END) break;
        if (zr == Z_BUF_ERROR && ipos >= n) break; /* truncated: flate2 zio::read returns Ok(short) */
        if (zr != Z_OK && zr != Z_BUF_ERROR) { rc = ORA_INVALID_INPUT; break; } /* "corrupt deflate stream" */
        if (zr == Z_OK && ipos >= n && zs.avail_out != 0) break;
    }
    inflateEnd(&zs);
    *out_len = total;
    if (rc) return rc;
    return overflow ? ORA_NOSPACE : ORA_OK;
}
int pna_oracle_decompress(int compression, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                          size_t* out_len) {
    *out_len = 0;
    switch (compression) {
        case 0: /* Compression::NO */
            *out_len = n;
            if (n > cap) return ORA_NOSPACE;
            memcpy(out, in, n);
            return ORA_OK;
        case 1: return zlib_decode(in, n, out, cap, out_len);
        case 2: return zstd_decode(in, n, out, cap, out_len);
        default: return ORA_UNSUPPORTED; /* XZ=4 and unknown: outside the hot path */
    }
}

/* ------------------------------------------------------------------ compress */
/* entry/write.rs:251-265: streaming encoders, size never pledged, 32 KiB-ish writes downstream.
 * level<0 selects the reference default (zstd 3: compress/zstandard.rs:46; deflate 6: deflate.rs:89). */
size_t pna_oracle_compress_bound(int compression, size_t n) {
    if (compression == 0) return n;
    return n + n / 8 + 1024;
}
int pna_oracle_compress(int compression, int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap,
                        size_t* out_len) {
    *out_len = 0;
    if (compression == 0) {
        if (n > cap) return ORA_NOSPACE;
        memcpy(out, in, n);
        *out_len = n;
        return ORA_OK;
    }
    if (compression == 2) {
        pthread_once(&z_once, z_load);
        if (!Z.h) return ORA_INTERNAL;
        void* cc = Z.createCCtx();
        if (!cc) return ORA_OOM;
        Z.setParameter(cc, 100 /* ZSTD_c_compressionLevel */, level < 0 ? 3 : level);
        ZIn zi = {in, n, 0};
        ZOut zo = {out, cap, 0};
        int rc = ORA_OK;
        /* io::Write::write_all in 128 KiB pieces then finish(): ZSTD_e_continue ... ZSTD_e_end */
        while (zi.pos < zi.size) {
            ZIn piece = {in, zi.pos + (n - zi.pos > (1u << 17) ? (1u << 17) : n - zi.pos), zi.pos};
            size_t r = Z.compressStream2(cc, &zo, &piece, 0);
            if (Z.isError(r)) { rc = ORA_INTERNAL; break; }
            if (piece.pos == zi.pos && zo.pos == zo.size) { rc = ORA_NOSPACE; break; }
            zi.pos = piece.pos;
        }
        while (rc == ORA_OK) {
            ZIn end = {in, n, n};
            size_t r = Z.compressStream2(cc, &zo, &end, 2 /* ZSTD_e_end */);
            if (Z.isError(r)) { rc = ORA_INTERNAL; break; }
            if (r == 0) break;
            if (zo.pos == zo.size) { rc = ORA_NOSPACE; break; }
        }
        Z.freeCCtx(cc);
        *out_len = zo.pos;
        return rc;
    }
    if (compression == 1) {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit(&zs, level < 0 ? 6 : level) != Z_OK) return ORA_OOM;
        size_t ipos = 0, opos = 0;
        int rc = ORA_OK;
        for (;;) {
            size_t ik = n - ipos > (1u << 30) ? (1u << 30) : n - ipos;
            size_t ok = cap - opos > (1u << 30) ? (1u << 30) : cap - opos;
            zs.next_in = (Bytef*)(in + ipos); zs.avail_in = (uInt)ik;
            zs.next_out = out + opos; zs.avail_out = (uInt)ok;
            int zr = deflate(&zs, ipos + ik >= n ? Z_FINISH : Z_NO_FLUSH);
            ipos += ik - zs.avail_in;
            opos += ok - zs.avail_out;
            if (zr == Z_STREAM_END) break;
            if (zr != Z_OK && zr != Z_BUF_ERROR) { rc = ORA_INTERNAL; break; }
            if (opos >= cap) { rc = ORA_NOSPACE; break; }
        }
        deflateEnd(&zs);
        *out_len = opos;
        return rc;
    }
    return ORA_UNSUPPORTED;
}

/* ------------------------------------------------------------------ whole-stream pipeline */
/* One entry's data stream = concat(FDAT/SDAT bodies) = [IV 16B if encrypted] || cipher(compress(plain)).
 * NormalEntry::reader  lib/src/entry.rs:1150  ->  decrypt_reader + decompress_reader. */
int pna_oracle_decode_stream(const uint8_t* stream, size_t n, int compression, int encryption, int cipher_mode,
                             const uint8_t key[32], uint8_t* out, size_t cap, size_t* out_len) {
    *out_len = 0;
    if (compression != 0 && compression != 1 && compression != 2) return ORA_UNSUPPORTED;
    if (encryption == 0) return pna_oracle_decompress(compression, stream, n, out, cap, out_len);
    if (encryption != 1 && encryption != 2) return ORA_UNSUPPORTED;
    if (cipher_mode != 0 && cipher_mode != 1 && cipher_mode != 2) return ORA_UNSUPPORTED;
    if (cipher_mode == 2) {                                           /* GCM STREAM: key = derived stream key */
        uint8_t* t2 = (uint8_t*)malloc(n ? n : 1);
        if (!t2) return ORA_OOM;
        size_t clen2 = 0;
        int rc2 = pna_oracle_gcm_decrypt_stream(encryption, key, stream, n, t2, &clen2);
        if (rc2 == ORA_OK) rc2 = pna_oracle_decompress(compression, t2, clen2, out, cap, out_len);
        free(t2);
        return rc2;
    }
    if (n < 16) return ORA_UNEXPECTED_EOF;                            /* read_exact(iv) entry/read.rs:80 */
    uint8_t* tmp = (uint8_t*)malloc(n ? n : 1);
    if (!tmp) return ORA_OOM;
    size_t clen = n - 16;
    int rc;
    if (cipher_mode == 1) rc = pna_oracle_ctr(encryption, key, stream, stream + 16, clen, tmp);
    else rc = pna_oracle_cbc_decrypt(encryption, key, stream, stream + 16, clen, tmp, &clen);
    if (rc == ORA_OK) rc = pna_oracle_decompress(compression, tmp, clen, out, cap, out_len);
    free(tmp);
    return rc;
}
/* FileEntryBuilder write path: get_writer = compression_writer(encryption_writer(w)); the IV is the
 * stream prefix (entry/write.rs:46-50,268-273; builder.rs:62-69). */
size_t pna_oracle_encode_bound(int compression, size_t n) { return pna_oracle_compress_bound(compression, n) + 48; }
int pna_oracle_encode_stream(const uint8_t* plain, size_t n, int compression, int level, int encryption,
                             int cipher_mode, const uint8_t key[32], const uint8_t iv[16], uint8_t* out, size_t cap,
                             size_t* out_len) {
    *out_len = 0;
    if (encryption == 0) return pna_oracle_compress(compression, level, plain, n, out, cap, out_len);
    if (encryption != 1 && encryption != 2) return ORA_UNSUPPORTED;
    if (cipher_mode != 0 && cipher_mode != 1) return ORA_UNSUPPORTED;
    size_t bound = pna_oracle_compress_bound(compression, n);
    uint8_t* tmp = (uint8_t*)malloc(bound ? bound : 1);
    if (!tmp) return ORA_OOM;
    size_t clen = 0;
    int rc = pna_oracle_compress(compression, level, plain, n, tmp, bound, &clen);
    if (rc == ORA_OK) {
        size_t need = 16 + (cipher_mode == 1 ? clen : (clen / 16 + 1) * 16);
        if (need > cap) rc = ORA_NOSPACE;
        else {
            memcpy(out, iv, 16);
            if (cipher_mode == 1) rc = pna_oracle_ctr(encryption, key, iv, tmp, clen, out + 16);
            else rc = pna_oracle_cbc_encrypt(encryption, key, iv, tmp, clen, out + 16, &clen);
            *out_len = 16 + clen;
        }
    }
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------ folding CRC-32 (PCLMULQDQ) */
/* The reference's chunk CRC is crc32fast 1.5.0 (lib/src/format/chunk.rs:8-11), whose x86-64 fast path
 * (crc32fast src/specialized/pclmulqdq.rs) is the carry-less-multiply folding of Gopal, Ozturk, Guilford et al., "Fast CRC
 * Computation for Generic Polynomials Using PCLMULQDQ Instruction" (Intel white paper, 2009): four 128-bit lanes folded
 * across 64 bytes per step, then 128 -> 64 -> 32 bits by Barrett reduction.  zlib 1.3's crc32() is a table method several
 * times slower, so timing the reference's serial CRC thread with it would flatter the GPU; this restatement of the published
 * method (constants: x^k mod P for the reflected polynomial 0xEDB88320, as tabulated in the paper) is what the CPU baseline
 * uses.  Pinned on the reference's CRC KATs and against zlib in tests/test_oracle.py. */
#if defined(__x86_64__)
#include <immintrin.h>
#include <cpuid.h>
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_fold_pclmul(const uint8_t* buf, size_t len, uint32_t state) {   /* len >= 64, len % 16 == 0; raw (pre-inverted) state */
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);   /* fold by 512 bits */
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);   /* fold by 128 bits */
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
    const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);   /* mu, P */
    const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(buf + 0)), x2 = _mm_loadu_si128((const __m128i*)(buf + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(buf + 32)), x4 = _mm_loadu_si128((const __m128i*)(buf + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)state));
    buf += 64; len -= 64;
    while (len >= 64) {
        __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11); x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11); x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i*)(buf + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i*)(buf + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i*)(buf + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i*)(buf + 48)));
        buf += 64; len -= 64;
    }
    __m128i t;
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00); x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), t);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00); x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), t);
    t = _mm_clmulepi64_si128(x1, k3k4, 0x00); x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), t);
    while (len >= 16) {
        t = _mm_clmulepi64_si128(x1, k3k4, 0x00); x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, _mm_loadu_si128((const __m128i*)buf)), t);
        buf += 16; len -= 16;
    }
    /* 128 -> 64 bits */
    x2 = _mm_clmulepi64_si128(x1, k3k4, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_clmulepi64_si128(_mm_and_si128(x1, mask32), k5, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    /* Barrett reduction to 32 bits */
    x2 = _mm_clmulepi64_si128(_mm_and_si128(x1, mask32), poly, 0x10);
    x2 = _mm_clmulepi64_si128(_mm_and_si128(x2, mask32), poly, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
static int have_pclmul(void) {
    static int cached = -1;
    if (cached < 0) { unsigned a, b, c, d; cached = __get_cpuid(1, &a, &b, &c, &d) && (c & (1u << 1)) && (c & (1u << 19)); }
    return cached;
}
#else
static int have_pclmul(void) { return 0; }
#endif
int pna_oracle_have_pclmul(void) { return have_pclmul(); }
/* same contract as zlib's crc32(): running value in, running value out */
uint32_t pna_oracle_crc32_fold(uint32_t crc, const uint8_t* buf, size_t len) {
#if defined(__x86_64__)
    if (have_pclmul() && len >= 64) {
        const size_t bulk = len & ~(size_t)15;
        crc = ~crc32_fold_pclmul(buf, bulk, ~crc);
        buf += bulk; len -= bulk;
    }
#endif
    while (len) { uInt k = len > (1u << 30) ? (1u << 30) : (uInt)len; crc = (uint32_t)crc32(crc, buf, k); buf += k; len -= k; }
    return crc;
}
/* crc_impl: 1 = zlib table CRC, 2 = folding CRC (what crc32fast does on x86-64) */
static uint32_t chunk_crc_impl(int crc_impl, const char ty[4], const uint8_t* s, size_t l) {
    uint32_t z = (uint32_t)crc32(0L, (const Bytef*)ty, 4);
    if (crc_impl == 2) return pna_oracle_crc32_fold(z, s, l);
    while (l) { uInt k = l > (1u << 30) ? (1u << 30) : (uInt)l; z = (uint32_t)crc32(z, s, k); s += k; l -= k; }
    return z;
}

/* ------------------------------------------------------------------ task-per-entry pool (CPU baseline) */
/* Mirrors the CLI's extract dataflow: ONE thread walks the archive and CRC-checks every chunk
 * (archive/read/slice.rs:42-66 -> bytes.rs:39-75), worker threads decode one entry each
 * (extract.rs:987 spawn_fifo -> extract_file_entry :1301).  Streams are passed already concatenated. */
typedef struct {
    const uint8_t* stream; uint64_t len;
    uint8_t compression, encryption, cipher_mode, _pad;
    uint8_t key[32];
    uint8_t* out; uint64_t cap; uint64_t out_len;
    int32_t status;
} ora_job;
typedef struct { ora_job* jobs; uint32_t n; volatile uint32_t next; int do_crc; } ora_pool;
static void* ora_worker(void* arg) {
    ora_pool* p = (ora_pool*)arg;
    for (;;) {
        uint32_t i = __sync_fetch_and_add(&p->next, 1);
        if (i >= p->n) break;
        ora_job* j = &p->jobs[i];
        size_t ol = 0;
        j->status = pna_oracle_decode_stream(j->stream, j->len, j->compression, j->encryption, j->cipher_mode,
                                             j->key, j->out, j->cap, &ol);
        j->out_len = ol;
    }
    return NULL;
}
/* crc_on_caller_thread: 0 = no CRC, 1 = zlib's table CRC, 2 = the folding (PCLMULQDQ) CRC the reference really runs */
int pna_oracle_decode_batch_mt(ora_job* jobs, uint32_t n, int nthreads, int crc_on_caller_thread,
                               uint32_t* crc_out) {
    ora_pool p = {jobs, n, 0, crc_on_caller_thread};
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return ORA_OOM;
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, ora_worker, &p);
    if (crc_on_caller_thread) /* the serial iterating thread of the reference */
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t z = chunk_crc_impl(crc_on_caller_thread, "FDAT", jobs[i].stream, jobs[i].len);
            if (crc_out) crc_out[i] = z;
        }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return ORA_OK;
}

/* create-side pool: workers compress+encrypt one entry each (core.rs:496-510,915), the caller thread CRCs
 * the produced FDAT bodies (archive/write.rs:368 -> io.rs:183 -> chunk/traits.rs:49). */
typedef struct {
    const uint8_t* plain; uint64_t len;
    uint8_t compression, encryption, cipher_mode, _pad; int32_t level;
    uint8_t key[32]; uint8_t iv[16];
    uint8_t* out; uint64_t cap; uint64_t out_len;
    int32_t status;
} ora_enc_job;
typedef struct { ora_enc_job* jobs; uint32_t n; volatile uint32_t next; } ora_enc_pool;
static void* ora_enc_worker(void* arg) {
    ora_enc_pool* p = (ora_enc_pool*)arg;
    for (;;) {
        uint32_t i = __sync_fetch_and_add(&p->next, 1);
        if (i >= p->n) break;
        ora_enc_job* j = &p->jobs[i];
        size_t ol = 0;
        j->status = pna_oracle_encode_stream(j->plain, j->len, j->compression, j->level, j->encryption,
                                             j->cipher_mode, j->key, j->iv, j->out, j->cap, &ol);
        j->out_len = ol;
    }
    return NULL;
}
static int g_encode_crc_impl = 2;   /* the writer's chunk CRC (io.rs:187) is crc32fast too */
void pna_oracle_set_encode_crc_impl(int impl) { g_encode_crc_impl = impl == 1 ? 1 : 2; }
int pna_oracle_encode_batch_mt(ora_enc_job* jobs, uint32_t n, int nthreads, uint32_t* crc_out) {
    ora_enc_pool p = {jobs, n, 0};
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return ORA_OOM;
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, ora_enc_worker, &p);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    if (crc_out)
        for (uint32_t i = 0; i < n; i++) crc_out[i] = chunk_crc_impl(g_encode_crc_impl, "FDAT", jobs[i].out, jobs[i].out_len);
    return ORA_OK;
}
