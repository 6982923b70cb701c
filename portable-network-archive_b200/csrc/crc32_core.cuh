// crc32_core.cuh -- CRC-32/ISO-HDLC (reflected 0xEDB88320, init/xorout 0xFFFFFFFF) arithmetic.
//
// Replaces format::chunk_crc (/root/reference/lib/src/format/chunk.rs:7-12, crate crc32fast 1.5.0)
// for batches of spans.  Strategy (no PCLMUL on a GPU):
//   * a span is cut into tiles of <= CRC_TILE bytes; one warp computes the RAW remainder (init 0, no
//     xorout) of a tile: rows of 512 B, lane l owns bytes [16l,16l+16) of every row (coalesced 16-byte
//     loads), and folds   state <- state * x^(8*512)  xor  raw16(chunk)   with 20 byte-indexed tables
//     (slicing-by-16 plus a 4-table "advance 512 bytes");
//   * rows are anchored at the 16-byte-aligned END of the tile, leading zeros are free for a raw CRC,
//     the <=15 trailing pad bytes are undone by multiplying with x^(-8z);
//   * lanes are combined with per-lane constants x^(8*16*(31-l)) and an XOR shuffle tree;
//   * tiles are combined per span with  acc <- acc * x^(8*len) xor raw  (GF(2) polynomial arithmetic
//     mod P in the reflected representation, same identities as zlib's crc32_combine).
#pragma once
#include "common.cuh"

namespace pna {

constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr uint32_t CRC_TILE = 64 * 1024;   // bytes per warp tile (multiple of 512)
constexpr int CRC_NTAB = 20;               // U_0..U_15, U_508..U_511

struct CrcConsts {
    uint32_t U[CRC_NTAB][256];  // U_k[b] = register after feeding byte b then k zero bytes (init 0)
    uint32_t x2n[32];           // x^(2^n) mod P
    uint32_t lane_k[32];        // x^(8*16*(31-l))
    uint32_t inv_z[16];         // x^(-8z)
    uint32_t x_tile;            // x^(8*CRC_TILE)
};

// a*b mod P, reflected representation (x^0 == 0x80000000)
PNA_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b >> 1) ^ (CRC_POLY & (0u - (b & 1u)));
    }
    return p;
}
// x^(n * 2^k) mod P
PNA_HD uint32_t crc_x2nmodp(const uint32_t* x2n, uint64_t n, unsigned k) {
    uint32_t p = 0x80000000u;
    while (n) {
        if (n & 1) p = crc_multmodp(x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

inline void crc_make_consts(CrcConsts* C) {
    uint32_t t0[256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CRC_POLY & (0u - (c & 1u)));
        t0[i] = c;
    }
    // U_k for k = 0..511, keep the 20 we need
    uint32_t cur[256];   // (not static: contexts are created concurrently by the host layer's worker threads)
    for (int i = 0; i < 256; i++) cur[i] = t0[i];
    for (int k = 0; k < 512; k++) {
        int slot = k < 16 ? k : (k >= 508 ? 16 + (k - 508) : -1);
        if (slot >= 0)
            for (int i = 0; i < 256; i++) C->U[slot][i] = cur[i];
        for (int i = 0; i < 256; i++) cur[i] = t0[cur[i] & 0xFF] ^ (cur[i] >> 8);
    }
    uint32_t p = 1u << 30;  // x^1
    C->x2n[0] = p;
    for (int n = 1; n < 32; n++) C->x2n[n] = p = crc_multmodp(p, p);
    for (int l = 0; l < 32; l++) C->lane_k[l] = crc_x2nmodp(C->x2n, (uint64_t)16 * (31 - l), 3);
    // P is irreducible of degree 32 => x^(2^32-1) == 1, so x^(-8z) = x^(2^32-1-8z)
    for (int z = 0; z < 16; z++) C->inv_z[z] = crc_x2nmodp(C->x2n, 0xFFFFFFFFull - 8ull * z, 0);
    C->x_tile = crc_x2nmodp(C->x2n, CRC_TILE, 3);
}

// raw remainder of one 16-byte chunk folded into a lane state that is 512 bytes "behind"
PNA_HD uint32_t crc_fold_row(uint32_t s, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, const uint32_t* U) {
    // U laid out [tab*256 + byte]; tabs 16..19 = U_508..U_511
    uint32_t r = U[19 * 256 + (s & 0xFF)] ^ U[18 * 256 + ((s >> 8) & 0xFF)] ^ U[17 * 256 + ((s >> 16) & 0xFF)] ^
                 U[16 * 256 + (s >> 24)];
    r ^= U[15 * 256 + (w0 & 0xFF)] ^ U[14 * 256 + ((w0 >> 8) & 0xFF)] ^ U[13 * 256 + ((w0 >> 16) & 0xFF)] ^ U[12 * 256 + (w0 >> 24)];
    r ^= U[11 * 256 + (w1 & 0xFF)] ^ U[10 * 256 + ((w1 >> 8) & 0xFF)] ^ U[9 * 256 + ((w1 >> 16) & 0xFF)] ^ U[8 * 256 + (w1 >> 24)];
    r ^= U[7 * 256 + (w2 & 0xFF)] ^ U[6 * 256 + ((w2 >> 8) & 0xFF)] ^ U[5 * 256 + ((w2 >> 16) & 0xFF)] ^ U[4 * 256 + (w2 >> 24)];
    r ^= U[3 * 256 + (w3 & 0xFF)] ^ U[2 * 256 + ((w3 >> 8) & 0xFF)] ^ U[1 * 256 + ((w3 >> 16) & 0xFF)] ^ U[0 * 256 + (w3 >> 24)];
    return r;
}

// byte mask helper: keep the bytes of the aligned 16-byte chunk at address `a` that lie in [S,E)
PNA_HD void crc_mask_chunk(uint64_t a, uint64_t S, uint64_t E, uint32_t w[4]) {
    if (a >= S && a + 16 <= E) return;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint64_t p = a + 4 * k + b;
            if (p >= S && p < E) m |= 0xFFu << (8 * b);
        }
        w[k] &= m;
    }
}

// One lane's share of a tile [S,E) of the image (E-S <= CRC_TILE).  img16 = image as aligned 16-byte words.
// Returns the lane's partial already multiplied by its lane constant; XOR over the 32 lanes, then
// multiply by inv_z[(16 - E%16)%16] to get the raw remainder of the tile.
PNA_HD uint32_t crc_tile_lane(const uint8_t* img, uint64_t S, uint64_t E, int lane, const uint32_t* U,
                              const uint32_t* lane_k) {
    uint64_t A = (E + 15) & ~(uint64_t)15;             // aligned end
    uint64_t span = A - (S & ~(uint64_t)15);           // aligned bytes covered
    uint64_t rows = (span + 511) / 512;
    uint32_t s = 0;
    for (uint64_t r = rows; r-- > 0;) {                // r = rows-1 is the FIRST (oldest) row
        int64_t a = (int64_t)A - (int64_t)(r + 1) * 512 + 16 * lane;
        uint32_t w[4] = {0, 0, 0, 0};
        if (a + 16 > (int64_t)S && a < (int64_t)E && a >= 0) {
            const uint32_t* q = (const uint32_t*)(img + a);
            w[0] = q[0]; w[1] = q[1]; w[2] = q[2]; w[3] = q[3];
            crc_mask_chunk((uint64_t)a, S, E, w);
        }
        s = crc_fold_row(s, w[0], w[1], w[2], w[3], U);
    }
    return crc_multmodp(s, lane_k[lane]);
}

}  // namespace pna
