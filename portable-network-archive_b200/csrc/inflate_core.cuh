// inflate_core.cuh -- zlib (RFC 1950) / DEFLATE (RFC 1951) decode, one stream per thread, PNA_HD.
//
// Replaces `flate2::bufread::ZlibDecoder` at /root/reference/lib/src/entry/read.rs:179
// (flate2 1.1.9 -> miniz_oxide 0.8.5, not vendored).  Error behaviour follows flate2's zio::read:
// a corrupt stream is io::ErrorKind::InvalidInput ("corrupt deflate stream"); a TRUNCATED stream
// returns the bytes produced so far without an error; bytes after the Adler-32 trailer are ignored.
//
// Sizing contract: when the output does not fit, decoding continues in count-only mode (the symbol
// stream never depends on output bytes), and ST_NOSPACE is returned with the exact decoded length.
#pragma once
#include "common.cuh"

namespace pna {
namespace inf {

constexpr int MAXBITS = 15, MAXL = 288, MAXD = 30;
constexpr int FAST_BITS = 9;

// Per-stream decoding tables (canonical Huffman): an FB-bit direct table in front of the
// count/symbol arrays (FB = 0: no direct table).  fast[] entry: (symbol << 4) | length, 0 = not
// resolvable in FB bits.  The literal/length code gets a 9-bit table, the distance code none, so that
// one stream's tables are 1.8 KB (an odd number of 32-bit words: the lanes of a warp start on different banks).
// Codes the direct table does not resolve are found without a bit-by-bit walk: the next 15 bits, MSB first, are compared
// against the 15 left-aligned length bounds lim[] (branch-free, no loop-carried dependency).
template <int NSYM, int FB>
struct Huff {
    uint16_t lim[MAXBITS + 1];    // lim[l], l = 1..15: first code of length l+1, left-aligned to 15 bits = exclusive bound of the codes of length <= l
    int16_t base[MAXBITS + 1];    // symbol[] index of a code of length l = base[l] + code
    uint16_t n01;                 // symbols with code length 0 or 1 (the callers' incomplete-code rule)
    uint16_t symbol[NSYM];
    uint16_t fast[FB ? (1 << FB) : 1];
};
struct Tables {
    Huff<MAXL, FAST_BITS> len;
    Huff<MAXD + 2, 0> dist;
    uint16_t _pad[3];   // 1804 bytes = 451 words: an odd number of 32-bit words per stream
};

// LSB-first bit reader.  On the device the next four input bytes are always loaded one refill AHEAD (`nx`), so the
// load's latency (L2 every 128 bytes, L1 otherwise) is off the symbol chain: a refill only shifts a register in.
// The arena behind every stream is padded, so the look-ahead may read (never use) a few bytes past the end.
struct Bits {
    const uint8_t* p;
    uint64_t n, pos;
    uint64_t buf;
    int cnt;
#if defined(__CUDA_ARCH__)
    uint32_t nx;   // bytes [pos, pos + 4) of the input, little endian
    PNA_HD void reload() { nx = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16) | ((uint32_t)p[pos + 3] << 24); }
    PNA_HD void init(const uint8_t* in, uint64_t len) { p = in; n = len; pos = 0; buf = 0; cnt = 0; nx = 0; if (len) reload(); }
    PNA_HD void fill(int need) {   // need <= 32
        if (cnt >= need) return;
        uint32_t w = nx;
        if (pos + 4 > n) {         // zeros beyond the end; overrun() tells
            const uint32_t valid = pos < n ? (uint32_t)(n - pos) : 0u;
            w = valid ? (w & (0xFFFFFFFFu >> (8u * (4u - valid)))) : 0u;
        }
        if (cnt <= 32) {
            buf |= (uint64_t)w << cnt;
            pos += 4;
            cnt += 32;
            reload();
        }
    }
    PNA_HD void align_byte() { pos = (consumed_bits() + 7) / 8; buf = 0; cnt = 0; reload(); }
    PNA_HD void skip_bytes(uint64_t k) { pos += k; reload(); }   // caller consumed k bytes at p + pos directly (stored block)
#else
    PNA_HD void reload() {}
    PNA_HD void init(const uint8_t* in, uint64_t len) { p = in; n = len; pos = 0; buf = 0; cnt = 0; }
    PNA_HD void fill(int need) {
        if (cnt >= need) return;
        if (cnt <= 32 && pos + 4 <= n) {   // four input bytes per refill: the loads are independent, one latency instead of four
            const uint32_t b0 = p[pos], b1 = p[pos + 1], b2 = p[pos + 2], b3 = p[pos + 3];
            buf |= (uint64_t)(b0 | (b1 << 8) | (b2 << 16) | (b3 << 24)) << cnt;
            pos += 4;
            cnt += 32;
            if (cnt >= need) return;
        }
        while (cnt < need) {
            if (pos < n) buf |= (uint64_t)p[pos] << cnt;   // zeros beyond the end; overrun() tells
            pos++;
            cnt += 8;
        }
    }
    PNA_HD void align_byte() { pos = (consumed_bits() + 7) / 8; buf = 0; cnt = 0; }
    PNA_HD void skip_bytes(uint64_t k) { pos += k; }
#endif
    PNA_HD uint64_t consumed_bits() const { return pos * 8 - (uint64_t)cnt; }
    PNA_HD bool overrun() const { return consumed_bits() > n * 8; }   // consumed bits that do not exist
    PNA_HD uint32_t get(int k) {
        if (k == 0) return 0;
        fill(k);
        uint32_t v = (uint32_t)buf & ((1u << k) - 1u);
        buf >>= k; cnt -= k;
        return v;
    }
};

PNA_HD uint32_t rev_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// canonical table from code lengths; returns 0 ok, <0 over-subscribed, >0 incomplete (puff semantics)
template <int NSYM, int FB>
PNA_HD int build(Huff<NSYM, FB>* h, const uint8_t* length, int n) {
    uint16_t count[MAXBITS + 1];
    for (int l = 0; l <= MAXBITS; l++) { count[l] = 0; h->lim[l] = 0; h->base[l] = 0; }
    for (int s = 0; s < n; s++) count[length[s]]++;
    h->n01 = (uint16_t)(count[0] + count[1]);
    if (FB) for (int i = 0; i < (1 << FB); i++) h->fast[i] = 0;
    if (count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= MAXBITS; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return left;
    }
    uint16_t offs[MAXBITS + 1];
    offs[1] = 0;
    for (int l = 1; l < MAXBITS; l++) offs[l + 1] = offs[l] + count[l];
    {   // length bounds and symbol bases of the canonical code
        uint32_t first = 0;
        for (int l = 1; l <= MAXBITS; l++) {
            h->base[l] = (int16_t)((int)offs[l] - (int)first);
            first += count[l];
            h->lim[l] = (uint16_t)(first << (MAXBITS - l));   // <= 2^15 for a code that is not over-subscribed
            first <<= 1;
        }
    }
    for (int s = 0; s < n; s++)
        if (length[s]) h->symbol[offs[length[s]]++] = (uint16_t)s;
    // fast table: canonical codes, bit-reversed because DEFLATE packs Huffman codes MSB first
    if (FB) {
        uint32_t code = 0;
        int idx = 0;
        for (int l = 1; l <= FB; l++) {
            for (int k = 0; k < count[l]; k++, idx++, code++) {
                uint32_t r = rev_bits(code, l);
                uint16_t e = (uint16_t)((h->symbol[idx] << 4) | l);
                for (uint32_t f = r; f < (1u << FB); f += (1u << l)) h->fast[f] = e;
            }
            code <<= 1;
        }
    }
    return left;
}

template <int NSYM, int FB>
PNA_HD int decode_sym(Bits& b, const Huff<NSYM, FB>* h) {
    b.fill(MAXBITS);
    if (FB) {
        uint16_t e = h->fast[(uint32_t)b.buf & ((1u << FB) - 1u)];
        if (e) {
            int l = e & 15;
            b.buf >>= l; b.cnt -= l;
            return e >> 4;
        }
    }
    // the next 15 bits MSB first; the code length is 1 + the number of length bounds the window has reached
#if defined(__CUDA_ARCH__)
    const uint32_t w = __brev((uint32_t)b.buf) >> 17;
#else
    const uint32_t w = rev_bits((uint32_t)b.buf & 0x7FFFu, MAXBITS);
#endif
    int l = 1;
#pragma unroll
    for (int k = 1; k < MAXBITS; k++) l += (int)(w >= h->lim[k]);
    if (w >= h->lim[MAXBITS]) return -1;             // no code: incomplete set, or none at all
    const int idx = (int)h->base[l] + (int)(w >> (MAXBITS - l));
    b.buf >>= l; b.cnt -= l;
    return h->symbol[idx];
}

PNA_HD uint32_t len_base(int s) {   // s = symbol - 257
    return s < 8 ? (uint32_t)(3 + s) : s == 28 ? 258u : (uint32_t)(((4 + (s & 3)) << ((s >> 2) - 1)) + 3);
}
PNA_HD int len_extra(int s) { return s < 8 ? 0 : s == 28 ? 0 : (s >> 2) - 1; }
PNA_HD uint32_t dist_base(int s) { return s < 4 ? (uint32_t)(1 + s) : (uint32_t)(((2 + (s & 1)) << ((s >> 1) - 1)) + 1); }
PNA_HD int dist_extra(int s) { return s < 4 ? 0 : (s >> 1) - 1; }

// Where the decoded symbols go.  DirectEmit writes the output itself (byte stores, match copies through the output,
// Adler-32 on the way): one thread owns one stream from bits to bytes.  TokenEmit only records WHAT the stream says --
// literal bytes into a literal buffer and one 8-byte (offset, literal run | match length << 16) record per match, the
// same records the zstd sequence stage produces -- so that the bit-serial part of DEFLATE runs one stream per LANE
// without any output traffic, and the LZ stage (zstd_lz_kernel, one CTA per stream) executes the copies in parallel.
// Both keep counting when the output no longer fits (the symbol stream never depends on output bytes).
struct DirectEmit {
    uint8_t* out;
    uint64_t cap, op;
    uint32_t a1, a2, a_pending;   // Adler-32 running sums, bytes since the last modulo
    bool counting;
    PNA_HD void init(uint8_t* o, uint64_t c) { out = o; cap = c; op = 0; a1 = 1; a2 = 0; a_pending = 0; counting = false; }
    PNA_HD void lit(uint8_t c) {
        if (op < cap) {
            out[op] = c; a1 += c; a2 += a1;
            if (++a_pending == 5552) { a1 %= 65521u; a2 %= 65521u; a_pending = 0; }
        } else counting = true;
        op++;
    }
    PNA_HD void match(uint32_t len, uint32_t dist) {
        if (op + len <= cap) {
            uint32_t k = 0;
            if (dist >= 8) {   // eight source bytes are loaded before the first of them is stored: the loads overlap
                for (; k + 8 <= len; k += 8) {
                    uint8_t t[8];
                    for (int q = 0; q < 8; q++) t[q] = out[op - dist + q];
                    for (int q = 0; q < 8; q++) lit(t[q]);
                }
            }
            for (; k < len; k++) lit(out[op - dist]);
        } else {
            for (uint32_t k = 0; k < len; k++) {
                if (op < cap) lit(out[op - dist]);
                else { counting = true; op++; }
            }
        }
    }
    // Adler-32 trailer check; want = stored value
    PNA_HD bool trailer_ok(uint32_t want) { a1 %= 65521u; a2 %= 65521u; return want == ((a2 << 16) | a1); }
    PNA_HD void no_trailer() {}
};
struct TokenRec { uint32_t x, y; };   // layout of zs::SeqRec: x = offset, y = literal run | match length << 16
constexpr uint32_t TOKEN_RUN_MAX = 65534;   // longest literal run one record carries (0xFFFF is the zstd escape value)
struct TokenEmit {
    uint8_t* lits;
    TokenRec* recs;
    uint64_t cap, op;
    uint32_t nlit, nrec, run;
    uint32_t want;        // stored Adler-32 (checked by the Adler pass over the finished output)
    bool counting, has_trailer;
    PNA_HD void init(uint8_t* l, TokenRec* r, uint64_t c) {
        lits = l; recs = r; cap = c; op = 0; nlit = 0; nrec = 0; run = 0; want = 0; counting = false; has_trailer = false;
    }
    PNA_HD void lit(uint8_t c) {
        if (op < cap) {
            if (run == TOKEN_RUN_MAX) { recs[nrec].x = 1u; recs[nrec].y = run; nrec++; run = 0; }   // run continues in the next record
            lits[nlit++] = c; run++;
        } else counting = true;
        op++;
    }
    PNA_HD void match(uint32_t len, uint32_t dist) {
        if (op + len <= cap) { recs[nrec].x = dist; recs[nrec].y = run | (len << 16); nrec++; run = 0; }
        else counting = true;
        op += len;
    }
    PNA_HD bool trailer_ok(uint32_t w) { want = w; has_trailer = true; return true; }
    PNA_HD void no_trailer() { has_trailer = false; }
};

// Decode one zlib stream into emitter `E`.  `t` is per-thread scratch.  Returns status; E.op = bytes produced
// (exact decoded length even when it exceeds the capacity -> ST_NOSPACE).
template <class Emit>
PNA_HD int32_t inflate_zlib_to(Emit& E, const uint8_t* in, uint64_t n, Tables* t) {
    if (n == 0) { E.no_trailer(); return ST_OK; }   // nothing in, nothing out (zio::read: eof with no data -> Ok(0))
    if (n < 2) { E.no_trailer(); return ST_OK; }    // truncated header: short read, no error
    uint32_t cmf = in[0], flg = in[1];
    if (((cmf << 8) | flg) % 31 != 0 || (cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 0x20)) return ST_INVALID_INPUT;
    Bits b;
    b.init(in + 2, n - 2);
#define PNA_INF_TRUNC() do { E.no_trailer(); return E.counting ? ST_NOSPACE : ST_OK; } while (0)
    int last = 0;
    while (!last) {
        last = (int)b.get(1);
        int type = (int)b.get(2);
        if (b.overrun()) PNA_INF_TRUNC();
        if (type == 0) {
            b.align_byte();
            if (b.pos + 4 > b.n) PNA_INF_TRUNC();
            uint32_t len = load_le16(b.p + b.pos), nlen = load_le16(b.p + b.pos + 2);
            b.skip_bytes(4);
            if (len != (~nlen & 0xFFFFu)) return ST_INVALID_INPUT;
            uint64_t avail = b.n - b.pos;
            uint32_t take = len <= avail ? len : (uint32_t)avail;
            for (uint32_t i = 0; i < take; i++) E.lit(b.p[b.pos + i]);
            b.skip_bytes(take);
            if (take < len) PNA_INF_TRUNC();
            continue;
        }
        if (type == 3) return ST_INVALID_INPUT;
        if (type == 1) {
            uint8_t lengths[MAXL];
            int s = 0;
            for (; s < 144; s++) lengths[s] = 8;
            for (; s < 256; s++) lengths[s] = 9;
            for (; s < 280; s++) lengths[s] = 7;
            for (; s < 288; s++) lengths[s] = 8;
            build(&t->len, lengths, 288);
            for (s = 0; s < 30; s++) lengths[s] = 5;
            build(&t->dist, lengths, 30);
        } else {
            uint8_t lengths[MAXL + MAXD + 2];
            int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
            if (b.overrun()) PNA_INF_TRUNC();
            if (nlen > 286 || ndist > 30) return ST_INVALID_INPUT;
            const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            int i = 0;
            for (; i < ncode; i++) lengths[order[i]] = (uint8_t)b.get(3);
            for (; i < 19; i++) lengths[order[i]] = 0;
            if (b.overrun()) PNA_INF_TRUNC();
            if (build(&t->len, lengths, 19) != 0) return ST_INVALID_INPUT;  // code-length code must be complete
            i = 0;
            while (i < nlen + ndist) {
                int sym = decode_sym(b, &t->len);
                if (b.overrun()) PNA_INF_TRUNC();
                if (sym < 0) return ST_INVALID_INPUT;
                if (sym < 16) lengths[i++] = (uint8_t)sym;
                else {
                    int rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) return ST_INVALID_INPUT;
                        val = lengths[i - 1];
                        rep = 3 + (int)b.get(2);
                    } else if (sym == 17) rep = 3 + (int)b.get(3);
                    else rep = 11 + (int)b.get(7);
                    if (b.overrun()) PNA_INF_TRUNC();
                    if (i + rep > nlen + ndist) return ST_INVALID_INPUT;
                    while (rep--) lengths[i++] = (uint8_t)val;
                }
            }
            if (lengths[256] == 0) return ST_INVALID_INPUT;
            int err = build(&t->len, lengths, nlen);
            if (err && (err < 0 || nlen != t->len.n01)) return ST_INVALID_INPUT;
            err = build(&t->dist, lengths + nlen, ndist);
            if (err && (err < 0 || ndist != t->dist.n01)) return ST_INVALID_INPUT;
        }
        // symbols.  Single-exit loop without returns inside: with one stream per LANE the literal and the match arm must
        // reconverge every iteration (an early return inside an arm leaves the lanes split for the rest of the block).
        int rc = -1;   // -1 continue, 0 end of block, 1 truncated, 2 invalid
        for (;;) {
            int sym = decode_sym(b, &t->len);
            if (b.overrun()) rc = 1;
            else if (sym < 0) rc = 2;
            else if (sym < 256) E.lit((uint8_t)sym);
            else if (sym == 256) rc = 0;
            else {
                sym -= 257;
                if (sym >= 29) rc = 2;
                else {
                    const uint32_t len = len_base(sym) + b.get(len_extra(sym));
                    const int ds = decode_sym(b, &t->dist);
                    if (b.overrun()) rc = 1;
                    else if (ds < 0 || ds >= 30) rc = 2;
                    else {
                        const uint32_t dist = dist_base(ds) + b.get(dist_extra(ds));
                        if (b.overrun()) rc = 1;
                        else if (dist > E.op) rc = 2;
                        else E.match(len, dist);
                    }
                }
            }
            if (rc >= 0) break;
        }
        if (rc == 1) PNA_INF_TRUNC();
        if (rc == 2) return ST_INVALID_INPUT;
    }
    if (E.counting) { E.no_trailer(); return ST_NOSPACE; }
    // Adler-32 trailer (big endian) after discarding to the byte boundary
    uint64_t tpos = (b.consumed_bits() + 7) / 8;
    if (tpos + 4 > b.n) { E.no_trailer(); return ST_OK; }  // truncated trailer: short read, no error
    const uint8_t* tr = b.p + tpos;
    uint32_t want = ((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3];
    if (!E.trailer_ok(want)) return ST_INVALID_INPUT;
    return ST_OK;
#undef PNA_INF_TRUNC
}

// One stream from bits to bytes by one thread (see DirectEmit).
PNA_HD int32_t inflate_zlib(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len, Tables* t) {
    DirectEmit E;
    E.init(out, cap);
    const int32_t st = inflate_zlib_to(E, in, n, t);
    *out_len = (st == ST_OK || st == ST_NOSPACE) ? E.op : E.op;
    return st;
}

}  // namespace inf
}  // namespace pna
