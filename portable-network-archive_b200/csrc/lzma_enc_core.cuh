// lzma_enc_core.cuh -- the XZ arm of the create seam: Compress::XZ -> liblzma's XzEncoder around the entry's writer
// (lib/src/entry/write.rs:263, lib/src/compress.rs:29; level: lib/src/compress/xz.rs:6-47).  PNA_HD: the same code runs in
// xz_encode_kernel and, compiled with g++, in the CPU test tier.
//
// LZMA adapts one probability per coded bit, so a stream is serial -- but LZMA2 cuts a stream into chunks that may reset the
// coder state and the dictionary, and that is where the parallelism comes from: every 32 KiB segment of the matcher
// (lz_match_kernel: the sequences never reach outside their segment) becomes ONE chunk with control byte 0xE0 (dictionary
// reset, state reset, new properties) or, when the range coder does not shrink it, 0x01 (uncompressed, dictionary reset).
// Chunks are therefore independent in both directions: segments are encoded by different warps, and a reader may decode them
// in parallel.  Container (xz-file-format-1.1.0): stream header, ONE block (LZMA2 filter, 64 KiB dictionary, no size fields),
// its chunks, end marker, padding, CRC32 check, index with one record, stream footer.  liblzma's encoder would write CRC64
// as the check; every .xz reader accepts CRC32, and this path already owns a CRC-32 engine.
#pragma once
#include "lzma_core.cuh"
#include "encode_core.cuh"
#include "crc32_core.cuh"

namespace pna {
namespace xz {

// lp = 0, pb = 2 as in every liblzma preset.  lc (how many bits of the previous byte select a literal's probability set) is
// liblzma's 3 at most; the literal sets are 3/4 of the probability arena at lc = 3, and the arena is what limits the coders in
// flight per SM, so the setting trades bytes for speed: lc = 0 costs 0.2-1.8 % of the stream (bench corpus / text) and runs 43
// segments per SM, lc = 2 costs <= 0.5 % and runs 22, lc = 3 runs 13.  xz_lc_for_effort picks it from the `level`.
constexpr uint32_t ENC_LC_MAX = 3, ENC_LP = 0, ENC_PB = 2;
PNA_HD constexpr uint32_t enc_props(uint32_t lc) { return (ENC_PB * 5 + ENC_LP) * 9 + lc; }
PNA_HD constexpr uint32_t enc_probs(uint32_t lc) { return LZMA_PROBS_FIXED + (0x300u << (lc + ENC_LP)); }
constexpr uint32_t ENC_PROBS = enc_probs(ENC_LC_MAX);
constexpr uint8_t ENC_DICT_CODE = 8;                                       // 64 KiB: distances stay below the 32 KiB segment
constexpr uint32_t XZ_HDR_SCRATCH = 96;                                    // per entry: 24 bytes in front, <= 47 behind

struct RangeEnc {
    uint64_t low;
    uint32_t range, cache, cache_size;
    uint8_t* p;
    uint8_t* limit;   // a payload that grows past this is given up (the chunk goes out uncompressed)
    uint8_t* end;     // end of the buffer
    bool over;
    PNA_HD void init(uint8_t* dst, uint8_t* lim, uint8_t* e) { low = 0; range = 0xFFFFFFFFu; cache = 0; cache_size = 1; p = dst; limit = lim; end = e; over = false; }
    // Bytes leave unchecked: every call of shift_low emits at most one byte of its own plus the 0xFF bytes earlier calls held
    // back, so the stream never grows faster than one byte per coded bit -- room() is asked once per batch of events.
    PNA_HD bool room(uint32_t n_events) { if (p > limit || p + cache_size + n_events + 8 > end) over = true; return !over; }
    PNA_HD void shift_low() {
        const uint32_t lo32 = (uint32_t)low, carry = (uint32_t)(low >> 32);
        if (lo32 < 0xFF000000u || carry != 0) {
            *p++ = (uint8_t)(cache + carry);
            if (__builtin_expect(cache_size > 1, 0)) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
                for (uint32_t k = 1; k < cache_size; k++) *p++ = (uint8_t)(0xFFu + carry);
            }
            cache = lo32 >> 24;
            cache_size = 0;
        }
        cache_size++;
        low = (uint64_t)(lo32 << 8);
    }
    PNA_HD void flush() { if (room(0)) for (int i = 0; i < 5; i++) shift_low(); }
};

// ---- The coded bit stream as a list of EVENTS.  Which probability a bit is coded with, and the bit itself, depend on the data and
// the parse only -- never on the adaptive state of the coder.  So the work splits: any thread can turn a symbol into its events
// (prob index | bit << 15, or a direct bit), and the one serial thing left, the range coder, is a single small loop over them
// (one copy of the coder in the instruction stream instead of one per call site: the first version of this encoder spent 3/4 of
// its stall samples waiting for instruction fetch).
constexpr uint32_t EV_BIT = 0x8000u, EV_DIRECT = 0x4000u, EV_INDEX = 0x1FFFu;
static_assert(ENC_PROBS <= EV_INDEX + 1, "prob index fits the event word");
constexpr uint32_t EV_PER_LITERAL = 9, EV_PER_MATCH_MAX = 40;
// a match: is_match, is_rep, length (<= 10), distance slot (6), and for distances below 2^17 at most 11 direct + 4 aligned bits
static_assert(enc::SEG <= (1u << 17), "EV_PER_MATCH_MAX counts on distances inside one segment");

PNA_HD uint32_t lit_state_next(uint32_t state) { return state < 4 ? 0 : state < 10 ? state - 3 : state - 6; }
PNA_HD uint32_t lit_state_after(uint32_t state, uint32_t k) {   // three literals take every state to 0
    for (uint32_t i = 0; i < k && i < 3; i++) state = lit_state_next(state);
    return state;
}
// the literal `byte` at position pos (state: the coder state in front of it; match_byte: the byte at the last distance, used when
// state >= 7): EV_PER_LITERAL events
PNA_HD void gen_literal_events(uint16_t* ev, uint32_t pos, uint32_t byte, uint32_t prev, uint32_t state, uint32_t match_byte, uint32_t lc) {
    const uint32_t pb_mask = (1u << ENC_PB) - 1u, lp_mask = (1u << ENC_LP) - 1u;
    ev[0] = (uint16_t)(Probs::IS_MATCH + state * 16 + (pos & pb_mask));
    const uint32_t lp = Probs::LITERAL + 0x300u * (((pos & lp_mask) << lc) + (prev >> (8 - lc)));
    uint32_t sym = 1;
    if (state < 7) {
        for (int i = 7; i >= 0; i--) { const uint32_t b = (byte >> i) & 1u; ev[8 - i] = (uint16_t)((lp + sym) | (b << 15)); sym = (sym << 1) | b; }
    } else {
        uint32_t offs = 0x100;
        for (int i = 7; i >= 0; i--) {
            match_byte <<= 1;
            const uint32_t match_bit = match_byte & offs, b = (byte >> i) & 1u;
            ev[8 - i] = (uint16_t)((lp + offs + match_bit + sym) | (b << 15));
            sym = (sym << 1) | b;
            offs &= b ? match_bit : ~match_bit;
        }
    }
}
struct EvOut {
    uint16_t* ev;
    uint32_t n;
    PNA_HD void bit(uint32_t idx, uint32_t b) { ev[n++] = (uint16_t)(idx | (b << 15)); }
    PNA_HD void tree(uint32_t base, int nbits, uint32_t v) {
        uint32_t m = 1;
        for (int i = nbits - 1; i >= 0; i--) { const uint32_t b = (v >> i) & 1u; bit(base + m, b); m = (m << 1) | b; }
    }
    PNA_HD void tree_reverse(uint32_t base, int nbits, uint32_t v) {
        uint32_t m = 1;
        for (int i = 0; i < nbits; i++) { const uint32_t b = (v >> i) & 1u; bit(base + m, b); m = (m << 1) | b; }
    }
    PNA_HD void direct(uint32_t v, int nbits) { for (int i = nbits - 1; i >= 0; i--) ev[n++] = (uint16_t)(EV_DIRECT | (((v >> i) & 1u) << 15)); }
    PNA_HD void len(uint32_t lc, uint32_t pos_state, uint32_t length) {
        const uint32_t l = length - 2;
        if (l < 8) { bit(lc + 0, 0); tree(lc + 2 + pos_state * 8, 3, l); }
        else if (l < 16) { bit(lc + 0, 1); bit(lc + 1, 0); tree(lc + 2 + 16 * 8 + pos_state * 8, 3, l - 8); }
        else { bit(lc + 0, 1); bit(lc + 1, 1); tree(lc + 2 + 2 * 16 * 8, 8, l - 16); }
    }
};
// which of the last four distances a match repeats (0..3), or 4: a new distance.  Updates the history the way the decoder does.
PNA_HD uint32_t rep_classify(uint32_t dist, uint32_t& rep0, uint32_t& rep1, uint32_t& rep2, uint32_t& rep3) {
    if (dist == rep0) return 0;
    if (dist == rep1) { rep1 = rep0; rep0 = dist; return 1; }
    if (dist == rep2) { rep2 = rep1; rep1 = rep0; rep0 = dist; return 2; }
    const uint32_t k = dist == rep3 ? 3u : 4u;
    rep3 = rep2; rep2 = rep1; rep1 = rep0; rep0 = dist;
    return k;
}
// a match of ml bytes at distance dist + 1, coded as repeat `kind` (rep_classify) at position pos in state `state`: at most
// EV_PER_MATCH_MAX events.  Returns their number.
PNA_HD uint32_t gen_match_events(uint16_t* ev, uint32_t pos, uint32_t state, uint32_t kind, uint32_t ml, uint32_t dist) {
    EvOut o{ev, 0};
    const uint32_t pos_state = pos & ((1u << ENC_PB) - 1u);
    o.bit(Probs::IS_MATCH + state * 16 + pos_state, 1);
    if (kind < 4) {
        o.bit(Probs::IS_REP + state, 1);
        if (kind == 0) { o.bit(Probs::IS_REP_G0 + state, 0); o.bit(Probs::IS_REP0_LONG + state * 16 + pos_state, 1); }
        else {
            o.bit(Probs::IS_REP_G0 + state, 1);
            if (kind == 1) o.bit(Probs::IS_REP_G1 + state, 0);
            else { o.bit(Probs::IS_REP_G1 + state, 1); o.bit(Probs::IS_REP_G2 + state, kind == 2 ? 0u : 1u); }
        }
        o.len(Probs::LEN_REP, pos_state, ml);
    } else {
        o.bit(Probs::IS_REP + state, 0);
        o.len(Probs::LEN_MATCH, pos_state, ml);
        const uint32_t dist_state = ml < 6 ? ml - 2 : 3;
        uint32_t slot = dist;
        if (dist >= 4) {
            uint32_t n = 31;
            while (!(dist >> n)) n--;
            slot = 2 * n + ((dist >> (n - 1)) & 1u);
        }
        o.tree(Probs::POS_SLOT + dist_state * 64, 6, slot);
        if (slot >= 4) {
            const int nb = (int)(slot >> 1) - 1;
            const uint32_t base = (2u | (slot & 1u)) << nb, rem = dist - base;
            if (slot < 14) o.tree_reverse(Probs::POS_SPECIAL + base - slot - 1, nb, rem);
            else { o.direct(rem >> 4, nb - 4); o.tree_reverse(Probs::POS_ALIGN, 4, rem & 15u); }
        }
    }
    return o.n;
}
PNA_HD uint32_t match_state_next(uint32_t state, uint32_t kind) { return kind < 4 ? (state < 7 ? 8u : 11u) : (state < 7 ? 7u : 10u); }

// ---- From events to the coder.  The probability an event is coded with is the value its slot has after all EARLIER events on the
// same slot -- a function of the event order per slot, not of range / low.  So that too comes off the serial thread: prob_step
// is the adaptation rule, the kernel resolves 32 events per round trip (lanes that hit the same slot take turns, by rank), and
// what is left for the one serial lane is code_values: bound, low, range, normalise.  A value word is v | bit << 15, or
// EV_DIRECT | bit << 15.
PNA_HD uint32_t prob_step(uint32_t v, uint32_t b) { return b ? v - (v >> 5) : v + ((2048u - v) >> 5); }
PNA_HD void resolve_events_serial(uint16_t* probs, const uint16_t* ev, uint16_t* val, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t e = ev[i], b = e >> 15;
        if (e & EV_DIRECT) { val[i] = (uint16_t)(EV_DIRECT | (b << 15)); continue; }
        const uint32_t v = probs[e & EV_INDEX];
        probs[e & EV_INDEX] = (uint16_t)prob_step(v, b);
        val[i] = (uint16_t)(v | (b << 15));
    }
}
// Straight-line per value (selects instead of branches: a warp with one active lane pays an instruction-fetch bubble for every
// taken branch); a direct bit is a coded bit whose bound is range / 2.
PNA_HD void code_value(RangeEnc& rc, uint32_t w) {
    const uint32_t b = w >> 15;
    const bool direct = (w & EV_DIRECT) != 0;
    const uint32_t bound = direct ? rc.range >> 1 : (rc.range >> 11) * (w & 0xFFFu);
    rc.low += b ? bound : 0u;
    rc.range = (b && !direct) ? rc.range - bound : bound;
    if (__builtin_expect(rc.range < (1u << 24), 0)) { rc.range <<= 8; rc.shift_low(); }   // one value in eight
}
PNA_HD void code_values(RangeEnc& rc, const uint16_t* val, uint32_t n) {
    if (!rc.room(n)) return;
    uint32_t i = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; i + 4 <= n; i += 4) {   // the values are there before the coder needs them: four loads ahead of the dependent chain
        const uint32_t w0 = val[i], w1 = val[i + 1], w2 = val[i + 2], w3 = val[i + 3];
        code_value(rc, w0); code_value(rc, w1); code_value(rc, w2); code_value(rc, w3);
    }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; i < n; i++) code_value(rc, val[i]);
}
// one thread doing all of it (CPU test tier)
PNA_HD void code_events(RangeEnc& rc, uint16_t* probs, uint16_t* ev, uint32_t n) {
    resolve_events_serial(probs, ev, ev, n);
    code_values(rc, ev, n);
}

// One segment d[0, len) with its parse (sequences: ll literals, then a match of ml bytes at distance off) as ONE LZMA chunk
// payload into dst[0, cap) (the buffer itself has room for `phys` bytes): fresh state, fresh probabilities (the caller has set probs[0, enc_probs(lc)) to PROB_INIT), positions
// counted from 0 -- exactly what a reader sees after control byte 0xE0.  Returns the payload size, or 0xFFFFFFFF when it does
// not fit into cap.  Matches are coded as repeats when their distance is one of the last four (the parse does not look for
// them; structured data produces them by itself), else as a new distance.  This is the one-thread form (CPU test tier);
// xz_encode_kernel runs the same generators with a lane per literal and the same coder loop on lane 0.
PNA_HD uint32_t lzma_encode_segment(const uint8_t* d, uint32_t len, const enc::Seq* sq, uint32_t nseq, uint16_t* probs, uint8_t* dst, uint32_t cap, uint32_t phys, uint32_t lc) {
    RangeEnc rc;
    rc.init(dst, dst + cap, dst + phys);
    uint16_t ev[EV_PER_MATCH_MAX];
    uint32_t state = 0, rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0, pos = 0;
    auto literals = [&](uint32_t n) {
        for (uint32_t k = 0; k < n && !rc.over; k++, pos++) {
            gen_literal_events(ev, pos, d[pos], pos ? d[pos - 1] : 0u, state, state >= 7 ? d[pos - rep0 - 1] : 0u, lc);
            code_events(rc, probs, ev, EV_PER_LITERAL);
            state = lit_state_next(state);
        }
    };
    for (uint32_t s = 0; s < nseq && !rc.over; s++) {
        const uint32_t ll = sq[s].llml & 0xFFFFu, ml = sq[s].llml >> 16, dist = sq[s].off - 1u;
        literals(ll);
        const uint32_t kind = rep_classify(dist, rep0, rep1, rep2, rep3);
        code_events(rc, probs, ev, gen_match_events(ev, pos, state, kind, ml, dist));
        state = match_state_next(state, kind);
        pos += ml;
    }
    literals(len - pos);
    rc.flush();
    return rc.over || (uint32_t)(rc.p - dst) > cap ? 0xFFFFFFFFu : (uint32_t)(rc.p - dst);
}

// LZMA2 chunk header for a segment: compressed (6 bytes) or uncompressed (3 bytes).  Returns its length.
PNA_HD uint32_t lzma2_chunk_header(uint8_t* h, uint32_t usize, uint32_t csize /* 0: uncompressed chunk */, uint32_t lc) {
    const uint32_t u = usize - 1;
    if (!csize) { h[0] = 0x01; h[1] = (uint8_t)(u >> 8); h[2] = (uint8_t)u; return 3; }
    const uint32_t c = csize - 1;
    h[0] = (uint8_t)(0xE0u | (u >> 16)); h[1] = (uint8_t)(u >> 8); h[2] = (uint8_t)u;
    h[3] = (uint8_t)(c >> 8); h[4] = (uint8_t)c; h[5] = (uint8_t)enc_props(lc);
    return 6;
}

// x^(8n) mod P (reflected): appending n bytes to a message multiplies its CRC by this
PNA_HD uint32_t crc_xpow_bytes(uint64_t n) {
    uint32_t p = 0x80000000u, sq = 0x00800000u;   // x^0, x^8
    while (n) {
        if (n & 1) p = crc_multmodp(sq, p);
        sq = crc_multmodp(sq, sq);
        n >>= 1;
    }
    return p;
}
// crc(A || B) from crc(A), crc(B), |B| (finalised values; zlib's crc32_combine identity)
PNA_HD uint32_t crc_concat(uint32_t crc_a, uint32_t crc_b, uint32_t xpow_b) { return crc_multmodp(xpow_b, crc_a) ^ crc_b; }
// plain CRC-32 of a slice, four bytes per step, no tables (a lane's 1 KiB share of a segment)
PNA_HD uint32_t crc_slice(const uint8_t* p, uint32_t n) {
    uint32_t c = 0xFFFFFFFFu, i = 0;
    for (; i + 4 <= n; i += 4) {
        c ^= (uint32_t)p[i] | (uint32_t)p[i + 1] << 8 | (uint32_t)p[i + 2] << 16 | (uint32_t)p[i + 3] << 24;
        for (int k = 0; k < 32; k++) c = (c >> 1) ^ (CRC_POLY & (0u - (c & 1u)));
    }
    for (; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CRC_POLY & (0u - (c & 1u)));
    }
    return ~c;
}

PNA_HD uint32_t xz_put_vli(uint8_t* p, uint64_t v) {
    uint32_t k = 0;
    while (v >= 0x80) { p[k++] = (uint8_t)(v | 0x80); v >>= 7; }
    p[k++] = (uint8_t)v;
    return k;
}
PNA_HD void xz_put_le32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// Everything of the container in front of the chunks: stream header (12 bytes) and, unless the stream is empty, the block header
// (12 bytes).  Returns the length.
PNA_HD uint32_t xz_write_front(uint8_t* h, bool empty) {
    h[0] = 0xFD; h[1] = '7'; h[2] = 'z'; h[3] = 'X'; h[4] = 'Z'; h[5] = 0;
    h[6] = 0x00; h[7] = 0x01;                                  // stream flags: check = CRC32
    xz_put_le32(h + 8, xz_crc32(h + 6, 2));
    if (empty) return 12;
    h[12] = 0x02;                                              // header size (12 bytes)
    h[13] = 0x00;                                              // one filter, no size fields
    h[14] = 0x21; h[15] = 0x01; h[16] = ENC_DICT_CODE;         // LZMA2, one property byte
    h[17] = h[18] = h[19] = 0;
    xz_put_le32(h + 20, xz_crc32(h + 12, 8));
    return 24;
}
// Everything behind the chunks: end marker, block padding, check, index, footer (no block at all for an empty stream, as
// liblzma writes it).  chunk_bytes: all chunk headers + payloads.  Returns the length (<= 47).
PNA_HD uint32_t xz_write_back(uint8_t* t, bool empty, uint64_t chunk_bytes, uint64_t plain_len, uint32_t crc) {
    uint32_t k = 0;
    if (!empty) {
        t[k++] = 0x00;                                         // LZMA2 end marker
        const uint64_t csz = chunk_bytes + 1;
        for (uint64_t q = csz; q & 3; q++) t[k++] = 0;
        xz_put_le32(t + k, crc); k += 4;
        const uint32_t ix = k;
        t[k++] = 0x00; t[k++] = 0x01;
        k += xz_put_vli(t + k, 12 + csz + 4);                  // unpadded size: header + data + check
        k += xz_put_vli(t + k, plain_len);
        while ((k - ix) & 3) t[k++] = 0;
        xz_put_le32(t + k, xz_crc32(t + ix, k - ix)); k += 4;
        uint8_t* f = t + k;
        xz_put_le32(f + 4, (k - ix) / 4 - 1);
        f[8] = 0x00; f[9] = 0x01; f[10] = 'Y'; f[11] = 'Z';
        xz_put_le32(f, xz_crc32(f + 4, 6));
        return k + 12;
    }
    t[0] = 0x00; t[1] = 0x00; t[2] = 0; t[3] = 0;
    xz_put_le32(t + 4, xz_crc32(t, 4));
    uint8_t* f = t + 8;
    xz_put_le32(f + 4, 8 / 4 - 1);
    f[8] = 0x00; f[9] = 0x01; f[10] = 'Y'; f[11] = 'Z';
    xz_put_le32(f, xz_crc32(f + 4, 6));
    return 20;
}

}  // namespace xz
}  // namespace pna
