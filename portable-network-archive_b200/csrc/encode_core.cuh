// encode_core.cuh -- bitstream writers of the encode seam (PNA_HD: shared by the sm_100a kernels and the
// g++ host test build): zstd compressed blocks (RFC 8878) and zlib/DEFLATE fixed-Huffman blocks (RFC 1950/1951)
// from the (literal length, match length, offset) sequences the LZ77 matcher (kernels_encode.cuh) produces.
//
// Replaces `compression_writer` /root/reference/lib/src/entry/write.rs:251-265 (zstd::Encoder level 3,
// flate2::ZlibEncoder level 6).  Encoded BYTES are not unique and are not pinned by any reference test
// (SURVEY 8c): the contract is that the reference's decoders (libzstd / zlib inflate) reproduce the plaintext,
// and the size ratio is reported against the reference's encoder at the same level.
//
// Stream shapes written here:
//   zstd   one frame per entry: magic, FHD 0x00 (no content size / checksum / dictionary -- like the
//          reference's streaming encoder), Window_Descriptor 0x38 (128 KiB), one block per 32 KiB segment.
//          Compressed block = Huffman-compressed (or raw) literals + sequences coded with the PREDEFINED FSE tables
//          (mode byte 0x00),
//          repeat offsets are used for the history slots the block has set itself (zstd_assign_repcodes).  A block that does not shrink becomes a Raw_Block.
//   zlib   0x78 0x9C, per segment one fixed-Huffman block followed by an empty stored block (the Z_SYNC_FLUSH
//          marker 00 00 FF FF) so that segments stay byte aligned and can be produced independently; the last
//          segment's block carries BFINAL; Adler-32 big endian.  A segment that does not shrink is a stored block.
#pragma once
#include "common.cuh"
#include "zstd_core.cuh"

namespace pna {
namespace enc {

constexpr uint32_t SEG = 32 * 1024;        // bytes per independently matched segment (= zstd block, deflate block)
constexpr uint32_t MIN_MATCH = 4, MAX_MATCH = 258;
constexpr uint32_t SEG_SEQ_MAX = SEG / MIN_MATCH;   // sequences a segment can hold

struct Seq { uint32_t off; uint32_t llml; };   // ll | ml << 16 (ll <= 32768, ml <= 258)

// FSE compression tables of the three predefined distributions (libzstd FSE_buildCTable semantics)
struct FseCTab {
    uint16_t state[64];     // tableU16
    struct Sym { uint32_t dnb; int32_t dfs; };   // deltaNbBits, deltaFindState: one 8-byte load per symbol (the per-block tables sit in
    Sym sym[56];                                   // local memory, where every scattered load is a sector)
    uint32_t log;
};
struct EncTables {
    FseCTab ll, of, ml;
    uint8_t ll_code[64];    // ll < 64 -> code
    uint8_t ml_code[128];   // (ml - 3) < 128 -> code
};

// FSE compression table from normalised counts (sum = 1 << log, -1 = "less than one"), libzstd FSE_buildCTable semantics.
// log <= 6, n <= 56.
PNA_HD void fse_build_ctab_norm(FseCTab* T, const int16_t* norm, int n, int log) {
    const int size = 1 << log;
    uint32_t cumul[57];
    uint8_t sym_of[64];
    int high = size - 1;
    cumul[0] = 0;
    for (int u = 1; u <= n; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; sym_of[high--] = (uint8_t)(u - 1); }
        else cumul[u] = cumul[u - 1] + (uint32_t)norm[u - 1];
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < norm[s]; i++) {
            sym_of[pos] = (uint8_t)s;
            pos = (pos + step) & mask;
            while (pos > high) pos = (pos + step) & mask;
        }
    for (int u = 0; u < size; u++) { int s = sym_of[u]; T->state[cumul[s]++] = (uint16_t)(size + u); }
    int total = 0;
    for (int s = 0; s < n; s++) {
        if (norm[s] == 0) { T->sym[s].dnb = ((uint32_t)(log + 1) << 16) - (1u << log); T->sym[s].dfs = 0; }
        else if (norm[s] == -1 || norm[s] == 1) { T->sym[s].dnb = ((uint32_t)log << 16) - (1u << log); T->sym[s].dfs = total - 1; total++; }
        else {
            const uint32_t max_bits = (uint32_t)log - (uint32_t)highbit32((uint32_t)norm[s] - 1);
            const uint32_t min_state_plus = (uint32_t)norm[s] << max_bits;
            T->sym[s].dnb = (max_bits << 16) - min_state_plus;
            T->sym[s].dfs = total - norm[s];
            total += norm[s];
        }
    }
    T->log = (uint32_t)log;
}
inline void fse_build_ctab(FseCTab* T, int kind) {
    const int log = kind == 1 ? 5 : 6, n = kind == 0 ? 36 : kind == 1 ? 29 : 53;
    int16_t norm[56];
    for (int s = 0; s < n; s++) norm[s] = (int16_t)zs::predef_norm(kind, s);
    fse_build_ctab_norm(T, norm, n, log);
}
// the table of an RLE-mode symbol: zero bits per symbol, state 0 (libzstd FSE_buildCTable_rle)
PNA_HD void fse_build_ctab_rle(FseCTab* T, uint32_t sym) {
    T->state[0] = 0; T->state[1] = 0;
    T->sym[sym].dnb = 0; T->sym[sym].dfs = 0;
    T->log = 0;
}

// ---- per-block ("FSE_Compressed_Mode") tables.  Blocks are written by one thread each, so everything here is small and
// serial: a histogram of the block's codes, counts normalised to 2^6 (every present symbol gets at least 1; the rounding
// remainder goes to / comes from the most frequent symbols), the table description (RFC 8878 4.1.1), and an integer cost
// model (1/256 bit units) that decides between Predefined, RLE and FSE_Compressed per table.
constexpr int FSE_DYN_LOG = 6;
// 256 * log2(n) for 1 <= n <= 64, piecewise linear in the mantissa (never above the true value by more than it is below: the
// error, <= 0.09 bit, only tilts close decisions)
PNA_HD uint32_t log2_256(uint32_t n) {
    const uint32_t hb = (uint32_t)highbit32(n);
    return (hb << 8) + (((n << 8) >> hb) - 256u);
}
// bits * 256 to code `count` symbols each of probability norm / 2^log
PNA_HD uint32_t fse_cost256(uint32_t count, int norm, int log) { return count * (((uint32_t)log << 8) - log2_256((uint32_t)(norm < 1 ? 1 : norm))); }

// counts[0..n) (sum = total >= 1, at most 64 non-zero) -> norm[0..n) with sum 2^FSE_DYN_LOG; returns the number of present symbols
PNA_HD int fse_normalize64(const uint16_t* counts, int n, uint32_t total, int16_t* norm) {
    const int size = 1 << FSE_DYN_LOG;
    int sum = 0, present = 0;
    for (int s = 0; s < n; s++) {
        int v = 0;
        if (counts[s]) {
            v = (int)(((uint64_t)counts[s] * (uint32_t)size + total / 2) / total);
            if (v < 1) v = 1;
            present++;
        }
        norm[s] = (int16_t)v;
        sum += v;
    }
    while (sum != size) {   // a few rounds: rounding leaves at most one unit per present symbol
        int big = 0;
        for (int s = 1; s < n; s++) if (norm[s] > norm[big]) big = s;
        if (sum < size) { norm[big] = (int16_t)(norm[big] + (size - sum)); break; }
        const int excess = sum - size, can = norm[big] - 1;
        if (can <= 0) break;   // cannot happen: at most 53 present symbols of weight 1 in a table of 64
        int take = can / 2 > 1 ? can / 2 : 1;   // a large excess is spread over several symbols
        if (take > excess) take = excess;
        norm[big] = (int16_t)(norm[big] - take);
        sum -= take;
    }
    return present;
}
// FSE table description; returns the byte count (< 56: 53 symbols of at most 7 bits, plus zero-run flags)
PNA_HD uint32_t fse_write_ncount(const int16_t* norm, int n, uint8_t* dst) {
    int last = n - 1;
    while (last > 0 && norm[last] == 0) last--;
    const int alphabet = last + 1, log = FSE_DYN_LOG;
    uint64_t acc = (uint64_t)(log - 5);
    uint32_t nb = 4, bytes = 0;
    int remaining = (1 << log) + 1, threshold = 1 << log, bits = log + 1, sym = 0;
    bool prev0 = false;
    auto flush = [&]() { while (nb >= 8) { dst[bytes++] = (uint8_t)acc; acc >>= 8; nb -= 8; } };
    while (sym < alphabet && remaining > 1) {
        if (prev0) {
            int start = sym;
            while (sym < alphabet && norm[sym] == 0) sym++;
            while (sym >= start + 3) { start += 3; acc |= (uint64_t)3 << nb; nb += 2; flush(); }
            acc |= (uint64_t)(sym - start) << nb; nb += 2;
            flush();
        }
        int count = norm[sym++];
        const int mx = (2 * threshold - 1) - remaining;
        remaining -= count < 0 ? -count : count;
        count++;
        if (count >= threshold) count += mx;
        acc |= (uint64_t)(uint32_t)count << nb;
        nb += (uint32_t)bits;
        if (count < mx) nb--;
        prev0 = count == 1;
        while (remaining < threshold) { bits--; threshold >>= 1; }
        flush();
    }
    if (nb) { dst[bytes++] = (uint8_t)acc; }
    return bytes;
}
inline void make_enc_tables(EncTables* E) {
    fse_build_ctab(&E->ll, 0); fse_build_ctab(&E->of, 1); fse_build_ctab(&E->ml, 2);
    for (uint32_t v = 0; v < 64; v++) {
        int c = 0;
        while (c + 1 < 36 && zs::ll_base(c + 1) <= v) c++;
        E->ll_code[v] = (uint8_t)c;
    }
    for (uint32_t v = 0; v < 128; v++) {   // v = ml - 3
        int c = 0;
        while (c + 1 < 53 && zs::ml_base(c + 1) <= v + 3) c++;
        E->ml_code[v] = (uint8_t)c;
    }
}

// forward little-endian bit writer into 4-byte aligned memory
struct BitOut {
    uint8_t* p;
    uint64_t acc;
    uint32_t n;        // bits in acc
    uint32_t bytes;    // bytes written
    PNA_HD void init(uint8_t* dst) { p = dst; acc = 0; n = 0; bytes = 0; }
    PNA_HD void add(uint32_t v, uint32_t nb) {   // nb <= 32; v may carry garbage above nb bits
        if (nb == 0) return;
        const uint64_t m = nb >= 32 ? 0xFFFFFFFFull : ((1ull << nb) - 1ull);
        acc |= ((uint64_t)v & m) << n;
        n += nb;
        if (n >= 32) {
            *reinterpret_cast<uint32_t*>(p + bytes) = (uint32_t)acc;
            bytes += 4;
            acc >>= 32;
            n -= 32;
        }
    }
    PNA_HD uint32_t finish() {   // flush whole bytes (zero padded); returns total bytes
        while (n > 0) { p[bytes++] = (uint8_t)acc; acc >>= 8; n = n > 8 ? n - 8 : 0; }
        return bytes;
    }
};

struct FseCState { uint32_t v; };
PNA_HD void fse_init_state(FseCState& s, const FseCTab& t, uint32_t sym) {
    const FseCTab::Sym e = t.sym[sym];
    const uint32_t nb = (e.dnb + (1u << 15)) >> 16;
    const uint32_t value = (nb << 16) - e.dnb;
    s.v = t.state[(value >> nb) + (uint32_t)e.dfs];
}
PNA_HD void fse_encode(BitOut& b, FseCState& s, const FseCTab& t, uint32_t sym) {
    const FseCTab::Sym e = t.sym[sym];
    const uint32_t nb = (s.v + e.dnb) >> 16;
    b.add(s.v, nb);
    s.v = t.state[(int32_t)(s.v >> nb) + e.dfs];
}

PNA_HD uint32_t ll_code_of(const EncTables& E, uint32_t ll) { return ll < 64 ? E.ll_code[ll] : (uint32_t)highbit32(ll) + 19u; }
PNA_HD uint32_t ml_code_of(const EncTables& E, uint32_t mlb) { return mlb < 128 ? E.ml_code[mlb] : (uint32_t)highbit32(mlb) + 36u; }

// Repeat offsets (RFC 8878 3.1.1.5), forward pass over a block's sequences: q.off (a real distance) becomes the
// Offset_Value to code -- SEQ_REPCODE | 1..3 for a repeat offset, left as the distance (coded as distance + 3) otherwise.  Blocks are written independently and in
// parallel, so the history a block inherits from its predecessor is unknown here: a history slot is only used once this
// block has set it itself (`known`), which is what makes the choice valid whatever the decoder's incoming history is.
constexpr uint32_t SEQ_REPCODE = 0x80000000u;   // Seq.off flag: the low two bits are a repeat-offset Offset_Value
PNA_HD void zstd_assign_repcodes(Seq* seqs, uint32_t nseq) {
    uint32_t r1 = 0, r2 = 0, r3 = 0;
    bool k1 = false, k2 = false, k3 = false;
    for (uint32_t i = 0; i < nseq; i++) {
        const uint32_t off = seqs[i].off, ll = seqs[i].llml & 0xFFFFu;
        uint32_t val;
        if (ll != 0 && k1 && off == r1) val = 1;                                    // history unchanged
        else if ((ll != 0 && k2 && off == r2) || (ll == 0 && k2 && off == r2)) {   // second slot: swap the first two
            val = ll != 0 ? 2u : 1u;
            const uint32_t t = r1; r1 = r2; r2 = t;
            const bool kt = k1; k1 = k2; k2 = kt;
        } else if (k3 && off == r3) {                                               // third slot: rotate
            val = ll != 0 ? 3u : 2u;
            const uint32_t t = r3; r3 = r2; r2 = r1; r1 = t;
            const bool kt = k3; k3 = k2; k2 = k1; k1 = kt;
        } else if (ll == 0 && k1 && r1 > 1 && off == r1 - 1) {                      // "first slot minus one"
            val = 3;
            r3 = r2; r2 = r1; r1 = off;
            k3 = k2; k2 = k1; k1 = true;
        } else {                                                                    // new offset: push
            val = off + 3;
            r3 = r2; r2 = r1; r1 = off;
            k3 = k2; k2 = k1; k1 = true;
        }
        if (val <= 3) seqs[i].off = SEQ_REPCODE | val;   // only repeat offsets are written back (one store per hit, none otherwise)
    }
}

// Sequences section of one block (nseq >= 1): Number_of_Sequences, Symbol_Compression_Modes, table descriptions, FSE bitstream.
// seqs[i].off is a distance, or SEQ_REPCODE | Offset_Value where zstd_assign_repcodes chose a repeat offset.
// dyn: choose per table between Predefined, RLE and FSE_Compressed (a table fitted to this block) by estimated cost; else all
// three Predefined (the fast setting).
// dst must be 4-byte aligned, cap bytes.  Returns (offset of first byte << 24) | length, or 0xFFFFFFFF when the
// section would not fit cap (the caller then emits the segment as a Raw_Block).
constexpr uint32_t SEQ_HDR_ROOM = 208;   // Number_of_Sequences (3) + modes (1) + three table descriptions (< 64 each), multiple of 4
PNA_HD uint32_t zstd_write_sequences(const EncTables& E, const Seq* seqs, uint32_t nseq, uint8_t* dst, uint32_t cap, bool dyn = false) {
    if (cap < SEQ_HDR_ROOM + 64) return 0xFFFFFFFFu;
    // ---- tables
    FseCTab dtab[3];
    const FseCTab* tab[3] = {&E.ll, &E.of, &E.ml};
    uint8_t desc[3][64];
    uint32_t desc_n[3] = {0, 0, 0}, mode[3] = {0, 0, 0};
    if (dyn && nseq >= 48) {
        uint16_t cnt[3][56];
        for (int t = 0; t < 3; t++) for (int k = 0; k < 56; k++) cnt[t][k] = 0;
        for (uint32_t i = 0; i < nseq; i++) {
            const Seq q = seqs[i];
            const uint32_t ll = q.llml & 0xFFFFu, mlb = (q.llml >> 16) - 3u, ob = (q.off & SEQ_REPCODE) ? (q.off & 3u) : q.off + 3u;
            cnt[0][ll_code_of(E, ll)]++; cnt[1][(uint32_t)highbit32(ob)]++; cnt[2][ml_code_of(E, mlb)]++;
        }
        for (int t = 0; t < 3; t++) {
            const int n = t == 0 ? 36 : t == 1 ? 29 : 53, plog = t == 1 ? 5 : 6;
            int16_t norm[56];
            int only = -1;
            uint32_t pre = 0;
            bool pre_ok = true;
            for (int k = 0; k < n; k++) {
                if (!cnt[t][k]) continue;
                if (cnt[t][k] == nseq) only = k;
                pre += fse_cost256(cnt[t][k], zs::predef_norm(t, k), plog);
            }
            for (int k = n; k < 56; k++) if (cnt[t][k]) pre_ok = false;   // (offset codes above 28 cannot occur with 32 KiB segments)
            if (only >= 0) { mode[t] = 1; desc[t][0] = (uint8_t)only; desc_n[t] = 1; fse_build_ctab_rle(&dtab[t], (uint32_t)only); tab[t] = &dtab[t]; continue; }
            fse_normalize64(cnt[t], n, nseq, norm);
            uint32_t fit = 0;
            for (int k = 0; k < n; k++) if (cnt[t][k]) fit += fse_cost256(cnt[t][k], norm[k], FSE_DYN_LOG);
            const uint32_t dn = fse_write_ncount(norm, n, desc[t]);
            if (!pre_ok || fit + (dn << 11) < pre) {   // description bytes * 8 bits * 256
                mode[t] = 2; desc_n[t] = dn;
                fse_build_ctab_norm(&dtab[t], norm, n, FSE_DYN_LOG);
                tab[t] = &dtab[t];
            }
        }
    }
    const FseCTab& TL = *tab[0];
    const FseCTab& TO = *tab[1];
    const FseCTab& TM = *tab[2];
    // header is written after the bitstream at its front; bitstream starts 4-byte aligned at dst + SEQ_HDR_ROOM
    BitOut b;
    b.init(dst + SEQ_HDR_ROOM);
    const uint32_t room = cap - SEQ_HDR_ROOM;
    FseCState sll, sof, sml;
    {
        const Seq q = seqs[nseq - 1];
        const uint32_t ll = q.llml & 0xFFFFu, mlb = (q.llml >> 16) - 3u, ob = (q.off & SEQ_REPCODE) ? (q.off & 3u) : q.off + 3u;
        const uint32_t cl = ll_code_of(E, ll), cm = ml_code_of(E, mlb), co = (uint32_t)highbit32(ob);
        fse_init_state(sml, TM, cm);
        fse_init_state(sof, TO, co);
        fse_init_state(sll, TL, cl);
        b.add(ll, (uint32_t)zs::ll_bits((int)cl));
        b.add(mlb, (uint32_t)zs::ml_bits((int)cm));
        b.add(ob, co);
    }
    for (uint32_t i = nseq - 1; i-- > 0;) {
        const Seq q = seqs[i];
        const uint32_t ll = q.llml & 0xFFFFu, mlb = (q.llml >> 16) - 3u, ob = (q.off & SEQ_REPCODE) ? (q.off & 3u) : q.off + 3u;
        const uint32_t cl = ll_code_of(E, ll), cm = ml_code_of(E, mlb), co = (uint32_t)highbit32(ob);
        fse_encode(b, sof, TO, co);
        fse_encode(b, sml, TM, cm);
        fse_encode(b, sll, TL, cl);
        b.add(ll, (uint32_t)zs::ll_bits((int)cl));
        b.add(mlb, (uint32_t)zs::ml_bits((int)cm));
        b.add(ob, co);
        if (b.bytes + 32 > room) return 0xFFFFFFFFu;
    }
    if (b.bytes + 32 > room) return 0xFFFFFFFFu;
    b.add(sml.v, TM.log);
    b.add(sof.v, TO.log);
    b.add(sll.v, TL.log);
    b.add(1, 1);
    const uint32_t bs = b.finish();
    // Number_of_Sequences (1-3 bytes) + modes byte + table descriptions (LL, OF, ML), right-aligned against the bitstream
    uint8_t h[4];
    uint32_t hn;
    if (nseq < 128) { h[0] = (uint8_t)nseq; hn = 1; }
    else if (nseq < 0x7F00) { h[0] = (uint8_t)((nseq >> 8) + 0x80); h[1] = (uint8_t)nseq; hn = 2; }
    else { h[0] = 0xFF; h[1] = (uint8_t)(nseq - 0x7F00); h[2] = (uint8_t)((nseq - 0x7F00) >> 8); hn = 3; }
    h[hn++] = (uint8_t)((mode[0] << 6) | (mode[1] << 4) | (mode[2] << 2));   // Symbol_Compression_Modes
    const uint32_t total_h = hn + desc_n[0] + desc_n[1] + desc_n[2];
    uint8_t* w = dst + SEQ_HDR_ROOM - total_h;
    for (uint32_t k = 0; k < hn; k++) *w++ = h[k];
    for (int t = 0; t < 3; t++) for (uint32_t k = 0; k < desc_n[t]; k++) *w++ = desc[t][k];
    return (SEQ_HDR_ROOM - total_h) << 24 | (total_h + bs);   // high byte: offset of the first valid byte inside dst
}

// Raw_Literals section header for `n` literal bytes; returns header length (1..3)
PNA_HD uint32_t zstd_raw_lit_header(uint32_t n, uint8_t* h) {
    if (n < 32) { h[0] = (uint8_t)(n << 3); return 1; }
    if (n < 4096) { h[0] = (uint8_t)((n << 4) | 4u); h[1] = (uint8_t)(n >> 4); return 2; }
    h[0] = (uint8_t)((n << 4) | 12u); h[1] = (uint8_t)(n >> 4); h[2] = (uint8_t)(n >> 12);
    return 3;
}
// ------------------------------------------------------------------------------------------------ Huffman literals
// Compressed_Literals_Block with four streams (RFC 8878 3.1.1.3.1.1-6, libzstd HUF_compress4X semantics):
//   tree description in the DIRECT form (header byte 127 + L, then L four-bit weights for symbols 0..L-1; the weight of
//   the last present symbol L is implied), which covers alphabets whose largest byte value is <= 128 -- text; anything
//   else stays Raw_Literals (the FSE-compressed weight form is the next step) -- code lengths limited to 11 bits,
//   canonical codes assigned like HUF_buildCTable (longest codes first, symbol order inside a length), every stream
//   written from its LAST symbol to its first with a closing 1 bit, a 6-byte jump table in front.
constexpr uint32_t HUF_ENC_SYMS = 129, HUF_ENC_MAXBITS = 11, HUF_ENC_MIN = 256;
struct ByteBits {   // forward little-endian bit writer into unaligned memory
    uint8_t* p;
    uint64_t acc;
    uint32_t n, bytes;
    PNA_HD void init(uint8_t* dst) { p = dst; acc = 0; n = 0; bytes = 0; }
    PNA_HD void add(uint32_t v, uint32_t nb) {   // nb <= 16
        acc |= (uint64_t)v << n;
        n += nb;
        if (n >= 32) {
            p[bytes] = (uint8_t)acc; p[bytes + 1] = (uint8_t)(acc >> 8); p[bytes + 2] = (uint8_t)(acc >> 16); p[bytes + 3] = (uint8_t)(acc >> 24);
            bytes += 4; acc >>= 32; n -= 32;
        }
    }
    PNA_HD uint32_t finish() { while (n > 0) { p[bytes++] = (uint8_t)acc; acc >>= 8; n = n > 8 ? n - 8 : 0; } return bytes; }
};
// Literals section of a block.  hdr receives the section header (<= 5 bytes, *hdr_len).  Returns the length of the
// compressed payload written to dst (tree + jump table + streams), or 0 when the literals stay raw (the payload is then
// the n literal bytes themselves).  dst needs n bytes of room.
PNA_HD uint32_t zstd_write_literals(const uint8_t* lits, uint32_t n, uint8_t* dst, uint8_t* hdr, uint32_t* hdr_len) {
    *hdr_len = zstd_raw_lit_header(n, hdr);
    if (n < HUF_ENC_MIN) return 0;
    uint32_t count[HUF_ENC_SYMS];
    for (uint32_t s = 0; s < HUF_ENC_SYMS; s++) count[s] = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t c = lits[i];
        if (c >= HUF_ENC_SYMS) return 0;
        count[c]++;
    }
    // present symbols sorted by ascending count (insertion sort: <= 129 keys)
    uint8_t order[HUF_ENC_SYMS];
    uint32_t m = 0, last_sym = 0;
    for (uint32_t s = 0; s < HUF_ENC_SYMS; s++) {
        if (!count[s]) continue;
        last_sym = s;
        uint32_t k = m++;
        while (k > 0 && count[order[k - 1]] > count[s]) { order[k] = order[k - 1]; k--; }
        order[k] = (uint8_t)s;
    }
    if (m < 2) return 0;
    // Huffman tree by the two-queue merge: leaves 0..m-1 (ascending), internal nodes m..2m-2 are born in ascending order
    uint32_t wt[2 * HUF_ENC_SYMS];
    uint16_t parent[2 * HUF_ENC_SYMS];
    for (uint32_t k = 0; k < m; k++) wt[k] = count[order[k]];
    uint32_t li = 0, ni = m, no = m;
    for (uint32_t k = 0; k + 1 < m; k++) {
        uint32_t pick[2];
        for (int q = 0; q < 2; q++) {
            if (li < m && (ni >= no || wt[li] <= wt[ni])) pick[q] = li++;
            else pick[q] = ni++;
        }
        wt[no] = wt[pick[0]] + wt[pick[1]];
        parent[pick[0]] = (uint16_t)no; parent[pick[1]] = (uint16_t)no;
        no++;
    }
    uint8_t depth[2 * HUF_ENC_SYMS];
    depth[no - 1] = 0;
    for (uint32_t k = no - 1; k-- > 0;) depth[k] = (uint8_t)(depth[parent[k]] + 1);
    // length limit (miniz's tdefl_huffman_enforce_max_code_size): fold the too-long codes into the limit, then pay the
    // Kraft excess back by splitting the longest code that is still shorter than the limit
    uint32_t num[33];
    for (uint32_t l = 0; l < 33; l++) num[l] = 0;
    for (uint32_t k = 0; k < m; k++) num[depth[k] > 32 ? 32 : depth[k]]++;
    for (uint32_t l = HUF_ENC_MAXBITS + 1; l < 33; l++) { num[HUF_ENC_MAXBITS] += num[l]; num[l] = 0; }
    uint32_t total = 0;
    for (uint32_t l = HUF_ENC_MAXBITS; l > 0; l--) total += num[l] << (HUF_ENC_MAXBITS - l);
    while (total != (1u << HUF_ENC_MAXBITS)) {
        num[HUF_ENC_MAXBITS]--;
        for (uint32_t l = HUF_ENC_MAXBITS - 1; l > 0; l--)
            if (num[l]) { num[l]--; num[l + 1] += 2; break; }
        total--;
    }
    // lengths: the most frequent symbols (end of `order`) take the shortest codes
    uint8_t len_of[HUF_ENC_SYMS];
    for (uint32_t s = 0; s < HUF_ENC_SYMS; s++) len_of[s] = 0;
    uint32_t max_bits = 0;
    {
        uint32_t k = m;
        for (uint32_t l = 1; l <= HUF_ENC_MAXBITS; l++)
            for (uint32_t c = 0; c < num[l]; c++) { len_of[order[--k]] = (uint8_t)l; max_bits = l; }
    }
    // canonical codes (HUF_buildCTable): longest codes start at 0, each shorter length continues at (next value) >> 1
    uint32_t val[HUF_ENC_MAXBITS + 2];
    {
        uint32_t mn = 0;
        for (uint32_t l = max_bits; l > 0; l--) { val[l] = mn; mn += num[l]; mn >>= 1; }
    }
    uint32_t ctab[HUF_ENC_SYMS];   // code | nbBits << 16
    for (uint32_t s = 0; s <= last_sym; s++) ctab[s] = len_of[s] ? (val[len_of[s]]++ | ((uint32_t)len_of[s] << 16)) : 0u;
    // exact size before a byte is written (the payload must not outgrow the n bytes the caller reserved): code bits by the
    // histogram, plus per stream the closing bit and the byte rounding, plus tree and jump table
    {
        uint64_t bits = 0;
        for (uint32_t s = 0; s <= last_sym; s++) bits += (uint64_t)count[s] * len_of[s];
        const uint64_t est = 1 + (last_sym + 1) / 2 + 6 + (bits + 7) / 8 + 8;
        if (est + 8 >= n) return 0;   // not smaller than the raw literals: keep them raw
    }
    // ---- payload: tree description, jump table, four streams
    uint32_t o = 0;
    dst[o++] = (uint8_t)(127 + last_sym);
    for (uint32_t s = 0; s < last_sym; s += 2) {
        const uint32_t w0 = len_of[s] ? max_bits + 1 - len_of[s] : 0u;
        const uint32_t w1 = (s + 1 < last_sym && len_of[s + 1]) ? max_bits + 1 - len_of[s + 1] : 0u;
        dst[o++] = (uint8_t)((w0 << 4) | w1);
    }
    const uint32_t jump = o;
    o += 6;
    const uint32_t seg = (n + 3) / 4;
    for (uint32_t q = 0; q < 4; q++) {
        const uint32_t b0 = q * seg, b1 = q < 3 ? b0 + seg : n;
        ByteBits bw;
        bw.init(dst + o);
        for (uint32_t i = b1; i-- > b0;) { const uint32_t c = ctab[lits[i]]; bw.add(c & 0xFFFFu, c >> 16); }
        bw.add(1, 1);
        const uint32_t sz = bw.finish();
        if (q < 3) { dst[jump + 2 * q] = (uint8_t)sz; dst[jump + 2 * q + 1] = (uint8_t)(sz >> 8); }
        o += sz;
    }
    // section header: type 2 (compressed), four streams; both sizes in 10 / 14 / 18 bits
    const uint32_t sf = (n < 1024 && o < 1024) ? 1u : (n < 16384 && o < 16384) ? 2u : 3u;
    const uint32_t bits = sf == 1 ? 10u : sf == 2 ? 14u : 18u;
    const uint64_t v = 2u | (sf << 2) | ((uint64_t)n << 4) | ((uint64_t)o << (4 + bits));
    const uint32_t hl = sf + 2;
    for (uint32_t k = 0; k < hl; k++) hdr[k] = (uint8_t)(v >> (8 * k));
    *hdr_len = hl;
    return o;
}

PNA_HD void zstd_block_header(uint32_t last, uint32_t type, uint32_t size, uint8_t* h) {
    const uint32_t v = last | (type << 1) | (size << 3);
    h[0] = (uint8_t)v; h[1] = (uint8_t)(v >> 8); h[2] = (uint8_t)(v >> 16);
}

// ------------------------------------------------------------------------------------------------ deflate
PNA_HD uint32_t rev_bits_n(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}
// fixed literal/length code (RFC 1951 3.2.6), returned bit-reversed for the LSB-first writer: value | nbits << 16
PNA_HD uint32_t fixed_litlen(uint32_t sym) {
    uint32_t code, nb;
    if (sym < 144) { code = 0x30 + sym; nb = 8; }
    else if (sym < 256) { code = 0x190 + (sym - 144); nb = 9; }
    else if (sym < 280) { code = sym - 256; nb = 7; }
    else { code = 0xC0 + (sym - 280); nb = 8; }
    return rev_bits_n(code, (int)nb) | (nb << 16);
}
PNA_HD void deflate_len_sym(uint32_t len, uint32_t* sym, uint32_t* xb, uint32_t* xv) {   // len 3..258
    if (len == 258) { *sym = 285; *xb = 0; *xv = 0; return; }
    const uint32_t l = len - 3;
    if (l < 8) { *sym = 257 + l; *xb = 0; *xv = 0; return; }
    const uint32_t hb = (uint32_t)highbit32(l);        // >= 3
    const uint32_t eb = hb - 2;
    *sym = 257 + 4 * eb + ((l >> eb) & 3) + 4;         // 265 + 4*(eb-1) + top two bits below the leading one
    *xb = eb; *xv = l & ((1u << eb) - 1u);
}
PNA_HD void deflate_dist_sym(uint32_t dist, uint32_t* sym, uint32_t* xb, uint32_t* xv) {   // dist 1..32768
    const uint32_t d = dist - 1;
    if (d < 4) { *sym = d; *xb = 0; *xv = 0; return; }
    const uint32_t hb = (uint32_t)highbit32(d);        // >= 2
    const uint32_t eb = hb - 1;
    *sym = 2 * hb + ((d >> eb) & 1);
    *xb = eb; *xv = d & ((1u << eb) - 1u);
}
// Length-limited Huffman code lengths for count[0..n) (n <= NMAX): len_of[s] = 0 for absent symbols.  Two-queue merge over
// the present symbols sorted by count, then miniz's length-limit repair (fold the too-long codes into the limit, pay the
// Kraft excess back by splitting the longest code below it).  Returns the number of present symbols (0 and 1: no code built).
// Scratch of the builder, owned by the CALLER as part of ONE object (DeflateScratch below).  The helpers of the deflate writer
// declare no arrays of their own, on purpose: with per-helper arrays inlined into one lane-per-segment frame, nvcc 12.9 (sm_100a,
// -O3 and -O1) let the frame slots of a later helper overlap an array of the caller that was still live -- the run-length tokens
// came back holding the sort order of the code-length alphabet -- so that unrelated edits (a table stride, an unrolled loop) made
// every dynamic-Huffman header corrupt on the GPU while the same source is clean under ASan / UBSan on the host, on the emulator,
// and under memcheck / racecheck / initcheck.  One enclosing object is one stack slot: nothing to merge.
struct HuffScratch {
    uint16_t order[286];
    uint32_t wt[2 * 286];       // weights while the tree is built, depths afterwards
    uint16_t parent[2 * 286];
    uint32_t num[34];
};
template <int NMAX>
PNA_HD uint32_t huff_lengths(const uint32_t* count, uint32_t stride, uint32_t n, uint32_t maxbits, uint8_t* len_of, HuffScratch& hs) {
    static_assert(NMAX <= 286, "scratch is sized for the literal/length alphabet");
    uint16_t* const order = hs.order;
    uint32_t m = 0;
    for (uint32_t s = 0; s < n; s++) {
        len_of[s] = 0;
        if (count[s * stride]) order[m++] = (uint16_t)s;
    }
    // ascending by count: shell sort (an insertion sort moves ~m^2/4 keys -- ten thousand steps for a literal alphabet)
    {
        for (int gi = 0; gi < 6; gi++) {
            const uint32_t gap = gi == 0 ? 132u : gi == 1 ? 57u : gi == 2 ? 23u : gi == 3 ? 10u : gi == 4 ? 4u : 1u;
            for (uint32_t i = gap; i < m; i++) {
                const uint16_t v = order[i];
                const uint32_t cv = count[v * stride];
                uint32_t k = i;
                while (k >= gap && count[order[k - gap] * stride] > cv) { order[k] = order[k - gap]; k -= gap; }
                order[k] = v;
            }
        }
    }
    if (m < 2) return m;
    uint32_t* const wt = hs.wt;
    uint16_t* const parent = hs.parent;
    for (uint32_t k = 0; k < m; k++) wt[k] = count[order[k] * stride];
    uint32_t li = 0, ni = m, no = m;
    for (uint32_t k = 0; k + 1 < m; k++) {
        uint32_t pick[2];
        for (int q = 0; q < 2; q++) {
            if (li < m && (ni >= no || wt[li] <= wt[ni])) pick[q] = li++;
            else pick[q] = ni++;
        }
        wt[no] = wt[pick[0]] + wt[pick[1]];
        parent[pick[0]] = (uint16_t)no; parent[pick[1]] = (uint16_t)no;
        no++;
    }
    wt[no - 1] = 0;
    for (uint32_t k = no - 1; k-- > 0;) wt[k] = wt[parent[k]] + 1;
    uint32_t* const num = hs.num;
    for (uint32_t l = 0; l < 34; l++) num[l] = 0;
    for (uint32_t k = 0; k < m; k++) num[wt[k] > 32 ? 32 : wt[k]]++;
    for (uint32_t l = maxbits + 1; l < 33; l++) { num[maxbits] += num[l]; num[l] = 0; }
    uint32_t total = 0;
    for (uint32_t l = maxbits; l > 0; l--) total += num[l] << (maxbits - l);
    while (total != (1u << maxbits)) {
        num[maxbits]--;
        for (uint32_t l = maxbits - 1; l > 0; l--)
            if (num[l]) { num[l]--; num[l + 1] += 2; break; }
        total--;
    }
    uint32_t k = m;
    for (uint32_t l = 1; l <= maxbits; l++)
        for (uint32_t c = 0; c < num[l]; c++) len_of[order[--k]] = (uint8_t)l;
    return m;
}
// canonical deflate codes (RFC 1951 3.2.2) for len_of[0..n), bit-reversed for the LSB-first writer: code | nbits << 16
PNA_HD void deflate_codes(const uint8_t* len_of, uint32_t n, uint32_t* code_of, uint32_t stride, uint32_t* bl_count /*[16]*/, uint32_t* next /*[16]*/) {
    for (int l = 0; l < 16; l++) bl_count[l] = 0;
    for (uint32_t s = 0; s < n; s++) bl_count[len_of[s]]++;
    bl_count[0] = 0;
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l < 16; l++) { code = (code + bl_count[l - 1]) << 1; next[l] = code; }
    for (uint32_t s = 0; s < n; s++) {
        const uint32_t l = len_of[s];
        code_of[s * stride] = l ? (rev_bits_n(next[l]++, (int)l) | (l << 16)) : 0u;
    }
}

// One segment as ONE deflate block (+ sync marker unless final): dynamic Huffman when `dyn` and it is smaller by the exact bit
// count, else fixed Huffman.  lits = the segment's literal bytes in order, n_lit_total of them (sequence literal runs first, the
// rest trail).  dst 4-byte aligned, room for 9/8 * len + 16.  Returns bytes written.
// ws: DEFLATE_WS 32-bit slots spaced `stride` words apart -- first the histograms, then the code tables: they are touched once or
// twice per literal (the kernel passes a thread-local array with stride 1; the stride exists for table layouts shared by a CTA).
constexpr uint32_t DEFLATE_WS = 286 + 30;
struct DeflateScratch {   // every array of the writer and its helpers, as ONE object (see HuffScratch)
    HuffScratch huff;
    uint8_t lens[286 + 30];
    uint8_t seq_len[286 + 30];
    uint16_t tok[286 + 30];      // symbol | extra value << 8
    uint32_t ccl[19];
    uint8_t cl_len[20];
    uint32_t cl_code[19];
    uint32_t bl_count[16], next[16];
};
PNA_HD uint32_t deflate_write_segment(const Seq* seqs, uint32_t nseq, const uint8_t* lits, uint32_t n_lit_total, bool final_seg,
                                      uint8_t* dst, bool dyn, uint32_t* ws, uint32_t stride, DeflateScratch& sc) {
    BitOut b;
    b.init(dst);
    uint32_t* const ll_code = ws;
    uint32_t* const d_code = ws + 286 * stride;
    bool use_dyn = false;
    if (dyn) {
        // ---- histograms
        uint32_t* const cl = ll_code;   // counts first, codes later
        uint32_t* const cd = d_code;
        for (uint32_t s = 0; s < DEFLATE_WS; s++) ws[s * stride] = 0;
        uint64_t extra = 0;
        for (uint32_t i = 0; i < n_lit_total; i++) cl[lits[i] * stride]++;
        for (uint32_t i = 0; i < nseq; i++) {
            uint32_t sym, xb, xv;
            deflate_len_sym(seqs[i].llml >> 16, &sym, &xb, &xv);
            cl[sym * stride]++; extra += xb;
            deflate_dist_sym(seqs[i].off, &sym, &xb, &xv);
            cd[sym * stride]++; extra += xb;
        }
        cl[256 * stride] = 1;
        {   // at least two distance codes, as zlib's deflate writes them (readers that reject a lone / absent distance code)
            uint32_t nd = 0;
            for (uint32_t s = 0; s < 30; s++) nd += cd[s * stride] != 0;
            if (nd < 2) { if (!cd[0]) cd[0] = 1; else cd[stride] = 1; if (nd == 0) cd[stride] = 1; }
        }
        uint8_t* const lens = sc.lens;
        huff_lengths<286>(cl, stride, 286, 15, lens, sc.huff);
        huff_lengths<30>(cd, stride, 30, 15, lens + 286, sc.huff);
        uint32_t nl = 286, nd = 30;
        while (nl > 257 && lens[nl - 1] == 0) nl--;
        while (nd > 1 && lens[286 + nd - 1] == 0) nd--;
        // ---- the two length sets as one run-length coded sequence over the code-length alphabet
        uint8_t* const seq_len = sc.seq_len;
        for (uint32_t i = 0; i < nl; i++) seq_len[i] = lens[i];
        for (uint32_t i = 0; i < nd; i++) seq_len[nl + i] = lens[286 + i];
        const uint32_t ntot = nl + nd;
        uint16_t* const tok = sc.tok;
        uint32_t ntok = 0;
        uint32_t* const ccl = sc.ccl;
        for (int s = 0; s < 19; s++) ccl[s] = 0;
        for (uint32_t i = 0; i < ntot;) {
            const uint32_t v = seq_len[i];
            uint32_t run = 1;
            while (i + run < ntot && seq_len[i + run] == v) run++;
            if (v == 0 && run >= 3) {
                const uint32_t r = run > 138 ? 138 : run;
                if (r <= 10) { tok[ntok++] = (uint16_t)(17 | ((r - 3) << 8)); ccl[17]++; }
                else { tok[ntok++] = (uint16_t)(18 | ((r - 11) << 8)); ccl[18]++; }
                i += r;
            } else if (v != 0 && run >= 4) {
                tok[ntok++] = (uint16_t)v; ccl[v]++;
                const uint32_t r = run - 1 > 6 ? 6 : run - 1;
                tok[ntok++] = (uint16_t)(16 | ((r - 3) << 8)); ccl[16]++;
                i += 1 + r;
            } else { tok[ntok++] = (uint16_t)v; ccl[v]++; i++; }
        }
        uint8_t* const cl_len = sc.cl_len;
        const uint32_t mcl = huff_lengths<19>(ccl, 1, 19, 7, cl_len, sc.huff);
        // order of the code-length code lengths in the header (RFC 1951 3.2.7), packed five bits per entry: no local table
        auto ord = [](uint32_t k) -> uint32_t {
            const uint64_t lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
            const uint64_t hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
            return (uint32_t)((k < 12 ? lo >> (5 * k) : hi >> (5 * (k - 12))) & 31u);
        };
        uint32_t ncl = 19;
        while (ncl > 4 && cl_len[ord(ncl - 1)] == 0) ncl--;
        // ---- exact sizes of both forms
        uint64_t bits_dyn = 3 + 5 + 5 + 4 + 3ull * ncl + extra, bits_fix = 3 + extra;
        for (uint32_t t = 0; t < ntok; t++) { const uint32_t sy = tok[t] & 0xFF; bits_dyn += cl_len[sy] + (sy == 16 ? 2 : sy == 17 ? 3 : sy == 18 ? 7 : 0); }
        for (uint32_t s = 0; s < 286; s++) { const uint64_t c = cl[s * stride]; bits_dyn += c * lens[s]; bits_fix += c * (fixed_litlen(s) >> 16); }
        for (uint32_t s = 0; s < 30; s++) { const uint64_t c = cd[s * stride]; bits_dyn += c * lens[286 + s]; bits_fix += c * 5u; }
        // (the forced extra distance code is counted in both and never emitted: an upper bound, equal for the comparison)
        use_dyn = mcl >= 2 && bits_dyn < bits_fix;   // (fewer than two code-length symbols cannot happen: EOB has a length, absent symbols have none)
        if (use_dyn) {
            uint32_t* const cl_code = sc.cl_code;
            deflate_codes(cl_len, 19, cl_code, 1, sc.bl_count, sc.next);
            deflate_codes(lens, 286, ll_code, stride, sc.bl_count, sc.next);
            deflate_codes(lens + 286, 30, d_code, stride, sc.bl_count, sc.next);
            b.add((final_seg ? 1u : 0u) | (2u << 1), 3);   // BFINAL, BTYPE=10
            b.add(nl - 257, 5); b.add(nd - 1, 5); b.add(ncl - 4, 4);
            for (uint32_t k = 0; k < ncl; k++) b.add(cl_len[ord(k)], 3);
            for (uint32_t t = 0; t < ntok; t++) {
                const uint32_t sy = tok[t] & 0xFF, xv = tok[t] >> 8;
                b.add(cl_code[sy] & 0xFFFFu, cl_code[sy] >> 16);
                if (sy == 16) b.add(xv, 2); else if (sy == 17) b.add(xv, 3); else if (sy == 18) b.add(xv, 7);
            }
        }
    }
    if (!use_dyn) {
        b.add((final_seg ? 1u : 0u) | (1u << 1), 3);   // BFINAL, BTYPE=01
        for (uint32_t s = 0; s < 286; s++) ll_code[s * stride] = fixed_litlen(s);
        for (uint32_t s = 0; s < 30; s++) d_code[s * stride] = rev_bits_n(s, 5) | (5u << 16);
    }
    uint32_t lp = 0;
    for (uint32_t i = 0; i < nseq; i++) {
        const Seq q = seqs[i];
        const uint32_t ll = q.llml & 0xFFFFu, ml = q.llml >> 16;
        for (uint32_t k = 0; k < ll; k++) { const uint32_t c = ll_code[lits[lp + k] * stride]; b.add(c & 0xFFFFu, c >> 16); }
        lp += ll;
        uint32_t sym, xb, xv;
        deflate_len_sym(ml, &sym, &xb, &xv);
        const uint32_t c = ll_code[sym * stride];
        b.add(c & 0xFFFFu, c >> 16);
        b.add(xv, xb);
        deflate_dist_sym(q.off, &sym, &xb, &xv);
        { const uint32_t dc = d_code[sym * stride]; b.add(dc & 0xFFFFu, dc >> 16); }
        b.add(xv, xb);
    }
    for (; lp < n_lit_total; lp++) { const uint32_t c = ll_code[lits[lp] * stride]; b.add(c & 0xFFFFu, c >> 16); }
    { const uint32_t c = ll_code[256 * stride]; b.add(c & 0xFFFFu, c >> 16); }
    if (!final_seg) {
        b.add(0, 3);                               // empty stored block, not final
        if (b.n & 7) b.add(0, 8 - (b.n & 7));      // to the byte boundary
        b.add(0x0000u, 16); b.add(0xFFFFu, 16);    // LEN = 0, NLEN = ~0
    }
    return b.finish();
}

}  // namespace enc
}  // namespace pna
