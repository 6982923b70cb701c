// host_core.cpp -- g++ build of the PNA_HD algorithmic cores (the very code the kernels run)
// so that the CPU-only test tier can pin them against the oracle.  Test infrastructure.
#include "../../portable-network-archive_b200/csrc/cipher_core.cuh"
using namespace pna;
static AesTables g_aes; static CamelliaTables g_cam; static bool g_init = false;
static void init() { if (!g_init) { aes_make_tables(&g_aes); camellia_make_tables(&g_cam); g_init = true; } }
extern "C" {
int hc_ecb(int encryption, int encrypt, const uint8_t* key, const uint8_t* in, size_t n, uint8_t* out) {
    init();
    if (encryption == 1) {
        AesKey K; aes256_expand_key(&g_aes, key, &K);
        TabView te{g_aes.te0, 1, 0}, td{g_aes.td0, 1, 0};
        for (size_t i = 0; i + 16 <= n; i += 16) {
            uint32_t s[4]; memcpy(s, in + i, 16);
            if (encrypt) aes256_encrypt_block(s, K.rk, te); else aes256_decrypt_block(s, K.dk, td, g_aes.inv_sbox);
            memcpy(out + i, s, 16);
        }
        return 0;
    }
    if (encryption == 2) {
        CamelliaKey K; camellia256_expand_key(&g_cam, key, &K);
        for (size_t i = 0; i + 16 <= n; i += 16) {
            uint32_t s[4]; memcpy(s, in + i, 16);
            camellia256_crypt_block(s, encrypt ? K.ek : K.dk, &g_cam.sp_hi[0][0], &g_cam.sp_lo[0][0]);
            memcpy(out + i, s, 16);
        }
        return 0;
    }
    return 4;
}
void hc_ctr_add(const uint8_t* iv, uint64_t add, uint8_t* out) {
    uint32_t a[4], o[4]; memcpy(a, iv, 16); ctr128be_add(a, add, o); memcpy(out, o, 16);
}
}

// ---- CRC emulation: the warp algorithm with the lanes looped on the host
#include "../../portable-network-archive_b200/csrc/crc32_core.cuh"
static CrcConsts g_crc; static bool g_crc_init = false;
extern "C" uint32_t hc_crc_span(const uint8_t* img, uint64_t S, uint64_t E) {
    if (!g_crc_init) { crc_make_consts(&g_crc); g_crc_init = true; }
    uint32_t acc = 0xFFFFFFFFu;
    for (uint64_t t = S; t < E || t == S; t += CRC_TILE) {
        uint64_t te = t + CRC_TILE < E ? t + CRC_TILE : E;
        uint32_t x = 0;
        for (int lane = 0; lane < 32; lane++) x ^= crc_tile_lane(img, t, te, lane, &g_crc.U[0][0], g_crc.lane_k);
        uint32_t raw = crc_multmodp(x, g_crc.inv_z[(16 - (te & 15)) & 15]);
        uint32_t sh = (te - t) == CRC_TILE ? g_crc.x_tile : crc_x2nmodp(g_crc.x2n, te - t, 3);
        acc = crc_multmodp(sh, acc) ^ raw;
        if (te >= E) break;
    }
    return ~acc;
}

// ---- zstd: the kernels' phases run serially on the host (scan -> parse -> resolve -> entropy -> prefix -> LZ)
#include "../../portable-network-archive_b200/csrc/zstd_core.cuh"
#include <vector>
extern "C" { int hc_site = 0; int hc_site_value() { return hc_site; } }
extern "C" int hc_zstd_decode(const uint8_t* in, uint64_t len, uint8_t* out, uint64_t cap, uint64_t* out_len,
                              uint32_t* stats /* [nblocks, nseq, nlit] optional */) {
    using namespace pna::zs;
    *out_len = 0;
    std::vector<uint8_t> arena(len + 64, 0);
    memcpy(arena.data(), in, len);
    const uint8_t* comp = arena.data();
    const uint32_t* words = (const uint32_t*)arena.data();
    uint32_t nb = 0;
    int32_t st = scan_entry(comp, 0, len, 0, nullptr, &nb);
    if (st != ST_OK) { hc_site = 1; return st; }
    std::vector<ZBlock> blocks(nb ? nb : 1);
    scan_entry(comp, 0, len, 0, blocks.data(), &nb);
    uint64_t lit_total = 0, seq_total = 0;
    for (uint32_t i = 0; i < nb; i++) {
        blocks[i].status = parse_block(comp, blocks[i]);
        if (blocks[i].status) { hc_site = 2; return blocks[i].status; }
        blocks[i].lit_off = lit_total; blocks[i].seq_off = seq_total;
        if (blocks[i].type == BT_COMPRESSED) {
            if (blocks[i].lit_type >= LT_COMPRESSED) lit_total += blocks[i].lit_regen;
            seq_total += blocks[i].nseq;
        }
    }
    st = resolve_sources(blocks.data(), 0, nb);
    if (st) { hc_site = 3; return st; }
    std::vector<uint8_t> lits(lit_total + 8);
    std::vector<SeqRec> seqs(seq_total + 1);
    std::vector<uint16_t> tab16(TAB16_TOTAL);
    uint32_t llb[36], mlb[53];
    for (int c = 0; c < 36; c++) { llb[c] = seq_pack_ll((uint32_t)c); if (ll_xbits(c) != (uint32_t)ll_bits(c)) return ST_INTERNAL; }
    for (int c = 0; c < 53; c++) { mlb[c] = seq_pack_ml((uint32_t)c); if (ml_xbits(c) != (uint32_t)ml_bits(c)) return ST_INTERNAL; }
    std::vector<uint16_t> huf(1 << HUF_LOG_MAX);
    for (uint32_t i = 0; i < nb; i++) {
        ZBlock& b = blocks[i];
        if (b.type != BT_COMPRESSED) continue;
        if (b.lit_type >= LT_COMPRESSED) {
            const ZBlock& hb = blocks[b.huf_src];
            uint8_t weights[257]; FseEntry fse[64]; int hlog = 0;
            int hdr = huf_read_table(words, comp, hb.src + hb.lit_pos, hb.lit_csize, huf.data(), &hlog, weights, fse);
            if (hdr < 0) { hc_site = 4; return ST_INVALID_DATA; }
            uint32_t skip = b.lit_type == LT_COMPRESSED ? (uint32_t)hdr : 0;
            if (skip > b.lit_csize) { hc_site = 5; return ST_INVALID_DATA; }
            uint64_t at = b.src + b.lit_pos + skip; uint32_t clen = b.lit_csize - skip;
            uint8_t* dst = lits.data() + b.lit_off;
            if (b.lit_streams == 1) {
                if (!huf_decode_stream_w(words, comp, at, clen, huf.data(), hlog, dst, b.lit_regen)) { hc_site = 6; return ST_INVALID_DATA; }
            } else {
                if (clen < 6) { hc_site = 7; return ST_INVALID_DATA; }
                uint32_t s1 = load_le16(comp + at), s2 = load_le16(comp + at + 2), s3 = load_le16(comp + at + 4);
                if ((uint64_t)s1 + s2 + s3 + 6 > clen) { hc_site = 8; return ST_INVALID_DATA; }
                uint32_t s4 = clen - 6 - s1 - s2 - s3;
                uint32_t seg = (b.lit_regen + 3) / 4;
                if (seg * 3 > b.lit_regen) { hc_site = 9; return ST_INVALID_DATA; }
                uint32_t sz[4] = {s1, s2, s3, s4};
                uint64_t o = at + 6;
                for (int k = 0; k < 4; k++) {
                    uint32_t cnt = k < 3 ? seg : b.lit_regen - 3 * seg;
                    if (!huf_decode_stream_w(words, comp, o, sz[k], huf.data(), hlog, dst + (uint64_t)k * seg, cnt)) { hc_site = 10; return ST_INVALID_DATA; }
                    o += sz[k];
                }
            }
        }
        if (b.nseq) {
            int16_t norm[64]; uint16_t nxt[64];
            Tab16 tll{tab16.data() + TAB16_LL, 1}, tof{tab16.data() + TAB16_OF, 1}, tml{tab16.data() + TAB16_ML, 1};
            int l0 = seq_tab16_for(comp, blocks.data(), b, 0, tll, norm, nxt);
            int l1 = seq_tab16_for(comp, blocks.data(), b, 1, tof, norm, nxt);
            int l2 = seq_tab16_for(comp, blocks.data(), b, 2, tml, norm, nxt);
            if (l0 < 0 || l1 < 0 || l2 < 0) { hc_site = 11; return ST_INVALID_DATA; }
            st = decode_sequences16(words, comp, b, tll, tof, tml, l0, l1, l2, llb, mlb, seqs.data() + b.seq_off, &b.esc_n,
                                    b.esc_idx, b.esc_ll, b.esc_ml);
            if (st) { hc_site = 12; return st; }
        }
    }
    uint64_t total = 0;
    st = prefix_entry(blocks.data(), 0, nb, 0, &total);
    if (st) { hc_site = 13; return st; }
    *out_len = total;
    if (stats) { stats[0] = nb; stats[1] = (uint32_t)seq_total; stats[2] = (uint32_t)lit_total; }
    if (total > cap) return ST_NOSPACE;
    for (uint32_t i = 0; i < nb; i++) {
        ZBlock& b = blocks[i];
        uint8_t* o = out + b.out_off;
        if (b.type == BT_RAW) { memcpy(o, comp + b.src, b.size); continue; }
        if (b.type == BT_RLE) { memset(o, comp[b.src], b.size); continue; }
        const uint8_t* lit; uint32_t stride = 1;
        if (b.lit_type == LT_RAW) lit = comp + b.src + b.lit_pos;
        else if (b.lit_type == LT_RLE) { lit = comp + b.src + b.lit_pos; stride = 0; }
        else lit = lits.data() + b.lit_off;
        uint64_t op = 0, lp = 0;
        for (uint32_t s = 0; s < b.nseq; s++) {
            const SeqRec r = seqs[b.seq_off + s];
            uint32_t ll = r.y & 0xFFFFu, ml = r.y >> 16;
            if (ll == SEQ_ESC || ml == SEQ_ESC)
                for (uint32_t q = 0; q < b.esc_n; q++) if (b.esc_idx[q] == s) { ll = b.esc_ll[q]; ml = b.esc_ml[q]; }
            uint32_t off = resolve_rep(r.x, b.rep_in);
            for (uint32_t k = 0; k < ll; k++) o[op + k] = lit[(lp + k) * stride];
            op += ll; lp += ll;
            if (off == 0 || off > (b.out_off + op) - b.frame_out) { hc_site = 14; return ST_INVALID_DATA; }
            for (uint32_t k = 0; k < ml; k++) o[op + k] = o[op + k - off];
            op += ml;
        }
        for (uint64_t k = lp; k < b.lit_regen; k++) o[op++] = lit[k * stride];
        if (op != b.out_size) return ST_INTERNAL;
    }
    return ST_OK;
}

// ---- inflate (one stream per thread on the GPU; same function here)
#include "../../portable-network-archive_b200/csrc/inflate_core.cuh"
extern "C" int hc_inflate(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len) {
    static pna::inf::Tables t;
    return pna::inf::inflate_zlib(in, n, out, cap, out_len, &t);
}

// ---- encode: the PNA_HD bitstream writers driven by a plain greedy matcher (test-side stand-in for the
// warp-synchronous matcher of kernels_encode.cuh); output must decode with libzstd / zlib (checked in Python)
#include "../../portable-network-archive_b200/csrc/encode_core.cuh"
#include "../../portable-network-archive_b200/csrc/lzma_enc_core.cuh"
static pna::enc::EncTables g_enc; static bool g_enc_init = false;
static int g_enc_dyn = 1;   // per-block FSE tables (the default effort) or predefined ones (the fast setting)
extern "C" void hc_set_enc_dyn(int v) { g_enc_dyn = v; }
static void greedy_segment(const uint8_t* d, uint32_t len, std::vector<pna::enc::Seq>& seqs, std::vector<uint8_t>& lits) {
    using namespace pna::enc;
    std::vector<int32_t> head(1 << 13, -1);
    uint32_t p = 0, run = 0;
    while (p < len) {
        uint32_t best = 0, bo = 0;
        if (p + 4 <= len) {
            uint32_t v; memcpy(&v, d + p, 4);
            uint32_t h = (v * 2654435761u) >> 19;
            int32_t c = head[h];
            head[h] = (int32_t)p;
            if (c >= 0) {
                uint32_t mx = len - p < MAX_MATCH ? len - p : MAX_MATCH, k = 0;
                while (k < mx && d[c + k] == d[p + k]) k++;
                if (k >= MIN_MATCH) { best = k; bo = p - (uint32_t)c; }
            }
        }
        if (best) { seqs.push_back({bo, run | (best << 16)}); run = 0; p += best; }
        else { lits.push_back(d[p]); run++; p++; }
    }
}
static uint32_t g_xz_lc = 2;
extern "C" void hc_set_xz_lc(int lc) { g_xz_lc = (uint32_t)lc; }
extern "C" uint64_t hc_encode(int compression, const uint8_t* in, uint64_t len, uint8_t* out) {
    using namespace pna::enc;
    if (!g_enc_init) { make_enc_tables(&g_enc); g_enc_init = true; }
    uint64_t o = 0;
    const uint64_t nseg = len ? (len + SEG - 1) / SEG : 0;
    if (compression == 4) {   // xz: the segment writer + container pieces of lzma_enc_core.cuh, stitched the way the kernels do
        using namespace pna::xz;
        o = xz_write_front(out, nseg == 0);
        uint64_t chunk_bytes = 0;
        uint32_t crc = 0;
        std::vector<uint16_t> probs(ENC_PROBS + 2);
        for (uint64_t s = 0; s < nseg; s++) {
            const uint8_t* d = in + s * SEG;
            const uint32_t n = (uint32_t)(len - s * SEG < SEG ? len - s * SEG : SEG);
            std::vector<Seq> seqs; std::vector<uint8_t> lits;
            greedy_segment(d, n, seqs, lits);
            for (auto& q : probs) q = (uint16_t)PROB_INIT;
            std::vector<uint8_t> body(SEG + SEG / 8 + 128);   // what the kernel has (TMP_SEG)
            const uint32_t cap = n > 4 ? n - 4 : 0;
            const uint32_t cs = cap ? lzma_encode_segment(d, n, seqs.data(), (uint32_t)seqs.size(), probs.data(), body.data(), cap, (uint32_t)body.size(), g_xz_lc) : 0xFFFFFFFFu;
            uint32_t hl;
            if (cs == 0xFFFFFFFFu) { hl = lzma2_chunk_header(out + o, n, 0, g_xz_lc); o += hl; memcpy(out + o, d, n); o += n; chunk_bytes += hl + n; }
            else { hl = lzma2_chunk_header(out + o, n, cs, g_xz_lc); o += hl; memcpy(out + o, body.data(), cs); o += cs; chunk_bytes += hl + cs; }
            // lanes' slices combined the way xz_encode_kernel does
            uint32_t sc = 0;
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t lo = l * 1024, nl = lo >= n ? 0 : (n - lo < 1024 ? n - lo : 1024);
                if (nl) sc = crc_concat(sc, crc_slice(d + lo, nl), crc_xpow_bytes(nl));
            }
            crc = crc_concat(crc, sc, crc_xpow_bytes(n));
        }
        o += xz_write_back(out + o, nseg == 0, chunk_bytes, len, crc);
        return o;
    }
    if (compression == 2) {
        const uint8_t fh[6] = {0x28, 0xB5, 0x2F, 0xFD, 0x00, 0x38};
        memcpy(out, fh, 6); o = 6;
        if (nseg == 0) { zstd_block_header(1, 0, 0, out + o); o += 3; return o; }
    } else {
        out[0] = 0x78; out[1] = 0x9C; o = 2;
        if (nseg == 0) { out[o++] = 0x03; out[o++] = 0x00; }
    }
    uint32_t a1 = 1, a2 = 0;
    for (uint64_t i = 0; i < len; i++) { a1 = (a1 + in[i]) % 65521u; a2 = (a2 + a1) % 65521u; }
    for (uint64_t s = 0; s < nseg; s++) {
        const uint8_t* d = in + s * SEG;
        const uint32_t n = (uint32_t)(len - s * SEG < SEG ? len - s * SEG : SEG);
        const bool last = s + 1 == nseg;
        std::vector<Seq> seqs; std::vector<uint8_t> lits;
        greedy_segment(d, n, seqs, lits);
        std::vector<uint8_t> tmp(2 * SEG + 64);
        uint8_t* t4 = tmp.data() + ((4 - ((uintptr_t)tmp.data() & 3)) & 3);
        if (compression == 2) {
            uint32_t ssz = 1, soff = 0;
            if (!seqs.empty()) { zstd_assign_repcodes(seqs.data(), (uint32_t)seqs.size()); uint32_t r = zstd_write_sequences(g_enc, seqs.data(), (uint32_t)seqs.size(), t4, 2 * SEG, g_enc_dyn != 0); soff = r >> 24; ssz = r & 0xFFFFFF; }
            else { t4[0] = 0; }
            uint8_t lh[5]; uint32_t lhn = 0;
            std::vector<uint8_t> cl(lits.size() + 64);
            const uint32_t clit = zstd_write_literals(lits.data(), (uint32_t)lits.size(), cl.data(), lh, &lhn);
            const uint32_t lpay = clit ? clit : (uint32_t)lits.size();
            uint32_t csize = lhn + lpay + ssz;
            if (csize >= n) { zstd_block_header(last, 0, n, out + o); o += 3; memcpy(out + o, d, n); o += n; }
            else {
                zstd_block_header(last, 2, csize, out + o); o += 3;
                memcpy(out + o, lh, lhn); o += lhn;
                memcpy(out + o, clit ? cl.data() : lits.data(), lpay); o += lpay;
                memcpy(out + o, t4 + soff, ssz); o += ssz;
            }
        } else {
            uint32_t ws[DEFLATE_WS];
            DeflateScratch scratch;
            uint32_t sz = deflate_write_segment(seqs.data(), (uint32_t)seqs.size(), lits.data(), (uint32_t)lits.size(), last, t4, g_enc_dyn != 0, ws, 1, scratch);
            if (sz >= n + 5) {   // stored block
                out[o++] = last ? 1 : 0; out[o++] = (uint8_t)n; out[o++] = (uint8_t)(n >> 8); out[o++] = (uint8_t)~n; out[o++] = (uint8_t)(~n >> 8);
                memcpy(out + o, d, n); o += n;
            } else { memcpy(out + o, t4, sz); o += sz; }
        }
    }
    if (compression != 2) { const uint32_t ad = (a2 << 16) | a1; out[o++] = ad >> 24; out[o++] = ad >> 16; out[o++] = ad >> 8; out[o++] = ad; }
    return o;
}

// ---- GCM STREAM: the tile algorithm of kernels_gcm.cuh with the lanes looped on the host (gcm_core.cuh is the same code)
#include "../../portable-network-archive_b200/csrc/gcm_core.cuh"
#include "../../portable-network-archive_b200/csrc/aead_host.hpp"
// one segment: CTR from J0+1 and the tag, GHASH evaluated tile by tile exactly like gcm_tiles_kernel / gcm_finish_kernel
// (first tile short, lane l takes blocks l, l+32, ... with H^32, 5-level lane tree, tiles chained with H^1024)
extern "C" int hc_gcm_segment(int encryption, const uint8_t* key, const uint8_t* nonce12, const uint8_t* ct, uint64_t n, uint8_t* plain,
                              uint8_t* tag) {
    init();
    using namespace pna::gcm;
    const uint32_t last4[16] = PNA_GCM_LAST4;
    AesKey AK; CamelliaKey CK;
    if (encryption == 1) aes256_expand_key(&g_aes, key, &AK); else camellia256_expand_key(&g_cam, key, &CK);
    auto E = [&](uint32_t s[4]) {
        if (encryption == 1) aes256_encrypt_block(s, AK.rk, TabView{g_aes.te0, 1, 0});
        else camellia256_crypt_block(s, CK.ek, &g_cam.sp_hi[0][0], &g_cam.sp_lo[0][0]);
    };
    uint32_t h[4] = {0, 0, 0, 0};
    E(h);
    GcmPow P;
    make_powers(from_le_words(h[0], h[1], h[2], h[3]), &P);
    uint32_t nw[3];
    memcpy(nw, nonce12, 12);
    const uint64_t nb = (n + 15) / 16, head = nb % 1024;
    G128 y = G128{{0, 0, 0, 0}};
    bool first_tile = true;
    for (uint64_t b0 = 0; b0 < nb;) {
        const uint32_t m = (b0 == 0 && head) ? (uint32_t)head : 1024u, pad = 1024u - m;
        G128 lanes[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
            G128 yl = G128{{0, 0, 0, 0}};
            for (uint32_t k = pad >> 5; k < 32; k++) {
                const uint32_t r = k * 32 + lane;
                yl = mul_table(yl, P.t[GCM_POW_STRIDE].e, last4);
                if (r < pad) continue;
                const uint64_t bi = b0 + (r - pad), off = bi * 16;
                const uint32_t have = (uint32_t)(n - off >= 16 ? 16 : n - off);
                uint32_t c[4] = {0, 0, 0, 0};
                memcpy(c, ct + off, have);
                uint32_t o[4] = {nw[0], nw[1], nw[2], bswap32((uint32_t)bi + 2u)};
                E(o);
                for (int q = 0; q < 4; q++) o[q] ^= c[q];
                memcpy(plain + off, o, have);
                gxor(yl, from_le_words(c[0], c[1], c[2], c[3]));
            }
            lanes[lane] = yl;
        }
        for (int l = 0; l < 5; l++) {
            G128 nx[32];
            for (uint32_t lane = 0; lane < 32; lane++) {
                G128 a = lanes[lane], b = lanes[lane ^ (1u << l)];
                if (!((lane >> l) & 1)) a = mul_table(a, P.t[l].e, last4); else b = mul_table(b, P.t[l].e, last4);
                gxor(a, b);
                nx[lane] = a;
            }
            memcpy(lanes, nx, sizeof nx);
        }
        const G128 part = mul_table(lanes[0], P.t[0].e, last4);
        if (!first_tile) y = mul_table(y, P.t[GCM_POW_TILE].e, last4);
        gxor(y, part);
        first_tile = false;
        b0 += m;
    }
    const uint64_t bits = n * 8;
    y.w[1] ^= (uint32_t)(bits >> 32); y.w[0] ^= (uint32_t)bits;
    y = mul_table(y, P.t[0].e, last4);
    uint32_t j0[4] = {nw[0], nw[1], nw[2], bswap32(1u)};
    E(j0);
    uint32_t t[4];
    to_le_words(y, t);
    for (int q = 0; q < 4; q++) t[q] ^= j0[q];
    memcpy(tag, t, 16);
    return 0;
}
extern "C" void hc_hkdf_sha256(const uint8_t* ikm, uint64_t n_ikm, const uint8_t* salt, uint64_t n_salt, const uint8_t* info, uint64_t n_info,
                               uint8_t* okm) {
    pna::aead::hkdf_sha256(ikm, n_ikm, salt, n_salt, info, n_info, okm);
}
extern "C" void hc_sha256(const uint8_t* d, uint64_t n, uint8_t* out) { pna::aead::sha256(d, n, nullptr, 0, out); }

// ---- xz: the .xz container + LZMA2 decoder the kernel runs one lane per stream
#include "../../portable-network-archive_b200/csrc/lzma_core.cuh"
extern "C" int hc_xz_decode(const uint8_t* in, uint64_t len, uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* /*stats*/) {
    std::vector<uint16_t> probs(pna::xz::LZMA_PROBS_MAX);
    return pna::xz::xz_decode(in, len, out, cap, out_len, probs.data());
}
extern "C" int hc_xz_size(const uint8_t* in, uint64_t len, uint64_t* out_len) { return pna::xz::xz_stream_size(in, len, out_len); }
// the chunk-parallel pass of kernels_xz.cuh with the windows looped on the host: every window decodes the chunks that start in it
// (xz_chunked_layout + lzma_chunk, CRC of what it produced), then xz_decode runs with the window records.  *used = 1 when the
// stream qualified and every window came back ok (the payloads were not decoded again).
extern "C" int hc_xz_decode_windows(const uint8_t* in, uint64_t len, uint8_t* out, uint64_t cap, uint64_t* out_len, int* used) {
    using namespace pna::xz;
    const uint32_t n_wins = (uint32_t)((cap + XZ_WIN - 1) / XZ_WIN);
    std::vector<XzWin> wins(n_wins);
    std::vector<uint16_t> probs(LZMA_PROBS_MAX);
    for (uint32_t w = 0; w < n_wins; w++) {
        XzWin r{0, 0, 0x80000000u, 2};
        uint64_t fi = 0, fo = 0;
        uint32_t nc = 0;
        bool ok = xz_chunked_layout(in, len, cap, w, &fi, &fo, &nc);
        uint64_t pos = fi, op = fo;
        for (uint32_t c = 0; ok && c < nc; c++) {
            const uint32_t ctl = in[pos];
            if (ctl == 1) {
                const uint32_t usize = (((uint32_t)in[pos + 1] << 8) | in[pos + 2]) + 1;
                memcpy(out + op, in + pos + 3, usize);
                pos += 3 + usize; op += usize;
            } else {
                const uint32_t usize = (((ctl & 0x1Fu) << 16) | ((uint32_t)in[pos + 1] << 8) | in[pos + 2]) + 1;
                const uint32_t csize = (((uint32_t)in[pos + 3] << 8) | in[pos + 4]) + 1;
                uint32_t props = in[pos + 5];
                LzmaState S;
                S.state = 0; S.rep0 = S.rep1 = S.rep2 = S.rep3 = 0;
                S.pb = props / 45; props -= S.pb * 45;
                S.lp = props / 9; S.lc = props - S.lp * 9;
                S.need_props = false; S.need_dict_reset = false;
                if (in[pos + 5] > (4 * 5 + 4) * 9 + 8 || S.lc + S.lp > 4) { ok = false; break; }
                lzma_reset_probs(probs.data(), S.lc, S.lp);
                if (lzma_chunk(S, probs.data(), in + pos + 6, csize, out, op, usize, op) != pna::ST_OK) { ok = false; break; }
                pos += 6 + csize; op += usize;
            }
        }
        if (ok) {
            // lanes' slices combined the way the kernel does
            const uint64_t L = op - fo, sl = (L + 31) / 32;
            uint32_t crc = 0;
            for (uint32_t l = 0; l < 32; l++) {
                const uint64_t lo = (uint64_t)l * sl, nl = lo >= L ? 0 : (L - lo < sl ? L - lo : sl);
                crc = xz_crc_mul(xz_crc_xpow(nl), crc) ^ (nl ? xz_crc32(out + fo + lo, nl) : 0u);
            }
            r = XzWin{crc, (uint32_t)L, xz_crc_xpow(L), 1};
        }
        wins[w] = r;
    }
    *used = 0;
    if (n_wins) {
        uint64_t fi = 0, fo = 0; uint32_t nc = 0;
        bool pre = xz_chunked_layout(in, len, cap, 0, &fi, &fo, &nc);
        for (uint32_t w = 0; pre && w < n_wins; w++) pre = wins[w].state == 1;
        *used = pre ? 1 : 0;
    }
    return xz_decode(in, len, out, cap, out_len, probs.data(), wins.data(), n_wins);
}

// ---- development aid: where do an encoded stream's bytes go?  stats[0..5] = literals, sequences, bytes of the literal
// sections, bytes of the sequence sections, empirical-entropy bytes of the (ll, ml, of) codes + extra bits per block, and
// the extra bits alone (what a per-block FSE table could reach at best, table headers not counted)
#include <cmath>
extern "C" void hc_encode_stats(const uint8_t* in, uint64_t len, double* stats) {
    using namespace pna::enc;
    if (!g_enc_init) { make_enc_tables(&g_enc); g_enc_init = true; }
    for (int i = 0; i < 24; i++) stats[i] = 0;
    const uint64_t nseg = (len + SEG - 1) / SEG;
    for (uint64_t s = 0; s < nseg; s++) {
        const uint8_t* d = in + s * SEG;
        const uint32_t n = (uint32_t)(len - s * SEG < SEG ? len - s * SEG : SEG);
        std::vector<Seq> seqs; std::vector<uint8_t> lits;
        greedy_segment(d, n, seqs, lits);
        stats[0] += lits.size(); stats[1] += seqs.size();
        std::vector<uint8_t> tmp(2 * SEG + 64), cl(lits.size() + 64);
        uint8_t* t4 = tmp.data() + ((4 - ((uintptr_t)tmp.data() & 3)) & 3);
        uint8_t lh[5]; uint32_t lhn = 0;
        const uint32_t clit = zstd_write_literals(lits.data(), (uint32_t)lits.size(), cl.data(), lh, &lhn);
        stats[2] += lhn + (clit ? clit : lits.size());
        if (seqs.empty()) continue;
        zstd_assign_repcodes(seqs.data(), (uint32_t)seqs.size());
        const uint32_t r = zstd_write_sequences(g_enc, seqs.data(), (uint32_t)seqs.size(), t4, 2 * SEG, g_enc_dyn != 0);
        stats[3] += r & 0xFFFFFF;
        double hl[36] = {0}, hm[53] = {0}, ho[32] = {0}, extra = 0;
        for (const Seq& q : seqs) {
            const uint32_t ll = q.llml & 0xFFFFu, mlb = (q.llml >> 16) - 3u, ob = (q.off & SEQ_REPCODE) ? (q.off & 3u) : q.off + 3u;
            const uint32_t c1 = ll_code_of(g_enc, ll), c2 = ml_code_of(g_enc, mlb), c3 = (uint32_t)pna::highbit32(ob);
            hl[c1]++; hm[c2]++; ho[c3]++;
            extra += pna::zs::ll_bits((int)c1) + pna::zs::ml_bits((int)c2) + c3;
        }
        double bits = extra;
        const double N = (double)seqs.size();
        for (double c : hl) if (c > 0) bits += -c * std::log2(c / N);
        for (double c : hm) if (c > 0) bits += -c * std::log2(c / N);
        for (double c : ho) if (c > 0) bits += -c * std::log2(c / N);
        stats[4] += bits / 8; stats[5] += extra / 8;
        // cost with counts normalised to 2^L (every present symbol >= 1), L = 5..9 -> stats[8 + L]
        for (int L = 5; L <= 9; L++) {
            double tb = extra;
            for (int t = 0; t < 3; t++) {
                double* h = t == 0 ? hl : t == 1 ? hm : ho;
                const int ns = t == 0 ? 36 : t == 1 ? 53 : 32;
                int norm[64]; int sum = 0, big = 0;
                const int size = 1 << L;
                int present = 0;
                for (int k = 0; k < ns; k++) if (h[k] > 0) present++;
                if (present > size) { tb += 1e9; continue; }
                for (int k = 0; k < ns; k++) { norm[k] = h[k] > 0 ? std::max(1, (int)std::lround(h[k] * size / N)) : 0; sum += norm[k]; if (norm[k] > norm[big]) big = k; }
                while (sum != size) {
                    big = 0; for (int k = 0; k < ns; k++) if (norm[k] > norm[big]) big = k;
                    if (sum > size) { norm[big]--; sum--; } else { norm[big]++; sum++; }
                }
                for (int k = 0; k < ns; k++) if (h[k] > 0) tb += -h[k] * std::log2((double)norm[k] / size);
                tb += 8 * (4 + present * 0.8);   // rough header: ~6 bits per present symbol
            }
            stats[8 + L] += tb / 8;
        }
    }
}
