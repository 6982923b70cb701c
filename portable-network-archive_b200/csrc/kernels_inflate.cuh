// kernels_inflate.cuh -- sm_100a kernels for zlib/deflate decode (K5).
//
// Two-stage path (every stream whose decoded size is known and < 2 GiB):
//   inflate_tokens_kernel  ONE LANE per stream.  DEFLATE's bit-serial part (dynamic/fixed Huffman tables in shared
//                          memory, 1804 B per stream, 128 streams per SM) only records what the stream says: literal
//                          bytes into a literal buffer and one 8-byte (distance, literal run | match length << 16) record
//                          per match -- the records the zstd sequence stage produces.  No output byte is touched, so a
//                          lane never waits on its own stores and the 32 lanes of a warp diverge only between the
//                          literal and the match arm of the symbol loop.  The kernel also writes the ZBlock / ZEntry that
//                          make each stream look like a one-block zstd frame.
//   zstd_lz_kernel         executes the copies, one CTA per stream (kernels_zstd_lz.cuh).
//   inflate_adler_kernel   warp per stream: Adler-32 of the finished output against the stored trailer.
// inflate_kernel (one stream per warp, lane 0 decodes from bits to bytes) remains for streams of 2 GiB and more.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "inflate_core.cuh"
#include "kernels_zstd_lz.cuh"

namespace pna {
namespace inf {

static_assert(sizeof(Tables) == 1804 && (sizeof(Tables) / 4) % 2 == 1, "one stream's tables: an odd number of 32-bit words");
static_assert(sizeof(TokenRec) == sizeof(zs::SeqRec), "token records are sequence records");
constexpr int INFLATE_CTA = 8;     // streams (= warps) per CTA of inflate_kernel
constexpr int TOKEN_CTA = 128;     // streams (= threads) per CTA of inflate_tokens_kernel
constexpr uint32_t TOKEN_SMEM_BYTES = (uint32_t)sizeof(Tables) * TOKEN_CTA;   // 230912

struct InfStream {      // one deflate stream of the two-stage path (host filled)
    uint32_t entry;     // index into EntryRec[]
    uint32_t _pad;
    uint64_t lit_off;   // its literal bytes inside the literal arena
    uint64_t rec_off;   // its first record inside the record arena
};
struct InfTrailer { uint32_t want, has; };
// worst-case records of a stream that decodes to `cap` bytes: a match yields >= 3 bytes, a literal run is cut every
// TOKEN_RUN_MAX bytes
__host__ __device__ inline uint64_t token_rec_bound(uint64_t cap) { return cap / 3 + cap / TOKEN_RUN_MAX + 2; }

// ---- the token stage as a warp-synchronous state machine -------------------------------------------------------
// inflate_zlib_to() (inflate_core.cuh) is the same decoder written as straight-line code with early exits; run one
// stream per lane it leaves the 32 lanes split after their first divergent branch (measured: 1.4 active lanes per
// instruction).  Here every lane advances its own stream through HEADER -> SYMBOLS | STORED -> TRAILER, and the warp
// reconverges after every state region and after every symbol, so a warp instruction of the symbol loop serves all
// lanes that are inside a block.  Accept / reject / truncate behaviour is that of inflate_zlib_to, case by case.
enum : int { S_HDR = 0, S_SYM = 1, S_STORED = 2, S_TRAILER = 3, S_DONE = 4 };

// code lengths of a dynamic block -> tables.  0 ok, 1 truncated input, 2 corrupt
__device__ __noinline__ int read_dynamic_tables(Bits& b, Tables* t) {
    uint8_t lengths[MAXL + MAXD + 2];
    int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
    if (b.overrun()) return 1;
    if (nlen > 286 || ndist > 30) return 2;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int i = 0;
    for (; i < ncode; i++) lengths[order[i]] = (uint8_t)b.get(3);
    for (; i < 19; i++) lengths[order[i]] = 0;
    if (b.overrun()) return 1;
    if (build(&t->len, lengths, 19) != 0) return 2;   // code-length code must be complete
    i = 0;
    while (i < nlen + ndist) {
        int sym = decode_sym(b, &t->len);
        if (b.overrun()) return 1;
        if (sym < 0) return 2;
        if (sym < 16) lengths[i++] = (uint8_t)sym;
        else {
            int rep, val = 0;
            if (sym == 16) {
                if (i == 0) return 2;
                val = lengths[i - 1];
                rep = 3 + (int)b.get(2);
            } else if (sym == 17) rep = 3 + (int)b.get(3);
            else rep = 11 + (int)b.get(7);
            if (b.overrun()) return 1;
            if (i + rep > nlen + ndist) return 2;
            while (rep--) lengths[i++] = (uint8_t)val;
        }
    }
    if (lengths[256] == 0) return 2;
    int err = build(&t->len, lengths, nlen);
    if (err && (err < 0 || nlen != t->len.n01)) return 2;
    err = build(&t->dist, lengths + nlen, ndist);
    if (err && (err < 0 || ndist != t->dist.n01)) return 2;
    return 0;
}
__device__ __noinline__ void build_fixed_tables(Tables* t) {
    uint8_t lengths[MAXL];
    int s = 0;
    for (; s < 144; s++) lengths[s] = 8;
    for (; s < 256; s++) lengths[s] = 9;
    for (; s < 280; s++) lengths[s] = 7;
    for (; s < 288; s++) lengths[s] = 8;
    build(&t->len, lengths, 288);
    for (s = 0; s < 30; s++) lengths[s] = 5;
    build(&t->dist, lengths, 30);
}

// All 32 lanes of the warp call this together (lanes without a stream pass n_valid = false).
__device__ __forceinline__ int32_t inflate_tokens_lane(TokenEmit& E, const uint8_t* in, uint64_t n, Tables* t, bool have) {
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31u;
    int state = S_DONE, last = 0;
    int32_t st = ST_OK;
    uint32_t stored_left = 0;
    Bits b;
    b.init(in, 0);
    if (have) {
        if (n < 2) E.no_trailer();   // nothing in, nothing out / truncated header: short read, no error
        else {
            const uint32_t cmf = in[0], flg = in[1];
            if (((cmf << 8) | flg) % 31 != 0 || (cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 0x20)) st = ST_INVALID_INPUT;
            else { b.init(in + 2, n - 2); state = S_HDR; }
        }
    }
    auto trunc = [&]() { E.no_trailer(); st = E.counting ? ST_NOSPACE : ST_OK; state = S_DONE; };
    auto bad = [&]() { st = ST_INVALID_INPUT; state = S_DONE; };
    for (;;) {
        if (!__any_sync(FULL, state != S_DONE)) break;
        // ---- block header (+ tables)
        if (state == S_HDR) {
            last = (int)b.get(1);
            const int type = (int)b.get(2);
            if (b.overrun()) trunc();
            else if (type == 0) {
                b.align_byte();
                if (b.pos + 4 > b.n) trunc();
                else {
                    const uint32_t len = load_le16(b.p + b.pos), nlen = load_le16(b.p + b.pos + 2);
                    b.skip_bytes(4);
                    if (len != (~nlen & 0xFFFFu)) bad();
                    else { stored_left = len; state = S_STORED; }
                }
            } else if (type == 3) bad();
            else {
                int rc = 0;
                if (type == 1) build_fixed_tables(t);
                else { Bits tb = b; rc = read_dynamic_tables(tb, t); b = tb; }   // only the copy's address escapes: b stays in registers
                if (rc == 1) trunc(); else if (rc == 2) bad(); else state = S_SYM;
            }
        }
        __syncwarp();
        // ---- symbols: up to 32 per round, the lanes inside a block in lock step
        {
            const uint32_t m = __ballot_sync(FULL, state == S_SYM);
            if ((m >> lane) & 1u) {
                for (int k = 0; k < 32; k++) {
                    if (state == S_SYM) {
                        int rc = -1;   // -1 continue, 0 end of block, 1 truncated, 2 invalid
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
                        if (rc == 0) state = last ? S_TRAILER : S_HDR;
                        else if (rc == 1) trunc();
                        else if (rc == 2) bad();
                    }
                    __syncwarp(m);
                }
            }
        }
        __syncwarp();
        // ---- stored block: up to 256 bytes per round
        if (state == S_STORED) {
            const uint64_t avail = b.n - b.pos;
            uint32_t take = stored_left < 256u ? stored_left : 256u;
            if (take > avail) take = (uint32_t)avail;
            for (uint32_t i = 0; i < take; i++) E.lit(b.p[b.pos + i]);
            b.skip_bytes(take);
            stored_left -= take;
            if (stored_left && b.pos >= b.n) trunc();
            else if (!stored_left) state = last ? S_TRAILER : S_HDR;
        }
        __syncwarp();
        // ---- Adler-32 trailer (big endian) after discarding to the byte boundary
        if (state == S_TRAILER) {
            if (E.counting) { E.no_trailer(); st = ST_NOSPACE; }
            else {
                const uint64_t tpos = (b.consumed_bits() + 7) / 8;
                if (tpos + 4 > b.n) E.no_trailer();   // truncated trailer: short read, no error
                else {
                    const uint8_t* tr = b.p + tpos;
                    E.trailer_ok(((uint32_t)tr[0] << 24) | ((uint32_t)tr[1] << 16) | ((uint32_t)tr[2] << 8) | tr[3]);
                }
            }
            state = S_DONE;
        }
    }
    return st;
}

// size_only: count the decoded length only (capacity 0: nothing is recorded); blocks / ze / trailers are not written.
__global__ void __launch_bounds__(TOKEN_CTA) inflate_tokens_kernel(const uint8_t* __restrict__ buf, EntryRec* entries,
                                                                  const InfStream* __restrict__ streams, uint32_t n,
                                                                  uint8_t* __restrict__ lits, zs::SeqRec* __restrict__ recs,
                                                                  zs::ZBlock* __restrict__ blocks, zs::ZEntry* __restrict__ ze,
                                                                  InfTrailer* __restrict__ trailers, int size_only) {
    extern __shared__ uint32_t tok_smem[];
    Tables* tabs = reinterpret_cast<Tables*>(tok_smem);
    const uint32_t i = blockIdx.x * TOKEN_CTA + threadIdx.x;
    const bool in_range = i < n;                       // lanes past the end still take part in the warp's rounds
    const InfStream s = in_range ? streams[i] : InfStream{0u, 0u, 0ull, 0ull};
    EntryRec& e = entries[s.entry];
    if (in_range && !size_only) {
        zs::ZEntry z;
        z.entry = s.entry; z.blk_begin = i; z.blk_count = 0; z.n_frames = 1; z.unit_begin = 0; z._pad = 0;
        z.lit_base = s.lit_off; z.seq_base = s.rec_off; z.lit_total = 0; z.seq_total = 0;
        ze[i] = z;
    }
    const bool have = in_range && (e.status == ST_OK || (size_only == 0 && e.status == ST_NOSPACE));
    TokenEmit E;
    E.init(lits + s.lit_off, reinterpret_cast<TokenRec*>(recs + s.rec_off), (size_only || !have) ? 0 : e.out_cap);
    int32_t st = inflate_tokens_lane(E, buf + e.comp_off, have ? e.comp_len : 0, tabs + threadIdx.x, have);
    if (!have) return;
    e.out_len = E.op;
    if (size_only && st == ST_NOSPACE) st = ST_OK;
    if (st != ST_OK) { atomicCAS(&e.status, ST_OK, st); return; }
    if (size_only) return;
    trailers[i].want = E.want; trailers[i].has = E.has_trailer ? 1u : 0u;
    zs::ZBlock b;
    memset(&b, 0, sizeof b);
    b.entry = s.entry;
    b.type = zs::BT_COMPRESSED; b.lit_type = zs::LT_COMPRESSED; b.first_in_frame = 1;
    b.out_size = (uint32_t)E.op;          // < 2^31 on this path
    b.lit_regen = E.nlit; b.nseq = E.nrec;
    b.rep_in[0] = 1; b.rep_in[1] = 4; b.rep_in[2] = 8;   // unused: offsets are absolute
    b.status = ST_OK;
    blocks[i] = b;
    ze[i].blk_count = 1;
}

// Adler-32 (RFC 1950) of every finished stream against its trailer: A = 1 + sum d_k, B = n + sum (n - k) d_k (mod 65521)
__global__ void __launch_bounds__(256) inflate_adler_kernel(EntryRec* entries, const InfStream* __restrict__ streams,
                                                            const InfTrailer* __restrict__ trailers, uint32_t n,
                                                            const uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    EntryRec& e = entries[streams[i].entry];
    if (e.status != ST_OK || !trailers[i].has) return;
    const uint64_t len = e.out_len;
    const uint8_t* p = out + e.out_off;   // 16-byte aligned
    uint64_t sa = 0, sb = 0;
    uint32_t since = 0;
    for (uint64_t k = 16ull * lane; k < len; k += 512) {
        const uint4 v = *reinterpret_cast<const uint4*>(p + k);   // the out arena is padded: reading past len is safe
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const uint64_t pos = k + q;
            const uint32_t d = pos < len ? (w[q >> 2] >> (8 * (q & 3))) & 0xFFu : 0u;
            sa += d;
            sb += (len - pos) * d;   // pos >= len contributes 0 (d = 0); len - pos wraps but is multiplied by 0
        }
        if (++since == (1u << 16)) { sa %= 65521u; sb %= 65521u; since = 0; }
    }
    sa %= 65521u; sb %= 65521u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xFFFFFFFFu, sa, o); sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o); }
    if (lane == 0) {
        const uint32_t A = (uint32_t)((1 + sa) % 65521u), B = (uint32_t)((len % 65521u + sb) % 65521u);
        if (((B << 16) | A) != trailers[i].want) atomicCAS(&e.status, ST_OK, ST_INVALID_INPUT);
    }
}

// list[i] = index into EntryRec[].  size_only: decode with cap 0 to learn the length (two-pass sizing).
__global__ void __launch_bounds__(32 * INFLATE_CTA) inflate_kernel(const uint8_t* __restrict__ buf, EntryRec* entries,
                                                              const uint32_t* __restrict__ list, uint32_t n,
                                                              uint8_t* __restrict__ out, int size_only) {
    extern __shared__ uint32_t smem_raw[];
    Tables* tabs = reinterpret_cast<Tables*>(smem_raw);
    if (threadIdx.x & 31) return;   // lane 0 of every warp works; the rest of the warp only keeps the schedulers' slots
    const uint32_t w = threadIdx.x >> 5;
    uint32_t i = blockIdx.x * INFLATE_CTA + w;
    if (i >= n) return;
    EntryRec& e = entries[list[i]];
    if (e.status != ST_OK && !(size_only == 0 && e.status == ST_NOSPACE)) return;
    uint64_t produced = 0;
    int32_t st = inflate_zlib(buf + e.comp_off, e.comp_len, out + e.out_off, size_only ? 0 : e.out_cap, &produced,
                              tabs + w);
    e.out_len = produced;
    if (size_only && st == ST_NOSPACE) st = ST_OK;
    if (st != ST_OK) atomicCAS(&e.status, ST_OK, st);
}

}  // namespace inf
}  // namespace pna
