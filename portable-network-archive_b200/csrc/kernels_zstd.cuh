// kernels_zstd.cuh -- sm_100a kernels for zstd frame decode (K4).  See zstd_core.cuh for the format
// logic and DESIGN.md for the phase plan:
//   zstd_count/fill  (thread per entry)  frame + block header walk
//   zstd_parse       (thread per block)  literal/sequence section headers, table description offsets
//   zstd_resolve     (thread per entry)  Repeat_Mode / treeless sources, per-entry literal+sequence offsets
//   zstd_entropy     (CTA per block)     Huffman literals (4 streams) and FSE sequences, block-parallel
//   zstd_prefix      (thread per entry)  output offsets, absolute repeat-offset history per block
//   zstd_lz          (warp per entry)    literal copy + match copy, 32 sequences per step
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "zstd_core.cuh"

namespace pna {
namespace zs {

struct ZEntry {          // per zstd entry, device resident
    uint32_t entry;      // index into EntryRec[]
    uint32_t blk_begin;  // first ZBlock
    uint32_t blk_count;
    uint32_t _pad;
    uint64_t lit_base;   // literal arena base of this entry
    uint64_t seq_base;   // sequence array base
    uint64_t lit_total;  // device-written by zstd_resolve
    uint64_t seq_total;
};

__device__ __forceinline__ void set_status(EntryRec* entries, uint32_t e, int32_t st) {
    if (st != ST_OK) atomicCAS(&entries[e].status, ST_OK, st);
}

__global__ void zstd_count_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, ZEntry* ze, uint32_t nz) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nz) return;
    const uint32_t e = ze[i].entry;
    uint32_t nb = 0;
    int32_t st = ST_OK;
    if (entries[e].status == ST_OK) st = scan_entry(buf, entries[e].comp_off, entries[e].comp_len, e, nullptr, &nb);
    set_status(entries, e, st);
    ze[i].blk_count = st == ST_OK && entries[e].status == ST_OK ? nb : 0;
}
// Re-walks the frames (the entry table is reset at the start of every run, so the status has to be
// re-established here too) and records the blocks of entries whose walk succeeded.
__global__ void zstd_fill_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, const ZEntry* ze, uint32_t nz,
                                 ZBlock* blocks) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nz) return;
    const uint32_t e = ze[i].entry;
    if (entries[e].status != ST_OK) return;
    uint32_t nb = 0;
    int32_t st = scan_entry(buf, entries[e].comp_off, entries[e].comp_len, e,
                            ze[i].blk_count ? blocks + ze[i].blk_begin : nullptr, &nb);
    set_status(entries, e, st);
}
__global__ void zstd_parse_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, ZBlock* blocks, uint32_t n_blocks) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) return;
    ZBlock b = blocks[i];
    b.status = parse_block(buf, b);
    blocks[i] = b;
    set_status(entries, b.entry, b.status);
}
__global__ void zstd_resolve_kernel(EntryRec* entries, ZEntry* ze, uint32_t nz, ZBlock* blocks) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nz) return;
    ZEntry z = ze[i];
    uint64_t lit = 0, seq = 0;
    if (entries[z.entry].status == ST_OK) {
        int32_t st = resolve_sources(blocks, z.blk_begin, z.blk_count);
        set_status(entries, z.entry, st);
        for (uint32_t k = z.blk_begin; k < z.blk_begin + z.blk_count; k++) {
            ZBlock& b = blocks[k];
            b.lit_off = lit;   // relative to the entry's bases
            b.seq_off = seq;
            if (b.type == BT_COMPRESSED) {
                if (b.lit_type >= LT_COMPRESSED) lit += (b.lit_regen + 15u) & ~15u;
                seq += b.nseq;
            }
        }
    }
    ze[i].lit_total = lit;
    ze[i].seq_total = seq;
}

// ------------------------------------------------------------------------------------------------
// Entropy stage: one CTA (64 threads) per block.  Warp 0 lane 0 builds the three FSE tables and
// decodes the sequences; warp 1 builds the Huffman table (lane 0) and decodes the 4 literal streams
// (lanes 0-3).  Both warps run concurrently.
struct EntropySmem {
    SeqEntry ll[512];
    SeqEntry of[256];
    SeqEntry ml[512];
    uint16_t huf[1 << HUF_LOG_MAX];
    uint8_t weights[260];
    FseEntry wfse[64];
    int16_t norm[64];
    uint16_t next_of[64];
    int huf_log;
    int huf_hdr;
    int lit_fail;
};

__global__ void __launch_bounds__(64) zstd_entropy_kernel(const uint8_t* __restrict__ buf, EntryRec* entries,
                                                          const ZEntry* __restrict__ ze_of_block_entry /*unused*/,
                                                          ZBlock* blocks, uint32_t n_blocks,
                                                          const uint64_t* __restrict__ lit_base_of_entry,
                                                          const uint64_t* __restrict__ seq_base_of_entry,
                                                          uint8_t* __restrict__ lits, uint32_t* __restrict__ sll,
                                                          uint32_t* __restrict__ sml, uint32_t* __restrict__ sof) {
    __shared__ EntropySmem S;
    const uint32_t bi = blockIdx.x;
    if (bi >= n_blocks) return;
    ZBlock& gb = blocks[bi];
    if (gb.type != BT_COMPRESSED || gb.status != ST_OK) return;
    if (entries[gb.entry].status != ST_OK) return;
    const uint32_t* words = reinterpret_cast<const uint32_t*>(buf);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0 && gb.nseq > 0) {
            ZBlock b = gb;
            int l0 = seq_table_for(buf, blocks, b, 0, S.ll, S.norm, S.next_of);
            int l1 = seq_table_for(buf, blocks, b, 1, S.of, S.norm, S.next_of);
            int l2 = seq_table_for(buf, blocks, b, 2, S.ml, S.norm, S.next_of);
            int32_t st = ST_INVALID_DATA;
            if (l0 >= 0 && l1 >= 0 && l2 >= 0) {
                const uint64_t so = seq_base_of_entry[b.entry] + b.seq_off;
                st = decode_sequences(words, buf, b, S.ll, S.of, S.ml, l0, l1, l2, sll + so, sml + so, sof + so);
            }
            if (st == ST_OK) {
                gb.out_size = b.out_size; gb.lit_used = b.lit_used;
                gb.rep_out[0] = b.rep_out[0]; gb.rep_out[1] = b.rep_out[1]; gb.rep_out[2] = b.rep_out[2];
            } else { gb.status = st; set_status(entries, b.entry, st); }
        }
        return;
    }
    // warp 1: literals
    if (gb.lit_type < LT_COMPRESSED) return;   // raw / RLE literals are read in place by the LZ stage
    if (lane == 0) {
        const ZBlock& hb = blocks[gb.huf_src];
        int hlog = 0;
        int hdr = huf_read_table(words, buf, hb.src + hb.lit_pos, hb.lit_csize, S.huf, &hlog, S.weights, S.wfse);
        S.huf_log = hlog; S.huf_hdr = hdr; S.lit_fail = hdr < 0 ? 1 : 0;
    }
    __syncwarp();
    bool ok = S.huf_hdr >= 0;
    const uint32_t skip = gb.lit_type == LT_COMPRESSED ? (uint32_t)(ok ? S.huf_hdr : 0) : 0;
    if (ok && skip > gb.lit_csize) ok = false;
    const uint64_t at = gb.src + gb.lit_pos + skip;
    const uint32_t clen = gb.lit_csize - (ok ? skip : 0);
    uint8_t* dst = lits + lit_base_of_entry[gb.entry] + gb.lit_off;
    if (ok) {
        if (gb.lit_streams == 1) {
            if (lane == 0) ok = huf_decode_stream(words, buf, at, clen, S.huf, S.huf_log, dst, gb.lit_regen);
        } else {
            uint32_t s1 = 0, s2 = 0, s3 = 0;
            if (clen < 6) ok = false;
            else { s1 = load_le16(buf + at); s2 = load_le16(buf + at + 2); s3 = load_le16(buf + at + 4); }
            if (ok && (uint64_t)s1 + s2 + s3 + 6 > clen) ok = false;
            const uint32_t seg = (gb.lit_regen + 3) / 4;
            if (ok && seg * 3 > gb.lit_regen) ok = false;
            if (ok && lane < 4) {
                uint32_t sz = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : clen - 6 - s1 - s2 - s3;
                uint64_t o = at + 6 + (lane > 0 ? s1 : 0) + (lane > 1 ? s2 : 0) + (lane > 2 ? s3 : 0);
                uint32_t cnt = lane < 3 ? seg : gb.lit_regen - 3 * seg;
                ok = huf_decode_stream(words, buf, o, sz, S.huf, S.huf_log, dst + (uint64_t)lane * seg, cnt);
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, !ok)) {
        if (lane == 0) { atomicCAS(&gb.status, ST_OK, ST_INVALID_DATA); set_status(entries, gb.entry, ST_INVALID_DATA); }
    }
}

__global__ void zstd_prefix_kernel(EntryRec* entries, const ZEntry* ze, uint32_t nz, ZBlock* blocks) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nz) return;
    const ZEntry z = ze[i];
    if (entries[z.entry].status != ST_OK) return;
    uint64_t total = 0;
    int32_t st = prefix_entry(blocks, z.blk_begin, z.blk_count, 0, &total);   // offsets relative to the entry
    set_status(entries, z.entry, st);
    entries[z.entry].out_len = total;
}

// ------------------------------------------------------------------------------------------------
// LZ execution.  One warp per entry, blocks in order, 32 sequences per step (lane = sequence).
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) {
    // bytes until dst is 16-aligned, then 16-byte stores with the source fetched by 4-byte words when possible
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    if (lane < (int)head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nv = n >> 4;
    if (((uintptr_t)src & 3) == 0) {
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint32_t* s = reinterpret_cast<const uint32_t*>(src) + i * 4;
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(s[0], s[1], s[2], s[3]);
        }
    } else {
        const uint32_t sh = (uint32_t)((uintptr_t)src & 3) * 8;
        const uint32_t* sa = reinterpret_cast<const uint32_t*>((uintptr_t)src & ~(uintptr_t)3);
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint32_t* s = sa + i * 4;
            uint32_t t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3], t4 = s[4];
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(__funnelshift_r(t0, t1, sh), __funnelshift_r(t1, t2, sh),
                                                           __funnelshift_r(t2, t3, sh), __funnelshift_r(t3, t4, sh));
        }
    }
    for (uint32_t i = nv * 16 + lane; i < n; i += 32) dst[i] = src[i];
}
__device__ __forceinline__ void warp_fill(uint8_t* dst, uint8_t v, uint32_t n, int lane) {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = v;
}
// overlapping-safe forward match copy done by the whole warp: dst[i] = dst[i - off], i in [0, n)
__device__ __forceinline__ void warp_match_copy(uint8_t* dst, uint32_t off, uint32_t n, int lane) {
    const uint8_t* src = dst - off;
    if (off >= n) { warp_copy(dst, src, n, lane); return; }
    if (off < 32) {   // periodic pattern: every byte comes from the first period, which is complete
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i % off];
        return;
    }
    for (uint32_t base = 0; base < n; base += 32) {   // off >= 32: a 32-byte step never reads what it writes
        uint32_t i = base + lane;
        uint8_t v = 0;
        if (i < n) v = src[i];
        if (i < n) dst[i] = v;
        __syncwarp();
    }
}

constexpr uint32_t LZ_SHORT = 32;   // per-lane copies up to this many bytes; longer ones are done by the whole warp

__global__ void __launch_bounds__(128) zstd_lz_kernel(const uint8_t* __restrict__ buf, EntryRec* entries,
                                                      const ZEntry* __restrict__ ze, uint32_t nz,
                                                      const ZBlock* __restrict__ blocks,
                                                      const uint8_t* __restrict__ lits, const uint32_t* __restrict__ sll,
                                                      const uint32_t* __restrict__ sml, const uint32_t* __restrict__ sof,
                                                      uint8_t* __restrict__ out) {
    const uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wi >= nz) return;
    const ZEntry z = ze[wi];
    EntryRec& er = entries[z.entry];
    if (er.status != ST_OK) return;
    if (er.out_len > er.out_cap) { if (lane == 0) atomicCAS(&er.status, ST_OK, ST_NOSPACE); return; }
    uint8_t* obase = out + er.out_off;
    int32_t fail = ST_OK;
    for (uint32_t k = z.blk_begin; k < z.blk_begin + z.blk_count && fail == ST_OK; k++) {
        const ZBlock& b = blocks[k];
        uint8_t* o = obase + b.out_off;
        if (b.type == BT_RAW) { warp_copy(o, buf + b.src, b.size, lane); __syncwarp(); continue; }
        if (b.type == BT_RLE) { warp_fill(o, buf[b.src], b.size, lane); __syncwarp(); continue; }
        const uint8_t* lit;
        uint32_t lstride = 1;
        if (b.lit_type == LT_RAW) lit = buf + b.src + b.lit_pos;
        else if (b.lit_type == LT_RLE) { lit = buf + b.src + b.lit_pos; lstride = 0; }
        else lit = lits + z.lit_base + b.lit_off;
        const uint64_t so = z.seq_base + b.seq_off;
        const uint32_t rep_in[3] = {b.rep_in[0], b.rep_in[1], b.rep_in[2]};
        const uint64_t frame_dist = b.out_off - b.frame_out;   // bytes of this frame before the block
        uint32_t op = 0, lp = 0;
        for (uint32_t base = 0; base < b.nseq; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < b.nseq;
            uint32_t ll = 0, ml = 0, off = 1;
            if (valid) { ll = sll[so + i]; ml = sml[so + i]; off = resolve_rep(sof[so + i], rep_in); }
            const uint32_t tot = ll + ml;
            const uint32_t incl = warp_incl_scan(tot, lane), lincl = warp_incl_scan(ll, lane);
            const uint32_t dst_lit = op + incl - tot, src_lit = lp + lincl - ll;
            const uint32_t dst_m = dst_lit + ll;
            const bool bad = valid && (off == 0 || (uint64_t)off > frame_dist + dst_m);
            if (__any_sync(0xFFFFFFFFu, bad)) { fail = ST_INVALID_DATA; break; }
            // ---- literals: short runs per lane, long runs by the whole warp
            {
                const uint32_t n = ll < LZ_SHORT ? ll : LZ_SHORT;
                for (uint32_t q = 0; q < n; q++) o[dst_lit + q] = lit[(size_t)(src_lit + q) * lstride];
                uint32_t longm = __ballot_sync(0xFFFFFFFFu, ll > LZ_SHORT);
                while (longm) {
                    const int src_lane = __ffs(longm) - 1;
                    longm &= longm - 1;
                    const uint32_t d = __shfl_sync(0xFFFFFFFFu, dst_lit, src_lane), s = __shfl_sync(0xFFFFFFFFu, src_lit, src_lane),
                                   n2 = __shfl_sync(0xFFFFFFFFu, ll, src_lane);
                    if (lstride) warp_copy(o + d + LZ_SHORT, lit + s + LZ_SHORT, n2 - LZ_SHORT, lane);
                    else warp_fill(o + d + LZ_SHORT, lit[0], n2 - LZ_SHORT, lane);
                }
            }
            __syncwarp();
            // ---- matches: wavefront.  Everything before the first pending lane's match start is complete;
            // a pending lane may run once its source lies below that frontier (its own overlap is fine).
            uint32_t pending = __ballot_sync(0xFFFFFFFFu, valid && ml > 0);
            const int64_t src_m = (int64_t)dst_m - (int64_t)off;     // relative to o, may be negative (earlier blocks)
            const int64_t src_need = src_m + (int64_t)(ml < off ? ml : off);   // exclusive end of what must exist
            bool mine = valid && ml > 0;
            while (pending) {
                const int first = __ffs(pending) - 1;
                const uint32_t frontier = __shfl_sync(0xFFFFFFFFu, dst_m, first);
                const uint32_t first_ml = __shfl_sync(0xFFFFFFFFu, ml, first);
                if (first_ml > LZ_SHORT) {   // long match at the frontier: whole warp
                    const uint32_t foff = __shfl_sync(0xFFFFFFFFu, off, first);
                    warp_match_copy(o + frontier, foff, first_ml, lane);
                    if (lane == first) mine = false;
                } else {
                    const bool go = mine && ml <= LZ_SHORT && (lane == first || src_need <= (int64_t)frontier);
                    if (go) {
                        uint8_t* d = o + dst_m;
                        const uint8_t* s = o + src_m;
                        for (uint32_t q = 0; q < ml; q++) d[q] = s[q];
                        mine = false;
                    }
                }
                __syncwarp();
                pending = __ballot_sync(0xFFFFFFFFu, mine);
            }
            op += __shfl_sync(0xFFFFFFFFu, incl, 31);
            lp += __shfl_sync(0xFFFFFFFFu, lincl, 31);
        }
        if (fail != ST_OK) break;
        // trailing literals
        if (lp > b.lit_regen || op + (b.lit_regen - lp) != b.out_size) { fail = ST_INVALID_DATA; break; }
        if (lstride) warp_copy(o + op, lit + lp, b.lit_regen - lp, lane);
        else warp_fill(o + op, lit[0], b.lit_regen - lp, lane);
        __syncwarp();
    }
    if (fail != ST_OK && lane == 0) atomicCAS(&er.status, ST_OK, fail);
}

}  // namespace zs
}  // namespace pna
