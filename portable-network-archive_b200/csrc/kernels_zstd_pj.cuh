// kernels_zstd_pj.cuh -- block-parallel LZ execution INSIDE one zstd frame (BASELINE config 5: a reference-written solid
// archive is ONE frame over every file, lib/src/entry.rs:567-583 decodes it on one thread).
//
// The window makes a frame's blocks order dependent, so the CTA-per-frame kernel (kernels_zstd_lz.cuh) leaves a 4 GiB frame
// on a single CTA.  Here every output byte gets a POINTER instead: a literal points at itself, a match byte at
// position - offset.  Chains of copies are then shortened by pointer jumping (ptr[i] <- ptr[ptr[i]]), which halves every
// chain per round regardless of how the copies nest -- log2(depth) fully parallel, HBM/L2-bound passes.  The frame is taken
// in segments of PJ_SEG_BLOCKS blocks (<= 128 MiB of output; pointers are 32-bit, relative to the segment start): bytes in
// front of the segment are final already (earlier segments), so a pointer that leaves the segment is a root as well.
//
//   pj_scan     warp per block    prefix sums of (literal length, total length) at every 32nd sequence; length checks
//   pj_expand   warp per 32 sequences: literals go to their final place, match bytes get their pointer (byte parallel)
//   pj_chase    thread per 4 bytes, one ascending sweep that follows every chain a few hops (resolves almost everything)
//   pj_jump     thread per 4 bytes, repeated until a round finds every pointer at a root (the launches that follow exit at once)
//   pj_gather   thread per 4 bytes: every non-root byte copies its root
//
// Algorithmic bytes: 8*nseq + L + U like the CTA-per-frame kernel; the pointer passes add 4 B per output byte per round.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "zstd_core.cuh"
#include "kernels_zstd_lz.cuh"   // ZEntry, warp_incl_scan

namespace pna {
namespace zs {

constexpr uint32_t PJ_SEG_BLOCKS = 1024;                    // blocks per segment: <= 128 MiB of output
constexpr uint32_t PJ_MAX_CHUNKS = 4096;                    // 32-sequence chunks per block: a block holds < 2^17 sequences
constexpr int PJ_MAX_ROUNDS = 40;
constexpr uint32_t PJ_EXPAND_X = 32;                        // CTAs along a block's bytes (8 warps x PJ_WARP_BYTES each)

struct PjSeg {               // one segment of one frame, device resident
    uint32_t ze;             // ZEntry
    uint32_t blk_begin, blk_count;
    uint32_t _pad;
};

__device__ __forceinline__ void pj_seq_lengths(const ZBlock& b, const SeqRec& r, uint32_t i, uint32_t& ll, uint32_t& ml) {
    ll = r.y & 0xFFFFu; ml = r.y >> 16;
    if (ll == SEQ_ESC || ml == SEQ_ESC)
        for (uint32_t q = 0; q < b.esc_n && q < (uint32_t)SEQ_ESC_MAX; q++)
            if (b.esc_idx[q] == i) { ll = b.esc_ll[q]; ml = b.esc_ml[q]; }
}

// prefix of (literals consumed, bytes produced) in front of every 32-sequence chunk of every compressed block of the segment
__global__ void __launch_bounds__(256) pj_scan_kernel(EntryRec* entries, const ZEntry* __restrict__ ze, const PjSeg* __restrict__ segs,
                                                      uint32_t seg_index, ZBlock* blocks, const SeqRec* __restrict__ seqs,
                                                      uint2* __restrict__ cpos /* [blk_local][PJ_MAX_CHUNKS] */) {
    const PjSeg sg = segs[seg_index];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t bl = blockIdx.x * 8u + warp;
    if (bl >= sg.blk_count) return;
    const ZEntry* zp = ze + sg.ze;
    EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK) return;
    ZBlock& b = blocks[sg.blk_begin + bl];
    if (b.type != BT_COMPRESSED) return;
    const uint32_t nseq = b.nseq;
    const SeqRec* sq = seqs + zp->seq_base + b.seq_off;
    uint2* cp = cpos + (size_t)bl * PJ_MAX_CHUNKS;
    uint32_t lit = 0, tot = 0;
    for (uint32_t c0 = 0; c0 < nseq && tot <= BLOCK_MAX; c0 += 128) {   // four chunks per round: four loads in flight per lane
        uint32_t a[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = c0 + 32u * k + lane;
            uint32_t ll = 0, ml = 0;
            if (i < nseq) pj_seq_lengths(b, sq[i], i, ll, ml);
            a[k] = ll; t[k] = ll + ml;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) { a[k] += __shfl_xor_sync(0xFFFFFFFFu, a[k], o); t[k] += __shfl_xor_sync(0xFFFFFFFFu, t[k], o); }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (c0 + 32u * k < nseq && lane == 0) cp[(c0 >> 5) + k] = make_uint2(lit, tot);
            lit += a[k]; tot += t[k];
        }
    }
    if (lane == 0) {
        // the same checks the CTA-per-frame kernel makes while it runs: literals not overrun, block size as announced
        const bool bad = lit > b.lit_regen || (uint64_t)tot + (b.lit_regen - (lit > b.lit_regen ? b.lit_regen : lit)) != b.out_size;
        if (bad) atomicCAS(&er.status, ST_OK, ST_INVALID_DATA);
    }
}

// Literals to their final place, match bytes to their pointers.  The work is cut by OUTPUT BYTES, not by sequences: every warp
// takes PJ_WARP_BYTES consecutive bytes of a block (grid = (PJ_EXPAND_X, blocks of the segment), 8 warps per CTA), finds the
// 32-sequence chunk its first byte falls into (warp-parallel search over the chunk prefix sums of pj_scan) and walks on from
// there -- a 64 KiB literal run or a long match is shared by many warps instead of serialising one.  The literals behind the
// last sequence, and raw / RLE blocks, are one more all-literal "sequence".
constexpr uint32_t PJ_WARP_BYTES = 512;
static_assert(PJ_WARP_BYTES * 8u * PJ_EXPAND_X >= BLOCK_MAX, "the grid covers a whole block");
__global__ void __launch_bounds__(256) pj_expand_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, const ZEntry* __restrict__ ze,
                                                        const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                        const uint8_t* __restrict__ lits, const SeqRec* __restrict__ seqs,
                                                        const uint2* __restrict__ cpos, uint8_t* out, int32_t* __restrict__ ptr) {
    const PjSeg sg = segs[seg_index];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t bl = blockIdx.y;
    if (bl >= sg.blk_count) return;
    const ZEntry* zp = ze + sg.ze;
    EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;
    const ZBlock& b = blocks[sg.blk_begin + bl];
    const uint32_t w0 = (blockIdx.x * 8u + warp) * PJ_WARP_BYTES;      // this warp's bytes of the block: [w0, w1)
    if (w0 >= b.out_size) return;
    const uint32_t w1 = w0 + PJ_WARP_BYTES < b.out_size ? w0 + PJ_WARP_BYTES : b.out_size;
    const uint64_t seg0 = blocks[sg.blk_begin].out_off;                 // entry-relative position of the segment's first byte
    uint8_t* const obase = out + er.out_off;                            // entry-relative positions index this
    const uint64_t bpos = b.out_off;                                    // this block's first byte
    const uint32_t brel = (uint32_t)(bpos - seg0);                      // segment-relative (< 2^27)
    const bool compressed = b.type == BT_COMPRESSED;
    const uint8_t* lit;
    uint32_t lstride = 1;
    if (!compressed) { lit = buf + b.src; lstride = b.type == BT_RAW ? 1u : 0u; }
    else if (b.lit_type == LT_RAW) lit = buf + b.src + b.lit_pos;
    else if (b.lit_type == LT_RLE) { lit = buf + b.src + b.lit_pos; lstride = 0; }
    else lit = lits + zp->lit_base + b.lit_off;
    const uint32_t nseq = compressed ? b.nseq : 0u, nch = (nseq + 31) >> 5;
    const uint32_t lit_total = compressed ? b.lit_regen : b.out_size;
    const uint32_t lit_used = nseq ? (b.lit_used > lit_total ? lit_total : b.lit_used) : 0u;
    const uint32_t seq_bytes = b.out_size - (lit_total - lit_used);     // bytes the sequences produce; the rest are trailing literals
    const SeqRec* sq = seqs + zp->seq_base + b.seq_off;
    const uint2* cp = cpos + (size_t)bl * PJ_MAX_CHUNKS;
    const uint32_t rep_in[3] = {b.rep_in[0], b.rep_in[1], b.rep_in[2]};
    const uint64_t fd64 = bpos - b.frame_out;                           // bytes of this frame in front of the block
    // chunk that holds byte w0: the last chunk whose start is <= w0 (chunk nch = the trailing literals, starting at seq_bytes)
    uint32_t ch = nch;
    if (w0 < seq_bytes) {
        uint32_t lo = 0, span = nch;
        while (span > 1) {
            const uint32_t step = (span + 31u) >> 5;
            const uint32_t c = lo + lane * step;
            const uint32_t v = (c < lo + span) ? cp[c].y : 0xFFFFFFFFu;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, v <= w0);     // lane 0 always qualifies (cp[lo].y <= w0)
            const uint32_t k = 31u - (uint32_t)__clz((int)(m | 1u));
            lo += k * step;
            span = (lo + step <= nch ? step : nch - lo);
        }
        ch = lo;
    }
    bool bad = false;
    for (; ch <= nch; ch++) {
        uint32_t ll = 0, ml = 0, off = 1;
        uint2 base;
        if (ch < nch) {
            const uint32_t i = ch * 32u + lane;
            if (i < nseq) { const SeqRec r = sq[i]; pj_seq_lengths(b, r, i, ll, ml); off = resolve_rep(r.x, rep_in); }
            base = cp[ch];
        } else { base = make_uint2(lit_used, seq_bytes); if (lane == 0) ll = lit_total - lit_used; }
        if (base.y >= w1) break;
        const uint32_t incl = warp_incl_scan(ll + ml, (int)lane);        // bytes up to and including this sequence
        const uint32_t lincl = warp_incl_scan(ll, (int)lane);
        const uint32_t start = base.y + incl - (ll + ml);                // block-relative first byte of this sequence
        const uint32_t lstart = base.x + lincl - ll;                     // its first literal
        if (ml) bad = bad || off == 0 || (uint64_t)off > fd64 + start + ll || off > 0x7FFFFFFFu;
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t qa = w0 > base.y ? w0 - base.y : 0u;              // this warp's part of the chunk: [qa, qb)
        const uint32_t qb = w1 - base.y < total ? w1 - base.y : total;
#pragma unroll 2
        for (uint32_t q0 = qa; q0 < qb; q0 += 32) {
            const uint32_t q = q0 + lane;
            // the sequence that holds byte q: the first lane whose inclusive sum exceeds q
            uint32_t lo = 0;
#pragma unroll
            for (int s2 = 16; s2 > 0; s2 >>= 1) {
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (int)(lo + s2 - 1));
                if (v <= q) lo += s2;
            }
            lo = lo > 31u ? 31u : lo;
            const uint32_t s_start = __shfl_sync(0xFFFFFFFFu, start, (int)lo), s_ll = __shfl_sync(0xFFFFFFFFu, ll, (int)lo);
            const uint32_t s_off = __shfl_sync(0xFFFFFFFFu, off, (int)lo), s_ls = __shfl_sync(0xFFFFFFFFu, lstart, (int)lo);
            if (q < qb) {
                const uint32_t p = base.y + q;                            // block-relative position of this byte
                const uint32_t r = p - s_start;                           // index inside the sequence
                if (r < s_ll) {
                    const uint32_t li = s_ls + r;
                    obase[bpos + p] = li < lit_total ? lit[(size_t)li * lstride] : 0;
                    ptr[brel + p] = (int32_t)(brel + p);
                } else ptr[brel + p] = (int32_t)((int64_t)brel + p - (int64_t)s_off);   // negative: in front of the segment
            }
        }
    }
    if (bad) atomicCAS(&er.status, ST_OK, ST_INVALID_DATA);
}

// One round of pointer jumping.  The segment is walked in tiles of 1024 pointers (one CTA iteration each); a tile whose
// pointers were all at a root when the previous pass left it is skipped -- deep chains cluster (runs, repeated records), so after
// the chase sweep a round touches a few tiles instead of re-reading half a gigabyte of pointers.  tiles_in / tiles_out: one
// byte per tile, ping-pong between rounds; flags[round] != 0 afterwards iff some pointer was still not at a root.  A launch
// whose predecessor found nothing to do returns at once.
__global__ void __launch_bounds__(256) pj_jump_kernel(const EntryRec* __restrict__ entries, const ZEntry* __restrict__ ze,
                                                      const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                      int32_t* ptr, uint32_t* flags, int round, const uint8_t* __restrict__ tiles_in,
                                                      uint8_t* __restrict__ tiles_out) {
    if (round > 0 && flags[round - 1] == 0) return;
    const PjSeg sg = segs[seg_index];
    const ZEntry* zp = ze + sg.ze;
    const EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;   // (too small an output: nothing was expanded)
    const ZBlock& last = blocks[sg.blk_begin + sg.blk_count - 1];
    const uint32_t n = (uint32_t)(last.out_off + last.out_size - blocks[sg.blk_begin].out_off);
    bool any_open = false;
    for (uint32_t tile = blockIdx.x; (uint64_t)tile * 1024u < n; tile += gridDim.x) {
        if (tiles_in[tile] == 0) { if (threadIdx.x == 0) tiles_out[tile] = 0; continue; }   // uniform per CTA
        const uint32_t i0 = tile * 1024u + threadIdx.x * 4u;
        bool open = false;
        if (i0 < n) {
            int32_t v[4];
            if (i0 + 4 <= n) { const int4 t = *reinterpret_cast<const int4*>(ptr + i0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
            else for (int k = 0; k < 4; k++) v[k] = i0 + k < n ? ptr[i0 + k] : -1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int32_t p = v[k];
                if (p < 0 || (uint32_t)p == i0 + k) continue;      // root: in front of the segment / a literal
                const int32_t w = ptr[p];                          // p < i0 + k: a match points backwards
                if (w == p) continue;                              // points at a literal: final
                ptr[i0 + k] = w;                                   // jump (w may leave the segment: root)
                open = true;
            }
        }
        const int tile_open = __syncthreads_or(open);
        if (threadIdx.x == 0) tiles_out[tile] = tile_open ? 1 : 0;
        any_open = any_open || tile_open;
    }
    if (any_open && threadIdx.x == 0) flags[round] = 1;
}

// First pass over the pointers: an ascending sweep (every CTA walks its tiles upwards, so the grid as a whole moves through the
// segment in windows of gridDim.x KiB) that follows each chain for up to PJ_CHASE_HOPS hops.  Sources are mostly recent bytes
// (the codec's window), i.e. pointers that this very sweep has just resolved: a typical chain ends at a root after two hops,
// and ONE pass leaves almost nothing for the doubling rounds that follow.  Any pointer read on the way is valid whether or
// not another thread has already shortened it (pointers only ever move towards their root).  tiles_out[tile] = 1 where a
// chain was cut off by the hop limit.
constexpr int PJ_CHASE_HOPS = 6;
__global__ void __launch_bounds__(256) pj_chase_kernel(const EntryRec* __restrict__ entries, const ZEntry* __restrict__ ze,
                                                       const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                       int32_t* ptr, uint8_t* __restrict__ tiles_out) {
    const PjSeg sg = segs[seg_index];
    const ZEntry* zp = ze + sg.ze;
    const EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;
    const ZBlock& last = blocks[sg.blk_begin + sg.blk_count - 1];
    const uint32_t n = (uint32_t)(last.out_off + last.out_size - blocks[sg.blk_begin].out_off);
    for (uint32_t tile = blockIdx.x; (uint64_t)tile * 1024u < n; tile += gridDim.x) {
        const uint32_t i0 = tile * 1024u + threadIdx.x * 4u;
        bool open = false;
        if (i0 < n) {
            int32_t v[4];
            if (i0 + 4 <= n) { const int4 t = *reinterpret_cast<const int4*>(ptr + i0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
            else for (int k = 0; k < 4; k++) v[k] = i0 + k < n ? ptr[i0 + k] : -1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int32_t p = v[k];
                if (p < 0 || (uint32_t)p == i0 + k) continue;
                int32_t r = p;
                bool at_root = false;
                for (int h = 0; h < PJ_CHASE_HOPS; h++) {
                    const int32_t w = ptr[r];
                    if (w == r) { at_root = true; break; }      // a literal: r is the root
                    r = w;
                    if (w < 0) { at_root = true; break; }       // left the segment: root
                }
                if (r != p) ptr[i0 + k] = r;
                open = open || !at_root;
            }
        }
        const int tile_open = __syncthreads_or(open);
        if (threadIdx.x == 0) tiles_out[tile] = tile_open ? 1 : 0;
    }
}

// every byte that is not its own root copies the root's byte
__global__ void __launch_bounds__(256) pj_gather_kernel(EntryRec* entries, const ZEntry* __restrict__ ze, const PjSeg* __restrict__ segs,
                                                        uint32_t seg_index, const ZBlock* __restrict__ blocks, const int32_t* __restrict__ ptr,
                                                        const uint32_t* __restrict__ flags, uint8_t* out) {
    const PjSeg sg = segs[seg_index];
    const ZEntry* zp = ze + sg.ze;
    EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK) return;
    if (er.out_len > er.out_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(&er.status, ST_OK, ST_NOSPACE); return; }
    if (flags[PJ_MAX_ROUNDS - 1] != 0) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(&er.status, ST_OK, ST_INTERNAL); return; }
    const uint64_t seg0 = blocks[sg.blk_begin].out_off;
    const ZBlock& last = blocks[sg.blk_begin + sg.blk_count - 1];
    const uint32_t n = (uint32_t)(last.out_off + last.out_size - seg0);
    uint8_t* const o = out + er.out_off + seg0;
    for (uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u; i0 < n; i0 += gridDim.x * 1024u) {
        for (int k = 0; k < 4; k++) {
            const uint32_t i = i0 + k;
            if (i >= n) break;
            const int32_t p = ptr[i];
            if ((uint32_t)p != i) o[i] = o[(int64_t)p];         // p < 0: a final byte of an earlier segment
        }
    }
}

}  // namespace zs
}  // namespace pna
