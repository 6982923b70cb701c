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
constexpr uint32_t PJ_EXPAND_X = 64;                        // CTAs along a block's chunks (each takes chunks x, x + 64, ...)

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

// literals to their final place, match bytes to their pointers.  grid = (PJ_EXPAND_X, blocks of the segment), 8 warps per CTA.
__global__ void __launch_bounds__(256) pj_expand_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, const ZEntry* __restrict__ ze,
                                                        const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                        const uint8_t* __restrict__ lits, const SeqRec* __restrict__ seqs,
                                                        const uint2* __restrict__ cpos, uint8_t* out, int32_t* __restrict__ ptr) {
    const PjSeg sg = segs[seg_index];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t bl = blockIdx.y;
    if (bl >= sg.blk_count) return;
    const ZEntry* zp = ze + sg.ze;
    EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;
    const ZBlock& b = blocks[sg.blk_begin + bl];
    const uint64_t seg0 = blocks[sg.blk_begin].out_off;                 // entry-relative position of the segment's first byte
    uint8_t* const obase = out + er.out_off;                            // entry-relative positions index this
    const uint64_t bpos = b.out_off;                                    // this block's first byte
    const uint32_t brel = (uint32_t)(bpos - seg0);                      // segment-relative (< 2^27)
    if (b.type == BT_RAW || b.type == BT_RLE) {                         // literal bytes only
        if (blockIdx.x != 0) return;
        const uint8_t* src = buf + b.src;
        for (uint32_t i = tid; i < b.size; i += 256) { obase[bpos + i] = b.type == BT_RAW ? src[i] : src[0]; ptr[brel + i] = (int32_t)(brel + i); }
        return;
    }
    const uint8_t* lit;
    uint32_t lstride = 1;
    if (b.lit_type == LT_RAW) lit = buf + b.src + b.lit_pos;
    else if (b.lit_type == LT_RLE) { lit = buf + b.src + b.lit_pos; lstride = 0; }
    else lit = lits + zp->lit_base + b.lit_off;
    const uint32_t nseq = b.nseq, nch = (nseq + 31) >> 5;
    const SeqRec* sq = seqs + zp->seq_base + b.seq_off;
    const uint2* cp = cpos + (size_t)bl * PJ_MAX_CHUNKS;
    const uint32_t rep_in[3] = {b.rep_in[0], b.rep_in[1], b.rep_in[2]};
    const uint64_t fd64 = bpos - b.frame_out;                           // bytes of this frame in front of the block
    bool bad = false;
    for (uint32_t ch = blockIdx.x * 8u + warp; ch < nch; ch += PJ_EXPAND_X * 8u) {
        const uint32_t i = ch * 32u + lane;
        uint32_t ll = 0, ml = 0, off = 1;
        if (i < nseq) { const SeqRec r = sq[i]; pj_seq_lengths(b, r, i, ll, ml); off = resolve_rep(r.x, rep_in); }
        const uint2 base = cp[ch];
        const uint32_t incl = warp_incl_scan(ll + ml, (int)lane);        // bytes up to and including this sequence
        const uint32_t lincl = warp_incl_scan(ll, (int)lane);
        const uint32_t start = base.y + incl - (ll + ml);                // block-relative first byte of this sequence
        const uint32_t lstart = base.x + lincl - ll;                     // its first literal
        if (i < nseq && ml) bad = bad || off == 0 || (uint64_t)off > fd64 + start + ll || off > 0x7FFFFFFFu;
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
#pragma unroll 4
        for (uint32_t q0 = 0; q0 < total; q0 += 32) {
            const uint32_t q = q0 + lane;
            // the sequence that holds byte q: the first lane whose inclusive sum exceeds q
            uint32_t lo = 0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (int)(lo + s - 1));
                if (v <= q) lo += s;
            }
            lo = lo > 31u ? 31u : lo;
            const uint32_t s_start = __shfl_sync(0xFFFFFFFFu, start, (int)lo), s_ll = __shfl_sync(0xFFFFFFFFu, ll, (int)lo);
            const uint32_t s_off = __shfl_sync(0xFFFFFFFFu, off, (int)lo), s_ls = __shfl_sync(0xFFFFFFFFu, lstart, (int)lo);
            if (q < total) {
                const uint32_t p = base.y + q;                            // block-relative position of this byte
                const uint32_t r = p - s_start;                           // index inside the sequence
                if (p < BLOCK_MAX) {
                    if (r < s_ll) {
                        const uint32_t li = s_ls + r;
                        obase[bpos + p] = li < b.lit_regen ? lit[(size_t)li * lstride] : 0;
                        ptr[brel + p] = (int32_t)(brel + p);
                    } else ptr[brel + p] = (int32_t)((int64_t)brel + p - (int64_t)s_off);   // negative: in front of the segment
                }
            }
        }
    }
    if (bad) atomicCAS(&er.status, ST_OK, ST_INVALID_DATA);
    // the literals behind the last sequence (the whole block when it has no sequences)
    if (blockIdx.x == 0) {
        uint32_t lit_used = b.lit_used, tot = b.out_size - (b.lit_regen - (b.lit_used > b.lit_regen ? b.lit_regen : b.lit_used));
        if (nseq == 0) { lit_used = 0; tot = 0; }
        for (uint32_t i = tid; lit_used + i < b.lit_regen && tot + i < BLOCK_MAX; i += 256) {
            obase[bpos + tot + i] = lit[(size_t)(lit_used + i) * lstride];
            ptr[brel + tot + i] = (int32_t)(brel + tot + i);
        }
    }
}

// one round of pointer jumping over n pointers; flags[round] != 0 afterwards iff some pointer was not at a root when the
// round began.  A launch whose predecessor found nothing to do returns at once.
__global__ void __launch_bounds__(256) pj_jump_kernel(const EntryRec* __restrict__ entries, const ZEntry* __restrict__ ze,
                                                      const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                      int32_t* ptr, uint32_t* flags, int round) {
    if (round > 0 && flags[round - 1] == 0) return;
    const PjSeg sg = segs[seg_index];
    const ZEntry* zp = ze + sg.ze;
    const EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;   // (too small an output: nothing was expanded)
    const ZBlock& last = blocks[sg.blk_begin + sg.blk_count - 1];
    const uint32_t n = (uint32_t)(last.out_off + last.out_size - blocks[sg.blk_begin].out_off);
    bool open = false;
    for (uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u; i0 < n; i0 += gridDim.x * 1024u) {
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
    if (open) flags[round] = 1;
}

// First pass over the pointers: an ascending sweep (every CTA walks its grid-stride positions upwards, so the grid as a whole
// moves through the segment in windows of gridDim.x KiB) that follows each chain for up to PJ_CHASE_HOPS hops.  Sources are
// mostly recent bytes (the codec's window), i.e. pointers that this very sweep has just resolved: a typical chain ends at a
// root after two hops, and ONE pass leaves almost nothing for the doubling rounds that follow.  Any pointer read on the way
// is valid whether or not another thread has already shortened it (pointers only ever move towards their root).
constexpr int PJ_CHASE_HOPS = 6;
__global__ void __launch_bounds__(256) pj_chase_kernel(const EntryRec* __restrict__ entries, const ZEntry* __restrict__ ze,
                                                       const PjSeg* __restrict__ segs, uint32_t seg_index, const ZBlock* __restrict__ blocks,
                                                       int32_t* ptr) {
    const PjSeg sg = segs[seg_index];
    const ZEntry* zp = ze + sg.ze;
    const EntryRec& er = entries[zp->entry];
    if (er.status != ST_OK || er.out_len > er.out_cap) return;
    const ZBlock& last = blocks[sg.blk_begin + sg.blk_count - 1];
    const uint32_t n = (uint32_t)(last.out_off + last.out_size - blocks[sg.blk_begin].out_off);
    for (uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 4u; i0 < n; i0 += gridDim.x * 1024u) {
        int32_t v[4];
        if (i0 + 4 <= n) { const int4 t = *reinterpret_cast<const int4*>(ptr + i0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else for (int k = 0; k < 4; k++) v[k] = i0 + k < n ? ptr[i0 + k] : -1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int32_t p = v[k];
            if (p < 0 || (uint32_t)p == i0 + k) continue;
            int32_t r = p;
            for (int h = 0; h < PJ_CHASE_HOPS; h++) {
                const int32_t w = ptr[r];
                if (w == r) break;                 // a literal: r is the root
                r = w;
                if (w < 0) break;                  // left the segment: root
            }
            if (r != p) ptr[i0 + k] = r;
        }
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
