// kernels_zstd.cuh -- sm_100a kernels for zstd frame decode (K4).  See zstd_core.cuh for the format
// logic and DESIGN.md for the phase plan:
//   zstd_count/fill  (thread per entry)  frame + block header walk
//   zstd_parse       (thread per block)  literal/sequence section headers, table description offsets
//   zstd_resolve     (thread per entry)  Repeat_Mode / treeless sources, per-entry literal+sequence offsets
//   zstd_order       (one CTA)           counting sort of the blocks by sequence / literal count (load balance)
//   zstd_seq         (LANE per block)    FSE sequence decode; 32 blocks per warp, 16-bit tables interleaved in smem
//   zstd_lit         (lane per stream)   Huffman literals, 8 blocks x 4 streams per warp, tables in smem
//   zstd_prefix      (thread per entry)  output offsets, absolute repeat-offset history per block
//   zstd_lz          (warp per entry)    LZ execution in a shared-memory window: 32 sequences per step with exact
//                                        dependency wavefronts, far sources prefetched one step ahead, 512 B flushes
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "zstd_core.cuh"
#include "kernels_crc_cipher.cuh"   // load16_any

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
// Block ordering: blocks with sequences sorted by descending sequence count (so the 32 lanes of a warp
// of zstd_seq_kernel finish together), blocks with Huffman literals by descending literal count.
// counts[0] = sequence blocks, counts[1] = literal blocks, counts[2..3] = work counters (zeroed here).
__global__ void __launch_bounds__(1024) zstd_order_kernel(const EntryRec* __restrict__ entries, const ZBlock* __restrict__ blocks,
                                                          uint32_t n_blocks, uint32_t* __restrict__ seq_order,
                                                          uint32_t* __restrict__ lit_order, uint32_t* __restrict__ counts) {
    __shared__ uint32_t hs[1024], hl[1024];
    __shared__ uint32_t carry[2];
    const uint32_t tid = threadIdx.x;
    hs[tid] = 0; hl[tid] = 0;
    __syncthreads();
    auto keys = [&](uint32_t i, uint32_t& ks, uint32_t& kl) {
        const ZBlock& b = blocks[i];
        ks = kl = 0xFFFFFFFFu;
        if (b.type != BT_COMPRESSED || b.status != ST_OK || entries[b.entry].status != ST_OK) return;
        if (b.nseq > 0) { uint32_t k = b.nseq >> 6; ks = 1023u - (k > 1023u ? 1023u : k); }
        if (b.lit_type >= LT_COMPRESSED) { uint32_t k = b.lit_regen >> 7; kl = 1023u - (k > 1023u ? 1023u : k); }
    };
    for (uint32_t i = tid; i < n_blocks; i += 1024) {
        uint32_t ks, kl;
        keys(i, ks, kl);
        if (ks != 0xFFFFFFFFu) atomicAdd(&hs[ks], 1u);
        if (kl != 0xFFFFFFFFu) atomicAdd(&hl[kl], 1u);
    }
    __syncthreads();
    // exclusive scan of both histograms (1024 bins, one per thread): warp scan + warp totals
    uint32_t vs = hs[tid], vl = hl[tid];
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t is = vs, il = vl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xFFFFFFFFu, is, o), b = __shfl_up_sync(0xFFFFFFFFu, il, o);
        if (lane >= o) { is += a; il += b; }
    }
    __shared__ uint32_t ws[32], wl[32];
    if (lane == 31) { ws[warp] = is; wl[warp] = il; }
    __syncthreads();
    if (warp == 0) {
        uint32_t a = ws[lane], b = wl[lane], ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t x = __shfl_up_sync(0xFFFFFFFFu, ia, o), y = __shfl_up_sync(0xFFFFFFFFu, ib, o);
            if (lane >= o) { ia += x; ib += y; }
        }
        ws[lane] = ia - a; wl[lane] = ib - b;
        if (lane == 31) { carry[0] = ia; carry[1] = ib; }
    }
    __syncthreads();
    hs[tid] = ws[warp] + is - vs;
    hl[tid] = wl[warp] + il - vl;
    __syncthreads();
    for (uint32_t i = tid; i < n_blocks; i += 1024) {
        uint32_t ks, kl;
        keys(i, ks, kl);
        if (ks != 0xFFFFFFFFu) seq_order[atomicAdd(&hs[ks], 1u)] = i;
        if (kl != 0xFFFFFFFFu) lit_order[atomicAdd(&hl[kl], 1u)] = i;
    }
    if (tid == 0) { counts[0] = carry[0]; counts[1] = carry[1]; counts[2] = 0; counts[3] = 0; counts[4] = 0; }
}

// ------------------------------------------------------------------------------------------------
// Sequence stage: ONE LANE per block.  The FSE chain (table lookup -> bit counts -> next state) is serial
// per block, so throughput = blocks in flight / chain latency, and blocks in flight = shared memory /
// table bytes.  16-bit cells (zstd_core.cuh Tab16) make a block's three tables 2.5 KB: 88 blocks fill the
// SM's 227 KB.  One CTA of 4 warps per SM (one warp per scheduler), 22 active lanes per warp, each warp's
// 22 tables interleaved by lane; the warps take batches of 22 blocks independently (no CTA barrier).
constexpr uint32_t SEQ_LANES = 22, SEQ_WARPS = 4;
constexpr uint32_t SEQ_SMEM_BYTES = TAB16_TOTAL * SEQ_LANES * SEQ_WARPS * sizeof(uint16_t);   // 225280
__global__ void __launch_bounds__(32 * SEQ_WARPS) zstd_seq_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, ZBlock* blocks,
                                                                  const uint32_t* __restrict__ order, uint32_t* counts,
                                                                  const uint64_t* __restrict__ seq_base_of_entry,
                                                                  SeqRec* __restrict__ seqs) {
    extern __shared__ uint16_t stab_all[];
    __shared__ uint32_t s_llb[36], s_mlb[53];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 36; c += blockDim.x) s_llb[c] = ll_base(c);
    for (int c = threadIdx.x; c < 53; c += blockDim.x) s_mlb[c] = ml_base(c);
    __syncthreads();
    const uint32_t n = counts[0];
    const uint32_t* words = reinterpret_cast<const uint32_t*>(buf);
    uint16_t* const stab = stab_all + (uint32_t)warp * TAB16_TOTAL * SEQ_LANES;
    const bool active = (uint32_t)lane < SEQ_LANES;
    const uint32_t l = active ? (uint32_t)lane : 0u;
    const Tab16 tll{stab + TAB16_LL * SEQ_LANES + l, SEQ_LANES}, tml{stab + TAB16_ML * SEQ_LANES + l, SEQ_LANES},
        tof{stab + TAB16_OF * SEQ_LANES + l, SEQ_LANES};
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&counts[2], SEQ_LANES);
        first = __shfl_sync(0xFFFFFFFFu, first, 0);
        if (first >= n) break;
        const uint32_t k = first + (uint32_t)lane;
        if (active && k < n) {
            const uint32_t bi = order[k];
            ZBlock& gb = blocks[bi];
            ZBlock b;                       // only the fields the decoder reads (the rest of the 200-byte record stays in HBM)
            b.src = gb.src; b.bs_pos = gb.bs_pos; b.bs_len = gb.bs_len; b.nseq = gb.nseq; b.lit_regen = gb.lit_regen;
            b.tsrc[0] = gb.tsrc[0]; b.tsrc[1] = gb.tsrc[1]; b.tsrc[2] = gb.tsrc[2];
            const uint32_t entry = gb.entry;
            int16_t norm[64];
            uint16_t next_of[64];
            const int l0 = seq_tab16_for(buf, blocks, b, 0, tll, norm, next_of);
            const int l1 = seq_tab16_for(buf, blocks, b, 1, tof, norm, next_of);
            const int l2 = seq_tab16_for(buf, blocks, b, 2, tml, norm, next_of);
            int32_t st = ST_INVALID_DATA;
            uint32_t esc_n = 0, esc_idx[SEQ_ESC_MAX], esc_ll[SEQ_ESC_MAX], esc_ml[SEQ_ESC_MAX];
            if (l0 >= 0 && l1 >= 0 && l2 >= 0) {
                const uint64_t so = seq_base_of_entry[entry] + gb.seq_off;
                st = decode_sequences16(words, buf, b, tll, tof, tml, l0, l1, l2, s_llb, s_mlb, seqs + so, &esc_n, esc_idx, esc_ll, esc_ml);
            }
            if (st == ST_OK) {
                gb.out_size = b.out_size; gb.lit_used = b.lit_used;
                gb.rep_out[0] = b.rep_out[0]; gb.rep_out[1] = b.rep_out[1]; gb.rep_out[2] = b.rep_out[2];
                gb.esc_n = esc_n;
                for (uint32_t q = 0; q < esc_n && q < (uint32_t)SEQ_ESC_MAX; q++) { gb.esc_idx[q] = esc_idx[q]; gb.esc_ll[q] = esc_ll[q]; gb.esc_ml[q] = esc_ml[q]; }
            } else { gb.status = st; set_status(entries, entry, st); }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Literal stage: one lane per Huffman stream, 8 blocks (x4 streams) per warp; each block's single-symbol
// table (<= 8 KB at the format's maximum log 12) sits in shared memory and is shared by its 4 lanes.
constexpr uint32_t LIT_SLOTS = 8;
constexpr uint32_t LIT_SMEM_BYTES = LIT_SLOTS * (1u << HUF_LOG_MAX) * sizeof(uint16_t);   // 65536
__global__ void __launch_bounds__(32) zstd_lit_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, ZBlock* blocks,
                                                      const uint32_t* __restrict__ order, uint32_t* counts,
                                                      const uint64_t* __restrict__ lit_base_of_entry, uint8_t* __restrict__ lits) {
    extern __shared__ uint16_t shuf[];
    __shared__ int s_log[LIT_SLOTS], s_hdr[LIT_SLOTS];
    const int lane = threadIdx.x, slot = lane >> 2, sub = lane & 3;
    const uint32_t n = counts[1];
    const uint32_t* words = reinterpret_cast<const uint32_t*>(buf);
    uint16_t* table = shuf + slot * (1u << HUF_LOG_MAX);
    for (;;) {
        uint32_t batch = 0;
        if (lane == 0) batch = atomicAdd(&counts[3], 1u);
        batch = __shfl_sync(0xFFFFFFFFu, batch, 0);
        if ((uint64_t)batch * LIT_SLOTS >= n) break;
        const uint32_t k = batch * LIT_SLOTS + slot;
        const bool have = k < n;
        const uint32_t bi = have ? order[k] : 0;
        if (have && sub == 0) {
            const ZBlock& gb = blocks[bi];
            const ZBlock& hb = blocks[gb.huf_src];
            uint8_t weights[260];
            FseEntry wfse[64];
            int hlog = 0;
            const int hdr = huf_read_table(words, buf, hb.src + hb.lit_pos, hb.lit_csize, table, &hlog, weights, wfse);
            s_log[slot] = hlog; s_hdr[slot] = hdr;
        }
        __syncwarp();
        if (have) {
            ZBlock& gb = blocks[bi];
            bool ok = s_hdr[slot] >= 0;
            const uint32_t lit_csize = gb.lit_csize, lit_regen = gb.lit_regen;
            const uint32_t skip = gb.lit_type == LT_COMPRESSED ? (uint32_t)(ok ? s_hdr[slot] : 0) : 0;
            if (ok && skip > lit_csize) ok = false;
            const uint64_t at = gb.src + gb.lit_pos + skip;
            const uint32_t clen = lit_csize - (ok ? skip : 0);
            uint8_t* dst = lits + lit_base_of_entry[gb.entry] + gb.lit_off;
            if (ok) {
                if (gb.lit_streams == 1) {
                    if (sub == 0) ok = huf_decode_stream_w(words, buf, at, clen, table, s_log[slot], dst, lit_regen);
                } else {
                    uint32_t s1 = 0, s2 = 0, s3 = 0;
                    if (clen < 6) ok = false;
                    else { s1 = load_le16(buf + at); s2 = load_le16(buf + at + 2); s3 = load_le16(buf + at + 4); }
                    if (ok && (uint64_t)s1 + s2 + s3 + 6 > clen) ok = false;
                    const uint32_t seg = (lit_regen + 3) / 4;
                    if (ok && seg * 3 > lit_regen) ok = false;
                    if (ok) {
                        const uint32_t sz = sub == 0 ? s1 : sub == 1 ? s2 : sub == 2 ? s3 : clen - 6 - s1 - s2 - s3;
                        const uint64_t o = at + 6 + (sub > 0 ? s1 : 0) + (sub > 1 ? s2 : 0) + (sub > 2 ? s3 : 0);
                        const uint32_t cnt = sub < 3 ? seg : lit_regen - 3 * seg;
                        ok = huf_decode_stream_w(words, buf, o, sz, table, s_log[slot], dst + (uint64_t)sub * seg, cnt);
                    }
                }
            }
            if (!ok) { atomicCAS(&gb.status, ST_OK, ST_INVALID_DATA); set_status(entries, gb.entry, ST_INVALID_DATA); }
        }
        __syncwarp();
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
// LZ execution.  One warp per entry (a frame's blocks are order-dependent through the window), many
// entries in flight.  The warp builds the output in a linear shared-memory window (LZ_BUF bytes; when
// it fills, the newest LZ_KEEP bytes slide to the front) and streams it to HBM in 512-byte rows of
// 16-byte stores.  A step takes up to 32 sequences (<= LZ_STEP_MAX output bytes):
//   1. front (one step AHEAD): sequences from the cp.async-staged chunk, positions by one packed warp
//      scan, offset validation, and -- for matches whose source is older than the window ("far",
//      offset > LZ_KEEP) -- the HBM fetch of the source words into registers;
//   2. setup: every lane writes, for each output byte of its sequence, a 16-bit SOURCE CODE into idx[]:
//      either a step-relative position < LZ_STEP_MAX (a byte produced by this same step) or
//      0x8000 | shared-memory offset of a byte that already exists (literal stage, older window bytes,
//      far scratch row);
//   3. resolve, byte-parallel and divergence-free: lane j chases idx[] until it hits an existing byte,
//      copies it into the window.  No ordering between sequences is needed: chains only ever follow
//      codes, never data, and codes strictly decrease.
// Sequences longer than LZ_SEQ_MAX (or far matches > 32 bytes) are rare and go one at a time through
// whole-warp copies.
constexpr uint32_t LZ_BUF = 16384, LZ_KEEP = 8192, LZ_STEP_MAX = 2048;
constexpr uint32_t LZ_SEQ_CH = 256;                                      // sequences per staged chunk (x2 buffers)
constexpr uint32_t LZ_LIT_RING = 4096, LZ_LIT_CH = 1024, LZ_LIT_GUARD = 256;
constexpr uint32_t LZ_SEQ_MAX = 255;                                     // ll and ml bound of the parallel path
constexpr uint32_t LZ_FAR_MAX = 32;                                      // far matches up to this length are prefetched
constexpr uint32_t LZ_FAR_ROW = 48;                                      // scratch bytes per lane (36 used)
constexpr int LZ_WARPS = 7;
struct LzSmem {
    uint8_t win[LZ_BUF];
    SeqRec seq[2][LZ_SEQ_CH];
    uint8_t lit[LZ_LIT_RING + LZ_LIT_GUARD];
    uint8_t far[32 * LZ_FAR_ROW];
    uint16_t idx[LZ_STEP_MAX];
};
constexpr uint32_t LZ_OFF_LIT = LZ_BUF + 2 * LZ_SEQ_CH * 8, LZ_OFF_FAR = LZ_OFF_LIT + LZ_LIT_RING + LZ_LIT_GUARD;
constexpr uint32_t LZ_SMEM_BYTES = (uint32_t)sizeof(LzSmem) * LZ_WARPS;
static_assert(sizeof(LzSmem) % 16 == 0, "per-warp shared block keeps 16-byte alignment");
static_assert(sizeof(LzSmem) <= 0x8000, "source codes address the warp's shared block with 15 bits");

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ uint32_t ldcg32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct LzW {                 // per-warp state (uniform across lanes)
    LzSmem* S;
    uint8_t* obase;          // entry's output in HBM (16-byte aligned)
    uint64_t bpos;           // entry-relative position of win[0] (multiple of 16)
    uint64_t cur;            // entry-relative position of the next output byte
    uint64_t flushed;        // HBM holds [0, flushed) of the entry (multiple of 16)
    int lane;

    __device__ __forceinline__ uint8_t* wptr(uint64_t pos) const { return S->win + (uint32_t)(pos - bpos); }
    // 512-byte rows; force: also the 16-byte groups and the byte tail below cur (rewritten later, same values)
    __device__ __forceinline__ void flush(bool force) {
        while (flushed + 512 <= cur) {
            const uint4 v = *reinterpret_cast<const uint4*>(wptr(flushed + 16 * lane));
            *reinterpret_cast<uint4*>(obase + flushed + 16 * lane) = v;
            flushed += 512;
        }
        if (force && flushed < cur) {
            const uint32_t n = (uint32_t)(cur - flushed);
            if (16u * lane + 16u <= n) {
                const uint4 v = *reinterpret_cast<const uint4*>(wptr(flushed + 16 * lane));
                *reinterpret_cast<uint4*>(obase + flushed + 16 * lane) = v;
            }
            const uint32_t full = n & ~15u;
            if (full + lane < n) obase[flushed + full + lane] = *wptr(flushed + full + lane);
            flushed += full;
        }
    }
    // make room for `need` more bytes: slide the newest LZ_KEEP bytes (16-byte granular) to the front
    __device__ __forceinline__ void reserve(uint32_t need) {
        if ((uint32_t)(cur - bpos) + need <= LZ_BUF) return;
        __syncwarp();
        flush(false);
        const uint32_t fill = (uint32_t)(cur - bpos);
        const uint32_t shift = (fill - LZ_KEEP) & ~15u;   // fill > LZ_BUF - need >= LZ_KEEP + 16
        const uint32_t nrows = (fill - shift + 511) / 512;
        for (uint32_t r = 0; r < nrows; r++) {            // reads run >= shift (>= 512) bytes ahead of writes
            const uint4 v = *reinterpret_cast<const uint4*>(S->win + shift + r * 512 + 16 * lane);
            *reinterpret_cast<uint4*>(S->win + r * 512 + 16 * lane) = v;
        }
        bpos += shift;
        __syncwarp();
    }
    // one output byte at entry position pos < cur: from the window when still there, else from HBM
    __device__ __forceinline__ uint8_t read_out(uint64_t pos) const {
        if (pos >= bpos) return *wptr(pos);
        const uint32_t w = ldcg32(reinterpret_cast<const uint32_t*>(obase + (pos & ~3ull)));
        return (uint8_t)(w >> (8 * (pos & 3)));
    }
    // whole-warp append of n bytes from HBM (raw blocks, long literal runs); stride 0 = one repeated byte.
    // 16-byte pieces from the (unaligned) source into 16-byte aligned window rows.
    __device__ __forceinline__ void emit_global(const uint8_t* src, uint32_t n, uint32_t stride) {
        uint32_t done = 0;
        const uint32_t rep = stride ? 0u : 0x01010101u * src[0];
        while (done < n) {
            const uint32_t chunk = n - done < 1024u ? n - done : 1024u;
            reserve(chunk);
            uint8_t* d = wptr(cur);
            uint32_t head = (16u - ((uint32_t)(cur - bpos) & 15u)) & 15u;
            if (head > chunk) head = chunk;
            if ((uint32_t)lane < head) d[lane] = stride ? src[done + lane] : (uint8_t)rep;
            const uint32_t body = (chunk - head) >> 4;     // 16-byte pieces, <= 64
            for (uint32_t p = lane; p < body; p += 32) {
                uint32_t w4[4] = {rep, rep, rep, rep};
                if (stride) load16_any(src + done + head + 16 * p, w4);
                *reinterpret_cast<uint4*>(d + head + 16 * p) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            const uint32_t t0 = head + body * 16;
            if (t0 + lane < chunk) d[t0 + lane] = stride ? src[done + t0 + lane] : (uint8_t)rep;
            __syncwarp();
            cur += chunk; done += chunk;
            flush(false);
        }
    }
    // whole-warp match copy of any length / offset (long matches); off validated by the caller
    __device__ __forceinline__ void emit_match(uint32_t off, uint32_t n) {
        __syncwarp();
        flush(true);   // HBM now holds everything below cur, so sources that left the window can be read back
        uint32_t done = 0;
        while (done < n) {
            const uint32_t step = off >= 32u ? 32u : off;      // never read what the same step writes
            const uint32_t m = n - done < step ? n - done : step;
            reserve(32);
            uint8_t v = 0;
            if ((uint32_t)lane < m) v = read_out(cur + lane - off);
            __syncwarp();
            if ((uint32_t)lane < m) *wptr(cur + lane) = v;
            __syncwarp();
            cur += m; done += m;
            flush(false);
        }
    }
};

__global__ void __launch_bounds__(32 * LZ_WARPS) zstd_lz_kernel(const uint8_t* __restrict__ buf, EntryRec* entries,
                                                                 const ZEntry* __restrict__ ze, uint32_t nz,
                                                                 const ZBlock* __restrict__ blocks, const uint8_t* __restrict__ lits,
                                                                 const SeqRec* __restrict__ seqs, uint8_t* out,
                                                                 uint32_t* counts) {
    extern __shared__ __align__(16) uint8_t lz_smem_raw[];
    const int lane = threadIdx.x & 31;
    LzW W;
    W.S = reinterpret_cast<LzSmem*>(lz_smem_raw) + (threadIdx.x >> 5);
    W.lane = lane;
    LzSmem* const S = W.S;
    const uint8_t* const sbytes = reinterpret_cast<const uint8_t*>(S);
    uint32_t* const row = reinterpret_cast<uint32_t*>(S->far + lane * LZ_FAR_ROW);
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(&counts[4], 1u);
        wi = __shfl_sync(0xFFFFFFFFu, wi, 0);
        if (wi >= nz) break;
        const ZEntry z = ze[wi];
        EntryRec& er = entries[z.entry];
        if (er.status != ST_OK) continue;
        if (er.out_len > er.out_cap) { if (lane == 0) atomicCAS(&er.status, ST_OK, ST_NOSPACE); continue; }
        W.obase = out + er.out_off;
        W.bpos = 0; W.cur = 0; W.flushed = 0;
        int32_t fail = ST_OK;
        for (uint32_t k = z.blk_begin; k < z.blk_begin + z.blk_count && fail == ST_OK; k++) {
            const ZBlock& b = blocks[k];
            // W.cur == b.out_off here (the prefix pass laid the blocks out back to back)
            if (b.type == BT_RAW) { W.emit_global(buf + b.src, b.size, 1); continue; }
            if (b.type == BT_RLE) { W.emit_global(buf + b.src, b.size, 0); continue; }
            const uint8_t* lit;
            uint32_t lstride = 1;
            if (b.lit_type == LT_RAW) lit = buf + b.src + b.lit_pos;
            else if (b.lit_type == LT_RLE) { lit = buf + b.src + b.lit_pos; lstride = 0; }
            else lit = lits + z.lit_base + b.lit_off;
            const uint32_t lit_regen = b.lit_regen, nseq = b.nseq;
            const SeqRec* sq = seqs + z.seq_base + b.seq_off;
            const uint32_t rep_in[3] = {b.rep_in[0], b.rep_in[1], b.rep_in[2]};
            const uint64_t bstart = W.cur;                            // == b.out_off
            const uint64_t frame_dist = b.out_off - b.frame_out;      // bytes of this frame before the block
            uint32_t lp = 0;                                          // literals consumed
            uint32_t lit_loaded = 0;                                  // literal bytes staged so far (multiple of LZ_LIT_CH)
            if (!lstride) {                                           // RLE literals: the stage is that byte everywhere
                const uint8_t v = lit[0];
                for (uint32_t i = lane; i < LZ_LIT_RING + LZ_LIT_GUARD; i += 32) S->lit[i] = v;
                lit_loaded = 0xFFFFFFFFu;
                __syncwarp();
            }
            uint32_t staged = 0;                                      // sequence chunks issued so far
            auto stage_seq = [&](uint32_t chunk) {
                const uint32_t s0 = chunk * LZ_SEQ_CH;
                SeqRec* dst = S->seq[chunk & 1];
#pragma unroll
                for (uint32_t j = 0; j < LZ_SEQ_CH / 32; j++) {
                    const uint32_t i = s0 + j * 32 + lane;
                    if (i < nseq) cp_async8(dst + j * 32 + lane, sq + i);
                }
                cp_async_commit();
            };
            // ---- pipeline registers of the NEXT step (its front part runs one step ahead)
            uint32_t n_off = 1, n_ll = 0, n_ml = 0, n_dl = 0, n_sl = 0, n_wtot = 0, n_wlit = 0, n_nw = 0, n_take = 0;
            uint32_t n_t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool n_far = false, n_bad = false;
            uint64_t n_start = W.cur;       // entry position where the next step begins
            auto front = [&](uint32_t wbase) {
                // sequence chunks: c and c+1 cover the 32 sequences from wbase; c+1 is issued on entering c
                const uint32_t c = wbase / LZ_SEQ_CH;
                if (staged <= c + 1 && staged * LZ_SEQ_CH < nseq) { stage_seq(staged); staged++; }
                if ((wbase & (LZ_SEQ_CH - 1)) + 32 > LZ_SEQ_CH || staged <= c + 1) cp_async_wait_all(); else cp_async_wait_1();
                __syncwarp();
                const uint32_t i = wbase + lane;
                const bool valid = i < nseq;
                SeqRec r{1u, 0u};
                if (valid) r = S->seq[(i / LZ_SEQ_CH) & 1][i & (LZ_SEQ_CH - 1)];
                n_off = resolve_rep(r.x, rep_in);
                n_ll = r.y & 0xFFFFu; n_ml = r.y >> 16;
                bool longf = valid && (n_ll > LZ_SEQ_MAX || n_ml > LZ_SEQ_MAX);
                const uint32_t ll_c = (valid && !longf) ? n_ll : 0u, tot_c = (valid && !longf) ? n_ll + n_ml : 0u;
                const uint32_t incl = warp_incl_scan((ll_c << 16) | tot_c, lane);
                n_dl = (incl & 0xFFFFu) - tot_c;          // step-relative start of this lane's literals
                n_sl = (incl >> 16) - ll_c;               // step-relative start in the literal stream
                const uint64_t dmp = n_start + n_dl + n_ll;           // entry position of the match
                n_bad = valid && !longf && (n_off == 0 || (uint64_t)n_off > frame_dist + (dmp - bstart));
                n_far = valid && !longf && !n_bad && n_ml > 0 && n_off > LZ_KEEP;
                longf = longf || (n_far && n_ml > LZ_FAR_MAX);
                const uint32_t stop = __ballot_sync(0xFFFFFFFFu, !valid || longf || (incl & 0xFFFFu) > LZ_STEP_MAX);
                n_take = stop ? (uint32_t)__ffs(stop) - 1u : 32u;
                const uint32_t last = __shfl_sync(0xFFFFFFFFu, incl, n_take ? n_take - 1 : 0);
                n_wtot = last & 0xFFFFu; n_wlit = last >> 16;
                const bool taken = (uint32_t)lane < n_take;
                n_bad = n_bad && taken;
                n_far = n_far && taken;
                if (n_far) {   // the source has left (or will have left) the window; it is in HBM already: fetch it now
                    const uint64_t sp = dmp - n_off;
                    const uint32_t* g = reinterpret_cast<const uint32_t*>(W.obase + (sp & ~3ull));
                    n_nw = ((uint32_t)(sp & 3) + n_ml + 3) >> 2;      // <= 9 words
#pragma unroll
                    for (int q = 0; q < 9; q++) if ((uint32_t)q < n_nw) n_t[q] = ldcg32(g + q);
                }
            };
            if (nseq) { stage_seq(0); staged = 1; n_start = W.cur; front(0); }
            uint32_t wbase = 0;
            while (wbase < nseq && fail == ST_OK) {
                // ---- take over the step prepared by front()
                const uint32_t off = n_off, ll = n_ll, ml = n_ml, dl = n_dl, sl = n_sl, wtot = n_wtot, wlit = n_wlit, nw = n_nw,
                               ntake = n_take;
                const bool far = n_far;
                const bool any_bad = __any_sync(0xFFFFFFFFu, n_bad);
                if (ntake == 0) {
                    // the first sequence is long (or a long far match): it goes alone, by whole-warp copies
                    const uint32_t o1 = __shfl_sync(0xFFFFFFFFu, off, 0);
                    uint32_t l1 = __shfl_sync(0xFFFFFFFFu, ll, 0), m1 = __shfl_sync(0xFFFFFFFFu, ml, 0);
                    if (l1 == SEQ_ESC || m1 == SEQ_ESC)
                        for (uint32_t q = 0; q < b.esc_n && q < (uint32_t)SEQ_ESC_MAX; q++)
                            if (b.esc_idx[q] == wbase) { l1 = b.esc_ll[q]; m1 = b.esc_ml[q]; }
                    if ((uint64_t)lp + l1 > lit_regen) { fail = ST_INVALID_DATA; break; }
                    if (l1) W.emit_global(lit + (size_t)lp * lstride, l1, lstride);
                    lp += l1;
                    if (o1 == 0 || (uint64_t)o1 > frame_dist + (W.cur - bstart)) { fail = ST_INVALID_DATA; break; }
                    if (m1) W.emit_match(o1, m1);
                    __syncwarp();
                    wbase += 1;
                    if (wbase < nseq) { n_start = W.cur; front(wbase); }
                    continue;
                }
                if (any_bad || (uint64_t)lp + wlit > lit_regen) { fail = ST_INVALID_DATA; break; }
                // ---- room in the window; stage the literals this step reads
                W.reserve(LZ_STEP_MAX);
                if (lstride && lit_loaded < lp + wlit) {
                    if (lit_loaded + LZ_LIT_RING < lp) lit_loaded = lp & ~(LZ_LIT_CH - 1);   // a long run was copied around the stage
                    __syncwarp();
                    while (lit_loaded < lp + wlit) {
#pragma unroll
                        for (int j = 0; j < 2; j++) {   // chunk [lit_loaded, +1024): two 16-byte pieces per lane, unaligned source
                            const uint32_t p = lit_loaded + j * 512 + 16 * lane;
                            if (p < lit_regen) {
                                uint32_t w4[4];
                                load16_any(lit + p, w4);
                                const uint32_t si = p & (LZ_LIT_RING - 1);
                                const uint4 v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                                *reinterpret_cast<uint4*>(S->lit + si) = v;
                                if (si < LZ_LIT_GUARD) *reinterpret_cast<uint4*>(S->lit + LZ_LIT_RING + si) = v;
                            }
                        }
                        lit_loaded += LZ_LIT_CH;
                    }
                }
                const uint64_t step_start = W.cur;
                const uint32_t sidx = (uint32_t)(step_start - W.bpos);       // window index of the step's first byte
                const bool taken = (uint32_t)lane < ntake;
                // this step's far words (fetched during the previous step) go to the lane's scratch row
                if (far) {
#pragma unroll
                    for (int q = 0; q < 9; q++) if ((uint32_t)q < nw) row[q] = n_t[q];
                }
                // ---- prepare the NEXT step now, so that its HBM fetches overlap this step's work
                wbase += ntake;
                if (wbase < nseq) { n_start = step_start + wtot; front(wbase); }
                // ---- setup: source codes of this lane's bytes (runs of consecutive codes)
                if (taken) {
                    uint16_t* ix = S->idx + dl;
                    uint32_t code = 0x8000u | (LZ_OFF_LIT + ((lp + sl) & (LZ_LIT_RING - 1)));   // guard covers the wrap
                    for (uint32_t q = 0; q < ll; q++) ix[q] = (uint16_t)(code + q);
                    ix += ll;
                    const int32_t srel = (int32_t)(dl + ll) - (int32_t)off;                      // step-relative source start
                    uint32_t nneg;                                                               // bytes that exist already
                    if (far) { nneg = ml; code = 0x8000u | (LZ_OFF_FAR + lane * LZ_FAR_ROW + (uint32_t)((step_start + dl + ll - off) & 3)); }
                    else { nneg = srel < 0 ? ((uint32_t)(-srel) < ml ? (uint32_t)(-srel) : ml) : 0u; code = 0x8000u | (uint32_t)((int32_t)sidx + srel); }
                    for (uint32_t q = 0; q < nneg; q++) ix[q] = (uint16_t)(code + q);
                    code = (uint32_t)(srel + (int32_t)nneg);                                     // >= 0: produced by this step
                    for (uint32_t q = nneg; q < ml; q++) ix[q] = (uint16_t)(code + (q - nneg));
                }
                __syncwarp();
                // ---- resolve: byte-parallel, each lane chases its byte's code down to a byte that exists
                {
                    uint8_t* const wd = S->win + sidx;
                    const uint16_t* const ix = S->idx;
                    for (uint32_t j = lane; j < wtot; j += 128) {
                        uint32_t c0 = ix[j], c1 = j + 32 < wtot ? ix[j + 32] : 0x8000u, c2 = j + 64 < wtot ? ix[j + 64] : 0x8000u,
                                 c3 = j + 96 < wtot ? ix[j + 96] : 0x8000u;
                        while (!((c0 & c1 & c2 & c3) & 0x8000u)) {
                            if (!(c0 & 0x8000u)) c0 = ix[c0];
                            if (!(c1 & 0x8000u)) c1 = ix[c1];
                            if (!(c2 & 0x8000u)) c2 = ix[c2];
                            if (!(c3 & 0x8000u)) c3 = ix[c3];
                        }
                        const uint8_t v0 = sbytes[c0 & 0x7FFFu], v1 = sbytes[c1 & 0x7FFFu], v2 = sbytes[c2 & 0x7FFFu], v3 = sbytes[c3 & 0x7FFFu];
                        wd[j] = v0;
                        if (j + 32 < wtot) wd[j + 32] = v1;
                        if (j + 64 < wtot) wd[j + 64] = v2;
                        if (j + 96 < wtot) wd[j + 96] = v3;
                    }
                }
                __syncwarp();
                W.cur = step_start + wtot;
                lp += wlit;
                W.flush(false);
            }
            if (fail != ST_OK) break;
            // trailing literals of the block
            if (lp > lit_regen || (W.cur - bstart) + (lit_regen - lp) != b.out_size) { fail = ST_INVALID_DATA; break; }
            __syncwarp();
            if (lit_regen > lp) W.emit_global(lit + (size_t)lp * lstride, lit_regen - lp, lstride);
        }
        __syncwarp();
        if (fail == ST_OK) W.flush(true);
        if (fail != ST_OK && lane == 0) atomicCAS(&er.status, ST_OK, fail);
        __syncwarp();
    }
}

}  // namespace zs
}  // namespace pna
