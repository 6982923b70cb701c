// kernels_xz.cuh -- decompress_reader's XZ arm (lib/src/entry/read.rs:182: liblzma's XzDecoder over the decrypted stream).
// LZMA is one adaptive range coder per stream: every bit decoded depends on the probabilities the bits before it left behind, so
// the only parallelism is ACROSS streams.  One warp per stream, lane 0 decoding (lzma_core.cuh), its 28 KB probability arena in
// shared memory (the dependent load-compare-update chain per bit then runs at shared-memory latency, not L2's); the other lanes
// only hold the scheduler slot.  Eight streams per SM.  The output buffer itself is the LZMA dictionary.
#pragma once
#include "common.cuh"
#include "lzma_core.cuh"

namespace pna {
namespace xz {

constexpr uint32_t XZ_SMEM_BYTES = (LZMA_PROBS_MAX * 2u + 15u) & ~15u;
// the chunk-parallel pass keeps arenas for lc + lp <= 2 only (what this library's writer emits: 10 KB, 22 windows in flight per SM
// instead of 8); a chunk with more literal-context bits sends its stream to the serial decoder
constexpr uint32_t XZ_WIN_LCLP_MAX = 2;
constexpr uint32_t XZ_WIN_SMEM_BYTES = ((LZMA_PROBS_FIXED + (0x300u << XZ_WIN_LCLP_MAX)) * 2u + 15u) & ~15u;

// list[i] = index into EntryRec[].  size_only: the decoded length from the chunk headers (two-pass sizing), no decoding.
// win_begin (n + 1 entries, null: no chunk-parallel pass ran) / wins: the window records of stream i are wins[win_begin[i] .. win_begin[i + 1])
__global__ void __launch_bounds__(32) xz_decode_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, const uint32_t* __restrict__ list,
                                                       uint32_t n, uint8_t* __restrict__ out, int size_only,
                                                       const uint32_t* __restrict__ win_begin, const XzWin* __restrict__ wins) {
    extern __shared__ uint32_t smem_raw[];
    if (threadIdx.x) return;
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    EntryRec& e = entries[list[i]];
    if (e.status != ST_OK && !(size_only == 0 && e.status == ST_NOSPACE)) return;
    uint64_t produced = 0;
    int32_t st;
    if (size_only) st = xz_stream_size(buf + e.comp_off, e.comp_len, &produced);
    else st = xz_decode(buf + e.comp_off, e.comp_len, out + e.out_off, e.out_cap, &produced, reinterpret_cast<uint16_t*>(smem_raw),
                        win_begin ? wins + win_begin[i] : nullptr, win_begin ? win_begin[i + 1] - win_begin[i] : 0u);
    e.out_len = produced;
    if (st != ST_OK) atomicCAS(&e.status, ST_OK, st);
}

// Chunk-parallel pass (lzma_core.cuh, "chunk-parallel decoding"): a warp per XZ_WIN bytes of output of a stream whose chunks all
// reset the dictionary.  The warp decodes the chunks that start in its window (lane 0 runs the range decoder, all lanes reset the
// probabilities, copy uncompressed chunks and take the CRC-32 of what was produced) and leaves (crc, length, x^(8 length), ok/bad)
// for the serial walk of xz_decode_kernel, which checks the container and combines the CRCs.  map[g] = (position in list[], window).
__global__ void __launch_bounds__(32) xz_window_kernel(const uint8_t* __restrict__ buf, const EntryRec* __restrict__ entries,
                                                       const uint32_t* __restrict__ list, const uint2* __restrict__ map, uint32_t n_wins,
                                                       uint8_t* __restrict__ out, XzWin* __restrict__ wins) {
    extern __shared__ uint32_t smem_raw[];
    uint16_t* const probs = reinterpret_cast<uint16_t*>(smem_raw);
    const uint32_t g = blockIdx.x, lane = threadIdx.x;
    if (g >= n_wins) return;
    const uint2 m = map[g];
    const EntryRec& e = entries[list[m.x]];
    const uint8_t* const in = buf + e.comp_off;
    uint8_t* const outp = out + e.out_off;
    uint64_t fi = 0, fo = 0;
    uint32_t nc = 0, ok = 0;
    if (lane == 0 && e.status == ST_OK) ok = xz_chunked_layout(in, e.comp_len, e.out_cap, m.y, &fi, &fo, &nc) ? 1u : 0u;
    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
    fi = __shfl_sync(0xFFFFFFFFu, fi, 0); fo = __shfl_sync(0xFFFFFFFFu, fo, 0); nc = __shfl_sync(0xFFFFFFFFu, nc, 0);
    uint64_t pos = fi, op = fo;
    for (uint32_t c = 0; ok && c < nc; c++) {
        const uint32_t ctl = in[pos];
        if (ctl == 1) {
            const uint32_t usize = (((uint32_t)in[pos + 1] << 8) | in[pos + 2]) + 1;
            for (uint32_t k = lane; k < usize; k += 32) outp[op + k] = in[pos + 3 + k];
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
            if (in[pos + 5] > (4 * 5 + 4) * 9 + 8 || S.lc + S.lp > XZ_WIN_LCLP_MAX) { ok = 0; break; }
            const uint32_t np = LZMA_PROBS_FIXED + (0x300u << (S.lc + S.lp));
            for (uint32_t k = lane; k < np; k += 32) probs[k] = (uint16_t)PROB_INIT;
            __syncwarp();
            int32_t st = ST_OK;
            if (lane == 0) st = lzma_chunk(S, probs, in + pos + 6, csize, outp, op, usize, op);
            st = __shfl_sync(0xFFFFFFFFu, st, 0);
            if (st != ST_OK) { ok = 0; break; }
            pos += 6 + csize; op += usize;
        }
    }
    __syncwarp();   // lane 0's bytes are visible to the lanes that checksum them (same warp, generic memory)
    __threadfence_block();
    XzWin r;
    r.crc = 0; r.len = 0; r.xpow = 0x80000000u; r.state = 2;
    if (ok) {
        const uint64_t L = op - fo, sl = (L + 31) / 32;
        const uint64_t lo = (uint64_t)lane * sl, nl = lo >= L ? 0 : (L - lo < sl ? L - lo : sl);
        const uint32_t my_crc = nl ? xz_crc32(outp + fo + lo, nl) : 0u, my_pow = xz_crc_xpow(nl);
        uint32_t crc = 0;
        for (int l = 0; l < 32; l++) crc = xz_crc_mul(__shfl_sync(0xFFFFFFFFu, my_pow, l), crc) ^ __shfl_sync(0xFFFFFFFFu, my_crc, l);
        r.crc = crc; r.len = (uint32_t)L; r.xpow = xz_crc_xpow(L); r.state = 1;
    }
    if (lane == 0) wins[g] = r;
}

}  // namespace xz
}  // namespace pna
