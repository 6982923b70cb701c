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

// list[i] = index into EntryRec[].  size_only: the decoded length from the chunk headers (two-pass sizing), no decoding.
__global__ void __launch_bounds__(32) xz_decode_kernel(const uint8_t* __restrict__ buf, EntryRec* entries, const uint32_t* __restrict__ list,
                                                       uint32_t n, uint8_t* __restrict__ out, int size_only) {
    extern __shared__ uint32_t smem_raw[];
    if (threadIdx.x) return;
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    EntryRec& e = entries[list[i]];
    if (e.status != ST_OK && !(size_only == 0 && e.status == ST_NOSPACE)) return;
    uint64_t produced = 0;
    int32_t st;
    if (size_only) st = xz_stream_size(buf + e.comp_off, e.comp_len, &produced);
    else st = xz_decode(buf + e.comp_off, e.comp_len, out + e.out_off, e.out_cap, &produced, reinterpret_cast<uint16_t*>(smem_raw));
    e.out_len = produced;
    if (st != ST_OK) atomicCAS(&e.status, ST_OK, st);
}

}  // namespace xz
}  // namespace pna
