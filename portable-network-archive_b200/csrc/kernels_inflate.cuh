// kernels_inflate.cuh -- sm_100a kernel for zlib/deflate decode (K5): one stream per WARP, decoded by the warp's
// lane 0 (Huffman tables in shared memory, 1732 B per stream).  DEFLATE decoding is control-flow heavy and every
// stream takes its own path: with a stream per lane the 32 paths of a warp serialise (measured 3 GB/s on 32k small
// files); with a stream per warp nothing diverges and 48 warps per SM hide the serial chain's latency.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "inflate_core.cuh"

namespace pna {
namespace inf {

static_assert(sizeof(Tables) == 1732, "one stream's tables: odd number of 32-bit words");
constexpr int INFLATE_CTA = 8;   // streams (= warps) per CTA

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
