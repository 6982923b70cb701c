// common.cuh -- shared definitions for the sm_100a kernels and their host-side test builds.
//
// Every algorithmic core in this directory is written as PNA_HD (host+device) code over plain
// pointers so that tests/host/ can compile the very same functions with g++ and check them on
// the CPU box (no GPU there); the __global__ wrappers live in the *_kernels.cuh files.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#if defined(__CUDACC__)
#define PNA_HD __host__ __device__ __forceinline__
#define PNA_D __device__ __forceinline__
#else
#define PNA_HD inline
#define PNA_D inline
#endif

namespace pna {

// per-entry status == reference io::ErrorKind class (mirrors include/pna_cuda.h)
enum : int32_t {
    ST_OK = 0,
    ST_INVALID_DATA = 1,
    ST_UNEXPECTED_EOF = 2,
    ST_INVALID_INPUT = 3,
    ST_UNSUPPORTED = 4,
    ST_NOSPACE = 5,
    ST_OOM = 6,
    ST_INTERNAL = 7,
};

PNA_HD uint32_t load_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
PNA_HD uint32_t load_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
PNA_HD uint32_t load_le24(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}
PNA_HD int highbit32(uint32_t v) {  // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
PNA_HD uint32_t rotl32(uint32_t v, int s) { return (v << s) | (v >> ((32 - s) & 31)); }
PNA_HD uint32_t bswap32(uint32_t v) {
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
}


// ---- asynchronous global -> shared copies (cp.async = LDGSTS).  Under the test-only SIMT emulator (tests/emu,
// PNA_EMU) the copy happens at issue time.
#if defined(__CUDACC__) || defined(PNA_EMU)
PNA_D void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(PNA_EMU)
    memcpy(smem_dst, gsrc, 16);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
#endif
}
PNA_D void cp_async8(void* smem_dst, const void* gsrc) {
#if defined(PNA_EMU)
    memcpy(smem_dst, gsrc, 8);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
#endif
}
PNA_D void cp_async_commit() {
#if !defined(PNA_EMU)
    asm volatile("cp.async.commit_group;");
#endif
}
PNA_D void cp_async_wait_all() {
#if !defined(PNA_EMU)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
PNA_D void cp_async_wait_1() {
#if !defined(PNA_EMU)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
PNA_D uint32_t ldcg32(const uint32_t* p) {
#if defined(PNA_EMU)
    return *p;
#else
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}
#endif

// One data stream (= one entry) as the kernels see it.  Host fills it (abi.cu); layout is shared.
struct Segment {        // one FDAT/SDAT body inside the uploaded image
    uint64_t img_off;   // byte offset in the device image
    uint64_t pos;       // byte offset of this body inside the entry's stream
};
struct EntryRec {
    uint64_t seg_begin;   // index of first Segment
    uint32_t n_segs;
    uint8_t compression, encryption, cipher_mode, _pad;
    uint64_t stream_len;  // sum of body lengths (IV included)
    uint64_t comp_off;    // where the decrypted (still compressed) stream goes in the comp arena (16B aligned)
    uint64_t comp_len;    // its length: host-computed for none/CTR; written by the cipher kernel for CBC (unpad)
    uint64_t out_off;     // where the decoded bytes go in the out arena
    uint64_t out_cap;     // capacity reserved there
    uint64_t out_len;     // decoded length (device-written)
    int32_t key_idx;      // index into the round-key table, -1 if unencrypted
    int32_t status;       // device-written
};

}  // namespace pna
