// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A fiber-based SIMT emulator: enough of the CUDA language and runtime
// to compile portable-network-archive_b200/csrc/*.cu* with g++ and execute the kernels' LOGIC on a box without a
// GPU (the development container has none; every GPU minute is budgeted).  It is never shipped, never linked into
// libpna_cuda.so and never loaded by the package: tests/emu/build_emu.py builds tests/emu/_gen/libpna_cuda.so and
// only `pytest --emu` (tests/conftest.py) points the ctypes loader at it.  Results obtained under the emulator are
// debugging aids, not parity evidence -- the parity tier proper is `pytest -m gpu` on a B200.
//
// Model: CTAs run one after another on the calling OS thread (launches are serialised by a global mutex); the threads
// of a CTA are fibers scheduled round-robin that switch only at __syncthreads* and at warp collectives (*_sync).
// Warp collectives gather the operands of all lanes named in the mask.  Data races are therefore invisible here
// (compute-sanitizer on the GPU box covers them); missing barriers between a lane's write and another lane's read
// ARE visible, because lanes never run in lockstep.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <chrono>
#include <functional>
#include <map>
#include <mutex>
#include <vector>

#define PNA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace emu {

extern "C" void pna_emu_switch(void** save_sp, void* load_sp);

struct Fiber {
    void* sp = nullptr;
    uint8_t* stack = nullptr;
    int state = 0;            // 0 ready, 1 waiting on *wait_gen != wait_val, 2 done
    const uint64_t* wait_gen = nullptr;
    uint64_t wait_val = 0;
    uint3 tid{0, 0, 0};
    unsigned linear = 0;
};
// One rendezvous per distinct mask: lanes of a sub-group may run through many collectives (e.g. __syncwarp(m) in a loop)
// while the other lanes of the warp wait at a full-mask one.
struct Rdv {
    uint32_t arrived = 0;
    uint64_t gen = 0;
    uint64_t slot[32];
    static constexpr int RING = 4;   // a lane resumed late still finds its generation's result
    uint64_t res[RING][32];
    uint32_t res_mask[RING];
    uint64_t pending_gen[32];
    Rdv() { for (auto& p : pending_gen) p = UINT64_MAX; }
};
struct Warp {
    uint32_t exited = 0;
    std::map<uint32_t, Rdv> rdv;
};
struct State {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    size_t stack_bytes = 256 * 1024;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    unsigned nthreads = 0, live = 0;
    // block barrier
    unsigned bar_count = 0;
    uint64_t bar_gen = 0;
    int bar_acc_or = 0, bar_acc_and = 1, bar_acc_cnt = 0, bar_res_or = 0, bar_res_and = 1, bar_res_cnt = 0;
    const std::function<void()>* body = nullptr;
    uint8_t* dyn = nullptr;
    size_t dyn_cap = 0;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    const char* kernel = "";
};
inline State& S() { static State s; return s; }
inline std::recursive_mutex& launch_mutex() { static std::recursive_mutex m; return m; }
inline void* dyn_smem() { return S().dyn; }

[[noreturn]] inline void die(const char* what) {
    State& s = S();
    fprintf(stderr, "[cuda_emu] %s: %s (block %u,%u,%u of %u, thread %u of %u)\n", s.kernel, what, s.bid.x, s.bid.y, s.bid.z, s.gdim.x, s.cur ? s.cur->linear : 0u, s.nthreads);
    abort();
}
inline void yield_wait(const uint64_t* gen, uint64_t val) {
    State& s = S();
    Fiber* f = s.cur;
    f->state = 1; f->wait_gen = gen; f->wait_val = val;
    pna_emu_switch(&f->sp, s.sched_sp);
}
inline void bar_release(State& s) {
    s.bar_res_or = s.bar_acc_or; s.bar_res_and = s.bar_acc_and; s.bar_res_cnt = s.bar_acc_cnt;
    s.bar_acc_or = 0; s.bar_acc_and = 1; s.bar_acc_cnt = 0;
    s.bar_count = 0; s.bar_gen++;
}
inline void barrier(int pred) {
    State& s = S();
    s.bar_acc_or |= pred != 0; s.bar_acc_and &= pred != 0; s.bar_acc_cnt += pred != 0;
    s.bar_count++;
    if (s.bar_count == s.live) { bar_release(s); return; }
    const uint64_t g = s.bar_gen;
    yield_wait(&s.bar_gen, g);
}
// gathers v of every lane in `mask` (minus exited lanes); returns the result array, valid until the lane's next collective
inline void rdv_complete(Warp& w, Rdv& r, uint32_t need) {
    const int k = (int)(r.gen % Rdv::RING);
    for (int i = 0; i < 32; i++)
        if (r.pending_gen[i] != UINT64_MAX && r.pending_gen[i] + Rdv::RING <= r.gen && !(w.exited >> i & 1))
            die("emulator limitation: a suspended lane's collective result was overwritten");
    for (int i = 0; i < 32; i++) if (need >> i & 1) r.res[k][i] = r.slot[i];
    r.res_mask[k] = need;
    r.arrived &= ~need;
    r.gen++;
}
inline const uint64_t* gather(uint32_t mask, uint64_t v, uint32_t* res_mask = nullptr) {
    State& s = S();
    Fiber* f = s.cur;
    Warp& w = s.warps[f->linear >> 5];
    const uint32_t lane = f->linear & 31u, bit = 1u << lane;
    if (!(mask & bit)) die("collective called by a lane outside its mask");
    Rdv& r = w.rdv[mask];
    r.slot[lane] = v;
    r.arrived |= bit;
    const uint32_t need = mask & ~w.exited;
    const uint64_t g = r.gen;
    if ((r.arrived & need) == need) rdv_complete(w, r, need);
    else {
        r.pending_gen[lane] = g;
        yield_wait(&r.gen, g);
        r.pending_gen[lane] = UINT64_MAX;
    }
    const int k = (int)(g % Rdv::RING);
    if (res_mask) *res_mask = r.res_mask[k];
    return r.res[k];
}
inline void fiber_exit() {
    State& s = S();
    Fiber* f = s.cur;
    f->state = 2;
    s.live--;
    Warp& w = s.warps[f->linear >> 5];
    w.exited |= 1u << (f->linear & 31u);
    if (s.live && s.bar_count == s.live) bar_release(s);   // the others were waiting for this thread only
    // pending collectives that only waited for this lane are complete now
    for (auto& kv : w.rdv) {
        Rdv& r = kv.second;
        const uint32_t need = kv.first & ~w.exited;
        if (r.arrived && (r.arrived & need) == need) rdv_complete(w, r, need);
    }
    pna_emu_switch(&f->sp, s.sched_sp);
    die("resumed a finished fiber");
}
extern "C" inline void pna_emu_fiber_main() {
    (*S().body)();
    fiber_exit();
}
void fiber_entry_asm();

inline void run_block(State& s) {
    const unsigned T = s.nthreads;
    s.live = T; s.bar_count = 0;
    for (auto& w : s.warps) { w.exited = 0; w.rdv.clear(); }
    for (unsigned t = 0; t < T; t++) {
        Fiber& f = s.fibers[t];
        if (!f.stack) {
            f.stack = (uint8_t*)mmap(nullptr, s.stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (f.stack == MAP_FAILED) die("mmap of a fiber stack failed");
        }
        uintptr_t top = ((uintptr_t)f.stack + s.stack_bytes) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                               // fake return address of the entry function
        *--sp = (void*)&pna_emu_fiber_main;            // popped by `ret` in pna_emu_switch
        for (int i = 0; i < 6; i++) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
        f.sp = sp;
        f.state = 0;
        f.linear = t;
        f.tid.x = t % s.bdim.x; f.tid.y = (t / s.bdim.x) % s.bdim.y; f.tid.z = t / (s.bdim.x * s.bdim.y);
    }
    unsigned done = 0;
    while (done < T) {
        bool progress = false;
        for (unsigned t = 0; t < T; t++) {
            Fiber& f = s.fibers[t];
            if (f.state == 2) continue;
            if (f.state == 1) { if (*f.wait_gen == f.wait_val) continue; f.state = 0; }
            s.cur = &f;
            pna_emu_switch(&s.sched_sp, f.sp);
            progress = true;
            if (f.state == 2) done++;
        }
        if (!progress) { s.cur = nullptr; die("deadlock: every live thread waits at a barrier or collective that cannot complete"); }
    }
    s.cur = nullptr;
}

template <class F>
inline void launch(const char* name, dim3 grid, dim3 block, size_t smem, F&& body) {
    std::lock_guard<std::recursive_mutex> g(launch_mutex());
    State& s = S();
    const unsigned T = block.x * block.y * block.z;
    if (T == 0 || T > 1024) die("bad block size");
    if (smem > 232448) die("more dynamic shared memory than an SM has");
    if (s.fibers.size() < T) s.fibers.resize(T);
    s.warps.assign((T + 31) / 32, Warp());
    s.nthreads = T; s.bdim = block; s.gdim = grid; s.kernel = name;
    if (s.dyn_cap < smem + 128) { free(s.dyn); s.dyn_cap = smem + 128; s.dyn = (uint8_t*)aligned_alloc(128, (s.dyn_cap + 127) & ~(size_t)127); }
    std::function<void()> fn = body;
    s.body = &fn;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                s.bid = uint3{x, y, z};
                memset(s.dyn, 0xA5, smem);   // shared memory starts undefined
                run_block(s);
            }
    s.body = nullptr;
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

// ---- synchronisation and warp collectives
static inline void __syncthreads() { emu::barrier(0); }
static inline int __syncthreads_or(int p) { emu::barrier(p); return emu::S().bar_res_or; }
static inline int __syncthreads_and(int p) { emu::barrier(p); return emu::S().bar_res_and; }
static inline int __syncthreads_count(int p) { emu::barrier(p); return emu::S().bar_res_cnt; }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::gather(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned emu_lane() { return emu::S().cur->linear & 31u; }
template <class T>
static inline uint64_t emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
static inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const unsigned lane = emu_lane();
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, emu_bits(v), &rm);
    const unsigned s = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    return (rm >> s & 1) ? emu_unbits<T>(r[s]) : v;
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const unsigned lane = emu_lane();
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, emu_bits(v), &rm);
    const unsigned base = lane & ~(unsigned)(width - 1);
    if (lane - base < delta) return v;
    const unsigned s = lane - delta;
    return (rm >> s & 1) ? emu_unbits<T>(r[s]) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const unsigned lane = emu_lane();
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, emu_bits(v), &rm);
    const unsigned base = lane & ~(unsigned)(width - 1);
    const unsigned s = lane + delta;
    if (s >= base + (unsigned)width) return v;
    return (rm >> s & 1) ? emu_unbits<T>(r[s]) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const unsigned lane = emu_lane();
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, emu_bits(v), &rm);
    const unsigned s = lane ^ (unsigned)lanemask;
    if ((s & ~(unsigned)(width - 1)) != (lane & ~(unsigned)(width - 1))) return v;
    return (rm >> s & 1) ? emu_unbits<T>(r[s]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, pred != 0, &rm);
    unsigned b = 0;
    for (int i = 0; i < 32; i++) if ((rm >> i & 1) && r[i]) b |= 1u << i;
    return b;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, pred != 0, &rm);
    for (int i = 0; i < 32; i++) if ((rm >> i & 1) && !r[i]) return 0;
    return 1;
}
template <class T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    const unsigned lane = emu_lane();
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, emu_bits(v), &rm);
    unsigned b = 0;
    for (int i = 0; i < 32; i++) if ((rm >> i & 1) && r[i] == r[lane]) b |= 1u << i;
    return b;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, v, &rm);
    unsigned b = 0;
    for (int i = 0; i < 32; i++) if (rm >> i & 1) b |= (unsigned)r[i];
    return b;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, v, &rm);
    unsigned b = 0;
    for (int i = 0; i < 32; i++) if (rm >> i & 1) b += (unsigned)r[i];
    return b;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, v, &rm);
    unsigned b = 0;
    for (int i = 0; i < 32; i++) if ((rm >> i & 1) && (unsigned)r[i] > b) b = (unsigned)r[i];
    return b;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    uint32_t rm;
    const uint64_t* r = emu::gather(mask, v, &rm);
    unsigned b = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++) if ((rm >> i & 1) && (unsigned)r[i] < b) b = (unsigned)r[i];
    return b;
}
static inline unsigned __activemask() { return 0xFFFFFFFFu & ~emu::S().warps[emu::S().cur->linear >> 5].exited; }   // convergence is not modelled

// ---- integer intrinsics
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) {
    v = (v >> 16) | (v << 16); v = ((v & 0xFF00FF00u) >> 8) | ((v & 0x00FF00FFu) << 8);
    v = ((v & 0xF0F0F0F0u) >> 4) | ((v & 0x0F0F0F0Fu) << 4); v = ((v & 0xCCCCCCCCu) >> 2) | ((v & 0x33333333u) << 2);
    return ((v & 0xAAAAAAAAu) >> 1) | ((v & 0x55555555u) << 1);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { s = s > 32 ? 32 : s; return (unsigned)((((uint64_t)hi << 32) | lo) >> s); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { s = s > 32 ? 32 : s; return (unsigned)(((((unsigned __int128)hi << 32) | lo) << s) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline void __nanosleep(unsigned) {}

// ---- atomics (one OS thread runs kernels at a time)
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicXor(T* p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- runtime
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0, cudaEventBlockingSync = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct emu_stream_t { int dev; };
typedef emu_stream_t* cudaStream_t;
struct emu_event_t { double t; };
typedef emu_event_t* cudaEvent_t;
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 148; char name[64] = "emulated sm_100 (tests/emu)"; size_t totalGlobalMem = (size_t)180 << 30; int pciBusID = 0, pciDeviceID = 0, pciDomainID = 0; };
namespace emu {
inline int& cur_device() { static thread_local int d = 0; return d; }
inline int device_count() { const char* e = getenv("PNA_EMU_DEVICES"); int n = e ? atoi(e) : 1; return n < 0 ? 0 : n; }
inline int sm_count() { const char* e = getenv("PNA_EMU_SMS"); int n = e ? atoi(e) : 148; return n > 0 ? n : 148; }
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu::device_count(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu::device_count()) return cudaErrorInvalidValue; emu::cur_device() = d; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = emu::cur_device(); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d) { *p = cudaDeviceProp(); p->multiProcessorCount = emu::sm_count(); p->pciBusID = d + 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    if (!*p) return cudaErrorMemoryAllocation;
    // device memory starts undefined: poison small allocations entirely, large ones only at the ends (speed)
    if (n <= ((size_t)64 << 20)) memset(*p, 0xCD, n); else { memset(*p, 0xCD, (size_t)1 << 20); memset((uint8_t*)*p + n - ((size_t)1 << 20), 0xCD, (size_t)1 << 20); }
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
typedef int cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 1 };
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = 0; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = aligned_alloc(4096, (n + 4095) & ~(size_t)4095); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaHostAlloc(p, n, 0); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::lock_guard<std::recursive_mutex> g(emu::launch_mutex());
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emu_stream_t{emu::cur_device()}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event_t{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu::now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaSetDeviceFlags(unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)8 << 30; *t = (size_t)180 << 30; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetPCIBusId(char* s, int len, int d) { snprintf(s, len, "0000:%02x:00.0", d + 1); return cudaSuccess; }
