// host_api.cpp -- implementation of include/pna_host.hpp: the container logic of the reference's data-chunk path in
// C++ above the C ABI of libpna_cuda.so.  Built by __graft_entry__.build() into libpna_host.so (g++, links libpna_cuda.so).
// Entry bytes are never touched by the CPU here; see the header for the reference file:line each piece follows.
#include "../../include/pna_host.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/random.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <mutex>
#include <functional>
#include <random>
#include <set>
#include <thread>
#include <sched.h>

namespace pna {

// uploads in flight per process: PCIe is the shared resource; two slots let one batch's host-side set-up run under the
// other's DMA while still staggering the workers
struct Slots {
    std::mutex mu;
    std::condition_variable cv;
    int free_;
    explicit Slots(int n) : free_(n) {}
    void acquire() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return free_ > 0; }); free_--; }
    void release() { { std::lock_guard<std::mutex> lk(mu); free_++; } cv.notify_one(); }
};
struct SlotGuard { Slots& s; explicit SlotGuard(Slots& x) : s(x) { s.acquire(); } ~SlotGuard() { s.release(); } };

// Pinned host buffers are expensive to make (cudaHostAlloc pins at a few GB/s): the windows and staging buffers of the
// file paths are kept for the life of the process and handed out again (best fit, at most 1.5x oversize).
struct PinnedPool {
    struct Blk { uint8_t* p; uint64_t bytes; };
    std::mutex mu;
    std::vector<Blk> free_list;
    uint8_t* get(pna_ctx* ctx, uint64_t bytes, uint64_t* got) {
        {
            std::lock_guard<std::mutex> g(mu);
            int best = -1;
            for (int i = 0; i < (int)free_list.size(); i++)
                if (free_list[i].bytes >= bytes && free_list[i].bytes <= bytes + bytes / 2 + ((uint64_t)64 << 20) &&
                    (best < 0 || free_list[i].bytes < free_list[best].bytes)) best = i;
            if (best >= 0) { Blk b = free_list[best]; free_list.erase(free_list.begin() + best); *got = b.bytes; return b.p; }
        }
        *got = bytes;
        return (uint8_t*)pna_cuda_host_alloc(ctx, bytes);
    }
    void put(pna_ctx* ctx, uint8_t* p, uint64_t bytes) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        if (free_list.size() >= 6) { pna_cuda_host_free(ctx, free_list.front().p); free_list.erase(free_list.begin()); }
        free_list.push_back({p, bytes});
    }
};
static PinnedPool g_pinned;
static const uint8_t SIGNATURE[8] = {0x89, 'P', 'N', 'A', 0x0D, 0x0A, 0x1A, 0x0A};
static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static inline bool ty_is(const RawChunk& c, const char* t) { return memcmp(c.ty, t, 4) == 0; }

// ---- contexts: one pna_ctx per worker thread and device, kept for the life of the process
static std::mutex g_pool_mu;
static std::map<int, std::vector<pna_ctx*>> g_pool;
struct CtxLease {
    pna_ctx* ctx = nullptr;
    int device;
    explicit CtxLease(int dev) : device(dev) {
        {
            std::lock_guard<std::mutex> g(g_pool_mu);
            auto& v = g_pool[dev];
            if (!v.empty()) { ctx = v.back(); v.pop_back(); }
        }
        if (!ctx && pna_cuda_init(&ctx, &dev, 1) != PNA_OK) throw Error(PNA_E_CUDA, "pna_cuda_init failed: no usable sm_100 device (there is no CPU fallback)");
    }
    ~CtxLease() { if (ctx) { std::lock_guard<std::mutex> g(g_pool_mu); g_pool[device].push_back(ctx); } }
};
static void ck(pna_ctx* ctx, int rc, const char* what) {
    if (rc != PNA_OK) throw Error(rc, std::string(what) + ": " + pna_cuda_strerror(rc) + " / " + pna_cuda_last_error(ctx));
}

// ---- index pass: walk chunk frames without touching chunk data (bytes::skip_chunk semantics, lib/src/bytes.rs:90)
static inline uint32_t ty32(const char* t) { uint32_t v; memcpy(&v, t, 4); return v; }
static inline uint32_t ty32(const RawChunk& c) { return ty32(c.ty); }
static const uint32_t T_AEND = ty32("AEND"), T_ANXT = ty32("ANXT"), T_FHED = ty32("FHED"), T_SHED = ty32("SHED"), T_FEND = ty32("FEND"),
                      T_SEND = ty32("SEND"), T_FDAT = ty32("FDAT"), T_SDAT = ty32("SDAT"), T_PHSF = ty32("PHSF"), T_fSIZ = ty32("fSIZ");
static inline bool ty_letters(uint32_t t) {   // every byte an ASCII letter (chunk/types.rs:204): bit 6 set, low five bits in 1..26
    const uint32_t low = t & 0x1F1F1F1Fu;
    return (t & 0xC0C0C0C0u) == 0x40404040u && !((low - 0x01010101u) & 0x80808080u) && !((0x1A1A1A1Au - low) & 0x80808080u);
}
// threads this process may use: its affinity mask (a rank pinned to a slice of the box must not plan for all of it)
static unsigned usable_cpus() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) return (unsigned)n; }
    return std::max(1u, std::thread::hardware_concurrency());
}
// serial walk from `pos` until `stop` (the first chunk that starts at or behind it is not taken); returns where it stopped
static size_t walk_chunks(const uint8_t* buf, size_t len, size_t pos, size_t stop, std::vector<RawChunk>& out, bool speculative, bool* broken) {
    while (pos < len && pos < stop) {
        RawChunk c;
        const bool ok_hdr = len - pos >= 12;
        if (ok_hdr) { c.len = be32(buf + pos); memcpy(c.ty, buf + pos + 4, 4); }
        const bool ok = ok_hdr && ty_letters(ty32(c.ty)) && len - pos - 12 >= c.len;
        if (!ok) {
            if (speculative) { *broken = true; return pos; }
            if (!ok_hdr) throw Error(PNA_E_UNEXPECTED_EOF, "truncated chunk");
            if (!ty_letters(ty32(c.ty))) throw Error(PNA_E_INVALID_DATA, "invalid chunk type");
            throw Error(PNA_E_UNEXPECTED_EOF, "truncated chunk body");
        }
        c.off = pos + 8;
        c.crc = be32(buf + pos + 8 + c.len);
        out.push_back(c);
        pos += 12 + (size_t)c.len;
    }
    return pos;
}
// The walk is a pointer chase (every header is a cache miss, ~70 ns a chunk): archives with millions of small chunks are
// indexed by several threads.  Thread k looks, from the start of its region, for an offset from which eight consecutive
// well-formed chunk headers follow, and walks on from there; its list is taken only when the previous thread's walk
// lands EXACTLY on that offset -- then the result is what the serial walk produces -- and re-walked serially otherwise.
static void index_chunks(const uint8_t* buf, size_t len, size_t pos, std::vector<RawChunk>& out) {
    const size_t span = len > pos ? len - pos : 0;
    unsigned nt = std::min(16u, std::max(1u, usable_cpus()));
    if (span < ((size_t)32 << 20) || getenv("PNA_INDEX_SERIAL")) nt = 1;
    if (nt == 1) {
        out.reserve(std::min<size_t>(len / 256 + 16, (size_t)1 << 24));
        walk_chunks(buf, len, pos, len, out, false, nullptr);
        return;
    }
    struct Part { size_t start = SIZE_MAX, end = 0; bool broken = false; std::vector<RawChunk> list; };
    std::vector<Part> parts(nt);
    std::vector<size_t> bound(nt + 1);
    for (unsigned k = 0; k <= nt; k++) bound[k] = pos + span / nt * k;
    bound[nt] = len;
    auto plausible = [&](size_t p) {
        for (int hop = 0; hop < 8 && p < len; hop++) {
            if (len - p < 12) return false;
            const uint32_t cl = be32(buf + p);
            if (!ty_letters(ty32((const char*)buf + p + 4)) || len - p - 12 < cl) return false;
            p += 12 + (size_t)cl;
        }
        return true;
    };
    auto run = [&](unsigned k) {
        Part& P = parts[k];
        if (k == 0) P.start = pos;
        else {
            const size_t lim = std::min(bound[k] + ((size_t)64 << 10), bound[k + 1]);
            for (size_t p = bound[k]; p < lim; p++) if (plausible(p)) { P.start = p; break; }
        }
        if (P.start == SIZE_MAX) return;
        P.list.reserve(span / nt / 256 + 16);
        P.end = walk_chunks(buf, len, P.start, bound[k + 1], P.list, true, &P.broken);
    };
    std::vector<std::thread> th;
    for (unsigned k = 1; k < nt; k++) th.emplace_back(run, k);
    run(0);
    for (auto& t : th) t.join();
    size_t total = 0;
    for (const Part& P : parts) total += P.list.size();
    out.reserve(total + 16);
    size_t at = pos;   // the authoritative walk stands here
    for (unsigned k = 0; k < nt; k++) {
        Part& P = parts[k];
        if (P.start == at) {
            out.insert(out.end(), P.list.begin(), P.list.end());
            at = P.end;
            if (P.broken) at = walk_chunks(buf, len, at, bound[k + 1], out, false, nullptr);   // reports the error the serial walk reports
        } else if (at < bound[k + 1]) at = walk_chunks(buf, len, at, bound[k + 1], out, false, nullptr);
    }
    if (at < len) walk_chunks(buf, len, at, len, out, false, nullptr);
}
static inline bool is_critical(const RawChunk& c) { return (c.ty[0] & 0x20) == 0; }

// next_raw_item (archive/read.rs:46-73) + TryFrom<RawEntry> (entry.rs:665-737, 757-886) for the entries that START in
// chunk range [i0, i1); `limit` = index of the first AEND / ANXT (nothing behind it is grouped).  Body spans go to
// pool[pool_at ...] (the slice reserved for this range); equal PHSF strings are shared.
static void group_range(const uint8_t* buf, const std::vector<RawChunk>& ch, size_t i0, size_t i1, size_t limit, std::vector<EntryInfo>& out,
                        pna_span* pool, size_t pool_at) {
    std::shared_ptr<const std::string> last_phsf;
    size_t i = i0;
    while (i < i1 && i < limit) {
        const RawChunk& c = ch[i];
        const uint32_t ct = ty32(c);
        const bool normal = ct == T_FHED, solid = ct == T_SHED;
        if (!normal && !solid) { i++; continue; }   // AHED and archive-level ancillary chunks
        out.emplace_back();
        EntryInfo& e = out.back();
        e.kind = solid ? 1 : 0;
        e.chunk_begin = (uint32_t)i;
        const uint8_t* h = buf + c.off;
        e.header_data = h; e.header_len = c.len;
        if (normal) {
            if (c.len < 6) throw Error(PNA_E_INVALID_DATA, "entry header too short");
            if (h[0] != 0 || h[1] != 0) throw Error(PNA_E_UNSUPPORTED, "entry version is not supported");
            e.data_kind = h[2]; e.compression = h[3]; e.encryption = h[4]; e.cipher_mode = h[5];
            e.name.assign((const char*)h + 6, c.len - 6);
        } else {
            if (c.len != 5) throw Error(PNA_E_INVALID_DATA, "solid header must be 5 bytes");
            if (h[0] != 0 || h[1] != 0) throw Error(PNA_E_UNSUPPORTED, "entry version is not supported");
            e.compression = h[2]; e.encryption = h[3]; e.cipher_mode = h[4];
        }
        const uint32_t end_ty = normal ? T_FEND : T_SEND, dat_ty = normal ? T_FDAT : T_SDAT;
        const size_t body_first = pool_at;
        size_t j = i + 1;
        bool closed = false;
        for (; j < limit; j++) {                      // ANXT / AEND at `limit`: the entry continues in the next part (archive/read.rs:118)
            const RawChunk& d = ch[j];
            const uint32_t dt = ty32(d);
            if (dt == end_ty) { closed = true; j++; break; }
            if (dt == dat_ty) { pool[pool_at++] = {buf + d.off, d.len}; e.compressed_size += d.len; }
            else if (dt == T_PHSF) {
                if (!last_phsf || last_phsf->size() != d.len || memcmp(last_phsf->data(), buf + d.off, d.len) != 0)
                    last_phsf = std::make_shared<const std::string>((const char*)buf + d.off, d.len);
                e.phsf_ = last_phsf; e.has_phsf = true;
            } else if (normal && dt == T_fSIZ) {
                uint64_t v = 0;
                const uint32_t n = d.len > 16 ? 16 : d.len;       // u128_from_be_bytes_last
                for (uint32_t k = d.len - n; k < d.len; k++) v = (v << 8) | buf[d.off + k];
                e.raw_file_size = v;
            } else if (is_critical(d)) throw Error(PNA_E_INVALID_DATA, "unknown critical chunk type");   // entry.rs:716,848
        }
        if (!closed) throw Error(PNA_E_UNEXPECTED_EOF, "entry without end chunk");
        e.bodies.p = pool + body_first;
        e.bodies.n = (uint32_t)(pool_at - body_first);
        e.chunk_end = (uint32_t)j;
        i = j;
    }
}
// Grouping reads every entry's header / size / PHSF chunk bodies (a few cache misses per entry): ranges of the chunk
// list that start at an entry header are grouped by several threads and concatenated in order.
static void group_entries(const uint8_t* buf, const std::vector<RawChunk>& ch, std::vector<EntryInfo>& out, std::vector<pna_span>& pool) {
    const size_t nc = ch.size();
    size_t limit = nc;
    for (size_t i = 0; i < nc; i++) { const uint32_t t = ty32(ch[i]); if (t == T_AEND || t == T_ANXT) { limit = i; break; } }
    unsigned nt = std::min(16u, std::max(1u, usable_cpus()));
    if (limit < 65536 || getenv("PNA_INDEX_SERIAL")) nt = 1;
    // range starts: the first entry header at or behind k * limit / nt.  An entry header can only be cut off from its
    // chunks by a range start that lies INSIDE the entry, so starts are moved forward to the next header that follows an
    // end chunk (FEND / SEND), i.e. a header at top level.
    std::vector<size_t> start(nt + 1, limit);
    start[0] = 0;
    for (unsigned k = 1; k < nt; k++) {
        size_t i = std::max(start[k - 1], limit / nt * k);
        while (i < limit) {
            const uint32_t t = ty32(ch[i]);
            if ((t == T_FHED || t == T_SHED) && i > 0 && (ty32(ch[i - 1]) == T_FEND || ty32(ch[i - 1]) == T_SEND)) break;
            i++;
        }
        start[k] = i;
    }
    std::vector<size_t> dat_before(nt + 1, 0);
    {
        unsigned k = 0;
        size_t cnt = 0;
        for (size_t i = 0; i < limit; i++) {
            while (k < nt && start[k + 1] <= i) { k++; dat_before[k] = cnt; }
            const uint32_t t = ty32(ch[i]);
            cnt += t == T_FDAT || t == T_SDAT;
        }
        while (k < nt) { k++; dat_before[k] = cnt; }
    }
    pool.assign(dat_before[nt], pna_span{nullptr, 0});
    std::vector<std::vector<EntryInfo>> parts(nt);
    std::vector<std::string> err(nt);
    std::vector<int> err_kind(nt, 0);
    auto run = [&](unsigned k) {
        try {
            parts[k].reserve((start[k + 1] - start[k]) / 4 + 4);
            group_range(buf, ch, start[k], start[k + 1], limit, parts[k], pool.data(), dat_before[k]);
        } catch (const Error& e) { err[k] = e.what(); err_kind[k] = e.kind ? e.kind : PNA_E_INTERNAL; }
    };
    std::vector<std::thread> th;
    for (unsigned k = 1; k < nt; k++) th.emplace_back(run, k);
    run(0);
    for (auto& t : th) t.join();
    size_t total = out.size();
    for (unsigned k = 0; k < nt; k++) {
        if (err_kind[k]) throw Error(err_kind[k], err[k]);   // the first failing range in archive order = the serial walk's error
        total += parts[k].size();
    }
    out.reserve(total);
    for (unsigned k = 0; k < nt; k++)
        for (EntryInfo& e : parts[k]) out.push_back(std::move(e));
}

Archive Archive::read_header_from_slice(const uint8_t* buf, size_t len) {
    Archive a;
    if (len < 8 || memcmp(buf, SIGNATURE, 8) != 0) throw Error(PNA_E_INVALID_DATA, "it is not PNA");
    a.buf_ = buf; a.len_ = len;
    const auto t0 = std::chrono::steady_clock::now();
    index_chunks(buf, len, 8, a.chunks_);
    const auto t1 = std::chrono::steady_clock::now();
    if (a.chunks_.empty() || !ty_is(a.chunks_[0], "AHED")) throw Error(PNA_E_INVALID_DATA, "expected `AHED` chunk");
    if (a.chunks_[0].len != 8) throw Error(PNA_E_INVALID_DATA, "bad archive header");
    const uint8_t* h = buf + a.chunks_[0].off;
    if (h[0] != 0) throw Error(PNA_E_UNSUPPORTED, "archive version is not supported");   // archive/header.rs:47
    a.archive_number_ = be32(h + 4);
    group_entries(buf, a.chunks_, a.entries_, a.body_pool_);
    if (getenv("PNA_HOST_TRACE"))
        fprintf(stderr, "[pna_host] index pass: %zu chunks walked in %.1f ms, %zu entries grouped in %.1f ms\n", a.chunks_.size(),
                std::chrono::duration<double, std::milli>(t1 - t0).count(), a.entries_.size(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    return a;
}

struct JoinCopy { uint8_t* dst; const uint8_t* src; uint64_t n; };
static void join_copy(const std::vector<JoinCopy>& copies);   // several threads; defined with the split writer
Archive Archive::read_multipart(const pna_span* parts, size_t n_parts, int pinned_device) {
    if (n_parts == 0) throw Error(PNA_E_INVALID_INPUT, "no archive part given");
    std::vector<std::vector<RawChunk>> lists(n_parts);
    uint64_t total = 8;
    for (size_t k = 0; k < n_parts; k++) {
        const uint8_t* b = parts[k].ptr;
        const size_t len = (size_t)parts[k].len;
        if (len < 8 || memcmp(b, SIGNATURE, 8) != 0) throw Error(PNA_E_INVALID_DATA, "it is not PNA");
        std::vector<RawChunk>& ch = lists[k];
        index_chunks(b, len, 8, ch);
        if (ch.empty() || !ty_is(ch[0], "AHED") || ch[0].len != 8) throw Error(PNA_E_INVALID_DATA, "expected `AHED` chunk");
        if (b[ch[0].off] != 0) throw Error(PNA_E_UNSUPPORTED, "archive version is not supported");
        if (be32(b + ch[0].off + 4) != (uint32_t)k) throw Error(PNA_E_INVALID_DATA, "part " + std::to_string(k) + " carries archive number " + std::to_string(be32(b + ch[0].off + 4)));
        if (ty32(ch.back()) != T_AEND) throw Error(PNA_E_UNEXPECTED_EOF, "part without `AEND`");
        bool has_next = false;
        for (const RawChunk& c : ch) { has_next |= ty32(c) == T_ANXT; total += 12 + (uint64_t)c.len; }
        if (has_next && k + 1 == n_parts) throw Error(PNA_E_UNEXPECTED_EOF, "next part missing");
        if (!has_next && k + 1 < n_parts) throw Error(PNA_E_INVALID_DATA, "part does not announce a next archive (`ANXT`)");
    }
    std::shared_ptr<uint8_t> mem;
    if (pinned_device >= 0) {              // joined stream in pinned memory from the process-wide pool: its upload is true DMA
        CtxLease L(pinned_device);
        pna_ctx* const pctx = L.ctx;
        uint64_t got = 0;
        uint8_t* p = g_pinned.get(pctx, total, &got);
        if (!p) throw Error(PNA_E_OOM, "pinned buffer for the joined parts");
        mem = std::shared_ptr<uint8_t>(p, [pctx, got](uint8_t* q) { g_pinned.put(pctx, q, got); });
    } else mem = std::shared_ptr<uint8_t>(new uint8_t[total], std::default_delete<uint8_t[]>());
    uint8_t* w = mem.get();
    memcpy(w, SIGNATURE, 8); w += 8;
    std::vector<JoinCopy> copies;          // runs of consecutive source frames; copied by several threads below
    auto put = [&](size_t k, const RawChunk& c) {
        const uint8_t* src = parts[k].ptr + c.off - 8;
        const uint64_t n = 12 + (uint64_t)c.len;
        if (!copies.empty() && copies.back().src + copies.back().n == src) copies.back().n += n;
        else copies.push_back({w, src, n});
        w += n;
    };
    auto archive_level = [](const RawChunk& c) { const uint32_t t = ty32(c); return t == T_AEND || t == T_ANXT || t == ty32("AHED"); };
    put(0, lists[0][0]);
    for (size_t k = 0; k < n_parts; k++)
        for (const RawChunk& c : lists[k]) if (!archive_level(c)) put(k, c);
    put(n_parts - 1, lists[n_parts - 1].back());
    for (size_t k = 0; k < n_parts; k++)              // behind AEND: indexed and CRC-checked, never grouped
        for (size_t i = 0; i < lists[k].size(); i++) {
            const RawChunk& c = lists[k][i];
            const bool placed = (k == 0 && i == 0) || (k + 1 == n_parts && i + 1 == lists[k].size());
            if (archive_level(c) && !placed) put(k, c);
        }
    join_copy(copies);
    Archive a = read_header_from_slice(mem.get(), (size_t)total);
    a.joined_ = std::move(mem);
    return a;
}

static bool cipher_supported(const EntryInfo& e) {
    return (e.encryption == PNA_ENCRYPTION_AES || e.encryption == PNA_ENCRYPTION_CAMELLIA) &&
           (e.cipher_mode == PNA_CIPHER_CBC || e.cipher_mode == PNA_CIPHER_CTR || e.cipher_mode == PNA_CIPHER_GCM);
}
// decrypt_reader's key path (entry/read.rs:45-78): returns false with a per-entry status when the key is unavailable
static int32_t fill_desc(const EntryInfo& e, const ReadOptions& opt, pna_decode_desc& d) {
    memset(&d, 0, sizeof d);
    d.bodies = e.bodies.data(); d.n_bodies = (uint32_t)e.bodies.size();
    d.compression = e.compression; d.encryption = e.encryption; d.cipher_mode = e.cipher_mode;
    d.raw_size_hint = e.raw_file_size;
    if (e.encryption != PNA_ENCRYPTION_NO && cipher_supported(e)) {
        if (!opt.has_password) return PNA_E_INVALID_INPUT;                 // "password was not provided"
        if (!e.has_phsf) return PNA_E_INVALID_DATA;                        // "`PHSF` chunk not found"
        auto it = opt.keys.find(e.phsf());
        if (it == opt.keys.end()) return PNA_E_INVALID_INPUT;
        memcpy(d.key, it->second.data(), 32);
        if (e.cipher_mode == PNA_CIPHER_GCM) {
            // GCM branch of decrypt_reader (entry/read.rs:105-139): stream header, key confirmation, per-stream key
            uint8_t hdr[75];
            uint64_t got = 0;
            for (const pna_span& b : e.bodies) {
                if (got >= sizeof hdr) break;
                const uint64_t k = std::min<uint64_t>(b.len, sizeof hdr - got);
                if (k) memcpy(hdr + got, b.ptr, k);
                got += k;
            }
            return pna_cuda_gcm_stream_key(it->second.data(), hdr, got, (const uint8_t*)(e.kind == 1 ? "SHED" : "FHED"), e.header_data,
                                           e.header_len, (const uint8_t*)e.phsf().data(), e.phsf().size(), d.key);
        }
    }
    return PNA_OK;
}

void Archive::prepare(const ReadOptions& opt, int device) {
    if (prepared_) return;
    CtxLease L(device);
    inner_.clear(); refs_.clear(); files_.clear();
    // 1. solid entries: decode the SDAT stream (sizing call, then the real one), index + verify the inner chunks
    for (size_t k = 0; k < entries_.size(); k++) {
        const EntryInfo& e = entries_[k];
        if (e.kind != 1) continue;
        Inner in;
        pna_decode_desc d;
        int32_t st = fill_desc(e, opt, d);
        if (st != PNA_OK) throw Error(st, "solid entry: key unavailable");
        d.raw_size_hint = UINT64_MAX;                                       // solid streams carry no size: the plan sizes itself
        pna_plan* plan = nullptr;
        // the SDAT bodies all lie inside the archive buffer: declared, so the tens of thousands of 32 KiB bodies of a large solid
        // entry travel in a few copies instead of one cudaMemcpyAsync each
        ck(L.ctx, pna_cuda_decode_plan_create_in_image(L.ctx, &d, 1, buf_, len_, nullptr, nullptr, nullptr, 0, &plan), "solid plan");
        in.plan = std::shared_ptr<pna_plan>(plan, [](pna_plan* p) { pna_cuda_plan_destroy(p); });
        ck(L.ctx, pna_cuda_decode_plan_run(plan), "solid decode");
        uint64_t dec_len = 0;
        ck(L.ctx, pna_cuda_decode_plan_lengths(plan, &dec_len, &st), "solid lengths");
        if (st != PNA_OK) throw Error(st, "solid entry: decode failed");
        {   // host copy for the index pass: a pinned buffer from the pool (true DMA, no zero fill)
            pna_ctx* const pctx = L.ctx;
            uint64_t got = 0;
            uint8_t* p = g_pinned.get(pctx, dec_len + 64, &got);
            if (!p) throw Error(PNA_E_OOM, "pinned buffer for the solid stream");
            in.mem = std::shared_ptr<uint8_t>(p, [pctx, got](uint8_t* q) { g_pinned.put(pctx, q, got); });
        }
        pna_buf ob{in.mem.get(), dec_len, 0};
        ck(L.ctx, pna_cuda_decode_plan_fetch(plan, &ob, &st), "solid fetch");
        if (st != PNA_OK) throw Error(st, "solid entry: decode failed");
        in.len = ob.len;
        index_chunks(in.data(), in.len, 0, in.chunks);       // entry.rs:401-423: chunks, CRC checked as read
        if (!in.chunks.empty()) {                                           // ... on the copy that is still in HBM
            std::vector<uint64_t> off(in.chunks.size()), len(in.chunks.size());
            std::vector<uint32_t> crc(in.chunks.size());
            for (size_t c = 0; c < in.chunks.size(); c++) { off[c] = in.chunks[c].off - 4; len[c] = (uint64_t)in.chunks[c].len + 4; }
            ck(L.ctx, pna_cuda_decode_plan_crc32_out(plan, 0, off.data(), len.data(), (uint32_t)off.size(), crc.data()), "solid inner crc");
            for (size_t c = 0; c < in.chunks.size(); c++)
                if (crc[c] != in.chunks[c].crc) throw Error(PNA_E_INVALID_DATA, "broken chunk (inside solid entry)");
        }
        group_entries(in.data(), in.chunks, in.entries, in.body_pool);
        inner_.push_back(std::move(in));
    }
    // 2. FILE entries in archive order
    uint32_t solid_no = 0;
    for (size_t k = 0; k < entries_.size(); k++) {
        const EntryInfo& e = entries_[k];
        if (e.kind == 1) {
            const Inner& in = inner_[solid_no];
            for (size_t q = 0; q < in.entries.size(); q++)
                if (in.entries[q].kind == 0 && in.entries[q].data_kind == (uint8_t)DataKind::File) refs_.push_back({solid_no + 1, (uint32_t)q});
            solid_no++;
        } else if (e.data_kind == (uint8_t)DataKind::File) refs_.push_back({0, (uint32_t)k});
    }
    // 3. sizes: fSIZ, else a sizing pass (cap 0 -> PNA_E_NOSPACE with the length)
    files_.resize(refs_.size());
    std::vector<uint32_t> unknown;
    for (size_t i = 0; i < refs_.size(); i++) {
        const EntryInfo& e = refs_[i].owner ? inner_[refs_[i].owner - 1].entries[refs_[i].entry] : entries_[refs_[i].entry];
        files_[i].name = e.name;
        // fSIZ is untrusted input and only a hint (the reference never sizes anything from it): values the library would
        // ignore are treated as absent, so the sizing pass below learns the real length
        const bool trusted = pna_cuda_size_hint_trusted(e.compression, e.compressed_size, e.raw_file_size) != 0;
        files_[i].size = trusted ? e.raw_file_size : UINT64_MAX;
        if (!trusted) unknown.push_back((uint32_t)i);
    }
    if (!unknown.empty()) {
        std::vector<pna_decode_desc> descs(unknown.size());
        std::vector<pna_buf> bufs(unknown.size(), pna_buf{nullptr, 0, 0});
        std::vector<int32_t> st(unknown.size(), 0), pre(unknown.size(), 0);
        for (size_t u = 0; u < unknown.size(); u++) {
            const FileRef r = refs_[unknown[u]];
            const EntryInfo& e = r.owner ? inner_[r.owner - 1].entries[r.entry] : entries_[r.entry];
            pre[u] = fill_desc(e, opt, descs[u]);
            if (pre[u] != PNA_OK) { descs[u].n_bodies = 0; descs[u].encryption = 0; descs[u].compression = 0; }
        }
        ck(L.ctx, pna_cuda_decode_batch(L.ctx, descs.data(), (uint32_t)descs.size(), bufs.data(), st.data()), "sizing pass");
        for (size_t u = 0; u < unknown.size(); u++) {
            FileOut& f = files_[unknown[u]];
            if (pre[u] != PNA_OK) { f.status = pre[u]; f.size = 0; }
            else if (st[u] == PNA_OK || st[u] == PNA_E_NOSPACE) f.size = bufs[u].len;
            else { f.status = st[u]; f.size = 0; }
        }
    }
    prepared_ = true;
}

void Archive::extract_files(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                            uint64_t group_bytes, bool verify) {
    extract_files(opt, out, offsets, status, std::vector<int>{device}, workers, group_bytes, verify);
}
void Archive::extract_files(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, const std::vector<int>& devices,
                            int workers, uint64_t group_bytes, bool verify) {
    if (devices.empty()) throw Error(PNA_E_BAD_ARG, "no device given");
    if (!prepared_) prepare(opt, devices[0]);
    crc_next_chunk_ = 0;
    extract_range(opt, out, offsets, status, devices, workers, group_bytes, verify, 0, refs_.size());
}
void Archive::extract_range(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                            uint64_t group_bytes, bool verify, size_t first, size_t last) {
    extract_range(opt, out, offsets, status, std::vector<int>{device}, workers, group_bytes, verify, first, last);
}

// Files [first, last) of files(): file i lands at out + (offsets[i] - offsets[first]).  Windows must be taken in order
// when verify is set: every call checks the chunks between the previous window's last entry and its own.
// Several devices: the entry groups are handed out dynamically (one shared counter) to `workers` threads PER DEVICE, each with its
// own pair of contexts on that device -- the partition by entry of SURVEY 8e with the load balance of a work queue; uploads take
// turns per device (every GPU has its own PCIe link).
void Archive::extract_range(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, const std::vector<int>& devices,
                            int workers, uint64_t group_bytes, bool verify, size_t first, size_t last) {
    if (devices.empty()) throw Error(PNA_E_BAD_ARG, "no device given");
    const int device = devices[0];
    if (!prepared_) prepare(opt, device);
    const size_t n = last;
    const uint64_t o0 = first < refs_.size() ? offsets[first] : 0;
    // groups of consecutive files of the same owner, cut when the compressed bytes reach group_bytes.  The first groups
    // ramp up (1/4, 1/2 of the size): the pipeline's first download starts that much earlier.
    struct Group { size_t lo, hi; };
    std::vector<Group> groups;
    {
        size_t lo = first;
        uint64_t acc = 0;
        for (size_t i = first; i < n; i++) {
            const EntryInfo& e = refs_[i].owner ? inner_[refs_[i].owner - 1].entries[refs_[i].entry] : entries_[refs_[i].entry];
            const uint64_t limit = groups.size() == 0 ? group_bytes / 4 : groups.size() == 1 ? group_bytes / 2 : group_bytes;
            const bool cut = i > lo && (refs_[i].owner != refs_[lo].owner || acc >= limit);
            if (cut) { groups.push_back({lo, i}); lo = i; acc = 0; }
            acc += e.compressed_size;
        }
        if (lo < n) groups.push_back({lo, n});
    }
    // chunk ranges of the top-level archive covered by each top-level group (contiguous, together all chunks)
    std::vector<std::pair<uint32_t, uint32_t>> crange(groups.size(), {0, 0});
    if (verify) {
        int64_t last_top_ref = -1;                 // the archive's last top-level FILE: its group also covers the trailing chunks
        for (size_t i = refs_.size(); i-- > 0;) if (refs_[i].owner == 0) { last_top_ref = (int64_t)i; break; }
        uint32_t prev_end = crc_next_chunk_;
        for (size_t g = 0; g < groups.size(); g++) {
            if (refs_[groups[g].lo].owner != 0) continue;
            const bool closes = (int64_t)groups[g].lo <= last_top_ref && last_top_ref < (int64_t)groups[g].hi;
            const uint32_t end = closes ? (uint32_t)chunks_.size() : entries_[refs_[groups[g].hi - 1].entry].chunk_end;
            crange[g] = {prev_end, end};
            prev_end = end;
        }
        crc_next_chunk_ = prev_end;
        if (last_top_ref < 0 && crc_next_chunk_ == 0 && !chunks_.empty()) {   // no top-level FILE entry at all: verify the archive's chunks in one call
            crc_next_chunk_ = (uint32_t)chunks_.size();
            CtxLease L(device);
            std::vector<uint64_t> off(chunks_.size()), len(chunks_.size());
            std::vector<uint32_t> crc(chunks_.size());
            for (size_t c = 0; c < chunks_.size(); c++) { off[c] = chunks_[c].off - 4; len[c] = (uint64_t)chunks_[c].len + 4; }
            ck(L.ctx, pna_cuda_crc32_image(L.ctx, buf_, len_, off.data(), len.data(), (uint32_t)off.size(), crc.data()), "archive crc");
            for (size_t c = 0; c < chunks_.size(); c++) if (crc[c] != chunks_[c].crc) throw Error(PNA_E_INVALID_DATA, "broken chunk");
        }
    }
    std::atomic<size_t> next{0};
    const auto t_begin = std::chrono::steady_clock::now();
    const bool trace = getenv("PNA_HOST_TRACE") != nullptr;
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_begin).count(); };
    std::mutex err_mu;
    std::vector<std::mutex> h2d_mus(devices.size());
    std::string err_msg;
    int err_kind = 0;
    // One group in flight per context; every worker thread keeps TWO contexts and software-pipelines its groups:
    //   create(g) [H2D]  ->  run(g) [kernels, asynchronous]  ->  fetch(g-1) [D2H of the previous group]
    // so the upload and the kernels of group g overlap the download of group g-1 (PCIe is full duplex, kernels are
    // on their own streams).  Everything a stage needs lives in Stage, which stays alive until its fetch.
    struct Stage {
        size_t g = 0;
        uint32_t m = 0;
        bool top = false;
        pna_plan* plan = nullptr;
        std::vector<pna_decode_desc> descs;
        std::vector<pna_buf> bufs;
        std::vector<int32_t> st, pre;
        std::vector<pna_span> spans;
        std::vector<uint32_t> expect;
        std::vector<int32_t> owner;
    };
    // files of a solid entry that are stored as they are (what the reference's solid writer produces): their bytes are
    // ranges of the decoded stream that is still in HBM -- copy them out from there
    auto store_ranges = [&](CtxLease& L, size_t g) -> bool {
        const Group G = groups[g];
        const uint32_t owner = refs_[G.lo].owner;
        if (owner == 0) return false;
        const Inner& in = inner_[owner - 1];
        if (!in.plan) return false;
        for (size_t i = G.lo; i < G.hi; i++) {
            const EntryInfo& e = in.entries[refs_[i].entry];
            if (files_[i].status || e.compression != PNA_COMPRESSION_NO || e.encryption != PNA_ENCRYPTION_NO) return false;
        }
        std::vector<uint64_t> src, len;
        std::vector<uint8_t*> dst;
        for (size_t i = G.lo; i < G.hi; i++) {
            const EntryInfo& e = in.entries[refs_[i].entry];
            const uint64_t cap = offsets[i + 1] - offsets[i];
            uint64_t at = 0;
            if (e.compressed_size > cap) { status[i] = PNA_E_NOSPACE; continue; }
            for (const pna_span& b : e.bodies) {
                src.push_back((uint64_t)(b.ptr - in.data())); len.push_back(b.len); dst.push_back(out + (offsets[i] - o0) + at);
                at += b.len;
            }
            status[i] = PNA_OK;
        }
        ck(L.ctx, pna_cuda_decode_plan_fetch_ranges(in.plan.get(), 0, src.data(), len.data(), dst.data(), (uint32_t)src.size()), "solid ranges");
        return true;
    };
    auto issue = [&](CtxLease& L, Stage& S, size_t g, std::mutex& h2d_mu) {
        const Group G = groups[g];
        const uint32_t m = (uint32_t)(G.hi - G.lo);
        S.g = g; S.m = m; S.plan = nullptr;
        if (store_ranges(L, g)) return;
        S.descs.assign(m, pna_decode_desc{}); S.bufs.resize(m); S.st.assign(m, 0); S.pre.assign(m, 0);
        for (uint32_t k = 0; k < m; k++) {
            const FileRef r = refs_[G.lo + k];
            const EntryInfo& e = r.owner ? inner_[r.owner - 1].entries[r.entry] : entries_[r.entry];
            S.pre[k] = files_[G.lo + k].status ? files_[G.lo + k].status : fill_desc(e, opt, S.descs[k]);
            if (S.pre[k] != PNA_OK) { memset(&S.descs[k], 0, sizeof S.descs[k]); S.descs[k].raw_size_hint = 0; }
            else S.descs[k].raw_size_hint = files_[G.lo + k].size;
            S.bufs[k] = pna_buf{out + (offsets[G.lo + k] - o0), offsets[G.lo + k + 1] - offsets[G.lo + k], 0};
        }
        S.top = refs_[G.lo].owner == 0;
        const auto t0 = std::chrono::steady_clock::now();
        {
            // one upload at a time: PCIe is the shared resource -- unless the group is thousands of small entries, whose plan
            // construction is host work (descriptors, chunk spans, CRC tiles) that dwarfs the upload: those build in parallel
            std::unique_lock<std::mutex> h2d(h2d_mu, std::defer_lock);
            if (m <= 2048) h2d.lock();
            if (verify && S.top && crange[g].second > crange[g].first) {
                const uint32_t c0 = crange[g].first, c1 = crange[g].second, nc = c1 - c0;
                S.spans.resize(nc); S.expect.resize(nc); S.owner.assign(nc, -1);
                for (uint32_t c = 0; c < nc; c++) { S.spans[c] = {buf_ + chunks_[c0 + c].off - 4, (uint64_t)chunks_[c0 + c].len + 4}; S.expect[c] = chunks_[c0 + c].crc; }
                for (uint32_t k = 0; k < m; k++) {
                    const EntryInfo& e = entries_[refs_[G.lo + k].entry];
                    for (uint32_t c = e.chunk_begin; c < e.chunk_end; c++) S.owner[c - c0] = (int32_t)k;
                }
                // spans and bodies all lie inside the archive buffer: declared, so neighbouring ones travel in one copy
                ck(L.ctx, pna_cuda_decode_plan_create_in_image(L.ctx, S.descs.data(), m, buf_, len_, S.spans.data(), S.expect.data(), S.owner.data(), nc, &S.plan), "plan_create_in_image");
            } else if (S.top) ck(L.ctx, pna_cuda_decode_plan_create_in_image(L.ctx, S.descs.data(), m, buf_, len_, nullptr, nullptr, nullptr, 0, &S.plan), "plan_create_in_image");
            else {
                const Inner& in = inner_[refs_[G.lo].owner - 1];
                ck(L.ctx, pna_cuda_decode_plan_create_in_image(L.ctx, S.descs.data(), m, in.data(), in.len, nullptr, nullptr, nullptr, 0, &S.plan), "plan_create_in_image");
            }
        }
        const auto t1 = std::chrono::steady_clock::now();
        const int rc = pna_cuda_decode_plan_run(S.plan);
        if (rc != PNA_OK) { pna_cuda_plan_destroy(S.plan); S.plan = nullptr; ck(L.ctx, rc, "decode_plan_run"); }
        if (trace) fprintf(stderr, "[pna_host] group %zu (%u entries): create/H2D %.1f-%.1f ms, run/launch -%.1f ms\n", g, m, ms_since(t0), ms_since(t1),
                           ms_since(std::chrono::steady_clock::now()));
    };
    auto collect = [&](CtxLease& L, Stage& S) {
        if (!S.plan) return;
        const Group G = groups[S.g];
        const auto t0 = std::chrono::steady_clock::now();
        int rc = pna_cuda_decode_plan_fetch(S.plan, S.bufs.data(), S.st.data());
        uint32_t broken = 0;
        if (rc == PNA_OK && verify && S.top) rc = pna_cuda_plan_crc_results(S.plan, nullptr, &broken);
        if (trace) {
            float ms[16] = {0};
            const int ns = pna_cuda_plan_stage_ms(S.plan, ms, 16);
            fprintf(stderr, "[pna_host] group %zu stages:", S.g);
            for (int q = 0; q < ns; q++) fprintf(stderr, " %s=%.2f", pna_cuda_stage_name((uint32_t)q), ms[q]);
            fprintf(stderr, "\n");
        }
        pna_cuda_plan_destroy(S.plan);
        S.plan = nullptr;
        if (trace) fprintf(stderr, "[pna_host] group %zu: fetch/wait+D2H %.1f-%.1f ms\n", S.g, ms_since(t0), ms_since(std::chrono::steady_clock::now()));
        ck(L.ctx, rc, "decode plan");
        bool any_entry_broken = false;
        for (uint32_t k = 0; k < S.m; k++) {
            status[G.lo + k] = S.pre[k] != PNA_OK ? S.pre[k] : S.st[k];
            if (S.st[k] == PNA_E_INVALID_DATA) any_entry_broken = true;
            // the decoded length is what the stream says, not what fSIZ said: callers slice / write files()[i].size bytes.
            // PNA_E_NOSPACE: size becomes the required length, so the caller can take the entry again with room for it.
            if (S.pre[k] == PNA_OK && (S.st[k] == PNA_OK || S.st[k] == PNA_E_NOSPACE)) files_[G.lo + k].size = S.bufs[k].len;
        }
        if (broken && !any_entry_broken) throw Error(PNA_E_INVALID_DATA, "broken chunk");   // an archive-level / non-file chunk
    };
    // groups in flight per worker: `depth` contexts in a ring -- issue(g) ... collect(g - depth + 1)  (PNA_PIPE_DEPTH, default 2)
    static const int depth = [] { const char* e = getenv("PNA_PIPE_DEPTH"); const int d = e ? atoi(e) : 2; return d < 2 ? 2 : d > 4 ? 4 : d; }();
    auto work = [&](size_t di) {
        std::vector<Stage> stage((size_t)depth);
        try {
            std::vector<std::unique_ptr<CtxLease>> L;
            for (int k = 0; k < depth; k++) L.emplace_back(new CtxLease(devices[di]));
            size_t issued = 0, collected = 0;
            for (;;) {
                const size_t g = next.fetch_add(1);
                if (g >= groups.size()) break;
                issue(*L[issued % depth], stage[issued % depth], g, h2d_mus[di]);
                issued++;
                if (issued - collected >= (size_t)depth) { collect(*L[collected % depth], stage[collected % depth]); collected++; }
            }
            for (; collected < issued; collected++) collect(*L[collected % depth], stage[collected % depth]);
        } catch (const Error& e) {
            for (auto& S : stage) if (S.plan) { pna_cuda_plan_destroy(S.plan); S.plan = nullptr; }
            std::lock_guard<std::mutex> g(err_mu);
            if (err_msg.empty()) { err_msg = e.what(); err_kind = e.kind; }
        }
    };
    const int nw = std::max(1, std::min<int>(workers * (int)devices.size(), (int)std::max<size_t>(groups.size(), 1)));
    std::vector<std::thread> th;
    for (int w = 1; w < nw; w++) th.emplace_back(work, (size_t)w % devices.size());
    work(0);
    for (auto& t : th) t.join();
    if (!err_msg.empty()) throw Error(err_kind, err_msg);
}

// ---------------------------------------------------------------------------------------------- create
static inline void wr_be32(uint8_t* p, uint32_t x) { p[0] = x >> 24; p[1] = x >> 16; p[2] = x >> 8; p[3] = x; }
static uint64_t entry_frame_bound(const std::string& name, uint64_t stream_bound, const std::string& phsf, bool enc, uint32_t mcs) {
    const uint64_t cap = mcs ? mcs : 0xFFFFFFFFull;
    const uint64_t nbody = stream_bound / cap + 2;
    // prefix chunk: 16 (IV) or 75 (GCM stream header); the GCM tags of callers that sized the stream for CBC/CTR fit in the
    // last term for segments of 64 KiB and more
    return 12 + 6 + name.size() + 12 + 16 + (enc ? 12 + phsf.size() + 12 + 75 : 0) + nbody * 12 + stream_bound + 12 + (enc ? 64 + (stream_bound >> 12) : 0);
}

// upper bound of one entry's data stream under `opt` (GCM: header + one tag per segment)
static uint64_t stream_bound_of(uint64_t plain_len, const WriteOptions& opt) {
    pna_encode_desc d;
    memset(&d, 0, sizeof d);
    uint8_t hdr[75] = {0};
    hdr[39] = (uint8_t)(opt.segment_size >> 24); hdr[40] = (uint8_t)(opt.segment_size >> 16);
    hdr[41] = (uint8_t)(opt.segment_size >> 8); hdr[42] = (uint8_t)opt.segment_size;
    d.plain.len = plain_len; d.compression = opt.compression; d.encryption = opt.encryption; d.cipher_mode = opt.cipher_mode;
    d.stream_header = hdr;
    return pna_cuda_encode_bound(&d);
}

// Archive::write_header + add_entry per file + finalize (archive/write.rs:92,368,545; wire order entry.rs:895-912), written
// straight into `out`: every worker encodes one group of files on its own pna_ctx, learns the produced stream lengths,
// takes its place behind the previous group, and lets the GPU copy the streams to their final position (D2H).
// on_region (optional): called by the worker that completed a group with the byte range of the archive that is final now
// (groups are contiguous and cover everything between the archive header and AEND) -- lets a caller write the file while later
// groups are still being encoded.
// fresh random bytes for IVs / salts (lib/src/random.rs:8 uses the OS CSPRNG)
static void os_random(uint8_t* p, size_t n) {
    size_t got = 0;
    while (got < n) {
        const ssize_t k = getrandom(p + got, n - got, 0);
        if (k <= 0) throw Error(PNA_E_INTERNAL, "getrandom failed");
        got += (size_t)k;
    }
}
static uint64_t create_archive_regions(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size,
                                       const std::vector<int>& devices, int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap,
                                       const std::function<void(uint64_t, uint64_t)>* on_region) {
    if (devices.empty()) throw Error(PNA_E_BAD_ARG, "no device given");
    const int device = devices[0];
    const size_t n = files.size();
    struct Group { size_t lo, hi; };
    std::vector<Group> groups;
    {
        size_t lo = 0;
        uint64_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            if (i > lo && acc >= group_bytes) { groups.push_back({lo, i}); lo = i; acc = 0; }
            acc += files[i].data.len;
        }
        if (lo < n) groups.push_back({lo, n});
    }
    const bool enc = opt.encryption != PNA_ENCRYPTION_NO;
    const bool gcm = enc && opt.cipher_mode == PNA_CIPHER_GCM;
    const uint64_t mcs = max_chunk_size ? max_chunk_size : 0xFFFFFFFFull, iv_len = enc ? (gcm ? 75 : 16) : 0;   // stream prefix = its own chunk
    // CBC / CTR: one fresh IV per entry (entry/write.rs:108-111 draws it inside the writer).  A builder without an IV of its own
    // gets one from the OS here -- never a shared default: under CTR that would reuse the keystream across the archive.
    std::vector<std::array<uint8_t, 16>> ivs(enc && !gcm ? n : 0);
    if (enc && !gcm) {
        std::vector<size_t> need;
        for (size_t i = 0; i < n; i++) { if (files[i].iv_set) memcpy(ivs[i].data(), files[i].iv, 16); else need.push_back(i); }
        if (!need.empty()) {
            std::vector<uint8_t> rnd(16 * need.size());
            os_random(rnd.data(), rnd.size());
            for (size_t q = 0; q < need.size(); q++) memcpy(ivs[need[q]].data(), rnd.data() + 16 * q, 16);
        }
    }
    // GCM (entry/write.rs:75-106): per entry a stream header (salt, nonce prefix, segment size, key confirmation) and a stream key
    // bound to the entry's FHED chunk -- host work, once per entry, through the library's key schedule
    std::vector<std::array<uint8_t, 75>> gcm_hdr(gcm ? n : 0);
    std::vector<std::array<uint8_t, 32>> gcm_key(gcm ? n : 0);
    if (cap < 8 + 20 + 12) throw Error(PNA_E_NOSPACE, "archive buffer too small");
    memcpy(out, SIGNATURE, 8);
    // ---- metadata chunks (type || data, contiguous) of EVERY entry plus AHED / AEND, CRCs in one GPU batch up front: the
    // device is idle now, later these tiny kernels would queue behind the encode kernels of whole groups
    const size_t metas_per_entry = enc ? 5 : 3;
    std::vector<uint8_t> meta;
    struct MetaRef { size_t off; uint32_t len; };
    std::vector<MetaRef> refs;
    refs.reserve(n * metas_per_entry + 2);
    auto add_meta = [&](const char* ty, const uint8_t* data, uint32_t len) {
        refs.push_back({meta.size(), len + 4});
        meta.insert(meta.end(), ty, ty + 4);
        if (len) meta.insert(meta.end(), data, data + len);
    };
    for (size_t i = 0; i < n; i++) {
        const FileEntryBuilder& f = files[i];
        const uint8_t h6[6] = {0, 0, (uint8_t)DataKind::File, opt.compression, opt.encryption, opt.cipher_mode};
        refs.push_back({meta.size(), (uint32_t)(4 + 6 + f.name.size())});
        meta.insert(meta.end(), {'F', 'H', 'E', 'D'});
        meta.insert(meta.end(), h6, h6 + 6);
        meta.insert(meta.end(), f.name.begin(), f.name.end());
        uint8_t sz[8];
        for (int b = 0; b < 8; b++) sz[b] = (uint8_t)(f.data.len >> (8 * (7 - b)));
        int skip = 0;
        while (skip < 8 && sz[skip] == 0) skip++;                       // entry.rs:901-903: minimal big-endian bytes
        add_meta("fSIZ", sz + skip, (uint32_t)(8 - skip));
        if (enc) {
            add_meta("PHSF", (const uint8_t*)opt.phsf.data(), (uint32_t)opt.phsf.size());
            if (gcm) {
                uint8_t salt[32], prefix[7];
                memcpy(salt, f.gcm_salt, 32); memcpy(prefix, f.gcm_nonce_prefix, 7);
                if (!f.gcm_params_set) { os_random(salt, 32); os_random(prefix, 7); }
                int32_t rc = pna_cuda_gcm_stream_header(opt.key, salt, prefix, opt.segment_size, gcm_hdr[i].data());
                if (rc == PNA_OK) {
                    std::vector<uint8_t> hd(h6, h6 + 6);
                    hd.insert(hd.end(), f.name.begin(), f.name.end());
                    rc = pna_cuda_gcm_stream_key(opt.key, gcm_hdr[i].data(), 75, (const uint8_t*)"FHED", hd.data(), hd.size(),
                                                 (const uint8_t*)opt.phsf.data(), opt.phsf.size(), gcm_key[i].data());
                }
                if (rc != PNA_OK) throw Error(rc, f.name + ": GCM stream parameters");
                add_meta("FDAT", gcm_hdr[i].data(), 75);                // the stream header is its own chunk, like the IV
            } else
            add_meta("FDAT", ivs[i].data(), 16);                        // the IV is its own chunk (builder.rs:62-69)
        }
        add_meta("FEND", nullptr, 0);
    }
    const uint8_t ahed8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    add_meta("AHED", ahed8, 8);
    add_meta("AEND", nullptr, 0);
    std::vector<uint32_t> mcrc(refs.size());
    {
        std::vector<pna_span> spans(refs.size());
        for (size_t r = 0; r < refs.size(); r++) spans[r] = {meta.data() + refs[r].off, refs[r].len};
        CtxLease L(device);
        ck(L.ctx, pna_cuda_crc32(L.ctx, spans.data(), (uint32_t)spans.size(), mcrc.data()), "metadata crc");
    }
    // group g starts at base[g]; base[g+1] is published by the worker of group g as soon as it knows its size
    std::vector<uint64_t> base(groups.size() + 1, 0);
    std::vector<uint8_t> ready(groups.size() + 1, 0);
    base[0] = 8 + 20; ready[0] = 1;
    std::mutex mu;
    std::vector<std::unique_ptr<Slots>> h2d_slots;   // uploads take turns per device (each GPU has its own PCIe link)
    for (size_t d = 0; d < devices.size(); d++) h2d_slots.emplace_back(new Slots(2));
    std::condition_variable cv;
    std::atomic<size_t> next{0};
    std::string err_msg;
    int err_kind = 0;
    bool failed = false;
    const bool trace = getenv("PNA_HOST_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    // Every worker keeps TWO contexts and software-pipelines its groups like extract_files does:
    //   issue(g): upload the plaintext (taking turns on PCIe) + launch the encode kernels (asynchronous)
    //   finish(g-1): wait for the previous group's lengths, take its place behind its predecessor, D2H the streams to their
    //                final position, write the chunk frames around them
    // so a group's kernels run while the next group's plaintext is on its way and the previous group's streams leave.
    struct Stage {
        size_t g = 0;
        uint32_t m = 0;
        pna_plan* plan = nullptr;
        uint64_t crc_total = 0;
        std::vector<pna_encode_desc> descs;
        double t_issue0 = 0, t_issue1 = 0;
    };
    auto issue = [&](CtxLease& L, Stage& S, size_t g, Slots& h2d_slot) {
        const Group G = groups[g];
        const uint32_t m = (uint32_t)(G.hi - G.lo);
        S.g = g; S.m = m; S.plan = nullptr; S.crc_total = 0;
        S.descs.assign(m, pna_encode_desc{});
        for (uint32_t k = 0; k < m; k++) {
            const FileEntryBuilder& f = files[G.lo + k];
            pna_encode_desc& d = S.descs[k];
            memset(&d, 0, sizeof d);
            d.plain = f.data;
            d.compression = opt.compression; d.encryption = opt.encryption; d.cipher_mode = opt.cipher_mode; d.level = opt.level;
            memcpy(d.key, opt.key, 32);
            if (!ivs.empty()) memcpy(d.iv, ivs[G.lo + k].data(), 16);
            if (gcm) { memcpy(d.key, gcm_key[G.lo + k].data(), 32); d.stream_header = gcm_hdr[G.lo + k].data(); }
            d.max_chunk_size = max_chunk_size;
            S.crc_total += pna_cuda_encode_crc_count(&d);
        }
        {
            SlotGuard h2d(h2d_slot);   // uploads take turns: PCIe is the shared resource
            S.t_issue0 = ms_now();
            ck(L.ctx, pna_cuda_encode_plan_create(L.ctx, S.descs.data(), m, &S.plan), "encode_plan_create");
        }
        S.t_issue1 = ms_now();
        const int rc = pna_cuda_encode_plan_run(S.plan);
        if (rc != PNA_OK) { pna_cuda_plan_destroy(S.plan); S.plan = nullptr; ck(L.ctx, rc, "encode_plan_run"); }
    };
    auto finish = [&](CtxLease& L, Stage& S) {
        if (!S.plan) return;
        struct PlanGuard { pna_plan*& p; ~PlanGuard() { if (p) { pna_cuda_plan_destroy(p); p = nullptr; } } } guard{S.plan};
        const size_t g = S.g;
        const Group G = groups[g];
        const uint32_t m = S.m;
        const MetaRef* R = refs.data() + G.lo * metas_per_entry;        // this group's metadata chunks
        const uint32_t* RC = mcrc.data() + G.lo * metas_per_entry;
        std::vector<uint64_t> lens(m);
        std::vector<int32_t> st(m);
        const double tr2 = ms_now();
        ck(L.ctx, pna_cuda_encode_plan_lengths(S.plan, lens.data(), st.data()), "encode_plan_lengths");
        const double tr3 = ms_now();
        // layout of this group
        std::vector<uint64_t> entry_pos(m + 1, 0);
        for (uint32_t k = 0; k < m; k++) {
            if (st[k] != PNA_OK) throw Error(st[k], files[G.lo + k].name + ": encode failed");
            uint64_t sz = 0;
            for (size_t q = 0; q < metas_per_entry; q++) sz += 8 + R[k * metas_per_entry + q].len;   // len + (type||data) + crc
            const uint64_t D = lens[k] - iv_len, nb = (D + mcs - 1) / mcs;
            sz += nb * 12 + D;
            entry_pos[k + 1] = entry_pos[k] + sz;
        }
        uint64_t my_base;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready[g] || failed; });
            if (failed) return;
            my_base = base[g];
            base[g + 1] = my_base + entry_pos[m];
            ready[g + 1] = 1;
            cv.notify_all();
        }
        if (my_base + entry_pos[m] + 12 > cap) throw Error(PNA_E_NOSPACE, "archive buffer too small");
        // fetch: single-body streams go straight to their final place (the 16 IV bytes land on the frame bytes in front
        // of the data and are overwritten by them afterwards); multi-body streams pass through a staging buffer
        std::vector<pna_buf> bufs(m);
        std::vector<std::vector<uint8_t>> stage(m);
        std::vector<uint64_t> data_pos(m);
        for (uint32_t k = 0; k < m; k++) {
            uint64_t p = my_base + entry_pos[k];
            for (size_t q = 0; q + 1 < metas_per_entry; q++) p += 8 + R[k * metas_per_entry + q].len;
            data_pos[k] = p;   // first FDAT body frame starts here
            const uint64_t D = lens[k] - iv_len, nb = (D + mcs - 1) / mcs;
            if (nb <= 1) bufs[k] = pna_buf{out + p + 8 - iv_len, lens[k], 0};
            else { stage[k].resize(lens[k]); bufs[k] = pna_buf{stage[k].data(), lens[k], 0}; }
        }
        std::vector<uint32_t> crcs(S.crc_total + 1), ncrc(m, 0);
        const double tr4 = ms_now();
        ck(L.ctx, pna_cuda_encode_plan_fetch_region(S.plan, bufs.data(), out + my_base, entry_pos[m], crcs.data(), ncrc.data(), st.data()), "encode_plan_fetch");
        const double tr5 = ms_now();
        size_t cpos = 0;
        for (uint32_t k = 0; k < m; k++) {
            if (st[k] != PNA_OK) throw Error(st[k], files[G.lo + k].name + ": fetch failed");
            uint8_t* w = out + my_base + entry_pos[k];
            auto put_meta = [&](size_t r) {
                const MetaRef& mr = R[r];
                wr_be32(w, mr.len - 4); memcpy(w + 4, meta.data() + mr.off, mr.len); wr_be32(w + 4 + mr.len, RC[r]);
                w += 8 + mr.len;
            };
            const uint64_t D = lens[k] - iv_len, nb = (D + mcs - 1) / mcs;
            // bodies first where they were staged, then the frames around them (which also repair the IV overlap)
            uint8_t* body_w = out + data_pos[k];
            for (uint64_t b = 0; b < nb; b++) {
                const uint64_t o = b * mcs, l = std::min<uint64_t>(mcs, D - o);
                if (nb > 1) memcpy(body_w + 8, stage[k].data() + iv_len + o, l);
                wr_be32(body_w + 8 + l, crcs[cpos + b]);
                body_w += 12 + l;
            }
            for (size_t q = 0; q + 1 < metas_per_entry; q++) put_meta(k * metas_per_entry + q);   // FHED, fSIZ, [PHSF, FDAT(iv)]
            body_w = out + data_pos[k];
            for (uint64_t b = 0; b < nb; b++) {
                const uint64_t o = b * mcs, l = std::min<uint64_t>(mcs, D - o);
                wr_be32(body_w, (uint32_t)l); memcpy(body_w + 4, "FDAT", 4);
                body_w += 12 + l;
            }
            w = body_w;
            put_meta(k * metas_per_entry + metas_per_entry - 1);   // FEND
            cpos += ncrc[k];
        }
        if (on_region) (*on_region)(my_base, entry_pos[m]);
        if (trace) fprintf(stderr, "[pna_host] create group %zu (%u files): plan_create/H2D %.1f-%.1f, lengths(wait) %.1f-%.1f, fetch/D2H %.1f-%.1f, frames -%.1f ms\n",
                           g, m, S.t_issue0, S.t_issue1, tr2, tr3, tr4, tr5, ms_now());
    };
    auto work = [&](size_t di) {
        Stage stage[2];
        try {
            CtxLease L0(devices[di]), L1(devices[di]);
            CtxLease* L[2] = {&L0, &L1};
            int cur = 0;
            bool have_prev = false;
            for (;;) {
                const size_t g = next.fetch_add(1);
                if (g >= groups.size()) break;
                issue(*L[cur], stage[cur], g, *h2d_slots[di]);
                if (have_prev) finish(*L[cur ^ 1], stage[cur ^ 1]);
                have_prev = true;
                cur ^= 1;
            }
            if (have_prev) finish(*L[cur ^ 1], stage[cur ^ 1]);
        } catch (const Error& e) {
            for (auto& S : stage) if (S.plan) { pna_cuda_plan_destroy(S.plan); S.plan = nullptr; }
            std::lock_guard<std::mutex> lk(mu);
            if (err_msg.empty()) { err_msg = e.what(); err_kind = e.kind; }
            failed = true;
            cv.notify_all();
        }
    };
    const int nw = std::max(1, std::min<int>(workers * (int)devices.size(), (int)std::max<size_t>(groups.size(), 1)));
    std::vector<std::thread> th;
    for (int w = 1; w < nw; w++) th.emplace_back(work, (size_t)w % devices.size());
    work(0);
    for (auto& t : th) t.join();
    if (!err_msg.empty()) throw Error(err_kind, err_msg);
    // archive framing: AHED right behind the signature, AEND behind the last group (archive/write.rs:92,545)
    const uint64_t end = base[groups.size()];
    if (end + 12 > cap) throw Error(PNA_E_NOSPACE, "archive buffer too small");
    {
        const MetaRef& ra = refs[n * metas_per_entry], & re = refs[n * metas_per_entry + 1];
        wr_be32(out + 8, 8); memcpy(out + 12, meta.data() + ra.off, 12); wr_be32(out + 24, mcrc[n * metas_per_entry]);
        wr_be32(out + end, 0); memcpy(out + end + 4, meta.data() + re.off, 4); wr_be32(out + end + 8, mcrc[n * metas_per_entry + 1]);
    }
    return end + 12;
}

uint64_t create_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size, int device,
                             int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap) {
    return create_archive_regions(files, opt, max_chunk_size, std::vector<int>{device}, workers, group_bytes, out, cap, nullptr);
}
uint64_t create_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size,
                             const std::vector<int>& devices, int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap) {
    return create_archive_regions(files, opt, max_chunk_size, devices, workers, group_bytes, out, cap, nullptr);
}

std::vector<uint8_t> create_archive(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size,
                                    int device, int workers, uint64_t group_bytes) {
    uint64_t bound = 8 + 20 + 12;
    for (const auto& f : files) {
        pna_encode_desc d;
        memset(&d, 0, sizeof d);
        d.plain.len = f.data.len; d.compression = opt.compression; d.encryption = opt.encryption;
        bound += entry_frame_bound(f.name, stream_bound_of(f.data.len, opt), opt.phsf, opt.encryption != 0, max_chunk_size);
    }
    std::vector<uint8_t> out(bound);
    out.resize(create_archive_into(files, opt, max_chunk_size, device, workers, group_bytes, out.data(), out.size()));
    return out;
}

// ---------------------------------------------------------------------------------------------- file-system side
struct PinnedBuf {   // RAII lease
    pna_ctx* ctx;
    uint8_t* p = nullptr;
    uint64_t bytes = 0;
    PinnedBuf(pna_ctx* c, uint64_t n) : ctx(c) { p = g_pinned.get(c, n, &bytes); }
    ~PinnedBuf() { g_pinned.put(ctx, p, bytes); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
};

// ---------------------------------------------------------------------------------------------- solid create
// Archive::write_solid_header + add_entry... + finalize (lib/src/archive/write.rs:438-471, lib/src/entry/builder/solid.rs:68-320):
// the inner entries are written STORE, chunk by chunk, into ONE stream that is compressed + encrypted as a whole and cut into
// SDAT bodies (wire order lib/src/entry.rs:471-483).  Inner chunk CRCs, the encode and the outer chunk CRCs are three GPU calls.
// The stream this library writes is one zstd frame per MiB, so the solid entry decodes frame-parallel (DESIGN "Frame-parallel streams").
// max_chunk_size cuts the SDAT bodies; an inner entry's data stays in FDAT chunks as large as the format allows (the inner entries are
// built before they are added, lib/src/entry/builder/solid.rs: their chunking is independent of the solid stream's), which also keeps
// the reader's inner CRC checks and range copies few and large.
static constexpr uint64_t INNER_FDAT = 0x40000000ull;
uint64_t create_solid_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size, int device,
                                   uint8_t* out, uint64_t cap) {
    const uint64_t mcs = max_chunk_size ? max_chunk_size : 0xFFFFFFFFull;
    const bool enc = opt.encryption != PNA_ENCRYPTION_NO, gcm = enc && opt.cipher_mode == PNA_CIPHER_GCM;
    CtxLease L(device);
    // ---- inner archive stream: FHED, fSIZ, FDAT..., FEND per file
    uint64_t inner_len = 0;
    std::vector<uint64_t> at(files.size());
    for (size_t i = 0; i < files.size(); i++) {
        at[i] = inner_len;
        const uint64_t D = files[i].data.len, nb = (D + INNER_FDAT - 1) / INNER_FDAT;
        inner_len += (12 + 6 + files[i].name.size()) + (12 + 8) + nb * 12 + D + 12;
    }
    PinnedBuf inner_buf(L.ctx, inner_len + 64);
    uint8_t* const inner = inner_buf.p;
    if (!inner) throw Error(PNA_E_OOM, "pinned inner stream");
    std::vector<pna_span> spans;
    std::vector<uint64_t> crc_at;
    uint64_t w = 0;
    // bodies of a MiB and more are copied afterwards by a few threads (the copy is the host-side cost of solid mode)
    std::vector<uint64_t> big_dst, big_len;
    std::vector<const uint8_t*> big_src;
    auto put = [&](const char* ty, const uint8_t* data, uint64_t copy_len, uint64_t len) {
        wr_be32(inner + w, (uint32_t)len); memcpy(inner + w + 4, ty, 4);
        if (copy_len) memcpy(inner + w + 8, data, copy_len);
        spans.push_back({inner + w + 4, len + 4});
        crc_at.push_back(w + 8 + len);
        w += 12 + len;
    };
    for (size_t i = 0; i < files.size(); i++) {
        const FileEntryBuilder& f = files[i];
        std::vector<uint8_t> h = {0, 0, (uint8_t)DataKind::File, 0, 0, 0};
        h.insert(h.end(), f.name.begin(), f.name.end());
        put("FHED", h.data(), h.size(), h.size());
        uint8_t sz[8];
        for (int b = 0; b < 8; b++) sz[b] = (uint8_t)(f.data.len >> (8 * (7 - b)));
        int skip = 0;
        while (skip < 8 && sz[skip] == 0) skip++;
        put("fSIZ", sz + skip, (uint64_t)(8 - skip), (uint64_t)(8 - skip));
        for (uint64_t o = 0; o < f.data.len; o += INNER_FDAT) {
            const uint64_t len = std::min<uint64_t>(INNER_FDAT, f.data.len - o);
            if (len >= (1u << 20)) { big_dst.push_back(w + 8); big_src.push_back(f.data.ptr + o); big_len.push_back(len); put("FDAT", nullptr, 0, len); }
            else put("FDAT", f.data.ptr + o, len, len);
        }
        put("FEND", nullptr, 0, 0);
    }
    inner_len = w;
    if (!big_dst.empty()) {
        std::atomic<size_t> next{0};
        auto copier = [&]() { for (size_t k; (k = next.fetch_add(1)) < big_dst.size();) memcpy(inner + big_dst[k], big_src[k], big_len[k]); };
        const int nt = (int)std::min<size_t>(8, big_dst.size());
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(copier);
        copier();
        for (auto& t : th) t.join();
    }
    {
        std::vector<uint32_t> crcs(spans.size());
        for (size_t b0 = 0; b0 < spans.size(); b0 += 1u << 20) {   // batches of a million chunks
            const uint32_t m = (uint32_t)std::min<size_t>(1u << 20, spans.size() - b0);
            ck(L.ctx, pna_cuda_crc32(L.ctx, spans.data() + b0, m, crcs.data() + b0), "inner chunk crc");
        }
        for (size_t k = 0; k < spans.size(); k++) wr_be32(inner + crc_at[k], crcs[k]);
    }
    // ---- encode the inner stream as one entry
    const uint8_t shed[5] = {0, 0, opt.compression, opt.encryption, opt.cipher_mode};
    pna_encode_desc d;
    memset(&d, 0, sizeof d);
    d.plain = pna_span{inner, inner_len};
    d.compression = opt.compression; d.encryption = opt.encryption; d.cipher_mode = opt.cipher_mode; d.level = opt.level;
    memcpy(d.key, opt.key, 32);
    uint8_t gcm_hdr[75], gcm_key[32];
    static thread_local std::random_device rd;
    if (gcm) {
        uint8_t salt[32], prefix[8];
        for (int k = 0; k < 32; k += 4) { const uint32_t r = rd(); memcpy(salt + k, &r, 4); }
        for (int k = 0; k < 8; k += 4) { const uint32_t r = rd(); memcpy(prefix + k, &r, 4); }
        int32_t rc = pna_cuda_gcm_stream_header(opt.key, salt, prefix, opt.segment_size, gcm_hdr);
        if (rc == PNA_OK)
            rc = pna_cuda_gcm_stream_key(opt.key, gcm_hdr, 75, (const uint8_t*)"SHED", shed, 5, (const uint8_t*)opt.phsf.data(), opt.phsf.size(), gcm_key);
        if (rc != PNA_OK) throw Error(rc, "GCM stream parameters");
        memcpy(d.key, gcm_key, 32);
        d.stream_header = gcm_hdr;
    } else if (enc) {
        for (int k = 0; k < 16; k += 4) { const uint32_t r = rd(); memcpy(d.iv + k, &r, 4); }   // entry/write.rs:108-111
    }
    const uint64_t sbound = pna_cuda_encode_bound(&d);
    PinnedBuf stream_buf(L.ctx, sbound + 64);
    if (!stream_buf.p) throw Error(PNA_E_OOM, "pinned stream buffer");
    pna_buf ob{stream_buf.p, sbound, 0};
    int32_t st = 0;
    ck(L.ctx, pna_cuda_encode_batch(L.ctx, &d, 1, &ob, nullptr, nullptr, &st), "solid encode");
    if (st != PNA_OK) throw Error(st, "solid encode failed");
    // ---- frame: signature, AHED, SHED, [PHSF], SDAT(prefix), SDAT..., SEND, AEND
    const uint64_t prefix_len = enc ? (gcm ? 75 : 16) : 0, D = ob.len - prefix_len, nbody = (D + mcs - 1) / mcs;
    const uint64_t need = 8 + 20 + (12 + 5) + (enc ? 12 + opt.phsf.size() + 12 + prefix_len : 0) + nbody * 12 + D + 12 + 12;
    if (need > cap) throw Error(PNA_E_NOSPACE, "archive buffer too small");
    memcpy(out, SIGNATURE, 8);
    uint64_t o = 8;
    std::vector<pna_span> ospans;
    std::vector<uint64_t> ocrc_at;
    auto oput = [&](const char* ty, const uint8_t* data, uint64_t len) {
        wr_be32(out + o, (uint32_t)len); memcpy(out + o + 4, ty, 4);
        if (len) memcpy(out + o + 8, data, len);
        ospans.push_back({out + o + 4, len + 4});
        ocrc_at.push_back(o + 8 + len);
        o += 12 + len;
    };
    const uint8_t ahed8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    oput("AHED", ahed8, 8);
    oput("SHED", shed, 5);
    if (enc) {
        oput("PHSF", (const uint8_t*)opt.phsf.data(), opt.phsf.size());
        oput("SDAT", stream_buf.p, prefix_len);                      // IV / stream header: its own chunk (builder.rs:62-69)
    }
    for (uint64_t b = 0; b < nbody; b++) oput("SDAT", stream_buf.p + prefix_len + b * mcs, std::min<uint64_t>(mcs, D - b * mcs));
    oput("SEND", nullptr, 0);
    oput("AEND", nullptr, 0);
    {
        std::vector<uint32_t> crcs(ospans.size());
        for (size_t b0 = 0; b0 < ospans.size(); b0 += 1u << 20) {
            const uint32_t m = (uint32_t)std::min<size_t>(1u << 20, ospans.size() - b0);
            ck(L.ctx, pna_cuda_crc32(L.ctx, ospans.data() + b0, m, crcs.data() + b0), "outer chunk crc");
        }
        for (size_t k = 0; k < ospans.size(); k++) wr_be32(out + ocrc_at[k], crcs[k]);
    }
    return o;
}
uint64_t create_solid_archive_bound(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size) {
    const uint64_t mcs = max_chunk_size ? max_chunk_size : 0xFFFFFFFFull;
    uint64_t inner = 0;
    for (const auto& f : files) inner += (12 + 6 + f.name.size()) + (12 + 8) + ((f.data.len + INNER_FDAT - 1) / INNER_FDAT) * 12 + f.data.len + 12;
    const uint64_t sb = stream_bound_of(inner, opt);
    return 8 + 20 + 17 + 12 + opt.phsf.size() + 12 + 75 + (sb / mcs + 2) * 12 + sb + 24 + 64;
}
std::string sanitize_entry_name(const std::string& name) {
    std::string out;
    size_t i = 0;
    while (i <= name.size()) {
        size_t j = name.find('/', i);
        if (j == std::string::npos) j = name.size();
        const std::string c = name.substr(i, j - i);
        if (!c.empty() && c != "." && c != "..") { if (!out.empty()) out += '/'; out += c; }
        i = j + 1;
    }
    return out;
}
static void mkdirs(const std::string& path) {   // mkdir -p
    for (size_t i = 1; i <= path.size(); i++)
        if (i == path.size() || path[i] == '/') {
            const std::string d = path.substr(0, i);
            if (mkdir(d.c_str(), 0755) != 0 && errno != EEXIST) throw Error(PNA_E_INTERNAL, "mkdir " + d + ": " + strerror(errno));
        }
}
struct DirCache {   // directories already made (many files share few parents)
    std::mutex mu;
    std::set<std::string> made;
    void ensure_parent(const std::string& file) {
        const size_t k = file.rfind('/');
        if (k == std::string::npos || k == 0) return;
        const std::string d = file.substr(0, k);
        { std::lock_guard<std::mutex> g(mu); if (made.count(d)) return; }
        mkdirs(d);
        std::lock_guard<std::mutex> g(mu);
        made.insert(d);
    }
};
static void write_whole(const std::string& path, const uint8_t* p, uint64_t n) {
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw Error(PNA_E_INTERNAL, "open " + path + ": " + strerror(errno));
    uint64_t done = 0;
    while (done < n) {
        const ssize_t w = write(fd, p + done, (size_t)std::min<uint64_t>(n - done, (uint64_t)1 << 30));
        if (w < 0) { if (errno == EINTR) continue; const std::string m = strerror(errno); close(fd); throw Error(PNA_E_INTERNAL, "write " + path + ": " + m); }
        done += (uint64_t)w;
    }
    close(fd);
}
template <class F>
static void parallel_for(size_t n, int threads, F&& f) {   // f(i), first exception rethrown
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::string msg;
    int kind = 0;
    auto work = [&]() {
        try { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) break; f(i); } }
        catch (const Error& e) { std::lock_guard<std::mutex> g(mu); if (msg.empty()) { msg = e.what(); kind = e.kind ? e.kind : PNA_E_INTERNAL; } next.store(n); }
    };
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(threads, 1), n));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (kind) throw Error(kind, msg);
}
static double ms_between(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
}

IoStats extract_to_dir(Archive& a, const ReadOptions& opt, const std::string& out_dir, int device, int workers, uint64_t group_bytes,
                       uint64_t window_bytes, int io_threads, bool verify, int32_t* status) {
    IoStats st;
    const auto t_begin = std::chrono::steady_clock::now();
    a.prepare(opt, device);
    st.index_ms = ms_between(t_begin, std::chrono::steady_clock::now());   // solid streams are decoded here too
    mkdirs(out_dir);
    DirCache dirs;
    for (const EntryInfo& e : a.entries()) {
        if (e.kind != 0) continue;
        if (e.data_kind == (uint8_t)DataKind::Directory) {
            const std::string rel = sanitize_entry_name(e.name);
            if (!rel.empty()) { mkdirs(out_dir + "/" + rel); st.dirs++; }
        } else if (e.data_kind != (uint8_t)DataKind::File) st.skipped++;   // links: metadata work, not on the data path
    }
    const std::vector<FileOut>& files = a.files();
    const size_t n = files.size();
    std::vector<uint64_t> offsets(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        const uint64_t sz = files[i].status ? 0 : files[i].size;
        if (sz > ((uint64_t)1 << 60) || offsets[i] > ((uint64_t)1 << 60)) throw Error(PNA_E_OOM, "decoded sizes overflow the buffer layout");   // checked sum
        offsets[i + 1] = offsets[i] + (sz + 15) / 16 * 16;
    }
    std::vector<int32_t> stv(n, 0);
    // windows of whole files
    std::vector<std::pair<size_t, size_t>> windows;
    uint64_t biggest = 0;
    for (size_t lo = 0; lo < n;) {
        size_t hi = lo + 1;
        while (hi < n && offsets[hi + 1] - offsets[lo] <= window_bytes) hi++;
        windows.push_back({lo, hi});
        biggest = std::max(biggest, offsets[hi] - offsets[lo]);
        lo = hi;
    }
    CtxLease L(device);
    PinnedBuf win0(L.ctx, biggest + 64), win1(L.ctx, windows.size() > 1 ? biggest + 64 : 64);
    if (!win0.p || !win1.p) throw Error(PNA_E_OOM, "pinned window buffer");
    uint8_t* pinned[2] = {win0.p, win1.p};
    std::mutex io_mu;
    std::future<void> writing[2];
    a.restart_verify();
    for (size_t w = 0; w < windows.size(); w++) {
        const int b = (int)(w & 1);
        if (writing[b].valid()) writing[b].get();   // this buffer's previous window is on disk
        const size_t lo = windows[w].first, hi = windows[w].second;
        const auto t0 = std::chrono::steady_clock::now();
        a.extract_range(opt, pinned[b], offsets.data(), stv.data(), device, workers, group_bytes, verify, lo, hi);
        st.gpu_ms += ms_between(t0, std::chrono::steady_clock::now());
        uint8_t* base = pinned[b];
        writing[b] = std::async(std::launch::async, [&, lo, hi, base]() {
            const auto t1 = std::chrono::steady_clock::now();
            parallel_for(hi - lo, io_threads, [&](size_t k) {
                const size_t i = lo + k;
                if (stv[i] != PNA_OK || files[i].status) return;
                const std::string rel = sanitize_entry_name(files[i].name);
                if (rel.empty()) return;
                const std::string path = out_dir + "/" + rel;
                dirs.ensure_parent(path);
                write_whole(path, base + (offsets[i] - offsets[lo]), files[i].size);
            });
            const double dt = ms_between(t1, std::chrono::steady_clock::now());
            std::lock_guard<std::mutex> g(io_mu);
            st.io_ms += dt;                                           // write time of the windows (overlaps the next window's gpu_ms)
        });
    }
    for (auto& f : writing) if (f.valid()) f.get();
    // files whose stream decodes to MORE than their fSIZ said (the reference extracts them all the same: it never looks at fSIZ):
    // extract_range left the required length in files()[i].size -- take them again, one by one, with room for it
    for (size_t i = 0; i < n; i++) {
        if (stv[i] != PNA_E_NOSPACE || files[i].status) continue;
        std::vector<uint64_t> off2(n + 1, 0);
        off2[i + 1] = (files[i].size + 15) / 16 * 16;
        PinnedBuf one(L.ctx, off2[i + 1] + 64);
        if (!one.p) throw Error(PNA_E_OOM, "pinned retry buffer");
        a.extract_range(opt, one.p, off2.data(), stv.data(), device, 1, group_bytes, false, i, i + 1);
        if (stv[i] != PNA_OK) continue;
        const std::string rel = sanitize_entry_name(files[i].name);
        if (rel.empty()) continue;
        const std::string path = out_dir + "/" + rel;
        dirs.ensure_parent(path);
        write_whole(path, one.p, files[i].size);
    }
    for (size_t i = 0; i < n; i++) {
        const int32_t s = files[i].status ? files[i].status : stv[i];
        if (status) status[i] = s;
        if (s == PNA_OK) { st.files++; st.bytes += files[i].size; } else st.skipped++;
    }
    st.total_ms = ms_between(t_begin, std::chrono::steady_clock::now());
    return st;
}

IoStats create_from_files(const std::vector<std::pair<std::string, std::string>>& name_and_path, const WriteOptions& opt,
                          uint32_t max_chunk_size, const std::string& archive_path, int device, int workers, uint64_t group_bytes,
                          int io_threads) {
    IoStats st;
    const auto t_begin = std::chrono::steady_clock::now();
    const size_t n = name_and_path.size();
    std::vector<uint64_t> sizes(n), offs(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        struct stat sb;
        if (stat(name_and_path[i].second.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) throw Error(PNA_E_INVALID_INPUT, "not a regular file: " + name_and_path[i].second);
        sizes[i] = (uint64_t)sb.st_size;
        offs[i + 1] = offs[i] + (sizes[i] + 15) / 16 * 16;
    }
    CtxLease L(device);
    PinnedBuf plain_buf(L.ctx, offs[n] + 64);
    uint8_t* const plain = plain_buf.p;
    if (!plain) throw Error(PNA_E_OOM, "pinned plaintext buffer");
    const auto t_lease = std::chrono::steady_clock::now();
    // core.rs:889-913: small files are read whole; here every file is, by io_threads readers, into pinned memory
    parallel_for(n, io_threads, [&](size_t i) {
        const int fd = open(name_and_path[i].second.c_str(), O_RDONLY);
        if (fd < 0) throw Error(PNA_E_INTERNAL, "open " + name_and_path[i].second + ": " + strerror(errno));
        uint64_t done = 0;
        while (done < sizes[i]) {
            const ssize_t r = read(fd, plain + offs[i] + done, (size_t)std::min<uint64_t>(sizes[i] - done, (uint64_t)1 << 30));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) { close(fd); throw Error(PNA_E_UNEXPECTED_EOF, "short read: " + name_and_path[i].second); }
            done += (uint64_t)r;
        }
        close(fd);
    });
    const auto t_read = std::chrono::steady_clock::now();
    st.io_ms = ms_between(t_begin, t_read);
    if (getenv("PNA_HOST_TRACE")) fprintf(stderr, "[pna_host] create_from_files: stat+pinned lease %.1f ms, read %.1f ms\n", ms_between(t_begin, t_lease), ms_between(t_lease, t_read));
    std::vector<FileEntryBuilder> files(n);
    std::random_device rd;
    uint64_t bound = 8 + 20 + 12;
    for (size_t i = 0; i < n; i++) {
        files[i].name = name_and_path[i].first;
        files[i].data = pna_span{plain + offs[i], sizes[i]};
        pna_encode_desc d;
        memset(&d, 0, sizeof d);
        d.plain.len = sizes[i]; d.compression = opt.compression; d.encryption = opt.encryption;
        bound += entry_frame_bound(files[i].name, stream_bound_of(sizes[i], opt), opt.phsf, opt.encryption != 0, max_chunk_size);
        st.bytes += sizes[i];
    }
    PinnedBuf arch_buf(L.ctx, bound + 64);
    uint8_t* const arch = arch_buf.p;
    if (!arch) throw Error(PNA_E_OOM, "pinned archive buffer");
    // the archive file is written WHILE later groups are encoded: every worker writes the group it has just completed (its next
    // group's kernels are already running on its other context); the header and AEND follow at the end
    const int fd = open(archive_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw Error(PNA_E_INTERNAL, "open " + archive_path + ": " + strerror(errno));
    auto write_range = [&](uint64_t o, uint64_t len) {
        const uint64_t end = o + len;
        while (o < end) {
            const ssize_t w = pwrite(fd, arch + o, (size_t)(end - o), (off_t)o);
            if (w < 0) { if (errno == EINTR) continue; throw Error(PNA_E_INTERNAL, "write " + archive_path + ": " + strerror(errno)); }
            o += (uint64_t)w;
        }
    };
    const std::function<void(uint64_t, uint64_t)> on_region = write_range;
    uint64_t alen = 0;
    try {
        alen = create_archive_regions(files, opt, max_chunk_size, std::vector<int>{device}, workers, group_bytes, arch, bound, &on_region);
        write_range(0, 8 + 20);
        write_range(alen - 12, 12);
    } catch (...) { close(fd); throw; }
    close(fd);
    (void)io_threads;
    const auto t_gpu = std::chrono::steady_clock::now();
    st.gpu_ms = ms_between(t_read, t_gpu);   // encode with the file writes riding on it
    st.files = n;
    st.total_ms = ms_between(t_begin, std::chrono::steady_clock::now());
    return st;
}

// ---- split writer (archive/split_parts.rs)
namespace {
struct CopyTask { uint8_t* dst; const uint8_t* src; uint64_t n; };
// large copies into fresh (untouched) memory are page-fault bound on one thread: pieces of 4 MiB over up to 8 threads
void parallel_copy(const std::vector<CopyTask>& tasks) {
    std::vector<CopyTask> pieces;
    uint64_t total = 0;
    for (const CopyTask& t : tasks)
        for (uint64_t o = 0; o < t.n; o += (uint64_t)4 << 20) { pieces.push_back({t.dst + o, t.src + o, std::min<uint64_t>((uint64_t)4 << 20, t.n - o)}); total += pieces.back().n; }
    const unsigned nt = total < ((uint64_t)16 << 20) ? 1u : std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<size_t> next{0};
    auto work = [&]() { for (size_t i; (i = next.fetch_add(1)) < pieces.size();) memcpy(pieces[i].dst, pieces[i].src, (size_t)pieces[i].n); };
    std::vector<std::thread> th;
    for (unsigned k = 1; k < nt; k++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}
// one step of the layout: a run of source frames copied verbatim, or one freshly framed chunk (length, type, data, CRC to come)
struct SplitOp { uint32_t part; bool fresh; char ty[4]; uint64_t src, n, dst; };   // src: offset in the archive (UINT64_MAX: data made here)
struct SplitPlan { std::vector<SplitOp> ops; std::vector<uint64_t> part_len; };

SplitPlan split_layout(const uint8_t* archive, size_t len, uint64_t max_part_bytes) {
    if (max_part_bytes < MIN_SPLIT_PART_BYTES)       // split_parts.rs:91-96
        throw Error(PNA_E_INVALID_INPUT, "max_part_bytes must be at least " + std::to_string(MIN_SPLIT_PART_BYTES) + " bytes");
    if (len < 8 || memcmp(archive, SIGNATURE, 8) != 0) throw Error(PNA_E_INVALID_DATA, "it is not PNA");
    std::vector<RawChunk> ch;
    index_chunks(archive, len, 8, ch);
    if (ch.empty() || !ty_is(ch[0], "AHED")) throw Error(PNA_E_INVALID_DATA, "expected `AHED` chunk");
    size_t end = ch.size();
    for (size_t i = 0; i < ch.size(); i++) if (ty32(ch[i]) == T_AEND || ty32(ch[i]) == T_ANXT) { end = i; break; }
    if (end == ch.size()) throw Error(PNA_E_UNEXPECTED_EOF, "archive without `AEND`");
    const uint64_t MINC = 12, budget = max_part_bytes - 52;   // signature 8 + AHED 20 + ANXT 12 + AEND 12
    SplitPlan P;
    uint64_t remaining = 0;
    auto fresh = [&](const char* ty, uint64_t src, uint64_t n) {
        SplitOp o{(uint32_t)P.part_len.size() - 1, true, {ty[0], ty[1], ty[2], ty[3]}, src, n, P.part_len.back()};
        P.ops.push_back(o);
        P.part_len.back() += MINC + n;
    };
    auto open_part = [&]() {
        if (P.part_len.size() > 0xFFFFFFFFull) throw Error(PNA_E_INVALID_INPUT, "too many archive parts");
        P.part_len.push_back(8);
        fresh("AHED", UINT64_MAX, 8);
        remaining = budget;
    };
    auto roll_over = [&]() { fresh("ANXT", UINT64_MAX, 0); fresh("AEND", UINT64_MAX, 0); open_part(); };
    auto verbatim = [&](const RawChunk& c) {
        const uint32_t part = (uint32_t)P.part_len.size() - 1;
        const uint64_t n = MINC + c.len;
        SplitOp* last = P.ops.empty() ? nullptr : &P.ops.back();
        if (last && !last->fresh && last->part == part && last->src + last->n == c.off - 8) last->n += n;
        else P.ops.push_back(SplitOp{part, false, {0, 0, 0, 0}, c.off - 8, n, P.part_len.back()});
        P.part_len.back() += n;
        remaining -= n;
    };
    auto does_not_fit = [&](uint64_t n) {
        return Error(PNA_E_INVALID_INPUT, "a " + std::to_string(n) + " byte chunk does not fit within the maximum part size of " + std::to_string(max_part_bytes) + " bytes");
    };
    open_part();
    for (size_t i = 1; i < end; i++) {
        const RawChunk& c = ch[i];
        const uint64_t clen = MINC + c.len;
        const bool stream = ty32(c) == T_FDAT || ty32(c) == T_SDAT;
        if (clen <= remaining) { verbatim(c); continue; }                       // put_chunk, split_parts.rs:140-163
        if (!stream) {
            if (clen > budget) throw does_not_fit(clen);
            roll_over(); verbatim(c); continue;
        }
        if (clen <= budget && remaining <= MINC) { roll_over(); verbatim(c); continue; }
        uint64_t at = c.off, left = c.len;
        for (;;) {                                                              // put_stream, split_parts.rs:165-188
            if (MINC + left <= remaining) { fresh(c.ty, at, left); remaining -= MINC + left; break; }
            if (remaining > MINC) {
                const uint64_t take = remaining - MINC;
                fresh(c.ty, at, take);
                remaining -= MINC + take; at += take; left -= take;
            } else if (budget <= MINC) throw does_not_fit(MINC + left);
            roll_over();
        }
    }
    fresh("AEND", UINT64_MAX, 0);
    return P;
}

// lays the parts out at out[k] (part_len[k] bytes each): copies by several threads, fresh CRCs from one pna_cuda_crc32 batch
void split_write(const uint8_t* archive, const SplitPlan& P, uint8_t* const* out, int device) {
    std::vector<CopyTask> copies;
    std::vector<pna_span> spans;
    for (size_t k = 0; k < P.part_len.size(); k++) memcpy(out[k], SIGNATURE, 8);
    for (const SplitOp& o : P.ops) {
        uint8_t* d = out[o.part] + o.dst;
        if (!o.fresh) { copies.push_back({d, archive + o.src, o.n}); continue; }
        wr_be32(d, (uint32_t)o.n);
        memcpy(d + 4, o.ty, 4);
        if (o.src != UINT64_MAX) copies.push_back({d + 8, archive + o.src, o.n});
        else if (o.n == 8) { memset(d + 8, 0, 4); wr_be32(d + 12, o.part); }     // ArchiveHeader::new(0, 0, archive_number)
        spans.push_back(pna_span{d + 4, o.n + 4});
    }
    parallel_copy(copies);
    std::vector<uint32_t> crc(spans.size());
    {
        CtxLease L(device);
        ck(L.ctx, pna_cuda_crc32(L.ctx, spans.data(), (uint32_t)spans.size(), crc.data()), "split writer: chunk CRCs");
    }
    for (size_t f = 0; f < spans.size(); f++) wr_be32(const_cast<uint8_t*>(spans[f].ptr) + spans[f].len, crc[f]);
}
}  // namespace
static void join_copy(const std::vector<JoinCopy>& copies) {
    std::vector<CopyTask> t;
    for (const JoinCopy& c : copies) t.push_back({c.dst, c.src, c.n});
    parallel_copy(t);
}

std::vector<std::vector<uint8_t>> split_archive(const uint8_t* archive, size_t len, uint64_t max_part_bytes, int device) {
    const SplitPlan P = split_layout(archive, len, max_part_bytes);
    std::vector<std::vector<uint8_t>> parts(P.part_len.size());
    std::vector<uint8_t*> ptr(parts.size());
    for (size_t k = 0; k < parts.size(); k++) { parts[k].resize((size_t)P.part_len[k]); ptr[k] = parts[k].data(); }
    split_write(archive, P, ptr.data(), device);
    return parts;
}
uint64_t split_archive_into(const uint8_t* archive, size_t len, uint64_t max_part_bytes, int device, uint8_t* out, uint64_t cap,
                            std::vector<uint64_t>& part_lens, uint64_t max_parts) {
    const SplitPlan P = split_layout(archive, len, max_part_bytes);
    part_lens = P.part_len;
    uint64_t total = 0;
    for (uint64_t n : P.part_len) total += n;
    if (total > cap || !out || P.part_len.size() > max_parts) return total;    // sizing: nothing copied, no GPU work
    std::vector<uint8_t*> ptr(P.part_len.size());
    uint64_t at = 0;
    for (size_t k = 0; k < ptr.size(); k++) { ptr[k] = out + at; at += P.part_len[k]; }
    split_write(archive, P, ptr.data(), device);
    return total;
}

}  // namespace pna

// ---------------------------------------------------------------------------------------------- flat C view
struct pnah_archive {
    pna::Archive a;
    pna::ReadOptions opt;
    void* map = nullptr;     // mmap of the archive file when opened by path
    size_t map_len = 0;
    ~pnah_archive() { if (map) munmap(map, map_len); }
};
static int fail(const pna::Error& e, char* err, uint64_t cap) {
    if (err && cap) { strncpy(err, e.what(), cap - 1); err[cap - 1] = 0; }
    return e.kind ? e.kind : PNA_E_INTERNAL;
}
extern "C" {
int pnah_open(const uint8_t* buf, uint64_t len, pnah_archive** out, char* err, uint64_t errcap) {
    try {
        pnah_archive* h = new pnah_archive{pna::Archive::read_header_from_slice(buf, len), {}};
        *out = h;
        return PNA_OK;
    } catch (const pna::Error& e) { *out = nullptr; return fail(e, err, errcap); }
}
int pnah_open_multipart(const uint8_t* const* parts, const uint64_t* lens, uint32_t n_parts, int pinned_device, pnah_archive** out, char* err, uint64_t errcap) {
    *out = nullptr;
    try {
        std::vector<pna_span> sp(n_parts);
        for (uint32_t k = 0; k < n_parts; k++) sp[k] = pna_span{parts[k], lens[k]};
        *out = new pnah_archive{pna::Archive::read_multipart(sp.data(), sp.size(), pinned_device), {}};
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
void pnah_close(pnah_archive* a) { delete a; }
int pnah_split(const uint8_t* archive, uint64_t len, uint64_t max_part_bytes, int device, uint8_t* out, uint64_t cap, uint64_t* total,
               uint64_t* part_lens, uint32_t max_parts, uint32_t* n_parts, char* err, uint64_t errcap) {
    try {
        std::vector<uint64_t> lens;
        *total = pna::split_archive_into(archive, (size_t)len, max_part_bytes, device, out, cap, lens, max_parts);
        *n_parts = (uint32_t)lens.size();
        for (size_t k = 0; k < lens.size() && k < max_parts; k++) part_lens[k] = lens[k];   // the layout, also on a sizing call
        if (*total > cap || !out || lens.size() > max_parts) return PNA_E_NOSPACE;
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_open_file(const char* path, pnah_archive** out, char* err, uint64_t errcap) {
    *out = nullptr;
    try {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) throw pna::Error(PNA_E_INTERNAL, std::string("open ") + path + ": " + strerror(errno));
        struct stat sb;
        if (fstat(fd, &sb) != 0) { close(fd); throw pna::Error(PNA_E_INTERNAL, "fstat failed"); }
        void* m = sb.st_size ? mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;   // utils/mmap.rs:36-45
        close(fd);
        if (sb.st_size && m == MAP_FAILED) throw pna::Error(PNA_E_OOM, "mmap failed");
        pnah_archive* h = nullptr;
        try { h = new pnah_archive{pna::Archive::read_header_from_slice((const uint8_t*)m, (size_t)sb.st_size), {}}; }
        catch (...) { if (m) munmap(m, (size_t)sb.st_size); throw; }
        h->map = m; h->map_len = (size_t)sb.st_size;
        *out = h;
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
static void put_stats(const pna::IoStats& s, pnah_io_stats* o) {
    if (!o) return;
    o->files = s.files; o->dirs = s.dirs; o->skipped = s.skipped; o->bytes = s.bytes;
    o->index_ms = s.index_ms; o->gpu_ms = s.gpu_ms; o->io_ms = s.io_ms; o->total_ms = s.total_ms;
}
int pnah_extract_to_dir(pnah_archive* a, const char* out_dir, int device, int workers, uint64_t group_bytes, uint64_t window_bytes,
                        int io_threads, int verify, pnah_io_stats* stats, int32_t* status, char* err, uint64_t errcap) {
    try {
        put_stats(pna::extract_to_dir(a->a, a->opt, out_dir, device, workers, group_bytes, window_bytes, io_threads, verify != 0, status), stats);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_create_from_files(uint32_t n, const char* const* names, const char* const* paths, uint8_t compression, int32_t level,
                           uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf, uint32_t max_chunk_size,
                           const char* archive_path, int device, int workers, uint64_t group_bytes, int io_threads,
                           pnah_io_stats* stats, char* err, uint64_t errcap) {
    try {
        std::vector<std::pair<std::string, std::string>> np(n);
        for (uint32_t i = 0; i < n; i++) np[i] = {names[i], paths[i]};
        pna::WriteOptions opt;
        opt.compression = compression; opt.level = level; opt.encryption = encryption; opt.cipher_mode = cipher_mode;
        if (key) memcpy(opt.key, key, 32);
        if (phsf) opt.phsf = phsf;
        put_stats(pna::create_from_files(np, opt, max_chunk_size, archive_path, device, workers, group_bytes, io_threads), stats);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
uint32_t pnah_entry_count(pnah_archive* a) { return (uint32_t)a->a.entries().size(); }
uint32_t pnah_chunk_count(pnah_archive* a) { return (uint32_t)a->a.chunks().size(); }
int pnah_entry_get(pnah_archive* a, uint32_t i, pnah_entry_info* info) {
    if (i >= a->a.entries().size()) return PNA_E_BAD_ARG;
    const pna::EntryInfo& e = a->a.entries()[i];
    info->kind = e.kind; info->data_kind = e.data_kind; info->compression = e.compression; info->encryption = e.encryption;
    info->cipher_mode = e.cipher_mode; info->n_bodies = (uint32_t)e.bodies.size(); info->compressed_size = e.compressed_size;
    info->raw_file_size = e.raw_file_size; info->name = e.name.c_str(); info->phsf = e.has_phsf ? e.phsf().c_str() : nullptr;
    return PNA_OK;
}
int pnah_set_key(pnah_archive* a, const char* phsf, const uint8_t key[32]) { a->opt.set_key(phsf, key); return PNA_OK; }
int pnah_prepare(pnah_archive* a, int device, char* err, uint64_t errcap) {
    try { a->a.prepare(a->opt, device); return PNA_OK; } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
uint32_t pnah_file_count(pnah_archive* a) { return (uint32_t)a->a.files().size(); }
int pnah_file_get(pnah_archive* a, uint32_t i, const char** name, uint64_t* size) {
    if (i >= a->a.files().size()) return PNA_E_BAD_ARG;
    *name = a->a.files()[i].name.c_str(); *size = a->a.files()[i].size;
    return a->a.files()[i].status;
}
int pnah_file_sizes(pnah_archive* a, uint64_t* sizes, int32_t* status) {
    const auto& f = a->a.files();
    for (size_t i = 0; i < f.size(); i++) { sizes[i] = f[i].size; if (status) status[i] = f[i].status; }
    return PNA_OK;
}
int pnah_extract_files(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                       uint64_t group_bytes, int verify, char* err, uint64_t errcap) {
    try { a->a.extract_files(a->opt, out, offsets, status, device, workers, group_bytes, verify != 0); return PNA_OK; }
    catch (const pna::Error& e) { return fail(e, err, errcap); }
}
uint64_t pnah_create_bound(uint32_t n, const char* const* names, const uint64_t* lens, uint8_t compression, uint8_t encryption,
                           const char* phsf, uint32_t max_chunk_size) {
    uint64_t t = 8 + 20 + 12;
    for (uint32_t i = 0; i < n; i++) {
        pna_encode_desc d;
        memset(&d, 0, sizeof d);
        d.plain.len = lens[i]; d.compression = compression; d.encryption = encryption;
        t += pna::entry_frame_bound(names[i], pna_cuda_encode_bound(&d), phsf ? phsf : "", encryption != 0, max_chunk_size);
    }
    return t;
}
uint64_t pnah_create_solid_bound(uint32_t n, const char* const* names, const uint64_t* lens, uint8_t compression, uint8_t encryption,
                                 uint8_t cipher_mode, const char* phsf, uint32_t max_chunk_size) {
    std::vector<pna::FileEntryBuilder> files(n);
    for (uint32_t i = 0; i < n; i++) { files[i].name = names[i]; files[i].data = pna_span{nullptr, lens[i]}; }
    pna::WriteOptions opt;
    opt.compression = compression; opt.encryption = encryption; opt.cipher_mode = cipher_mode;
    if (phsf) opt.phsf = phsf;
    return pna::create_solid_archive_bound(files, opt, max_chunk_size);
}
int pnah_create_solid(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, uint8_t compression, int32_t level,
                      uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf, uint32_t max_chunk_size, int device,
                      uint8_t* out, uint64_t cap, uint64_t* out_len, char* err, uint64_t errcap) {
    try {
        std::vector<pna::FileEntryBuilder> files(n);
        for (uint32_t i = 0; i < n; i++) { files[i].name = names[i]; files[i].data = pna_span{data[i], lens[i]}; }
        pna::WriteOptions opt;
        opt.compression = compression; opt.level = level; opt.encryption = encryption; opt.cipher_mode = cipher_mode;
        if (key) memcpy(opt.key, key, 32);
        if (phsf) opt.phsf = phsf;
        *out_len = pna::create_solid_archive_into(files, opt, max_chunk_size, device, out, cap);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_extract_range(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers, uint64_t group_bytes,
                       int verify, uint64_t first, uint64_t last, char* err, uint64_t errcap) {
    try {
        if (first > last || last > a->a.files().size()) throw pna::Error(PNA_E_BAD_ARG, "file range out of bounds");
        a->a.extract_range(a->opt, out, offsets, status, device, workers, group_bytes, verify != 0, (size_t)first, (size_t)last);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_extract_files_on(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, const int* devices, uint32_t n_devices,
                          int workers_per_device, uint64_t group_bytes, int verify, char* err, uint64_t errcap) {
    try {
        if (!devices || !n_devices) throw pna::Error(PNA_E_BAD_ARG, "no device given");
        a->a.extract_files(a->opt, out, offsets, status, std::vector<int>(devices, devices + n_devices), workers_per_device, group_bytes, verify != 0);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_create_on(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, const uint8_t* ivs,
                   uint8_t compression, int32_t level, uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf,
                   uint32_t max_chunk_size, const int* devices, uint32_t n_devices, int workers_per_device, uint64_t group_bytes, uint8_t* out,
                   uint64_t cap, uint64_t* out_len, char* err, uint64_t errcap) {
    try {
        if (!devices || !n_devices) throw pna::Error(PNA_E_BAD_ARG, "no device given");
        std::vector<pna::FileEntryBuilder> files(n);
        for (uint32_t i = 0; i < n; i++) {
            files[i].name = names[i];
            files[i].data = pna_span{data[i], lens[i]};
            if (ivs) files[i].set_iv(ivs + 16 * (size_t)i);
        }
        pna::WriteOptions opt;
        opt.compression = compression; opt.level = level; opt.encryption = encryption; opt.cipher_mode = cipher_mode;
        if (key) memcpy(opt.key, key, 32);
        if (phsf) opt.phsf = phsf;
        *out_len = pna::create_archive_into(files, opt, max_chunk_size, std::vector<int>(devices, devices + n_devices), workers_per_device, group_bytes, out, cap);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
int pnah_create(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, const uint8_t* ivs,
                uint8_t compression, int32_t level, uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf,
                uint32_t max_chunk_size, int device, int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap, uint64_t* out_len,
                char* err, uint64_t errcap) {
    try {
        std::vector<pna::FileEntryBuilder> files(n);
        for (uint32_t i = 0; i < n; i++) {
            files[i].name = names[i];
            files[i].data = pna_span{data[i], lens[i]};
            if (ivs) files[i].set_iv(ivs + 16 * (size_t)i);   // NULL: the writer draws a fresh IV per entry
        }
        pna::WriteOptions opt;
        opt.compression = compression; opt.level = level; opt.encryption = encryption; opt.cipher_mode = cipher_mode;
        if (key) memcpy(opt.key, key, 32);
        if (phsf) opt.phsf = phsf;
        *out_len = pna::create_archive_into(files, opt, max_chunk_size, device, workers, group_bytes, out, cap);
        return PNA_OK;
    } catch (const pna::Error& e) { return fail(e, err, errcap); }
}
}
