// pna_host.hpp -- C++ host layer above the C ABI (include/pna_cuda.h): the reference's container logic for the
// data-chunk path, mirroring `libpna` names and behaviour (reference = ChanTsune/Portable-Network-Archive v0.37.0,
// Rust; this image has no Rust toolchain, so the host side is C++ as the reference is compiled code).
//
//   pna::Archive::read_header_from_slice   lib/src/archive/read/slice.rs:17  (signature lib/src/format/signature.rs:6,
//                                          AHED lib/src/archive/header.rs:27-56, chunk walk lib/src/bytes.rs:39-111)
//   pna::NormalEntry / pna::SolidEntry     lib/src/entry.rs:457-483, 665-737, 757-886  (FHED/SHED bytes
//                                          lib/src/entry/header.rs:123-162,274-296)
//   pna::ReadOptions (key cache)           lib/src/entry/options.rs:79-112,1344-1390   (KDF itself is host work supplied
//                                          by the caller, lib/src/hash.rs:45-85)
//   pna::Archive::extract_files            the CLI extract loop cli/src/command/extract.rs:868-1019 folded into GPU
//                                          batches: chunk CRC check + decrypt + decompress via pna_cuda_decode_plan_*,
//                                          several pna_ctx (one per worker thread) so that H2D, kernels and D2H of
//                                          consecutive entry groups overlap
//   pna::FileEntryBuilder / Archive::create   lib/src/entry/builder/file.rs:41-140, lib/src/entry.rs:895-912,
//                                          lib/src/archive/write.rs:92,368,545, lib/src/io.rs:183-197
//
// No entry byte is processed on the CPU here: the host only walks 12-byte chunk frames, groups them and moves bytes.
#ifndef PNA_HOST_HPP
#define PNA_HOST_HPP
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "pna_cuda.h"

namespace pna {

struct Error : std::runtime_error {   // io::Error mirror: kind = PNA_E_* (== io::ErrorKind class)
    int kind;
    Error(int k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

struct RawChunk {        // lib/src/chunk.rs RawChunk, data borrowed from the archive slice
    char ty[4];
    uint64_t off;        // offset of the data field
    uint32_t len;
    uint32_t crc;        // stored CRC (big endian on the wire)
};

enum class DataKind : uint8_t { File = 0, Directory = 1, SymbolicLink = 2, HardLink = 3 };

struct EntryInfo {       // NormalEntry (kind 0) or SolidEntry (kind 1) as the index pass sees it
    uint8_t kind = 0, data_kind = 0, compression = 0, encryption = 0, cipher_mode = 0;
    std::string name;
    std::shared_ptr<const std::string> phsf_;   // PHSF string, shared by all entries that carry the same one
    bool has_phsf = false;
    const std::string& phsf() const { static const std::string none; return phsf_ ? *phsf_ : none; }
    uint64_t raw_file_size = UINT64_MAX;     // fSIZ
    uint64_t compressed_size = 0;
    struct Bodies {                          // FDAT / SDAT bodies, in order: a view into the owning archive's span pool
        const pna_span* p = nullptr;
        uint32_t n = 0;
        const pna_span* data() const { return p; }
        size_t size() const { return n; }
        const pna_span* begin() const { return p; }
        const pna_span* end() const { return p + n; }
    } bodies;
    uint32_t chunk_begin = 0, chunk_end = 0; // chunk index range [begin, end) incl. FHED..FEND
    const uint8_t* header_data = nullptr;    // Data field of the FHED / SHED chunk (GCM binds it into the stream key, aead.rs:166)
    uint32_t header_len = 0;
};

class ReadOptions {      // options.rs:1344; keys: PHSF string -> derived 32-byte key (KeyCache, options.rs:79)
public:
    bool has_password = false;
    std::map<std::string, std::array<uint8_t, 32>> keys;
    void set_key(const std::string& phsf, const uint8_t key[32]) { has_password = true; std::array<uint8_t, 32> k; for (int i = 0; i < 32; i++) k[i] = key[i]; keys[phsf] = k; }
};

struct FileOut {         // one FILE entry to extract: where it lives and where its bytes go
    std::string name;
    uint64_t size = UINT64_MAX;   // decoded size when known (fSIZ or sizing pass)
    int32_t status = 0;
};

class Archive {
public:
    Archive() = default;
    Archive(Archive&&) = default;               // entries hold views into body_pool_: movable, not copyable
    Archive& operator=(Archive&&) = default;
    Archive(const Archive&) = delete;
    Archive& operator=(const Archive&) = delete;
    static Archive read_header_from_slice(const uint8_t* buf, size_t len);
    // Split archive (archive/read.rs:105-165 read_next_archive; writer archive/split_parts.rs): part k is a complete chunk
    // stream with archive_number k that ends ... [ANXT] AEND, and an entry's chunks simply continue in the next part.  The
    // parts are joined into one chunk stream the Archive owns (every chunk frame of every part verbatim, stored CRCs
    // included: the entry chunks in order, then AEND, the other parts' archive-level chunks behind it), so the CRC check
    // and the entry groups run as for a single archive.
    // pinned_device >= 0: the joined stream is leased from the pinned pool of that device's contexts (DMA uploads); -1: pageable.
    static Archive read_multipart(const pna_span* parts, size_t n_parts, int pinned_device = -1);
    const std::vector<RawChunk>& chunks() const { return chunks_; }
    const std::vector<EntryInfo>& entries() const { return entries_; }
    uint32_t archive_number() const { return archive_number_; }

    // Decodes solid entries (their inner archives stay resident in host memory), lists every FILE entry in archive order
    // and learns the decoded size of the ones that carry no fSIZ (sizing pass on the GPU).
    void prepare(const ReadOptions& opt, int device);
    const std::vector<FileOut>& files() const { return files_; }
    // Extract all FILE entries into out[offsets[i] .. offsets[i+1]) (offsets from files()[i].size, caller-computed).
    // verify: chunk CRCs of the whole archive are checked on the GPU, fused with the decode batches.
    // After the call files()[i].size is the DECODED length of every file that came out (fSIZ is only a hint: a smaller real
    // length is reported here; a larger one fails that file with PNA_E_NOSPACE and leaves the required length in size).
    void extract_files(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                       uint64_t group_bytes, bool verify);
    // Same over several GPUs of the box (the partition by entry of the reference's CLI, cli/src/command/extract.rs:868-1019):
    // entry groups go to `workers` threads per device from one shared queue; no collective, no peer traffic.
    void extract_files(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, const std::vector<int>& devices,
                       int workers_per_device, uint64_t group_bytes, bool verify);
    // Same for the files [first, last) only (file i lands at out + offsets[i] - offsets[first]): windows of a large archive.
    // With verify, windows must be taken in archive order (each call checks the chunks up to its last entry).
    void extract_range(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                       uint64_t group_bytes, bool verify, size_t first, size_t last);
    void extract_range(const ReadOptions& opt, uint8_t* out, const uint64_t* offsets, int32_t* status, const std::vector<int>& devices,
                       int workers_per_device, uint64_t group_bytes, bool verify, size_t first, size_t last);
    void restart_verify() { crc_next_chunk_ = 0; }

private:
    struct FileRef { uint32_t owner; uint32_t entry; };   // owner: 0 = top-level archive, k+1 = inner archive of solid k
    // inner archive of a solid entry: host copy for the index pass, and the decode plan whose output (the same bytes)
    // stays resident in HBM for the inner chunk CRC check and for range copies of STORE entries
    struct Inner {
        std::shared_ptr<uint8_t> mem;   // the decoded stream on the host (pinned, from the process-wide pool)
        uint64_t len = 0;
        const uint8_t* data() const { return mem.get(); }
        std::vector<RawChunk> chunks; std::vector<EntryInfo> entries; std::vector<pna_span> body_pool; std::shared_ptr<pna_plan> plan;
    };
    const uint8_t* buf_ = nullptr;
    size_t len_ = 0;
    std::shared_ptr<uint8_t> joined_;          // read_multipart: the joined chunk stream buf_ points into
    uint32_t archive_number_ = 0;
    std::vector<RawChunk> chunks_;
    std::vector<EntryInfo> entries_;
    std::vector<pna_span> body_pool_;          // every FDAT / SDAT body of the top-level archive (EntryInfo::bodies point here)
    std::vector<Inner> inner_;
    std::vector<FileRef> refs_;
    std::vector<FileOut> files_;
    bool prepared_ = false;
    uint32_t crc_next_chunk_ = 0;              // top-level chunks already handed to a CRC check by extract_range
};

struct WriteOptions {    // options.rs:1035
    uint8_t compression = 0, encryption = 0, cipher_mode = 1;
    int32_t level = -1;
    uint8_t key[32] = {0};
    std::string phsf;    // recorded in the PHSF chunk when encryption != 0
    uint32_t segment_size = 1u << 20;   // GCM datastream segment size (options.rs:1200); unused by CBC/CTR
};
struct FileEntryBuilder {   // builder/file.rs:41: plaintext is borrowed until Archive::create returns
    std::string name;
    pna_span data{nullptr, 0};
    // CBC / CTR IV.  The reference's writer draws a fresh one per entry (entry/write.rs:108-111, random.rs:8) and so does
    // this one, from the OS CSPRNG, unless the caller supplies its own with set_iv (tests that pin ciphertext bytes).
    uint8_t iv[16] = {0};
    bool iv_set = false;
    void set_iv(const uint8_t v[16]) { for (int i = 0; i < 16; i++) iv[i] = v[i]; iv_set = true; }
    // GCM: salt and nonce prefix of the stream header (entry/write.rs:81-85); drawn by the writer unless set_gcm_params was called
    uint8_t gcm_salt[32] = {0};
    uint8_t gcm_nonce_prefix[7] = {0};
    bool gcm_params_set = false;
    void set_gcm_params(const uint8_t salt[32], const uint8_t prefix[7]) {
        for (int i = 0; i < 32; i++) gcm_salt[i] = salt[i];
        for (int i = 0; i < 7; i++) gcm_nonce_prefix[i] = prefix[i];
        gcm_params_set = true;
    }
};
// Archive::write_header + add_entry per file + finalize, with one GPU encode batch per worker group.
std::vector<uint8_t> create_archive(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size,
                                    int device, int workers, uint64_t group_bytes);
// Solid mode (lib/src/archive/write.rs:438-471): all files as STORE entries inside ONE compressed (+ encrypted) stream, SDAT bodies of
// max_chunk_size.  Returns the archive length; create_solid_archive_bound sizes `out`.
uint64_t create_solid_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size, int device,
                                   uint8_t* out, uint64_t cap);
uint64_t create_solid_archive_bound(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size);
// Same, written straight into a caller buffer (pinned memory makes the stream copies true DMA); returns the archive length.
uint64_t create_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size, int device,
                             int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap);
// Same over several GPUs: file groups go to `workers` threads per device from one shared queue (entries are independent).
uint64_t create_archive_into(const std::vector<FileEntryBuilder>& files, const WriteOptions& opt, uint32_t max_chunk_size,
                             const std::vector<int>& devices, int workers_per_device, uint64_t group_bytes, uint8_t* out, uint64_t cap);

// Split writer (lib/src/archive/split_parts.rs:90-188 SplitParts::{new, put_chunk, put_stream, roll_over}): a finished archive is cut
// into parts of at most max_part_bytes.  Part k = signature, AHED(0, 0, k), chunks, [ANXT], AEND; a chunk that fits the part's
// remaining budget is copied verbatim, FDAT / SDAT chunks that do not are re-cut at the budget boundary (their new CRCs, like those
// of the new AHED / ANXT / AEND chunks, come from ONE pna_cuda_crc32 batch over the laid-out parts), other chunks move to the next
// part.  Errors as the reference: InvalidInput below MIN_SPLIT_PART_BYTES (64) or when a non-stream chunk exceeds a part.
constexpr uint64_t MIN_SPLIT_PART_BYTES = 64;
std::vector<std::vector<uint8_t>> split_archive(const uint8_t* archive, size_t len, uint64_t max_part_bytes, int device);
// Same, parts written back to back into out (copies by several threads); part_lens always receives the layout, the return value is
// the total length.  When out is null, cap too small or there are more than max_parts parts, nothing is written (sizing call).
uint64_t split_archive_into(const uint8_t* archive, size_t len, uint64_t max_part_bytes, int device, uint8_t* out, uint64_t cap,
                            std::vector<uint64_t>& part_lens, uint64_t max_parts = UINT64_MAX);

// ---- the file-system side of the path (SURVEY 8f "next", item 1): the CLI's extract / create data flow around the kernels
struct IoStats {
    uint64_t files = 0, dirs = 0, skipped = 0, bytes = 0;   // skipped: links and other non-file kinds, entries that failed
    double index_ms = 0, gpu_ms = 0, io_ms = 0, total_ms = 0;
};
// lib/src/entry/name.rs:148 `sanitize`: only normal path components survive ("", ".", ".." and roots are dropped)
std::string sanitize_entry_name(const std::string& name);
// `pna extract` (cli/src/command/extract.rs:868-1019 run_extract_archive over an mmap, :1301-1366 extract_file_entry): index
// the archive (caller: read_header_from_slice over the mapping), decode on the GPU in windows of at most window_bytes of
// output (pinned, double-buffered) and write the files with io_threads writers while the next window decodes.
// status (optional) receives one code per files() entry.
IoStats extract_to_dir(Archive& a, const ReadOptions& opt, const std::string& out_dir, int device, int workers, uint64_t group_bytes,
                       uint64_t window_bytes, int io_threads, bool verify, int32_t* status);
// `pna create` (cli/src/command/core.rs:889-913 write_from_path, create.rs:575 create_archive_file): read the files with
// io_threads readers into pinned memory, one GPU encode pass, write the archive file.
IoStats create_from_files(const std::vector<std::pair<std::string, std::string>>& name_and_path, const WriteOptions& opt,
                          uint32_t max_chunk_size, const std::string& archive_path, int device, int workers, uint64_t group_bytes,
                          int io_threads);

}  // namespace pna
extern "C" {
#endif

/* flat C view for bindings (tests/bench use it through ctypes) */
typedef struct pnah_archive pnah_archive;
typedef struct {
    uint8_t kind, data_kind, compression, encryption, cipher_mode;
    uint32_t n_bodies;
    uint64_t compressed_size, raw_file_size;
    const char* name;
    const char* phsf; /* NULL when absent */
} pnah_entry_info;
typedef struct { uint64_t files, dirs, skipped, bytes; double index_ms, gpu_ms, io_ms, total_ms; } pnah_io_stats;
int pnah_open(const uint8_t* buf, uint64_t len, pnah_archive** out, char* err, uint64_t errcap);
int pnah_open_multipart(const uint8_t* const* parts, const uint64_t* lens, uint32_t n_parts, int pinned_device /* -1: pageable */,
                        pnah_archive** out, char* err, uint64_t errcap);   /* split archive: parts in order; the handle owns a joined copy */
/* split writer: parts are written back to back into out (cap bytes), their lengths into part_lens (max_parts slots).  PNA_E_NOSPACE
 * when either is too small (or out is NULL: sizing call, no copy and no GPU work); *total, *n_parts and the first max_parts part_lens
 * always carry the layout. */
int pnah_split(const uint8_t* archive, uint64_t len, uint64_t max_part_bytes, int device, uint8_t* out, uint64_t cap, uint64_t* total,
               uint64_t* part_lens, uint32_t max_parts, uint32_t* n_parts, char* err, uint64_t errcap);
int pnah_open_file(const char* path, pnah_archive** out, char* err, uint64_t errcap);   /* mmap; the handle owns the mapping */
int pnah_extract_to_dir(pnah_archive* a, const char* out_dir, int device, int workers, uint64_t group_bytes, uint64_t window_bytes,
                        int io_threads, int verify, pnah_io_stats* stats, int32_t* status /* per file, may be NULL */, char* err,
                        uint64_t errcap);
int pnah_create_from_files(uint32_t n, const char* const* names, const char* const* paths, uint8_t compression, int32_t level,
                           uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf, uint32_t max_chunk_size,
                           const char* archive_path, int device, int workers, uint64_t group_bytes, int io_threads,
                           pnah_io_stats* stats, char* err, uint64_t errcap);
void pnah_close(pnah_archive* a);
uint32_t pnah_entry_count(pnah_archive* a);
int pnah_entry_get(pnah_archive* a, uint32_t i, pnah_entry_info* info);
uint32_t pnah_chunk_count(pnah_archive* a);
int pnah_set_key(pnah_archive* a, const char* phsf, const uint8_t key[32]);
int pnah_prepare(pnah_archive* a, int device, char* err, uint64_t errcap);
uint32_t pnah_file_count(pnah_archive* a);
int pnah_file_get(pnah_archive* a, uint32_t i, const char** name, uint64_t* size);
int pnah_file_sizes(pnah_archive* a, uint64_t* sizes, int32_t* status /* may be NULL */);   /* bulk form of pnah_file_get */
int pnah_extract_files(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers,
                       uint64_t group_bytes, int verify, char* err, uint64_t errcap);
/* files [first, last) only; file i lands at out + offsets[i] - offsets[first] (windows of a large archive; a second take of a file
 * that failed with PNA_E_NOSPACE because its fSIZ understated it -- pnah_file_get then reports the required size) */
int pnah_extract_range(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, int device, int workers, uint64_t group_bytes,
                       int verify, uint64_t first, uint64_t last, char* err, uint64_t errcap);
/* over several GPUs of the box: entries are partitioned by entry (a shared queue of entry groups), no collective */
int pnah_extract_files_on(pnah_archive* a, uint8_t* out, const uint64_t* offsets, int32_t* status, const int* devices, uint32_t n_devices,
                          int workers_per_device, uint64_t group_bytes, int verify, char* err, uint64_t errcap);
int pnah_create_on(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, const uint8_t* ivs /* NULL: drawn */,
                   uint8_t compression, int32_t level, uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf,
                   uint32_t max_chunk_size, const int* devices, uint32_t n_devices, int workers_per_device, uint64_t group_bytes, uint8_t* out,
                   uint64_t cap, uint64_t* out_len, char* err, uint64_t errcap);
uint64_t pnah_create_solid_bound(uint32_t n, const char* const* names, const uint64_t* lens, uint8_t compression, uint8_t encryption,
                                 uint8_t cipher_mode, const char* phsf, uint32_t max_chunk_size);
int pnah_create_solid(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, uint8_t compression, int32_t level,
                      uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf, uint32_t max_chunk_size, int device,
                      uint8_t* out, uint64_t cap, uint64_t* out_len, char* err, uint64_t errcap);
int pnah_create(uint32_t n, const char* const* names, const uint8_t* const* data, const uint64_t* lens, const uint8_t* ivs,
                uint8_t compression, int32_t level, uint8_t encryption, uint8_t cipher_mode, const uint8_t key[32], const char* phsf,
                uint32_t max_chunk_size, int device, int workers, uint64_t group_bytes, uint8_t* out, uint64_t cap, uint64_t* out_len,
                char* err, uint64_t errcap);
uint64_t pnah_create_bound(uint32_t n, const char* const* names, const uint64_t* lens, uint8_t compression, uint8_t encryption,
                           const char* phsf, uint32_t max_chunk_size);

#ifdef __cplusplus
}
#endif
#endif
