/*
 * pna_cuda.h -- C ABI of libpna_cuda.so, the B200 (sm_100a) implementation of PNA's per-entry
 * data-chunk pipeline.  This is the drop-in boundary: the three internal seams of `libpna`
 * (reference = ChanTsune/Portable-Network-Archive v0.37.0, paths relative to /root/reference)
 * are replaced by the batch calls below; everything is POD, `extern "C"`, no exceptions cross.
 *
 *   seam 1  CRC     format::chunk_crc / validate_chunk_crc        lib/src/format/chunk.rs:7,16
 *                   (callers io::read_chunk lib/src/io.rs:141, bytes::read_chunk lib/src/bytes.rs:64,
 *                    Chunk::crc lib/src/chunk/traits.rs:49)                      -> pna_cuda_crc32
 *   seam 2  decode  decrypt_reader + decompress_reader             lib/src/entry/read.rs:59,171
 *                   (NormalEntry::reader lib/src/entry.rs:1150,
 *                    SolidEntry::entries lib/src/entry.rs:567)                   -> pna_cuda_decode_batch
 *   seam 3  encode  get_writer = compression_writer(encryption_writer)  lib/src/entry/write.rs:189,251,268
 *                   (FileEntryBuilder lib/src/entry/builder/file.rs:65-140, EntryBuilderCore::build
 *                    lib/src/entry/builder.rs:171, write_chunk CRC lib/src/io.rs:183) -> pna_cuda_encode_batch
 *
 * Ownership: the caller owns every host buffer for the duration of a call; the library owns all
 * device memory.  Nothing returned must be freed except pna_ctx / pna_plan handles and
 * pna_cuda_host_alloc() memory.
 * Errors: the function return is a context-level error (bad argument, CUDA failure); per-entry
 * results are in status[] and map 1:1 to the reference's io::ErrorKind classes (see enum).
 * Threading: a single-device pna_ctx serialises batch calls internally; use one ctx per host thread for concurrency.
 * Multi-GPU: a pna_ctx created over several devices shards every batch call by entry across them (greedy
 * longest-processing-time on stream bytes, one host thread per GPU; entries are independent, the path has no collective
 * and no peer traffic -- the reference's counterpart is the per-entry task fan-out of cli/src/command/extract.rs:868-1019).
 * There is NO CPU fallback: every entry point fails with PNA_E_CUDA when no sm_100 device is usable.
 */
#ifndef PNA_CUDA_H
#define PNA_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-entry status == reference io::ErrorKind class */
enum {
    PNA_OK = 0,
    PNA_E_INVALID_DATA = 1,   /* "broken chunk" format/chunk.rs:18; bad PKCS#7 cipher/block/read.rs:101; corrupt zstd */
    PNA_E_UNEXPECTED_EOF = 2, /* partial CBC block block/read.rs:90; stream shorter than IV entry/read.rs:80; truncated zstd frame */
    PNA_E_INVALID_INPUT = 3,  /* "corrupt deflate stream" (flate2 zio); bad key length stream/read.rs:27 */
    PNA_E_UNSUPPORTED = 4,    /* unknown codes entry/read.rs:152-163,184-187 */
    PNA_E_NOSPACE = 5,        /* out.cap too small; out.len = required size (two-pass sizing contract) */
    PNA_E_OOM = 6,            /* util/io.rs:19 */
    PNA_E_INTERNAL = 7,
    PNA_E_CUDA = 8,           /* context-level: CUDA runtime/driver failure or no device */
    PNA_E_BAD_ARG = 9
};

/* header byte codes, lib/src/entry/options.rs:237-247,483-491,596-604 */
enum { PNA_COMPRESSION_NO = 0, PNA_COMPRESSION_DEFLATE = 1, PNA_COMPRESSION_ZSTD = 2, PNA_COMPRESSION_XZ = 4 };
enum { PNA_ENCRYPTION_NO = 0, PNA_ENCRYPTION_AES = 1, PNA_ENCRYPTION_CAMELLIA = 2 };
enum { PNA_CIPHER_CBC = 0, PNA_CIPHER_CTR = 1, PNA_CIPHER_GCM = 2 };

typedef struct pna_ctx pna_ctx;   /* one or several devices: streams, device arenas, pinned staging */
typedef struct pna_plan pna_plan; /* a batch resident in HBM (used for kernel-only timing and re-runs) */

typedef struct { const uint8_t* ptr; uint64_t len; } pna_span;        /* borrowed host memory */
typedef struct { uint8_t* ptr; uint64_t cap; uint64_t len; } pna_buf; /* caller-owned output; library sets len */

/* One entry's data stream: the FDAT (or SDAT) bodies in order.  The 16-byte IV is the stream prefix when
 * encryption != 0 and may straddle bodies (lib/src/entry/read.rs:79-103; chunk boundaries are arbitrary,
 * lib/src/chunk/types.rs:315-320).  cipher_mode == PNA_CIPHER_GCM: the prefix is the 75-byte stream header followed by
 * { ciphertext || tag } segments (lib/src/cipher/gcm.rs:206-290), and `key` is the per-stream key from
 * pna_cuda_gcm_stream_key; a tag that does not verify, or a malformed / truncated layout, is PNA_E_INVALID_DATA. */
typedef struct {
    const pna_span* bodies;
    uint32_t n_bodies;
    uint8_t compression, encryption, cipher_mode, _pad; /* FHED bytes [3],[4],[5] / SHED [2],[3],[4] */
    uint8_t key[32];                                    /* KDF output (host side, lib/src/hash.rs:45); ignored if encryption==0 */
    uint64_t raw_size_hint;                             /* fSIZ (lib/src/entry.rs:817) or UINT64_MAX.  A HINT: untrusted archive
                                                         * input the reference never sizes anything from.  Values no stream of
                                                         * this length can decode to (pna_cuda_decode_size_bound) are ignored and
                                                         * the exact sizing pass runs instead; the decoded length always comes
                                                         * back in pna_buf.len / pna_cuda_decode_plan_lengths. */
} pna_decode_desc;

/* One entry to build: plaintext in, IV || cipher(compress(plain)) out (lib/src/entry/write.rs:268-273; the prefix chunk is
 * split off by the caller: builder.rs:62-69). */
typedef struct {
    pna_span plain;
    uint8_t compression, encryption, cipher_mode, _pad;
    int32_t level;          /* <0: reference default (zstd 3 compress/zstandard.rs:46, deflate 6 compress/deflate.rs:89).
                             * Selects the encoder setting: zstd 1-2 / deflate 1-3 fast (greedy parse, Predefined FSE tables);
                             * zstd 3-5 / deflate 4-6 default (per-block FSE tables chosen by cost); zstd >= 6 / deflate 7-9
                             * high (lazy parse); deflate 0 stored blocks; xz (0..9, default 6 compress/xz.rs:10-16) 0-3 greedy parse, lc = 0;
                             * 4-9 lazy parse, lc = 2; always one LZMA2 chunk per 32 KiB segment and a CRC32 check.  Sizes differ from the reference's at the same
                             * level (another encoder); every setting decodes with the reference's codecs. */
    uint8_t key[32];
    uint8_t iv[16];         /* caller-drawn (lib/src/entry/write.rs:108-111, random.rs:8); unused by GCM */
    uint32_t max_chunk_size; /* FDAT body cap for CRC emission; 0 = u32::MAX - one body (lib/src/util/io.rs:24-33) */
    /* cipher_mode == PNA_CIPHER_GCM only: the 75-byte stream header (pna_cuda_gcm_stream_header; salt and nonce prefix
     * caller-drawn, lib/src/entry/write.rs:81-99) -- it becomes the stream prefix where CBC/CTR put the IV -- and `key` is the
     * per-stream key (pna_cuda_gcm_stream_key).  Output: header || { ciphertext(segment_size) || tag(16) }..., the last segment
     * flagged final (lib/src/cipher/gcm.rs:44-90). */
    const uint8_t* stream_header;
} pna_encode_desc;

/* ---- context ---- */
/* device_ids: the CUDA devices this context owns (distinct, each an sm_100 part); NULL = devices 0 .. n_devices-1
 * (n_devices == 0: every visible device).  One device: an ordinary context.  Several: every batch / plan call below is
 * sharded by entry across them and results come back in caller order. */
int pna_cuda_init(pna_ctx** out, const int* device_ids, int n_devices);
int pna_cuda_device_count(pna_ctx* ctx);          /* devices owned by the context */
int pna_cuda_device_id(pna_ctx* ctx, int i);      /* CUDA ordinal of the i-th, -1 when out of range */
void pna_cuda_destroy(pna_ctx* ctx);
const char* pna_cuda_strerror(int32_t status);
const char* pna_cuda_last_error(pna_ctx* ctx);   /* text of the last context-level failure */
/* pinned host memory for end-to-end paths (cudaHostAlloc); pageable buffers are accepted everywhere too */
void* pna_cuda_host_alloc(pna_ctx* ctx, uint64_t bytes);
void pna_cuda_host_free(pna_ctx* ctx, void* p);
/* the CUDA stream every kernel of this ctx (its first device) is launched on (cudaStream_t), for CUDA-event timing by callers */
void* pna_cuda_stream(pna_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py "gpu_launches") */
uint64_t pna_cuda_launch_count(pna_ctx* ctx);

/* Transfer yardstick for end-to-end measurements: pinned -> HBM copy of h2d_bytes and HBM -> pinned copy of d2h_bytes, each
 * alone and both at once on two streams, timed with CUDA events (milliseconds).  The buffers are the caller's (pinned). */
int pna_cuda_transfer_probe(pna_ctx* ctx, const uint8_t* h2d_src, uint64_t h2d_bytes, uint8_t* d2h_dst, uint64_t d2h_bytes,
                            float* h2d_ms, float* d2h_ms, float* both_ms);

/* ---- seam 1: chunk CRC ---- */
/* crc_out[i] = CRC-32/ISO-HDLC over spans[i] (the caller passes type||data, lib/src/format/chunk.rs:7-12). */
int pna_cuda_crc32(pna_ctx* ctx, const pna_span* type_and_data, uint32_t n, uint32_t* crc_out);
/* Whole-archive variant for the index pass: `image` is uploaded once, spans are (offset,len) pairs inside it. */
int pna_cuda_crc32_image(pna_ctx* ctx, const uint8_t* image, uint64_t image_len, const uint64_t* span_off,
                         const uint64_t* span_len, uint32_t n, uint32_t* crc_out);

/* ---- seam 2: decode ---- */
/* largest size a compressed stream of stream_len bytes can decode to (zstd: 128 KiB per 4 bytes, deflate: 1032x, store: 1x) */
uint64_t pna_cuda_decode_size_bound(uint8_t compression, uint64_t stream_len);
/* 1 when the library lays the output out from this fSIZ value, 0 when it ignores it and sizes the entry exactly (beyond the
 * bound above, or above 1 GiB and more than 256x the stream): host layers apply the same rule before they size buffers */
int pna_cuda_size_hint_trusted(uint8_t compression, uint64_t stream_len, uint64_t hint);
int pna_cuda_decode_batch(pna_ctx* ctx, const pna_decode_desc* descs, uint32_t n, pna_buf* out, int32_t* status);
/* staged form: upload once, run the kernels any number of times on HBM-resident input, fetch once */
int pna_cuda_decode_plan_create(pna_ctx* ctx, const pna_decode_desc* descs, uint32_t n, pna_plan** plan);
/* Same, with seam 1 fused in: crc_spans are the chunks' type||data ranges (normally 4 bytes before each body in the
 * same archive buffer, so they ride on the same upload), crc_expect the stored CRCs, crc_entry[i] the entry the
 * chunk belongs to (-1: archive-level chunk).  A mismatch marks the entry PNA_E_INVALID_DATA ("broken chunk",
 * lib/src/format/chunk.rs:16-21) before it is decoded; pna_cuda_plan_crc_results returns all computed CRCs. */
int pna_cuda_decode_plan_create_crc(pna_ctx* ctx, const pna_decode_desc* descs, uint32_t n, const pna_span* crc_spans,
                                    const uint32_t* crc_expect, const int32_t* crc_entry, uint32_t n_spans, pna_plan** plan);
/* Same again for callers whose spans and bodies all lie inside ONE host allocation [image, image + image_len) -- an
 * archive buffer or mapping.  Only then may the library bridge the few bytes of chunk framing between neighbouring spans
 * and upload whole ranges in one copy (a million small files must not mean a million copies); without the declaration
 * (the two calls above) every span is copied on its own, because the gap between two independent allocations is not the
 * caller's memory.  crc_spans may be NULL with n_spans == 0 (no CRC check, e.g. after the caller verified the archive). */
int pna_cuda_decode_plan_create_in_image(pna_ctx* ctx, const pna_decode_desc* descs, uint32_t n, const uint8_t* image,
                                         uint64_t image_len, const pna_span* crc_spans, const uint32_t* crc_expect,
                                         const int32_t* crc_entry, uint32_t n_spans, pna_plan** plan);
int pna_cuda_plan_crc_results(pna_plan* plan, uint32_t* crc_out, uint32_t* n_broken);
int pna_cuda_decode_plan_run(pna_plan* plan);                 /* asynchronous on pna_cuda_stream(ctx) after first call */
int pna_cuda_decode_plan_fetch(pna_plan* plan, pna_buf* out, int32_t* status);
/* decoded length and status of every entry after pna_cuda_decode_plan_run (waits for the run; PNA_E_NOSPACE + length
 * when the entry was given a too small raw_size_hint): lets the caller size host buffers before fetching */
int pna_cuda_decode_plan_lengths(pna_plan* plan, uint64_t* out_len, int32_t* status);
/* Solid entries (lib/src/entry.rs:401-423): the decoded SDAT stream is itself a chunk sequence whose CRCs the reference
 * checks as it reads, and whose STORE entries are slices of that stream.  Both work on the decoded bytes where they
 * already are (HBM): CRC-32 of (offset, length) spans inside entry `entry`'s output, and device-to-host copies of
 * ranges of it straight to their destinations.  pna_cuda_decode_plan_lengths must have been called. */
int pna_cuda_decode_plan_crc32_out(pna_plan* plan, uint32_t entry, const uint64_t* span_off, const uint64_t* span_len,
                                   uint32_t n, uint32_t* crc_out);
int pna_cuda_decode_plan_fetch_ranges(pna_plan* plan, uint32_t entry, const uint64_t* src_off, const uint64_t* len,
                                      uint8_t* const* dst, uint32_t n);
/* stream / decoded byte totals of a plan (algorithmic bytes for the roofline: C and U) */
int pna_cuda_plan_stats(pna_plan* plan, uint64_t* stream_bytes, uint64_t* plain_bytes, uint64_t* launches_per_run);
/* zstd work of a prepared plan: blocks, sequences (8-byte records between the entropy and LZ stages) and
 * Huffman-coded literal bytes (16-byte padded per block) -- the per-stage algorithmic bytes in DESIGN.md */
int pna_cuda_plan_counts(pna_plan* plan, uint64_t* n_blocks, uint64_t* n_sequences, uint64_t* literal_bytes);
/* per-stage device time (CUDA events on pna_cuda_stream) of the plan's last run; returns the number of stages */
int pna_cuda_plan_stage_ms(pna_plan* plan, float* ms, uint32_t cap);
const char* pna_cuda_stage_name(uint32_t stage);
void pna_cuda_plan_destroy(pna_plan* plan);

/* ---- seam 3: encode ---- */
/* out[i] receives IV || cipher(compress(plain)) (no IV when encryption==0).  fdat_crc_out (optional) receives,
 * per entry, crc32("FDAT" || body) for each body of at most max_chunk_size bytes; crc_count_out[i] = number
 * written for entry i, laid out consecutively (caller sizes it with pna_cuda_encode_crc_count). */
uint64_t pna_cuda_encode_bound(const pna_encode_desc* desc);
uint64_t pna_cuda_encode_crc_count(const pna_encode_desc* desc);
int pna_cuda_encode_batch(pna_ctx* ctx, const pna_encode_desc* descs, uint32_t n, pna_buf* out,
                          uint32_t* fdat_crc_out, uint32_t* crc_count_out, int32_t* status);
int pna_cuda_encode_plan_create(pna_ctx* ctx, const pna_encode_desc* descs, uint32_t n, pna_plan** plan);
int pna_cuda_encode_plan_run(pna_plan* plan);
/* produced stream length of every entry (waits for the run): lets the caller place the streams before fetching them */
int pna_cuda_encode_plan_lengths(pna_plan* plan, uint64_t* out_len, int32_t* status);
int pna_cuda_encode_plan_fetch(pna_plan* plan, pna_buf* out, uint32_t* fdat_crc_out, uint32_t* crc_count_out,
                               int32_t* status);
/* The same, for a caller that owns one contiguous `region` into which every out[i] points (an archive being assembled, the
 * chunk frames between the streams still to be written): when the streams are many, small and close together they come down
 * in ONE copy of the span they cover, laid out on the device first -- the bytes of that span BETWEEN the streams are
 * unspecified afterwards.  Falls back to one copy per stream whenever that does not apply.  region == NULL: plain fetch. */
int pna_cuda_encode_plan_fetch_region(pna_plan* plan, pna_buf* out, uint8_t* region, uint64_t region_len, uint32_t* fdat_crc_out,
                                      uint32_t* crc_count_out, int32_t* status);

/* stage names of an encode plan for pna_cuda_plan_stage_ms (first 5 entries: lz_match, block_write, layout, cipher, crc) */
const char* pna_cuda_encode_stage_name(uint32_t stage);

/* ---- GCM STREAM key schedule (cipher mode 2; host only, once per entry like the password KDF) ---- */
/* decrypt_reader's GCM branch (lib/src/entry/read.rs:105-139) up to the cipher construction: parses the 75-byte stream header
 * (= the first bytes of the data stream, lib/src/cipher/aead.rs:92-150), checks the key confirmation against k_master
 * (aead.rs:162-164, :115-121) and derives the per-stream key HKDF-SHA-256(k_master, salt, entry context) (aead.rs:166-208) that
 * goes into pna_decode_desc.key / pna_encode_desc.key of a cipher_mode == PNA_CIPHER_GCM entry.  header_type = "FHED" or "SHED",
 * header_data / phsf = the Data fields of those chunks.  PNA_E_INVALID_DATA: short or out-of-range header, or key mismatch
 * (every AeadError maps to io::ErrorKind::InvalidData, lib/src/error.rs:67-74). */
int32_t pna_cuda_gcm_stream_key(const uint8_t k_master[32], const uint8_t* stream_header, uint64_t stream_header_len,
                                const uint8_t header_type[4], const uint8_t* header_data, uint64_t header_len,
                                const uint8_t* phsf, uint64_t phsf_len, uint8_t out_key[32]);
/* writer side (lib/src/entry/write.rs:81-106): salt(32) || nonce_prefix(7) || segment_size (BE) || key confirmation(32) */
int32_t pna_cuda_gcm_stream_header(const uint8_t k_master[32], const uint8_t salt[32], const uint8_t nonce_prefix[7],
                                   uint32_t segment_size, uint8_t out_header[75]);

/* ---- block-cipher primitives (test hooks for the KATs in lib/src/cipher.rs:256-292) ---- */
/* ECB over n 16-byte blocks with the same key schedule the stream kernels use. */
int pna_cuda_ecb(pna_ctx* ctx, int encryption, int encrypt, const uint8_t key[32], const uint8_t* in, uint64_t n_bytes,
                 uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
