//! Raw bindings to `include/pna_cuda.h` -- the C ABI that replaces three internal seams of `libpna`
//! (ChanTsune/Portable-Network-Archive v0.37.0):
//!
//! | seam | reference code | entry points |
//! |---|---|---|
//! | chunk CRC | `format::chunk_crc` / `validate_chunk_crc` (`lib/src/format/chunk.rs:7,16`) | [`pna_cuda_crc32`], [`pna_cuda_crc32_image`] |
//! | decode | `decrypt_reader` + `decompress_reader` (`lib/src/entry/read.rs:59-190`) | [`pna_cuda_decode_batch`], `pna_cuda_decode_plan_*` |
//! | encode | `get_writer` (`lib/src/entry/write.rs:189-273`) | [`pna_cuda_encode_batch`], `pna_cuda_encode_plan_*` |
//!
//! Every struct is `#[repr(C)]` and plain data; the function list is exactly the header's (checked by
//! `tests/test_rust_binding.py` in the repository that carries the CUDA sources).  `rust/patches/*.patch` show where
//! `libpna` calls into [`Gpu`].
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::io;
use std::os::raw::{c_char, c_int, c_void};

/// per-entry status == reference `io::ErrorKind` class
pub const PNA_OK: i32 = 0;
pub const PNA_E_INVALID_DATA: i32 = 1;
pub const PNA_E_UNEXPECTED_EOF: i32 = 2;
pub const PNA_E_INVALID_INPUT: i32 = 3;
pub const PNA_E_UNSUPPORTED: i32 = 4;
pub const PNA_E_NOSPACE: i32 = 5;
pub const PNA_E_OOM: i32 = 6;
pub const PNA_E_INTERNAL: i32 = 7;
pub const PNA_E_CUDA: i32 = 8;
pub const PNA_E_BAD_ARG: i32 = 9;

pub const PNA_COMPRESSION_NO: u8 = 0;
pub const PNA_COMPRESSION_DEFLATE: u8 = 1;
pub const PNA_COMPRESSION_ZSTD: u8 = 2;
pub const PNA_COMPRESSION_XZ: u8 = 4;
pub const PNA_ENCRYPTION_NO: u8 = 0;
pub const PNA_ENCRYPTION_AES: u8 = 1;
pub const PNA_ENCRYPTION_CAMELLIA: u8 = 2;
pub const PNA_CIPHER_CBC: u8 = 0;
pub const PNA_CIPHER_CTR: u8 = 1;
pub const PNA_CIPHER_GCM: u8 = 2;

/// opaque: one or several devices (streams, device arenas, pinned staging)
#[repr(C)]
pub struct pna_ctx {
    _private: [u8; 0],
}
/// opaque: a batch resident in HBM
#[repr(C)]
pub struct pna_plan {
    _private: [u8; 0],
}

/// borrowed host memory
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct pna_span {
    pub ptr: *const u8,
    pub len: u64,
}

/// caller-owned output; the library sets `len`
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct pna_buf {
    pub ptr: *mut u8,
    pub cap: u64,
    pub len: u64,
}

/// one entry's data stream: the FDAT (or SDAT) bodies in order
#[repr(C)]
#[derive(Clone, Copy)]
pub struct pna_decode_desc {
    pub bodies: *const pna_span,
    pub n_bodies: u32,
    pub compression: u8,
    pub encryption: u8,
    pub cipher_mode: u8,
    pub _pad: u8,
    pub key: [u8; 32],
    pub raw_size_hint: u64,
}

/// one entry to build: plaintext in, `IV || cipher(compress(plain))` out
#[repr(C)]
#[derive(Clone, Copy)]
pub struct pna_encode_desc {
    pub plain: pna_span,
    pub compression: u8,
    pub encryption: u8,
    pub cipher_mode: u8,
    pub _pad: u8,
    pub level: i32,
    pub key: [u8; 32],
    pub iv: [u8; 16],
    pub max_chunk_size: u32,
    pub stream_header: *const u8,
}

extern "C" {
    pub fn pna_cuda_init(out: *mut *mut pna_ctx, device_ids: *const c_int, n_devices: c_int) -> c_int;
    pub fn pna_cuda_device_count(ctx: *mut pna_ctx) -> c_int;
    pub fn pna_cuda_device_id(ctx: *mut pna_ctx, i: c_int) -> c_int;
    pub fn pna_cuda_destroy(ctx: *mut pna_ctx);
    pub fn pna_cuda_strerror(status: i32) -> *const c_char;
    pub fn pna_cuda_last_error(ctx: *mut pna_ctx) -> *const c_char;
    pub fn pna_cuda_host_alloc(ctx: *mut pna_ctx, bytes: u64) -> *mut c_void;
    pub fn pna_cuda_host_free(ctx: *mut pna_ctx, p: *mut c_void);
    pub fn pna_cuda_stream(ctx: *mut pna_ctx) -> *mut c_void;
    pub fn pna_cuda_launch_count(ctx: *mut pna_ctx) -> u64;
    pub fn pna_cuda_transfer_probe(ctx: *mut pna_ctx, h2d_src: *const u8, h2d_bytes: u64, d2h_dst: *mut u8, d2h_bytes: u64, h2d_ms: *mut f32, d2h_ms: *mut f32, both_ms: *mut f32) -> c_int;
    pub fn pna_cuda_crc32(ctx: *mut pna_ctx, type_and_data: *const pna_span, n: u32, crc_out: *mut u32) -> c_int;
    pub fn pna_cuda_crc32_image(ctx: *mut pna_ctx, image: *const u8, image_len: u64, span_off: *const u64, span_len: *const u64, n: u32, crc_out: *mut u32) -> c_int;
    pub fn pna_cuda_decode_size_bound(compression: u8, stream_len: u64) -> u64;
    pub fn pna_cuda_size_hint_trusted(compression: u8, stream_len: u64, hint: u64) -> c_int;
    pub fn pna_cuda_decode_batch(ctx: *mut pna_ctx, descs: *const pna_decode_desc, n: u32, out: *mut pna_buf, status: *mut i32) -> c_int;
    pub fn pna_cuda_decode_plan_create(ctx: *mut pna_ctx, descs: *const pna_decode_desc, n: u32, plan: *mut *mut pna_plan) -> c_int;
    pub fn pna_cuda_decode_plan_create_crc(ctx: *mut pna_ctx, descs: *const pna_decode_desc, n: u32, crc_spans: *const pna_span, crc_expect: *const u32, crc_entry: *const i32, n_spans: u32, plan: *mut *mut pna_plan) -> c_int;
    pub fn pna_cuda_decode_plan_create_in_image(ctx: *mut pna_ctx, descs: *const pna_decode_desc, n: u32, image: *const u8, image_len: u64, crc_spans: *const pna_span, crc_expect: *const u32, crc_entry: *const i32, n_spans: u32, plan: *mut *mut pna_plan) -> c_int;
    pub fn pna_cuda_plan_crc_results(plan: *mut pna_plan, crc_out: *mut u32, n_broken: *mut u32) -> c_int;
    pub fn pna_cuda_decode_plan_run(plan: *mut pna_plan) -> c_int;
    pub fn pna_cuda_decode_plan_fetch(plan: *mut pna_plan, out: *mut pna_buf, status: *mut i32) -> c_int;
    pub fn pna_cuda_decode_plan_lengths(plan: *mut pna_plan, out_len: *mut u64, status: *mut i32) -> c_int;
    pub fn pna_cuda_decode_plan_crc32_out(plan: *mut pna_plan, entry: u32, span_off: *const u64, span_len: *const u64, n: u32, crc_out: *mut u32) -> c_int;
    pub fn pna_cuda_decode_plan_fetch_ranges(plan: *mut pna_plan, entry: u32, src_off: *const u64, len: *const u64, dst: *const *mut u8, n: u32) -> c_int;
    pub fn pna_cuda_plan_stats(plan: *mut pna_plan, stream_bytes: *mut u64, plain_bytes: *mut u64, launches_per_run: *mut u64) -> c_int;
    pub fn pna_cuda_plan_counts(plan: *mut pna_plan, n_blocks: *mut u64, n_sequences: *mut u64, literal_bytes: *mut u64) -> c_int;
    pub fn pna_cuda_plan_stage_ms(plan: *mut pna_plan, ms: *mut f32, cap: u32) -> c_int;
    pub fn pna_cuda_stage_name(stage: u32) -> *const c_char;
    pub fn pna_cuda_plan_destroy(plan: *mut pna_plan);
    pub fn pna_cuda_encode_bound(desc: *const pna_encode_desc) -> u64;
    pub fn pna_cuda_encode_crc_count(desc: *const pna_encode_desc) -> u64;
    pub fn pna_cuda_encode_batch(ctx: *mut pna_ctx, descs: *const pna_encode_desc, n: u32, out: *mut pna_buf, fdat_crc_out: *mut u32, crc_count_out: *mut u32, status: *mut i32) -> c_int;
    pub fn pna_cuda_encode_plan_create(ctx: *mut pna_ctx, descs: *const pna_encode_desc, n: u32, plan: *mut *mut pna_plan) -> c_int;
    pub fn pna_cuda_encode_plan_run(plan: *mut pna_plan) -> c_int;
    pub fn pna_cuda_encode_plan_lengths(plan: *mut pna_plan, out_len: *mut u64, status: *mut i32) -> c_int;
    pub fn pna_cuda_encode_plan_fetch(plan: *mut pna_plan, out: *mut pna_buf, fdat_crc_out: *mut u32, crc_count_out: *mut u32, status: *mut i32) -> c_int;
    pub fn pna_cuda_encode_plan_fetch_region(plan: *mut pna_plan, out: *mut pna_buf, region: *mut u8, region_len: u64, fdat_crc_out: *mut u32, crc_count_out: *mut u32, status: *mut i32) -> c_int;
    pub fn pna_cuda_encode_stage_name(stage: u32) -> *const c_char;
    pub fn pna_cuda_gcm_stream_key(k_master: *const u8, stream_header: *const u8, stream_header_len: u64, header_type: *const u8, header_data: *const u8, header_len: u64, phsf: *const u8, phsf_len: u64, out_key: *mut u8) -> i32;
    pub fn pna_cuda_gcm_stream_header(k_master: *const u8, salt: *const u8, nonce_prefix: *const u8, segment_size: u32, out_header: *mut u8) -> i32;
    pub fn pna_cuda_ecb(ctx: *mut pna_ctx, encryption: c_int, encrypt: c_int, key: *const u8, in_: *const u8, n_bytes: u64, out: *mut u8) -> c_int;
}

/// `status[]` code -> the `io::ErrorKind` the reference returns at the same place (`include/pna_cuda.h` enum)
pub fn error_kind(status: i32) -> io::ErrorKind {
    match status {
        PNA_E_INVALID_DATA => io::ErrorKind::InvalidData,
        PNA_E_UNEXPECTED_EOF => io::ErrorKind::UnexpectedEof,
        PNA_E_INVALID_INPUT | PNA_E_BAD_ARG => io::ErrorKind::InvalidInput,
        PNA_E_UNSUPPORTED => io::ErrorKind::Unsupported,
        PNA_E_OOM => io::ErrorKind::OutOfMemory,
        _ => io::ErrorKind::Other,
    }
}

/// A context over one or several B200s.  `Send + Sync`: a single-device context serialises batch calls internally, a
/// multi-device one shards every batch by entry across its devices (no collective; entries are independent).
pub struct Gpu {
    ctx: *mut pna_ctx,
}
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}

impl Gpu {
    /// `devices`: CUDA ordinals; empty = every visible device.  Fails when no sm_100 device is usable -- there is no CPU fallback.
    pub fn new(devices: &[i32]) -> io::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe {
            if devices.is_empty() {
                pna_cuda_init(&mut ctx, std::ptr::null(), 0)
            } else {
                pna_cuda_init(&mut ctx, devices.as_ptr(), devices.len() as c_int)
            }
        };
        if rc != PNA_OK {
            return Err(io::Error::new(io::ErrorKind::Other, "pna_cuda_init: no usable sm_100 device (there is no CPU fallback)"));
        }
        Ok(Self { ctx })
    }

    fn check(&self, rc: c_int) -> io::Result<()> {
        if rc == PNA_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(pna_cuda_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(io::Error::new(error_kind(rc), msg))
    }

    /// seam 1: `crc32(type || data)` of every span in one batch (`format::chunk_crc`, lib/src/format/chunk.rs:7-12)
    pub fn chunk_crcs(&self, type_and_data: &[&[u8]]) -> io::Result<Vec<u32>> {
        let spans: Vec<pna_span> = type_and_data.iter().map(|s| pna_span { ptr: s.as_ptr(), len: s.len() as u64 }).collect();
        let mut out = vec![0u32; spans.len()];
        self.check(unsafe { pna_cuda_crc32(self.ctx, spans.as_ptr(), spans.len() as u32, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// seam 2: decode a batch of entries.  Sizes come from the stream, never from `fSIZ`: the first call runs with empty
    /// buffers where no trustworthy hint exists (PNA_E_NOSPACE + required length), the second with room for every entry.
    pub fn decode(&self, descs: &[pna_decode_desc]) -> io::Result<Vec<io::Result<Vec<u8>>>> {
        let n = descs.len();
        let mut bufs = vec![pna_buf { ptr: std::ptr::null_mut(), cap: 0, len: 0 }; n];
        let mut status = vec![0i32; n];
        self.check(unsafe { pna_cuda_decode_batch(self.ctx, descs.as_ptr(), n as u32, bufs.as_mut_ptr(), status.as_mut_ptr()) })?;
        let mut outs: Vec<Vec<u8>> = bufs.iter().zip(&status).map(|(b, &s)| if s == PNA_OK || s == PNA_E_NOSPACE { vec![0u8; b.len as usize] } else { Vec::new() }).collect();
        for (b, o) in bufs.iter_mut().zip(outs.iter_mut()) {
            b.ptr = o.as_mut_ptr();
            b.cap = o.len() as u64;
            b.len = 0;
        }
        self.check(unsafe { pna_cuda_decode_batch(self.ctx, descs.as_ptr(), n as u32, bufs.as_mut_ptr(), status.as_mut_ptr()) })?;
        Ok(outs
            .into_iter()
            .zip(bufs.iter().zip(&status))
            .map(|(mut o, (b, &s))| {
                if s == PNA_OK {
                    o.truncate(b.len as usize);
                    Ok(o)
                } else {
                    Err(io::Error::new(error_kind(s), unsafe { CStr::from_ptr(pna_cuda_strerror(s)) }.to_string_lossy().into_owned()))
                }
            })
            .collect())
    }

    /// seam 3: encode a batch of entries; returns per entry the stream and the CRC of every FDAT body of `max_chunk_size`.
    pub fn encode(&self, descs: &[pna_encode_desc]) -> io::Result<Vec<io::Result<(Vec<u8>, Vec<u32>)>>> {
        let n = descs.len();
        let mut outs: Vec<Vec<u8>> = descs.iter().map(|d| vec![0u8; unsafe { pna_cuda_encode_bound(d) } as usize]).collect();
        let mut bufs: Vec<pna_buf> = outs.iter_mut().map(|o| pna_buf { ptr: o.as_mut_ptr(), cap: o.len() as u64, len: 0 }).collect();
        let total_crc: u64 = descs.iter().map(|d| unsafe { pna_cuda_encode_crc_count(d) }).sum();
        let mut crcs = vec![0u32; total_crc as usize + 1];
        let mut counts = vec![0u32; n];
        let mut status = vec![0i32; n];
        self.check(unsafe {
            pna_cuda_encode_batch(self.ctx, descs.as_ptr(), n as u32, bufs.as_mut_ptr(), crcs.as_mut_ptr(), counts.as_mut_ptr(), status.as_mut_ptr())
        })?;
        let mut at = 0usize;
        Ok((0..n)
            .map(|i| {
                let c = counts[i] as usize;
                let mine = crcs[at..at + c].to_vec();
                at += c;
                if status[i] == PNA_OK {
                    let mut o = std::mem::take(&mut outs[i]);
                    o.truncate(bufs[i].len as usize);
                    Ok((o, mine))
                } else {
                    Err(io::Error::new(error_kind(status[i]), "encode failed"))
                }
            })
            .collect())
    }

    pub fn raw(&self) -> *mut pna_ctx {
        self.ctx
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { pna_cuda_destroy(self.ctx) }
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use std::mem::size_of;

    #[test]
    fn layouts_match_the_header() {
        assert_eq!(size_of::<pna_span>(), 16);
        assert_eq!(size_of::<pna_buf>(), 24);
        assert_eq!(size_of::<pna_decode_desc>(), 56);
        assert_eq!(size_of::<pna_encode_desc>(), 88);
    }

    /// lib/src/format/chunk.rs:31 and lib/src/io.rs:179, through the GPU
    #[test]
    fn reference_crc_kats() {
        let gpu = Gpu::new(&[0]).expect("needs a B200");
        let crcs = gpu.chunk_crcs(&[&b"FDAT\xAA\xBB\xCC\xDD"[..], &b"AEND"[..]]).unwrap();
        assert_eq!(crcs, vec![0x47F3_2B10, 0x6BF6_486D]);
    }
}
