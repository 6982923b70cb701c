//! lib/src/gpu.rs (new file, `mod gpu;` in lib/src/lib.rs next to `mod format;`): the process-wide GPU context the three
//! seam patches call.  One context over every visible B200 -- batch calls shard by entry inside `libpna_cuda`
//! (entries are independent units: own IV, own compressed stream, own chunks).
use std::io;
use std::sync::OnceLock;

use pna_cuda_sys::Gpu;

static GPU: OnceLock<Result<Gpu, String>> = OnceLock::new();

/// The context, created on first use.  There is no CPU fallback: without a usable sm_100 device every data-chunk
/// operation of the library fails with this error.
pub(crate) fn get() -> io::Result<&'static Gpu> {
    GPU.get_or_init(|| {
        let devices: Vec<i32> = std::env::var("PNA_CUDA_DEVICES")
            .ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect())
            .unwrap_or_default(); // empty = every visible device
        Gpu::new(&devices).map_err(|e| e.to_string())
    })
    .as_ref()
    .map_err(|e| io::Error::new(io::ErrorKind::Other, e.clone()))
}
