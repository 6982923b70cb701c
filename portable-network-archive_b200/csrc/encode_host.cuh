// encode_host.cuh -- host side of seam 3 (included at the end of abi.cu).
extern "C" uint64_t pna_cuda_encode_bound(const pna_encode_desc* d) { return d ? d->plain.len + d->plain.len / 8 + 1024 + 48 : 0; }
extern "C" uint64_t pna_cuda_encode_crc_count(const pna_encode_desc* d) {
    if (!d) return 0;
    uint64_t cap = d->max_chunk_size ? d->max_chunk_size : 0xFFFFFFFFull;
    uint64_t b = pna_cuda_encode_bound(d);
    return (b + cap - 1) / cap + 1;
}
extern "C" int pna_cuda_encode_batch(pna_ctx*, const pna_encode_desc*, uint32_t, pna_buf*, uint32_t*, uint32_t*, int32_t*) { return PNA_E_INTERNAL; }
extern "C" int pna_cuda_encode_plan_create(pna_ctx*, const pna_encode_desc*, uint32_t, pna_plan**) { return PNA_E_INTERNAL; }
extern "C" int pna_cuda_encode_plan_run(pna_plan*) { return PNA_E_INTERNAL; }
extern "C" int pna_cuda_encode_plan_fetch(pna_plan*, pna_buf*, uint32_t*, uint32_t*, int32_t*) { return PNA_E_INTERNAL; }
