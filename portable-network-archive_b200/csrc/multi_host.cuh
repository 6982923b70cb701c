// multi_host.cuh -- a pna_ctx over several GPUs of one box (pna_cuda_init with n_devices > 1; included at the end of abi.cu).
//
// Entries are independent units (own IV, own compressed stream, own chunks), so every batch call shards BY ENTRY: greedy
// longest-processing-time on stream bytes, one single-device context and one host thread per GPU, no peer traffic and no
// collective (the reference's counterpart is the per-entry task fan-out of cli/src/command/extract.rs:868-1019).
// A multi-device plan is a list of ordinary single-device plans plus the index maps that put results back in caller order.
#pragma once
#include <thread>

namespace pna { namespace multi {

static int crc32(pna_ctx* root, const pna_span* spans, uint32_t n, uint32_t* crc_out) {
    std::vector<uint64_t> w(n);
    for (uint32_t i = 0; i < n; i++) w[i] = spans[i].len;
    const auto part = shard(w, root->devs.size());
    const int rc = for_devices(root->devs.size(), [&](size_t d) -> int {
        const auto& idx = part[d];
        if (idx.empty()) return PNA_OK;
        std::vector<pna_span> s(idx.size());
        std::vector<uint32_t> c(idx.size());
        for (size_t k = 0; k < idx.size(); k++) s[k] = spans[idx[k]];
        const int r = pna_cuda_crc32(root->devs[d], s.data(), (uint32_t)s.size(), c.data());
        if (r == PNA_OK) for (size_t k = 0; k < idx.size(); k++) crc_out[idx[k]] = c[k];
        return r;
    });
    if (rc) carry_error(root);
    return rc;
}

struct Plan {   // hangs off pna_plan::multi
    std::vector<pna_plan*> sub;                   // one per device (null: no entry landed there)
    std::vector<std::vector<uint32_t>> idx;       // sub[d] entry k == caller entry idx[d][k]
    std::vector<std::vector<uint32_t>> crc_idx;   // sub[d] CRC span k == caller span crc_idx[d][k]
    std::vector<uint32_t> where_dev, where_loc;   // caller entry -> (device, local index)
    uint32_t n_crc = 0;
};

static int decode_plan_create(pna_ctx* root, const pna_decode_desc* descs, uint32_t n, const uint8_t* image, uint64_t image_len,
                              const pna_span* crc_spans, const uint32_t* crc_expect, const int32_t* crc_entry, uint32_t n_spans, bool with_crc,
                              pna_plan** plan) {
    const size_t nd = root->devs.size();
    std::vector<uint64_t> w(n);
    for (uint32_t i = 0; i < n; i++) { uint64_t b = 0; for (uint32_t k = 0; k < descs[i].n_bodies; k++) b += descs[i].bodies[k].len; w[i] = b; }
    Plan* M = new Plan();
    M->idx = shard(w, nd);
    M->sub.assign(nd, nullptr);
    M->crc_idx.assign(nd, {});
    M->where_dev.assign(n, 0); M->where_loc.assign(n, 0);
    M->n_crc = n_spans;
    for (size_t d = 0; d < nd; d++) for (size_t k = 0; k < M->idx[d].size(); k++) { M->where_dev[M->idx[d][k]] = (uint32_t)d; M->where_loc[M->idx[d][k]] = (uint32_t)k; }
    // a chunk travels with its entry; archive-level chunks (entry -1) are spread round-robin
    for (uint32_t c = 0, rr = 0; c < n_spans; c++) M->crc_idx[crc_entry[c] >= 0 ? M->where_dev[crc_entry[c]] : (rr++ % nd)].push_back(c);
    const int rc = for_devices(nd, [&](size_t d) -> int {
        const auto& idx = M->idx[d];
        const auto& cidx = M->crc_idx[d];
        if (idx.empty() && cidx.empty()) return PNA_OK;
        std::vector<pna_decode_desc> dd(idx.size());
        for (size_t k = 0; k < idx.size(); k++) dd[k] = descs[idx[k]];
        std::vector<pna_span> cs(cidx.size());
        std::vector<uint32_t> ce(cidx.size());
        std::vector<int32_t> co(cidx.size());
        for (size_t k = 0; k < cidx.size(); k++) {
            cs[k] = crc_spans[cidx[k]]; ce[k] = crc_expect[cidx[k]];
            co[k] = crc_entry[cidx[k]] >= 0 ? (int32_t)M->where_loc[crc_entry[cidx[k]]] : -1;
        }
        if (image) return pna_cuda_decode_plan_create_in_image(root->devs[d], dd.data(), (uint32_t)dd.size(), image, image_len, cs.data(), ce.data(),
                                                               co.data(), (uint32_t)cs.size(), &M->sub[d]);
        if (with_crc) return pna_cuda_decode_plan_create_crc(root->devs[d], dd.data(), (uint32_t)dd.size(), cs.data(), ce.data(), co.data(),
                                                             (uint32_t)cs.size(), &M->sub[d]);
        return pna_cuda_decode_plan_create(root->devs[d], dd.data(), (uint32_t)dd.size(), &M->sub[d]);
    });
    if (rc) {
        carry_error(root);
        for (pna_plan* s : M->sub) if (s) pna_cuda_plan_destroy(s);
        delete M;
        return rc;
    }
    pna_plan* P = new pna_plan();
    P->ctx = root; P->kind = 0; P->n = n; P->multi = M;
    *plan = P;
    return PNA_OK;
}

template <class F>
static int each_sub(pna_plan* P, F&& f) {
    Plan* M = P->multi;
    const int rc = for_devices(M->sub.size(), [&](size_t d) -> int { return M->sub[d] ? f(d, M->sub[d]) : PNA_OK; });
    if (rc) carry_error(P->ctx);
    return rc;
}
static int decode_plan_run(pna_plan* P) { return each_sub(P, [](size_t, pna_plan* s) { return pna_cuda_decode_plan_run(s); }); }
static int decode_plan_fetch(pna_plan* P, pna_buf* out, int32_t* status) {
    Plan* M = P->multi;
    return each_sub(P, [&](size_t d, pna_plan* s) -> int {
        const auto& idx = M->idx[d];
        std::vector<pna_buf> b(idx.size());
        std::vector<int32_t> st(idx.size());
        for (size_t k = 0; k < idx.size(); k++) b[k] = out[idx[k]];
        const int r = pna_cuda_decode_plan_fetch(s, b.data(), st.data());
        if (r == PNA_OK) for (size_t k = 0; k < idx.size(); k++) { out[idx[k]].len = b[k].len; status[idx[k]] = st[k]; }
        return r;
    });
}
static int decode_plan_lengths(pna_plan* P, uint64_t* out_len, int32_t* status) {
    Plan* M = P->multi;
    return each_sub(P, [&](size_t d, pna_plan* s) -> int {
        const auto& idx = M->idx[d];
        std::vector<uint64_t> l(idx.size());
        std::vector<int32_t> st(idx.size());
        const int r = pna_cuda_decode_plan_lengths(s, l.data(), st.data());
        if (r == PNA_OK) for (size_t k = 0; k < idx.size(); k++) { out_len[idx[k]] = l[k]; status[idx[k]] = st[k]; }
        return r;
    });
}
static int plan_crc_results(pna_plan* P, uint32_t* crc_out, uint32_t* n_broken) {
    Plan* M = P->multi;
    std::vector<uint32_t> broken(M->sub.size(), 0);
    const int rc = each_sub(P, [&](size_t d, pna_plan* s) -> int {
        const auto& cidx = M->crc_idx[d];
        std::vector<uint32_t> c(cidx.size());
        const int r = pna_cuda_plan_crc_results(s, crc_out ? c.data() : nullptr, &broken[d]);
        if (r == PNA_OK && crc_out) for (size_t k = 0; k < cidx.size(); k++) crc_out[cidx[k]] = c[k];
        return r;
    });
    if (n_broken) { *n_broken = 0; for (uint32_t b : broken) *n_broken += b; }
    return rc;
}
static int route(pna_plan* P, uint32_t entry, pna_plan** sub, uint32_t* local) {
    Plan* M = P->multi;
    if (entry >= P->n || !M->sub[M->where_dev[entry]]) return PNA_E_BAD_ARG;
    *sub = M->sub[M->where_dev[entry]];
    *local = M->where_loc[entry];
    return PNA_OK;
}
static void sum_stats(pna_plan* P, uint64_t* a, uint64_t* b, uint64_t* c, bool counts) {
    uint64_t ta = 0, tb = 0, tc = 0;
    for (pna_plan* s : P->multi->sub) {
        if (!s) continue;
        uint64_t x = 0, y = 0, z = 0;
        if (counts) pna_cuda_plan_counts(s, &x, &y, &z); else pna_cuda_plan_stats(s, &x, &y, &z);
        ta += x; tb += y; tc += z;
    }
    if (a) *a = ta;
    if (b) *b = tb;
    if (c) *c = tc;
}
static int stage_ms(pna_plan* P, float* ms, uint32_t cap) {   // per stage: the slowest device
    int ns = 0;
    for (uint32_t i = 0; i < cap; i++) ms[i] = 0.f;
    for (pna_plan* s : P->multi->sub) {
        if (!s) continue;
        float t[32] = {0};
        const int k = pna_cuda_plan_stage_ms(s, t, 32);
        if (k < 0) return k;
        ns = std::max(ns, k);
        for (int i = 0; i < k && (uint32_t)i < cap; i++) ms[i] = std::max(ms[i], t[i]);
    }
    return ns;
}
static void plan_destroy(pna_plan* P) {
    Plan* M = P->multi;
    for (pna_plan* s : M->sub) if (s) pna_cuda_plan_destroy(s);
    delete M;
    P->multi = nullptr;
    delete P;
}

// ---- encode
static int encode_plan_create(pna_ctx* root, const pna_encode_desc* descs, uint32_t n, pna_plan** plan) {
    const size_t nd = root->devs.size();
    std::vector<uint64_t> w(n);
    for (uint32_t i = 0; i < n; i++) w[i] = descs[i].plain.len;
    Plan* M = new Plan();
    M->idx = shard(w, nd);
    M->sub.assign(nd, nullptr);
    M->crc_idx.assign(nd, {});
    M->where_dev.assign(n, 0); M->where_loc.assign(n, 0);
    for (size_t d = 0; d < nd; d++) for (size_t k = 0; k < M->idx[d].size(); k++) { M->where_dev[M->idx[d][k]] = (uint32_t)d; M->where_loc[M->idx[d][k]] = (uint32_t)k; }
    const int rc = for_devices(nd, [&](size_t d) -> int {
        const auto& idx = M->idx[d];
        if (idx.empty()) return PNA_OK;
        std::vector<pna_encode_desc> dd(idx.size());
        for (size_t k = 0; k < idx.size(); k++) dd[k] = descs[idx[k]];
        return pna_cuda_encode_plan_create(root->devs[d], dd.data(), (uint32_t)dd.size(), &M->sub[d]);
    });
    if (rc) {
        carry_error(root);
        for (pna_plan* s : M->sub) if (s) pna_cuda_plan_destroy(s);
        delete M;
        return rc;
    }
    pna_plan* P = new pna_plan();
    P->ctx = root; P->kind = 1; P->n = n; P->multi = M;
    P->h_crc_count_bound.resize(n);
    for (uint32_t i = 0; i < n; i++) P->h_crc_count_bound[i] = pna_cuda_encode_crc_count(&descs[i]);
    *plan = P;
    return PNA_OK;
}
static int encode_plan_run(pna_plan* P) { return each_sub(P, [](size_t, pna_plan* s) { return pna_cuda_encode_plan_run(s); }); }
static int encode_plan_lengths(pna_plan* P, uint64_t* out_len, int32_t* status) {
    Plan* M = P->multi;
    return each_sub(P, [&](size_t d, pna_plan* s) -> int {
        const auto& idx = M->idx[d];
        std::vector<uint64_t> l(idx.size());
        std::vector<int32_t> st(idx.size());
        const int r = pna_cuda_encode_plan_lengths(s, l.data(), st.data());
        if (r == PNA_OK) for (size_t k = 0; k < idx.size(); k++) { out_len[idx[k]] = l[k]; status[idx[k]] = st[k]; }
        return r;
    });
}
static int encode_plan_fetch(pna_plan* P, pna_buf* out, uint32_t* fdat_crc_out, uint32_t* crc_count_out, int32_t* status) {
    Plan* M = P->multi;
    // per-device CRC lists, re-packed in caller order afterwards (the flat layout is consecutive per entry)
    std::vector<std::vector<uint32_t>> crcs(M->sub.size()), counts(M->sub.size());
    const int rc = each_sub(P, [&](size_t d, pna_plan* s) -> int {
        const auto& idx = M->idx[d];
        std::vector<pna_buf> b(idx.size());
        std::vector<int32_t> st(idx.size());
        uint64_t bound = 0;
        for (size_t k = 0; k < idx.size(); k++) { b[k] = out[idx[k]]; bound += P->h_crc_count_bound[idx[k]]; }
        crcs[d].resize(bound + 1); counts[d].resize(idx.size());
        const int r = pna_cuda_encode_plan_fetch(s, b.data(), crcs[d].data(), counts[d].data(), st.data());
        if (r == PNA_OK) for (size_t k = 0; k < idx.size(); k++) { out[idx[k]].len = b[k].len; status[idx[k]] = st[k]; }
        return r;
    });
    if (rc) return rc;
    std::vector<uint64_t> at(M->sub.size(), 0);   // read cursor per device: entries come back in local (= caller) order
    std::vector<std::vector<uint64_t>> start(M->sub.size());
    for (size_t d = 0; d < M->sub.size(); d++) {
        start[d].resize(M->idx[d].size());
        uint64_t pos = 0;
        for (size_t k = 0; k < M->idx[d].size(); k++) { start[d][k] = pos; pos += M->sub[d] ? counts[d][k] : 0; }
    }
    uint64_t pos = 0;
    for (uint32_t i = 0; i < P->n; i++) {
        const uint32_t d = M->where_dev[i], k = M->where_loc[i];
        const uint32_t c = M->sub[d] ? counts[d][k] : 0;
        if (crc_count_out) crc_count_out[i] = c;
        if (fdat_crc_out) for (uint32_t q = 0; q < c; q++) fdat_crc_out[pos + q] = crcs[d][start[d][k] + q];
        pos += c;
    }
    return PNA_OK;
}

}}  // namespace pna::multi
