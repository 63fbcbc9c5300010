// matcher.cu -- orbm_* C ABI: host wrappers around the matching kernels.
// Mirrors ORB_SLAM2::ORBmatcher (R21/include/ORBmatcher.h:37-102, R21/src/ORBmatcher.cc).
#include "internal.h"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace orbcuda;

extern "C" {

// ORBmatcher::DescriptorDistance R21/src/ORBmatcher.cc:1647-1663 -- any exact 256-bit popcount is bit-identical
int orb_hamming256(const void* a, const void* b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

int orbm_knn2_device(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, int32_t* d_out,
                     int variant, void* stream) {
    if (!d_q || nq < 1 || !d_out || nm < 0 || (nm > 0 && !d_m)) { set_error("orbm_knn2_device: bad arguments"); return ORB_ERR_ARG; }
    if (variant < 0 || variant > 5) { set_error("orbm_knn2: unknown variant %d (0 = POPC, 1 = mma.sync integer MMA via shared memory, 2 = mma.sync streaming, 3 = tcgen05 integer MMA with TMEM accumulators, 4 = tcgen05 with the query operand in TMEM, 5 = tcgen05 on CTA pairs)", variant); return ORB_ERR_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_q) & 15) || (reinterpret_cast<uintptr_t>(d_m) & 15)) {
        set_error("orbm_knn2_device: descriptor arrays must be 16-byte aligned");
        return ORB_ERR_ARG;
    }
    if (launch_knn2(d_q, nq, d_m, nm, index_base, d_out, variant, (cudaStream_t)stream) < 0) {
        cuda_ok(cudaGetLastError(), "knn2 launch");
        return ORB_ERR_CUDA;
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_merge_top2_device(const int32_t* d_parts, int parts, int nq, int32_t* d_out, void* stream) {
    if (!d_parts || parts < 1 || nq < 1 || !d_out) return ORB_ERR_ARG;
    launch_merge_top2(d_parts, parts, nq, d_out, (cudaStream_t)stream);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbm_merge_top2_host(const int32_t* parts_rec, int parts, int nq, int32_t* out) {
    if (!parts_rec || parts < 1 || nq < 0 || !out) return ORB_ERR_ARG;
    host_merge_top2(parts_rec, parts, nq, out);
    return ORB_OK;
}

int orbm_ratio_test_host(const int32_t* rec, int nq, float ratio, int th, int strict, int32_t* out_match) {
    if (!rec || !out_match) return ORB_ERR_ARG;
    for (int i = 0; i < nq; i++) {
        const int d1 = rec[4 * i], i1 = rec[4 * i + 1], d2 = rec[4 * i + 2];
        const bool ok = (strict ? d1 < th : d1 <= th) && (float)d1 < ratio * (float)d2;   // R21 ORBmatcher.cc:228-230 / :598-600
        out_match[i] = ok ? i1 : -1;
    }
    return ORB_OK;
}

static bool knn2_args_ok(const char* who, const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int variant) {
    if (!d_q || nq < 1 || nm < 0 || (nm > 0 && !d_m)) { set_error("%s: bad arguments", who); return false; }
    if (variant < 0 || variant > 5) { set_error("%s: unknown variant %d", who, variant); return false; }
    if ((reinterpret_cast<uintptr_t>(d_q) & 15) || (reinterpret_cast<uintptr_t>(d_m) & 15)) { set_error("%s: descriptor arrays must be 16-byte aligned", who); return false; }
    return true;
}

// 2-NN search + the reference's ratio test (R21/src/ORBmatcher.cc:228-230 / :598-600) without leaving the device:
// d_match[q] = index of the accepted nearest neighbour or -1; d_rec (optional) also receives the {d1, i1, d2, i2} records.
int orbm_knn2_ratio_device(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, float ratio, int th, int strict,
                           int32_t* d_rec, int32_t* d_match, int variant, void* stream) {
    if (!d_match || !knn2_args_ok("orbm_knn2_ratio_device", d_q, nq, d_m, nm, variant)) return ORB_ERR_ARG;
    const RatioTest rt = {d_match, ratio, th, strict};
    if (launch_knn2(d_q, nq, d_m, nm, index_base, d_rec, variant, (cudaStream_t)stream, nullptr, &rt) < 0) {
        cuda_ok(cudaGetLastError(), "knn2 launch");
        return ORB_ERR_CUDA;
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// The ratio test on records that are already on the device (e.g. the merged result of a sharded search).
int orbm_ratio_test_device(const int32_t* d_rec, int nq, float ratio, int th, int strict, int32_t* d_match, void* stream) {
    if (!d_rec || nq < 1 || !d_match) { set_error("orbm_ratio_test_device: bad arguments"); return ORB_ERR_ARG; }
    const RatioTest rt = {d_match, ratio, th, strict};
    launch_ratio_test(d_rec, nq, rt, (cudaStream_t)stream);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// Host arrays in and out.  Runs on the calling thread's matcher workspace (MatchCtx: one stream, a grow-only device arena
// and a pinned staging arena per thread): no stream creation, no cudaMalloc and no pageable copy per call.  Maps larger
// than 8 MB are copied straight from the caller's memory (the driver stages them) so the pinned arena stays small.
int orbm_knn2(const uint8_t* queries, int nq, const uint8_t* map, int64_t nm, int64_t index_base, int32_t* best_idx,
              int32_t* best_dist, int32_t* second_dist, int32_t* second_idx, int variant, int device) {
    if (!queries || nq < 1 || nm < 0 || (nm > 0 && !map) || !best_idx || !best_dist || !second_dist) {
        set_error("orbm_knn2: bad arguments");
        return ORB_ERR_ARG;
    }
    if (variant < 0 || variant > 5) { set_error("orbm_knn2: unknown variant %d", variant); return ORB_ERR_ARG; }
    const size_t qb = (size_t)nq * 32, mb = (size_t)std::max<int64_t>(nm, 1) * 32, rb = (size_t)nq * 16;
    const bool big_map = mb > ((size_t)8 << 20);
    MatchCtx& cx = match_ctx();
    if (!cx.begin(device, qb + mb + rb + 4 * 256, qb + (big_map ? 0 : mb) + rb + 4 * 256)) return ORB_ERR_CUDA;
    const uint8_t* d_q = (const uint8_t*)cx.upload(queries, qb);
    uint8_t* d_m = nullptr;
    if (big_map) {
        d_m = (uint8_t*)cx.dalloc(mb);
        if (d_m && !cuda_ok(cudaMemcpyAsync(d_m, map, (size_t)nm * 32, cudaMemcpyHostToDevice, cx.s()), "cudaMemcpyAsync")) return ORB_ERR_CUDA;
    } else {
        d_m = (uint8_t*)cx.upload(map, nm > 0 ? (size_t)nm * 32 : 0);
    }
    int32_t* d_out = (int32_t*)cx.dalloc(rb);
    if (!d_q || !d_m || !d_out) return ORB_ERR_CUDA;
    int rc = orbm_knn2_device(d_q, nq, d_m, nm, index_base, d_out, variant, cx.s());
    if (rc) { cudaStreamSynchronize(cx.s()); return rc; }
    std::vector<int32_t> rec((size_t)nq * 4);
    if (!cx.download(rec.data(), d_out, rb) || !cx.finish()) return ORB_ERR_CUDA;
    for (int i = 0; i < nq; i++) {
        best_dist[i] = rec[4 * i]; best_idx[i] = rec[4 * i + 1]; second_dist[i] = rec[4 * i + 2];
        if (second_idx) second_idx[i] = rec[4 * i + 3];
    }
    return ORB_OK;
}

}  // extern "C"
