// bow.cu -- Frame::ComputeBoW / KeyFrame::ComputeBoW (R21/src/Frame.cc:400-407, R21/src/KeyFrame.cc:60-69):
//   mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)
// i.e. DBoW2 TemplatedVocabulary::transform (third-party, not vendored by the reference; algorithm restated in
// oracle/bow_oracle.cc).  The vocabulary tree (k = 10, L = 6, ~1.1 M nodes, 35 MB of node descriptors for ORBvoc) is
// uploaded once per handle and stays in HBM; a warp walks one descriptor down the tree: lane c takes child c of the
// current node (Hamming distance to the node descriptor), a shuffle arg-min with the lowest-child tie rule (the
// reference keeps the first child on equal distance) picks the next node.  The per-feature (word, node, weight)
// triples come back to the host, where the two std::map-shaped results are built in the reference's accumulation
// order (the double-precision weight sums and the L1 norm are order dependent).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "internal.h"
#include "orbcuda.h"

namespace orbcuda {

struct Vocabulary {
    int device = 0;
    int n_nodes = 0, depth = 0;
    int* d_child_ptr = nullptr;
    int* d_child_idx = nullptr;
    uint32_t* d_desc = nullptr;
    int* d_word = nullptr;
    double* d_weight = nullptr;
};

__global__ void __launch_bounds__(128) bow_descend_kernel(const uint32_t* __restrict__ desc, int n, const int* __restrict__ child_ptr,
                                                          const int* __restrict__ child_idx, const uint32_t* __restrict__ node_desc,
                                                          const int* __restrict__ word_id, const double* __restrict__ weight, int nid_level,
                                                          int* __restrict__ out_word, int* __restrict__ out_node, double* __restrict__ out_weight) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= n) return;
    uint32_t q[8];
#pragma unroll
    for (int w = 0; w < 8; w++) q[w] = desc[(size_t)f * 8 + w];
    int node = 0, nid = 0, level = 0;
    while (true) {
        const int b = child_ptr[node], e = child_ptr[node + 1];
        if (e <= b) break;                      // leaf
        ++level;
        // (distance, child position) arg-min over the children, 32 at a time; the lowest position wins ties
        unsigned best = 0xffffffffu;
        for (int c0 = b; c0 < e; c0 += 32) {
            unsigned key = 0xffffffffu;
            if (c0 + lane < e) {
                const uint32_t* nd = node_desc + (size_t)child_idx[c0 + lane] * 8;
                int d = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) d += __popc(q[w] ^ nd[w]);
                key = ((unsigned)d << 20) | (unsigned)(c0 - b + lane);      // <= 2^20 children per node
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, off));
            best = min(best, key);
        }
        node = child_idx[b + (int)(best & 0xfffffu)];
        if (level == nid_level) nid = node;
    }
    if (lane == 0) { out_word[f] = word_id[node]; out_node[f] = nid; out_weight[f] = weight[node]; }
}

}  // namespace orbcuda

using namespace orbcuda;

extern "C" {

int orbv_create(const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc, const int32_t* word_id,
                const double* weight, int n_nodes, int depth_L, int device, orbv_handle_t* out) {
    if (!child_ptr || !node_desc || !word_id || !weight || n_nodes <= 0 || depth_L <= 0 || !out || child_ptr[0] != 0) { set_error("orbv_create: bad arguments"); return ORB_ERR_ARG; }
    const int n_edges = child_ptr[n_nodes];
    if (n_edges < 0 || (n_edges > 0 && !child_idx)) { set_error("orbv_create: bad child lists"); return ORB_ERR_ARG; }
    for (int i = 0; i < n_nodes; i++)
        if (child_ptr[i + 1] < child_ptr[i] || child_ptr[i + 1] - child_ptr[i] >= (1 << 20)) { set_error("orbv_create: bad child_ptr at node %d", i); return ORB_ERR_ARG; }
    for (int c = 0; c < n_edges; c++)
        if (child_idx[c] <= 0 || child_idx[c] >= n_nodes) { set_error("orbv_create: child index %d out of range", child_idx[c]); return ORB_ERR_ARG; }
    ORB_CUDA_TRY(cudaSetDevice(device));
    Vocabulary* v = new Vocabulary;
    v->device = device; v->n_nodes = n_nodes; v->depth = depth_L;
    auto up = [&](void** d, const void* h, size_t bytes) {
        if (!cuda_ok(cudaMalloc(d, std::max<size_t>(bytes, 16)), "cudaMalloc")) return false;
        return bytes == 0 || cuda_ok(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice), "cudaMemcpy");
    };
    if (!up((void**)&v->d_child_ptr, child_ptr, (size_t)(n_nodes + 1) * 4) || !up((void**)&v->d_child_idx, child_idx, (size_t)n_edges * 4) ||
        !up((void**)&v->d_desc, node_desc, (size_t)n_nodes * 32) || !up((void**)&v->d_word, word_id, (size_t)n_nodes * 4) ||
        !up((void**)&v->d_weight, weight, (size_t)n_nodes * 8)) {
        orbv_destroy(reinterpret_cast<orbv_handle_t>(v));
        return ORB_ERR_CUDA;
    }
    *out = reinterpret_cast<orbv_handle_t>(v);
    return ORB_OK;
}

int orbv_destroy(orbv_handle_t h) {
    Vocabulary* v = reinterpret_cast<Vocabulary*>(h);
    if (!v) return ORB_OK;
    cudaSetDevice(v->device);
    cudaFree(v->d_child_ptr); cudaFree(v->d_child_idx); cudaFree(v->d_desc); cudaFree(v->d_word); cudaFree(v->d_weight);
    delete v;
    return ORB_OK;
}

int orbv_transform(orbv_handle_t h, const uint8_t* desc, int n, int levelsup, int32_t* out_word, int32_t* out_node, double* out_weight) {
    Vocabulary* v = reinterpret_cast<Vocabulary*>(h);
    if (!v || n < 0 || (n && (!desc || !out_word || !out_node || !out_weight))) { set_error("orbv_transform: bad arguments"); return ORB_ERR_ARG; }
    if (n == 0) return ORB_OK;
    MatchCtx& cx = match_ctx();
    const size_t need = (size_t)n * (32 + 4 + 4 + 8) + 8 * 256;
    if (!cx.begin(v->device, need, need)) return ORB_ERR_CUDA;
    const uint32_t* d_desc = (const uint32_t*)cx.upload(desc, (size_t)n * 32);
    int* d_w = (int*)cx.dalloc((size_t)n * 4); int* d_n = (int*)cx.dalloc((size_t)n * 4); double* d_wt = (double*)cx.dalloc((size_t)n * 8);
    if (!d_desc || !d_w || !d_n || !d_wt) return ORB_ERR_CUDA;
    bow_descend_kernel<<<(n + 3) / 4, 128, 0, cx.s()>>>(d_desc, n, v->d_child_ptr, v->d_child_idx, v->d_desc, v->d_word, v->d_weight,
                                                          v->depth - levelsup, d_w, d_n, d_wt);
    ORB_CUDA_TRY(cudaGetLastError());
    if (!cx.download(out_word, d_w, (size_t)n * 4) || !cx.download(out_node, d_n, (size_t)n * 4) || !cx.download(out_weight, d_wt, (size_t)n * 8) ||
        !cx.finish()) return ORB_ERR_CUDA;
    return ORB_OK;
}

int orbv_bow_vectors(const int32_t* word, const int32_t* node, const double* weight, int n, int normalize_l1, int32_t* bow_words,
                     double* bow_values, int* n_words, int32_t* fv_nodes, int32_t* fv_ptr, int32_t* fv_idx, int* n_fv_nodes) {
    if (n < 0 || !n_words || !n_fv_nodes || !fv_ptr || (n && (!word || !node || !weight || !bow_words || !bow_values || !fv_nodes || !fv_idx))) {
        set_error("orbv_bow_vectors: bad arguments");
        return ORB_ERR_ARG;
    }
    // BowVector::addWeight / FeatureVector::addFeature in feature order, then BowVector::normalize(L1)
    std::map<int32_t, double> v;
    std::map<int32_t, std::vector<int32_t> > fv;
    for (int i = 0; i < n; i++) {
        if (weight[i] > 0) {
            v[word[i]] += weight[i];      // a new key starts from 0.0: 0.0 + w == w exactly
            fv[node[i]].push_back(i);
        }
    }
    if (!v.empty() && normalize_l1) {
        double norm = 0.0;
        for (const auto& kv : v) norm += std::fabs(kv.second);
        if (norm > 0.0)
            for (auto& kv : v) kv.second /= norm;
    }
    int k = 0;
    for (const auto& kv : v) { bow_words[k] = kv.first; bow_values[k] = kv.second; k++; }
    int m = 0, at = 0;
    for (const auto& kv : fv) {
        fv_nodes[m] = kv.first; fv_ptr[m] = at;
        for (int32_t i : kv.second) fv_idx[at++] = i;
        m++;
    }
    fv_ptr[m] = at;
    *n_words = k; *n_fv_nodes = m;
    return ORB_OK;
}

}  // extern "C"
