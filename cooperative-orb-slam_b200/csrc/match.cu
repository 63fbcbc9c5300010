// match.cu -- K6: brute-force 2-NN Hamming search over 256-bit descriptors with the reference's
// best/second-best rule (ORBmatcher::DescriptorDistance R21/src/ORBmatcher.cc:1647-1663 inside the
// candidate loops :216-225), and K10's merge of per-shard records.
//
// Record per query: {d1, i1, d2, i2}; (d1,i1) is the lexicographically smallest (distance, index),
// (d2,i2) the second smallest counting duplicates, both start at (256, -1).  This equals the
// reference's scan (strict '<' in ascending index order) and is associative, so shards of the map can
// be searched independently and merged in any grouping.
#include "internal.h"

#include <algorithm>

namespace orbcuda {

constexpr int kKnnThreads = 256;   // one query per thread
constexpr int kKnnTile = 256;      // map descriptors staged in shared memory per step (8 KB)

struct Top2 { int d1, i1, d2, i2; };

__device__ __forceinline__ void top2_push(Top2& t, int d, int i) {
    // (d,i) arrives with i larger than every index already seen by this scan
    if (d < t.d1) { t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = i; }
    else if (d < t.d2) { t.d2 = d; t.i2 = i; }
}

// merge two records whose index ranges may interleave: lexicographic (d, i)
__device__ __host__ __forceinline__ bool lex_less(int da, int ia, int db, int ib) {
    // index -1 marks "nothing": it must lose against any real entry of the same distance
    const unsigned ua = (unsigned)ia, ub = (unsigned)ib;
    return da < db || (da == db && ua < ub);
}
__device__ __host__ __forceinline__ void top2_merge(int& d1, int& i1, int& d2, int& i2, int e1, int j1, int e2, int j2) {
    // candidates: (d1,i1) <= (d2,i2) and (e1,j1) <= (e2,j2)
    if (lex_less(e1, j1, d1, i1)) {
        // new best is e1; second is min(d1, e2)
        if (lex_less(e2, j2, d1, i1)) { d2 = e2; i2 = j2; } else { d2 = d1; i2 = i1; }
        d1 = e1; i1 = j1;
    } else {
        if (lex_less(e1, j1, d2, i2)) { d2 = e1; i2 = j1; }
    }
}

__global__ void __launch_bounds__(kKnnThreads) knn2_popc_kernel(const uint4* __restrict__ q, int nq,
                                                               const uint4* __restrict__ m, long long nm,
                                                               long long per_split, long long index_base,
                                                               int4* __restrict__ partial) {
    __shared__ uint4 tile[kKnnTile * 2];
    const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
    const bool valid = qi < nq;
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (valid) { qa = q[2 * (size_t)qi]; qb = q[2 * (size_t)qi + 1]; }
    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    Top2 t = {256, -1, 256, -1};
    for (long long base = lo; base < hi; base += kKnnTile) {
        const int cnt = (int)min((long long)kKnnTile, hi - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += kKnnThreads) tile[i] = m[2 * base + i];
        __syncthreads();
        const int ib = (int)(index_base + base);
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
            const int d = __popc(a.x ^ qa.x) + __popc(a.y ^ qa.y) + __popc(a.z ^ qa.z) + __popc(a.w ^ qa.w) +
                          __popc(b.x ^ qb.x) + __popc(b.y ^ qb.y) + __popc(b.z ^ qb.z) + __popc(b.w ^ qb.w);
            if (d < t.d2) top2_push(t, d, ib + j);
        }
    }
    if (valid) partial[(size_t)blockIdx.y * nq + qi] = make_int4(t.d1, t.i1, t.d2, t.i2);
}

__global__ void merge_top2_kernel(const int4* __restrict__ parts, int nparts, int nq, int4* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d1 = 256, i1 = -1, d2 = 256, i2 = -1;
    for (int p = 0; p < nparts; p++) {
        const int4 r = parts[(size_t)p * nq + qi];
        top2_merge(d1, i1, d2, i2, r.x, r.y, r.z, r.w);
    }
    out[qi] = make_int4(d1, i1, d2, i2);
}

int launch_knn2(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, int32_t* d_out, int variant,
                cudaStream_t s) {
    (void)variant;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int qblocks = (nq + kKnnThreads - 1) / kKnnThreads;
    // enough map splits for >= 2 CTAs per SM, each at least one tile
    int splits = (2 * sms + qblocks - 1) / qblocks;
    const int64_t max_splits = (nm + kKnnTile - 1) / kKnnTile;
    if (splits > max_splits) splits = (int)std::max<int64_t>(max_splits, 1);
    int64_t per_split = (nm + splits - 1) / splits;
    per_split = (per_split + kKnnTile - 1) / kKnnTile * kKnnTile;
    splits = nm > 0 ? (int)((nm + per_split - 1) / per_split) : 1;
    // per-call scratch from the stream-ordered allocator: matcher entry points are re-entrant
    const size_t need = (size_t)splits * nq * sizeof(int4);
    int4* partial = nullptr;
    if (cudaMallocAsync((void**)&partial, need, s) != cudaSuccess) return -1;
    knn2_popc_kernel<<<dim3(qblocks, splits), kKnnThreads, 0, s>>>((const uint4*)d_q, nq, (const uint4*)d_m, nm, per_split,
                                                                   index_base, partial);
    merge_top2_kernel<<<(nq + 255) / 256, 256, 0, s>>>(partial, splits, nq, (int4*)d_out);
    cudaFreeAsync(partial, s);
    return 2;
}

int launch_merge_top2(const int32_t* d_parts, int parts, int nq, int32_t* d_out, cudaStream_t s) {
    merge_top2_kernel<<<(nq + 255) / 256, 256, 0, s>>>((const int4*)d_parts, parts, nq, (int4*)d_out);
    return 1;
}

void host_merge_top2(const int32_t* parts, int nparts, int nq, int32_t* out) {
    for (int qi = 0; qi < nq; qi++) {
        int d1 = 256, i1 = -1, d2 = 256, i2 = -1;
        for (int p = 0; p < nparts; p++) {
            const int32_t* r = parts + ((size_t)p * nq + qi) * 4;
            top2_merge(d1, i1, d2, i2, r[0], r[1], r[2], r[3]);
        }
        out[4 * qi] = d1; out[4 * qi + 1] = i1; out[4 * qi + 2] = d2; out[4 * qi + 3] = i2;
    }
}

}  // namespace orbcuda
