// octree.cu -- K4: deterministic parallel reformulation of ORBextractor::DistributeOctTree +
// ExtractorNode::DivideNode (R21/src/ORBextractor.cc:539-763, :481-537).  One CTA per (frame, level).
//
// The reference keeps nodes in a std::list, push_front'ing every new child, and (a) sweeps the list
// dividing every node that holds more than one key point until "size + 3*nToExpand > N", then (b)
// repeatedly sorts the expandable nodes by (size, node pointer), divides them from the back and stops
// the moment the list holds >= N nodes; finally the best-response key point of each node is kept, in
// list order.  Here nodes live in an array in *creation order* (surviving roots first, stored in
// reverse so that list order == descending array index), every round is level-synchronous:
//   1. count the key points of each expandable node per quadrant (shared-memory atomics),
//   2. order the nodes to divide (phase a: descending index; phase b: (size desc, index desc) by rank
//      counting) and cut the phase-b order where the running node count reaches N (inclusive scan),
//   3. scatter survivors (stable) and children (in division order, n1..n4) into the other node buffer,
//   4. relabel the key points.
// Pointer ties in the reference's sort are allocator dependent; the canonical rule
// "later created == larger pointer" is the one the CPU oracle is pinned to.
#include "internal.h"

#include <cstdlib>

namespace orbcuda {

// CTA size is a template parameter (<= 512: 16 warp slots in the scans).  512 threads give the shortest CTA (single-frame
// latency); 256 threads make each CTA ~1.3x slower but twice as many fit on an SM, which wins as soon as a launch has
// more CTAs than the GPU holds (batches): measured +4.4 % frames/s at 4 streams x 128 frames.
constexpr int kOctThreadsLatency = 512, kOctThreadsBatch = 256;
constexpr int kOctBatchFrames = 16;     // launches with at least this many frames use the small CTA

struct OctSmem {
    short4* box[2];     // (ulx, uly, brx, bry)
    int* cnt[2];
    int* ccnt;          // [cap][4] children key point counts of nodes that may divide this round
    int* a0;            // scratch: divide flag / scan
    int* a1;            // scratch: survivor position / child base
    int* order;         // phase-b division order
    unsigned long long* key;
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// in-place exclusive scan of a[0..n) by the whole block; returns the total.  `wsum` is 17 ints of smem.
template <int kOctThreads> __device__ int block_excl_scan(int* a, int n, int* wsum) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int running = 0;
    for (int base = 0; base < n; base += kOctThreads) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        const int inc = warp_incl_scan(v, lane);
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = lane < kOctThreads / 32 ? wsum[lane] : 0;
            const int winc = warp_incl_scan(w, lane);
            if (lane < kOctThreads / 32) wsum[lane] = winc - w;
            if (lane == kOctThreads / 32 - 1) wsum[16] = winc;
        }
        __syncthreads();
        if (i < n) a[i] = running + wsum[wid] + inc - v;
        running += wsum[16];
        __syncthreads();
    }
    return running;
}

__device__ __forceinline__ int quadrant_of(uint32_t kp, short4 b) {
    const int x = kp & 0xfff, y = (kp >> 12) & 0xfff;
    const int mx = b.x + ((b.z - b.x + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)   (R21 :483)
    const int my = b.y + ((b.w - b.y + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

// Shared by the batched kernel and the stand-alone entry point.
// kp[0..n): packed candidates in input order; node[0..n): scratch; out_sel[0..kp_cap): selected
// candidates in list order; returns the number of nodes (all threads).
template <int kOctThreads> __device__ int octree_run(const uint32_t* kp, uint16_t* node, int n, int width, int height, int N, int n_ini, float h_x,
                          uint32_t* out_sel, int32_t* out_idx, int cap, unsigned char* smem_raw, int order_cols, int w_cell,
                          int h_cell) {
    __shared__ int wsum[17];
    __shared__ int s_flag[4];
    const int tid = threadIdx.x;
    OctSmem S;
    {
        unsigned char* p = smem_raw;
        S.key = reinterpret_cast<unsigned long long*>(p); p += sizeof(unsigned long long) * cap;
        S.box[0] = reinterpret_cast<short4*>(p); p += sizeof(short4) * cap;
        S.box[1] = reinterpret_cast<short4*>(p); p += sizeof(short4) * cap;
        S.cnt[0] = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
        S.cnt[1] = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
        S.ccnt = reinterpret_cast<int*>(p); p += sizeof(int) * 4 * cap;
        S.a0 = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
        S.a1 = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
        S.order = reinterpret_cast<int*>(p);
    }
    int cur = 0;
    // ---- roots (R21 :543-592).  Root r is stored at index n_ini-1-r.
    for (int i = tid; i < n_ini; i += kOctThreads) {
        const int r = n_ini - 1 - i;
        short4 b;
        b.x = (short)(int)__fmul_rn(h_x, (float)r);
        b.z = (short)(int)__fmul_rn(h_x, (float)(r + 1));
        b.y = 0;
        b.w = (short)height;
        S.box[0][i] = b;
        S.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += kOctThreads) {
        const int x = kp[k] & 0xfff;
        int r = (int)__fdiv_rn((float)x, h_x);        // vpIniNodes[kp.pt.x/hX]  (:569)
        r = min(max(r, 0), n_ini - 1);
        node[k] = (uint16_t)(n_ini - 1 - r);
        atomicAdd(&S.cnt[0][n_ini - 1 - r], 1);
    }
    __syncthreads();
    // drop empty roots (:581-582) -- stable compaction
    {
        for (int i = tid; i < n_ini; i += kOctThreads) S.a0[i] = S.cnt[0][i] > 0 ? 1 : 0;
        __syncthreads();
        const int alive0 = block_excl_scan<kOctThreads>(S.a0, n_ini, wsum);
        if (alive0 != n_ini) {
            for (int i = tid; i < n_ini; i += kOctThreads)
                if (S.cnt[0][i] > 0) { S.box[1][S.a0[i]] = S.box[0][i]; S.cnt[1][S.a0[i]] = S.cnt[0][i]; }
            __syncthreads();
            for (int k = tid; k < n; k += kOctThreads) node[k] = (uint16_t)S.a0[node[k]];
            cur = 1;
        }
        n_ini = alive0;
    }
    int alive = n_ini;
    (void)width;
    bool phase_b = false;
    __syncthreads();

    for (int round = 0; round < 64; round++) {
        short4* box = S.box[cur];
        int* cnt = S.cnt[cur];
        short4* nbox = S.box[cur ^ 1];
        int* ncnt = S.cnt[cur ^ 1];
        const int prev = alive;
        // 1. children counts of every expandable node
        for (int i = tid; i < alive * 4; i += kOctThreads) S.ccnt[i] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += kOctThreads) {
            const int p = node[k];
            if (cnt[p] > 1) atomicAdd(&S.ccnt[4 * p + quadrant_of(kp[k], box[p])], 1);
        }
        __syncthreads();
        // 2. which nodes divide.  a0[i] = divide flag
        if (!phase_b) {
            for (int i = tid; i < alive; i += kOctThreads) S.a0[i] = cnt[i] > 1 ? 1 : 0;
            __syncthreads();
        } else {
            // rank expandable nodes by (size desc, index desc)  == walking the sorted vector from the back (:684-685)
            for (int i = tid; i < alive; i += kOctThreads)
                S.key[i] = cnt[i] > 1 ? (((unsigned long long)cnt[i] << 32) | (unsigned)i) : 0ull;
            __syncthreads();
            int m = 0;
            for (int i = tid; i < alive; i += kOctThreads) {
                const unsigned long long ki = S.key[i];
                if (ki) {
                    int rank = 0;
                    for (int j = 0; j < alive; j++) rank += S.key[j] > ki ? 1 : 0;
                    S.order[rank] = i;
                }
            }
            for (int i = tid; i < alive; i += kOctThreads) m += cnt[i] > 1 ? 1 : 0;
            // m = number of expandable nodes (block reduce)
            m = warp_incl_scan(m, tid & 31);
            if ((tid & 31) == 31) wsum[tid >> 5] = m;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < kOctThreads / 32; w++) t += wsum[w]; s_flag[0] = t; }
            __syncthreads();
            m = s_flag[0];
            // running list size after dividing the j-th node of the order: alive + sum_{i<=j} (nc_i - 1)
            for (int j = tid; j < m; j += kOctThreads) {
                const int p = S.order[j];
                const int nc = (S.ccnt[4 * p] > 0) + (S.ccnt[4 * p + 1] > 0) + (S.ccnt[4 * p + 2] > 0) + (S.ccnt[4 * p + 3] > 0);
                S.a1[j] = nc - 1;
            }
            __syncthreads();
            block_excl_scan<kOctThreads>(S.a1, m, wsum);   // exclusive: a1[j] = growth before node j
            // cut = first j with alive + a1[j] + (nc_j-1) >= N ; divide j <= cut   (:731-732)
            if (tid == 0) s_flag[1] = m;   // default: divide all
            __syncthreads();
            for (int j = tid; j < m; j += kOctThreads) {
                const int p = S.order[j];
                const int nc = (S.ccnt[4 * p] > 0) + (S.ccnt[4 * p + 1] > 0) + (S.ccnt[4 * p + 2] > 0) + (S.ccnt[4 * p + 3] > 0);
                if (alive + S.a1[j] + nc - 1 >= N) atomicMin(&s_flag[1], j + 1);
            }
            __syncthreads();
            const int ndiv = s_flag[1];
            for (int i = tid; i < alive; i += kOctThreads) S.a0[i] = 0;
            __syncthreads();
            for (int j = tid; j < ndiv; j += kOctThreads) S.a0[S.order[j]] = 1;
            __syncthreads();
        }
        // 3. positions.  survivors keep their relative order at the front, children follow in division order.
        //    a1[i] <- survivor position (for non-divided) ; key[i] <- child base (for divided)
        for (int i = tid; i < alive; i += kOctThreads) S.a1[i] = S.a0[i] ? 0 : 1;
        __syncthreads();
        const int n_surv = block_excl_scan<kOctThreads>(S.a1, alive, wsum);
        int n_children;
        int* cbase = reinterpret_cast<int*>(S.key);   // reuse (phase-b keys are dead by now)
        if (!phase_b) {
            // division order = descending index: child base of node i = #children of divided nodes with index > i
            int* tmp = cbase;
            for (int i = tid; i < alive; i += kOctThreads) {
                const int p = alive - 1 - i;   // reversed
                tmp[i] = S.a0[p] ? (S.ccnt[4 * p] > 0) + (S.ccnt[4 * p + 1] > 0) + (S.ccnt[4 * p + 2] > 0) + (S.ccnt[4 * p + 3] > 0) : 0;
            }
            __syncthreads();
            n_children = block_excl_scan<kOctThreads>(tmp, alive, wsum);   // tmp[i] = base of node alive-1-i
        } else {
            int* tmp = cbase;
            const int ndiv = s_flag[1];
            for (int j = tid; j < ndiv; j += kOctThreads) {
                const int p = S.order[j];
                tmp[j] = (S.ccnt[4 * p] > 0) + (S.ccnt[4 * p + 1] > 0) + (S.ccnt[4 * p + 2] > 0) + (S.ccnt[4 * p + 3] > 0);
            }
            __syncthreads();
            n_children = block_excl_scan<kOctThreads>(tmp, ndiv, wsum);    // tmp[j] = base of node order[j]
            // scatter to node-indexed form through a0 (flag) -> store base+1 in a0 for divided nodes
            __syncthreads();
            for (int j = tid; j < ndiv; j += kOctThreads) S.a0[S.order[j]] = tmp[j] + 1;
            __syncthreads();
        }
        // 4. build the next node array
        for (int i = tid; i < alive; i += kOctThreads) {
            if (!S.a0[i]) {
                nbox[S.a1[i]] = box[i];
                ncnt[S.a1[i]] = cnt[i];
            } else {
                const int base = n_surv + (phase_b ? S.a0[i] - 1 : cbase[alive - 1 - i]);
                const short4 b = box[i];
                const short mx = (short)(b.x + ((b.z - b.x + 1) >> 1));
                const short my = (short)(b.y + ((b.w - b.y + 1) >> 1));
                int r = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = S.ccnt[4 * i + q];
                    if (c > 0) {
                        short4 cb;
                        cb.x = (q & 1) ? mx : b.x;  cb.z = (q & 1) ? b.z : mx;
                        cb.y = (q & 2) ? my : b.y;  cb.w = (q & 2) ? b.w : my;
                        nbox[base + r] = cb;
                        ncnt[base + r] = c;
                        r++;
                    }
                }
            }
        }
        // 5. relabel key points (reads the old arrays, which stay intact until the swap)
        for (int k = tid; k < n; k += kOctThreads) {
            const int p = node[k];
            int nid;
            if (!S.a0[p]) {
                nid = S.a1[p];
            } else {
                const int base = n_surv + (phase_b ? S.a0[p] - 1 : cbase[alive - 1 - p]);
                const int q = quadrant_of(kp[k], box[p]);
                int r = 0;
#pragma unroll
                for (int qq = 0; qq < 3; qq++)
                    if (qq < q && S.ccnt[4 * p + qq] > 0) r++;
                nid = base + r;
            }
            node[k] = (uint16_t)nid;
        }
        __syncthreads();
        cur ^= 1;
        alive = n_surv + n_children;
        // 6. termination (R21 :667-738)
        if (alive >= N || alive == prev) break;
        if (!phase_b) {
            // nToExpand = nodes created this sweep holding more than one key point
            int c = 0;
            const int* cn = S.cnt[cur];
            for (int i = tid; i < alive; i += kOctThreads) c += cn[i] > 1 ? 1 : 0;
            c = warp_incl_scan(c, tid & 31);
            if ((tid & 31) == 31) wsum[tid >> 5] = c;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < kOctThreads / 32; w++) t += wsum[w]; s_flag[2] = t; }
            __syncthreads();
            if (alive + 3 * s_flag[2] > N) phase_b = true;
        }
        __syncthreads();
    }
    // ---- keep the best key point of each node (R21 :741-760): max response, first in input order on ties.
    // "Input order" is cell-major, raster inside a cell (:789-826).  With order_cols > 0 that rank is computed
    // from the coordinates (the list itself may be in any order); with order_cols == 0 it is the list index.
    unsigned long long* best = S.key;
    for (int i = tid; i < alive; i += kOctThreads) best[i] = 0ull;
    __syncthreads();
    auto order_of = [&](int k, uint32_t v) -> unsigned {
        if (order_cols <= 0) return (unsigned)k;
        const int xr = (int)(v & 0xfff) - 3, yr = (int)((v >> 12) & 0xfff) - 3;   // relative to the first cell interior (19,19)
        const int cx = xr / w_cell, cyy = yr / h_cell;
        return ((unsigned)(cyy * order_cols + cx) << 12) | (unsigned)((yr - cyy * h_cell) * w_cell + (xr - cx * w_cell));
    };
    for (int k = tid; k < n; k += kOctThreads) {
        const uint32_t v = kp[k];
        atomicMax(&best[node[k]], ((unsigned long long)(v >> 24) << 32) | (0xffffffffu - order_of(k, v)));
    }
    __syncthreads();
    for (int k = tid; k < n; k += kOctThreads) {
        const uint32_t v = kp[k];
        const int nd = node[k];
        if (best[nd] == (((unsigned long long)(v >> 24) << 32) | (0xffffffffu - order_of(k, v)))) {
            out_sel[alive - 1 - nd] = v;   // list order == descending creation index
            if (out_idx) out_idx[alive - 1 - nd] = (int32_t)k;
        }
    }
    return alive;
}

size_t octree_smem_bytes(int cap) {
    return (size_t)cap * (sizeof(unsigned long long) + 2 * sizeof(short4) + 2 * sizeof(int) + 4 * sizeof(int) +
                          3 * sizeof(int));
}

template <int kOctThreads> __global__ void __launch_bounds__(kOctThreads) octree_kernel(DevPtrs d, FrameLayout fl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_total;
    // grid = (frames, levels): CTAs are dispatched level 0 first for every frame, i.e. longest first (a level-0 CTA runs
    // about 4x longer than a level-7 one), so the tail of the launch is made of the short ones
    const int level = blockIdx.y, frame = blockIdx.x, tid = threadIdx.x;
    const LevelGeom g = d.geom[level];
    const int32_t* cc = d.cell_count + (size_t)frame * fl.n_cells + g.cell_base;
    const uint32_t* cand = d.cand + (size_t)frame * fl.cand_entries + g.cand_off;
    uint32_t* kp = d.oct_scratch + (size_t)frame * fl.cand_entries + g.cand_off;
    uint16_t* node = d.oct_node + (size_t)frame * fl.cand_entries + g.cand_off;
    // Filter the level's NMS survivors into the quadtree's input list (vToDistributeKeys, R21 :789-826): a
    // cell whose flag is up holds a survivor reaching iniThFAST and keeps only those (== cv::FAST(cell,
    // iniThFAST)); otherwise it keeps all (== the minThFAST retry, :809-816).  The list order is arbitrary:
    // nothing in the quadtree depends on it except the "first maximum wins" rule, which octree_run resolves
    // with an explicit (cell, raster) order key instead of the list position.
    int running = 0;
    {
        const int nraw = d.level_raw[(size_t)frame * kMaxLevels + level];
        const int lane = tid & 31, wid = tid >> 5;
        int* wcnt = reinterpret_cast<int*>(smem_raw);   // [16] warp counts, [16..31] exclusive bases
        for (int base = 0; base < nraw; base += kOctThreads) {
            const int i = base + tid;
            uint32_t v = 0;
            bool pass = false;
            if (i < nraw) {
                v = cand[i];
                const int xr = (int)(v & 0xfff) - 3, yr = (int)((v >> 12) & 0xfff) - 3;
                const int cell = (yr / g.h_cell) * g.n_cols + xr / g.w_cell;
                pass = cc[cell] == 0 || (int)(v >> 24) >= fl.ini_th;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) wcnt[wid] = __popc(m);
            __syncthreads();
            if (wid == 0) {
                const int c = lane < kOctThreads / 32 ? wcnt[lane] : 0;
                const int inc = warp_incl_scan(c, lane);
                if (lane < kOctThreads / 32) wcnt[16 + lane] = inc - c;
                if (lane == kOctThreads / 32 - 1) wcnt[32] = inc;
            }
            __syncthreads();
            if (pass) kp[running + wcnt[16 + wid] + __popc(m & ((1u << lane) - 1))] = v;
            running += wcnt[32];
            __syncthreads();
        }
    }
    if (tid == 0) s_total = running;
    __syncthreads();
    const int n = s_total;
    int count = 0;
    if (n > 0) {
        __threadfence_block();
        count = octree_run<kOctThreads>(kp, node, n, g.w - 2 * kMinBorder, g.h - 2 * kMinBorder, g.n_feat, g.n_ini, g.h_x,
                           d.sel + (size_t)frame * fl.kp_cap + g.kp_slot, nullptr, fl.node_cap, smem_raw, g.n_cols, g.w_cell,
                           g.h_cell);
    }
    if (tid == 0) d.level_count[(size_t)frame * kMaxLevels + level] = count;
}

template <int kOctThreads> static int launch_octree_t(const DevPtrs& d, const FrameLayout& fl, int n_frames, cudaStream_t s) {
    const size_t smem = octree_smem_bytes(fl.node_cap) > (size_t)kOctThreads * 4 ? octree_smem_bytes(fl.node_cap)
                                                                                 : (size_t)kOctThreads * 4;
    // the opt-in is per device and per instantiation (and grows with the largest node capacity seen on it)
    static thread_local size_t configured[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (smem > 48 * 1024 && smem > configured[dev]) {
        if (cudaFuncSetAttribute(octree_kernel<kOctThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        configured[dev] = smem;
    }
    octree_kernel<kOctThreads><<<dim3(n_frames, fl.nlevels), kOctThreads, smem, s>>>(d, fl);
    return 1;
}

int launch_octree(const DevPtrs& d, const FrameLayout& fl, int n_frames, cudaStream_t s) {
    // ORBCUDA_OCT_THREADS=128|256|512 forces one CTA size (A/B and soak switch; results are identical)
    static const int forced = [] { const char* e = getenv("ORBCUDA_OCT_THREADS"); return e ? atoi(e) : 0; }();
    const int threads = forced ? forced : (n_frames >= kOctBatchFrames ? kOctThreadsBatch : kOctThreadsLatency);
    switch (threads) {
        case 128: return launch_octree_t<128>(d, fl, n_frames, s);
        case 256: return launch_octree_t<256>(d, fl, n_frames, s);
        case 512: return launch_octree_t<512>(d, fl, n_frames, s);
        default: return -1;
    }
}

template <int kOctThreads> __global__ void __launch_bounds__(kOctThreads) octree_single_kernel(const uint32_t* kp, int n, int width, int height,
                                                                    int n_feat, int n_ini, float h_x, uint16_t* node,
                                                                    uint32_t* sel, int32_t* sel_idx, int32_t* count,
                                                                    int node_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = n > 0 ? octree_run<kOctThreads>(kp, node, n, width, height, n_feat, n_ini, h_x, sel, sel_idx, node_cap, smem_raw, 0, 1, 1) : 0;
    if (threadIdx.x == 0) *count = c;
}

int launch_octree_single(const uint32_t* d_cand, int n, int width, int height, int n_feat, int n_ini, float h_x,
                         int32_t* d_sel_idx, uint16_t* d_node, uint32_t* d_sel, int32_t* d_count, int kp_cap,
                         int node_cap, cudaStream_t s) {
    (void)kp_cap;
    const size_t smem = octree_smem_bytes(node_cap);
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(octree_single_kernel<kOctThreadsLatency>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
    }
    octree_single_kernel<kOctThreadsLatency><<<1, kOctThreadsLatency, smem, s>>>(d_cand, n, width, height, n_feat, n_ini, h_x, d_node, d_sel,
                                                      d_sel_idx, d_count, node_cap);
    return 1;
}

}  // namespace orbcuda
