// match.cu -- K6: brute-force 2-NN Hamming search over 256-bit descriptors with the reference's
// best/second-best rule (ORBmatcher::DescriptorDistance R21/src/ORBmatcher.cc:1647-1663 inside the
// candidate loops :216-225), and K10's merge of per-shard records.
//
// Record per query: {d1, i1, d2, i2}; (d1,i1) is the lexicographically smallest (distance, index),
// (d2,i2) the second smallest counting duplicates, both start at (256, -1).  This equals the
// reference's scan (strict '<' in ascending index order) and is associative, so shards of the map can
// be searched independently and merged in any grouping.
#include "internal.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

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

// ---------------------------------------------------------------------------------------------
// K7: tensor-core variant.  d(a,b) = popc(a) + popc(b) - 2*popc(a & b); popc(a & b) is the dot product of
// the two descriptors seen as 256-element {0,1} vectors, i.e. a 16x8x256 AND-popc contraction per warp
// tile.  PTX's own `mma.sync ... b1 ... and.popc` is software-emulated on sm_100a (each m16n8k256 expands
// the bits and issues 8 IMMA.16832 plus ~100 ALU ops), so this kernel issues the integer MMAs
// directly (mma.sync.m16n8k32.s32.u8.u8.s32) and hoists the bit -> byte expansion out of the inner loop:
// query fragments are expanded once into registers (64 regs: 32 queries x 256 elements per warp), map
// descriptors once per CTA into a shared-memory tile laid out in fragment order (conflict-free LDS.64).
// The k dimension is permuted freely (element e of k-step s = bit e of descriptor word s) because the
// same permutation is applied to both operands.  Top-2 bookkeeping as in the POPC kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 256;            // 8 warps x 32 queries
constexpr int kMmaQPerCta = 256;
constexpr int kMmaTile = 128;               // map descriptors per shared-memory tile (32 KB expanded)

__device__ __forceinline__ uint32_t expand_nibble(uint32_t w, int shift) {
    // 4 bits -> 4 bytes of 0/1
    return (((w >> shift) & 0xfu) * 0x00204081u) & 0x01010101u;
}

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kMmaThreads, 2) knn2_mma_kernel(const uint32_t* __restrict__ q, int nq,
                                                                 const uint4* __restrict__ m, long long nm,
                                                                 long long per_split, long long index_base,
                                                                 int4* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2* s_b = reinterpret_cast<uint2*>(smem_raw);                            // [tile/8][8 k-steps][32 lanes]
    int* s_pb = reinterpret_cast<int*>(smem_raw + (size_t)kMmaTile * 256);        // popc of each staged descriptor
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int q0 = blockIdx.x * kMmaQPerCta + warp * 32;
    // ---- A fragments: rows gid, gid+8 of two 16-query tiles; k-step s = descriptor word s
    uint32_t A[2][8][4];
    int pa[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int r0 = q0 + t * 16 + gid, r1 = r0 + 8;
        int p0 = 0, p1 = 0;
#pragma unroll
        for (int sx = 0; sx < 8; sx++) {
            const uint32_t w0 = r0 < nq ? q[(size_t)r0 * 8 + sx] : 0u;
            const uint32_t w1 = r1 < nq ? q[(size_t)r1 * 8 + sx] : 0u;
            p0 += __popc(w0); p1 += __popc(w1);
            A[t][sx][0] = expand_nibble(w0, 4 * tig);          // row gid,   k = 4*tig .. +3
            A[t][sx][1] = expand_nibble(w1, 4 * tig);          // row gid+8
            A[t][sx][2] = expand_nibble(w0, 16 + 4 * tig);     // row gid,   k = 16 + 4*tig .. +3
            A[t][sx][3] = expand_nibble(w1, 16 + 4 * tig);
        }
        pa[t][0] = p0; pa[t][1] = p1;
    }
    Top2 best[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int h = 0; h < 2; h++) best[t][h] = Top2{256, -1, 256, -1};

    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    for (long long base = lo; base < hi; base += kMmaTile) {
        const int cnt = (int)min((long long)kMmaTile, hi - base);
        __syncthreads();
        // ---- stage + expand the map tile: thread handles (descriptor j, k-step s) pairs
        for (int it = tid; it < kMmaTile * 8; it += kMmaThreads) {
            const int j = it >> 3, sx = it & 7;
            const uint32_t w = j < cnt ? reinterpret_cast<const uint32_t*>(m + 2 * (base + j))[sx] : 0u;
            // B fragment of column n = j%8 at k-step sx: lane (gid=n, tig) holds k = 4*tig..+3 and 16+4*tig..+3
            uint2* dstp = s_b + ((size_t)(j >> 3) * 8 + sx) * 32 + (j & 7) * 4;
#pragma unroll
            for (int tg = 0; tg < 4; tg++) dstp[tg] = make_uint2(expand_nibble(w, 4 * tg), expand_nibble(w, 16 + 4 * tg));
        }
        for (int j = tid; j < kMmaTile; j += kMmaThreads) {
            int p = 0;
            if (j < cnt) {
                const uint4 x = m[2 * (base + j)], y = m[2 * (base + j) + 1];
                p = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
            }
            s_pb[j] = p;
        }
        __syncthreads();
        const int ngroups = (cnt + 7) >> 3;
        for (int g8 = 0; g8 < ngroups; g8++) {
            int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
            const uint2* bp = s_b + (size_t)g8 * 8 * 32 + lane;
#pragma unroll
            for (int sx = 0; sx < 8; sx++) {
                const uint2 b = bp[sx * 32];
                mma_u8(c0, A[0][sx], b.x, b.y);
                mma_u8(c1, A[1][sx], b.x, b.y);
            }
            const int col = g8 * 8 + 2 * tig;
            const int pb0 = s_pb[col], pb1 = s_pb[col + 1];
            const int idx0 = (int)(index_base + base) + col;
            const bool v0 = col < cnt, v1 = col + 1 < cnt;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int* c = t ? c1 : c0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int d0 = pa[t][h] + pb0 - 2 * c[2 * h], d1 = pa[t][h] + pb1 - 2 * c[2 * h + 1];
                    if (v0 && d0 < best[t][h].d2) top2_push(best[t][h], d0, idx0);
                    if (v1 && d1 < best[t][h].d2) top2_push(best[t][h], d1, idx0 + 1);
                }
            }
        }
    }
    // ---- merge the four lanes of a row group (they saw disjoint column subsets), lane tig==0 writes
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            Top2 b = best[t][h];
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const int e1 = __shfl_xor_sync(0xffffffffu, b.d1, o), j1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
                const int e2 = __shfl_xor_sync(0xffffffffu, b.d2, o), j2 = __shfl_xor_sync(0xffffffffu, b.i2, o);
                top2_merge(b.d1, b.i1, b.d2, b.i2, e1, j1, e2, j2);
            }
            const int row = q0 + t * 16 + gid + 8 * h;
            if (tig == 0 && row < nq) partial[(size_t)blockIdx.y * nq + row] = make_int4(b.d1, b.i1, b.d2, b.i2);
        }
}

// K7b: same contraction without shared memory.  Every warp streams the map itself: the four lanes of a
// fragment column load that column's packed descriptor (one 32-byte read, coalesced across the group) and
// expand their own nibbles in registers, so there is no staging pass and no block barrier; the next group's
// descriptor is prefetched while the current one is multiplied.
__global__ void __launch_bounds__(kMmaThreads, 2) knn2_mma_stream_kernel(const uint32_t* __restrict__ q, int nq,
                                                                        const uint4* __restrict__ m, long long nm,
                                                                        long long per_split, long long index_base,
                                                                        int4* __restrict__ partial) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int q0 = blockIdx.x * kMmaQPerCta + warp * 32;
    uint32_t A[2][8][4];
    int pa[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int r0 = q0 + t * 16 + gid, r1 = r0 + 8;
        int p0 = 0, p1 = 0;
#pragma unroll
        for (int sx = 0; sx < 8; sx++) {
            const uint32_t w0 = r0 < nq ? q[(size_t)r0 * 8 + sx] : 0u;
            const uint32_t w1 = r1 < nq ? q[(size_t)r1 * 8 + sx] : 0u;
            p0 += __popc(w0); p1 += __popc(w1);
            A[t][sx][0] = expand_nibble(w0, 4 * tig);
            A[t][sx][1] = expand_nibble(w1, 4 * tig);
            A[t][sx][2] = expand_nibble(w0, 16 + 4 * tig);
            A[t][sx][3] = expand_nibble(w1, 16 + 4 * tig);
        }
        pa[t][0] = p0; pa[t][1] = p1;
    }
    Top2 best[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int h = 0; h < 2; h++) best[t][h] = Top2{256, -1, 256, -1};

    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    const long long last = hi - 1;
    auto fetch = [&](long long j, uint4& x, uint4& y) {
        const long long jj = min(j, last);     // clamp: columns past the end are masked in the epilogue
        x = m[2 * jj]; y = m[2 * jj + 1];
    };
    if (lo < hi) {
        uint4 cx, cy, nx, ny;
        fetch(lo + gid, cx, cy);
        for (long long base = lo; base < hi; base += 8) {
            fetch(base + 8 + gid, nx, ny);                  // prefetch the next group's column
            const uint32_t w[8] = {cx.x, cx.y, cx.z, cx.w, cy.x, cy.y, cy.z, cy.w};
            int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
            int pb = 0;
#pragma unroll
            for (int sx = 0; sx < 8; sx++) {
                const uint32_t b0 = expand_nibble(w[sx], 4 * tig), b1 = expand_nibble(w[sx], 16 + 4 * tig);
                pb += __popc(w[sx]);
                mma_u8(c0, A[0][sx], b0, b1);
                mma_u8(c1, A[1][sx], b0, b1);
            }
            // popcounts of columns 2*tig and 2*tig+1 live in the lanes of groups gid' = 2*tig, 2*tig+1
            const int pb0 = __shfl_sync(0xffffffffu, pb, (2 * tig) * 4), pb1 = __shfl_sync(0xffffffffu, pb, (2 * tig + 1) * 4);
            const long long col = base + 2 * tig;
            const int idx0 = (int)(index_base + col);
            const bool v0 = col < hi, v1 = col + 1 < hi;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int* c = t ? c1 : c0;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int d0 = pa[t][h] + pb0 - 2 * c[2 * h], d1 = pa[t][h] + pb1 - 2 * c[2 * h + 1];
                    if (v0 && d0 < best[t][h].d2) top2_push(best[t][h], d0, idx0);
                    if (v1 && d1 < best[t][h].d2) top2_push(best[t][h], d1, idx0 + 1);
                }
            }
            cx = nx; cy = ny;
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            Top2 b = best[t][h];
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const int e1 = __shfl_xor_sync(0xffffffffu, b.d1, o), j1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
                const int e2 = __shfl_xor_sync(0xffffffffu, b.d2, o), j2 = __shfl_xor_sync(0xffffffffu, b.i2, o);
                top2_merge(b.d1, b.i1, b.d2, b.i2, e1, j1, e2, j2);
            }
            const int row = q0 + t * 16 + gid + 8 * h;
            if (tig == 0 && row < nq) partial[(size_t)blockIdx.y * nq + row] = make_int4(b.d1, b.i1, b.d2, b.i2);
        }
}

// ---------------------------------------------------------------------------------------------
// K7c: the same contraction on the 5th-generation tensor cores: tcgen05.mma kind::i8 (M=128 queries x
// N=128 map descriptors x K=32 per instruction) with the int32 accumulators in tensor memory.  Operands are
// the expanded descriptors, written by the CTA itself into shared memory in the canonical K-major
// no-swizzle UMMA layout (8-row x 16-byte core matrices; leading byte offset 128 between the K chunks,
// stride byte offset 2048 between 8-row groups), so no TMA is needed.
//   * The query operand is signed (+1 for a set bit, -1 for a clear one), the map operand unsigned {0,1}: the
//     accumulator already is g = 2*popc(a&b) - popc(b) and d = popc(a) - g.  A column can only enter a
//     query's top-2 if g > popc(a) - d2, a per-thread scalar, so draining 32 columns costs one tcgen05.ld,
//     one 3-input max tree and -- rarely -- a detailed scan of the 8-column groups that beat the bound.
//   * The contraction index may be permuted freely as long as both operands agree, so bit b of a 32-bit
//     word goes to byte (b%8)*4 + b/8 of its 32-byte run: four bytes per shift+mask.
//   * One CTA = 256 queries (two 128-row A tiles) x one map split, warp specialised: 16 worker warps and
//     one MMA warp coupled only through mbarriers (no CTA-wide barrier in the loop).  A stage = one
//     expanded 128-descriptor B tile in shared memory + two 128-column accumulators in TMEM; two stages.
//     Workers: expand tile i into stage i%2 (raw words were loaded one tile earlier) -> arrive full[i%2];
//     prefetch the raw words of tile i+1; wait done[(i-1)%2]; drain tile i-1 -> arrive empty[(i-1)%2].
//     MMA warp: wait full[st] and empty[st]; one elected lane issues the 16 MMAs and commits to done[st].
//     The tensor pipe runs tile i while the workers drain tile i-1 and expand tile i+1.
//   * A warp may only touch TMEM lanes 32*(warp%4)..+31: worker w owns query rows 128*((w/4)%2) +
//     32*(w%4) + lane and the column half w/8 of every tile; the two halves are merged at the end.
// ---------------------------------------------------------------------------------------------
constexpr int kTcWorkers = 16;                       // worker warps
constexpr int kTcThreads = (kTcWorkers + 1) * 32;    // + the MMA warp
constexpr int kTcM = 128;     // rows of one A tile / one accumulator
constexpr int kTcQ = 256;     // queries per CTA (two A tiles)
constexpr int kTcN = 128;     // map descriptors per tile (2 stages x 2 A tiles x 128 = 512 TMEM columns)
constexpr int kTcBStages = 4; // expanded B tiles in shared memory (the workers run ahead of the tensor pipe)
constexpr int kTcLag = 3;     // a worker drains tile i - kTcLag after expanding tile i (< kTcBStages: see the worker loop)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    // K-major, SWIZZLE_NONE: start address, LBO = 128 B, SBO = 2048 B (all >> 4), version 1 (Blackwell)
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
}

// Operand encoding.  Bit b = 8*k + j of a 32-bit descriptor word goes to byte 4*j + k of the word's 32-byte run
// (plane j, byte k).  The map operand keeps the bit where it is: plane j < 7 holds x & (0x01010101 << j), i.e.
// bytes of 0 or 2^j (one LOP3 each); plane 7 holds (x >> 1) & 0x40404040 (0 or 64).  The query operand carries
// the compensating weight with the sign of its bit: +-2^(6-j) for j < 7, +-1 for plane 7.  Every product is
// +-64 or 0, so the accumulator is 64 * (2*popc(a&b) - popc(b)) = 64 * g, |.| <= 16384: it fits 16 bits.
constexpr int kTcScaleLog2 = 6;
__device__ __forceinline__ void expand32_map(uint32_t x, uint4& c0, uint4& c1) {
    c0 = make_uint4(x & 0x01010101u, x & 0x02020202u, x & 0x04040404u, x & 0x08080808u);
    c1 = make_uint4(x & 0x10101010u, x & 0x20202020u, x & 0x40404040u, (x >> 1) & 0x40404040u);
}
__device__ __forceinline__ uint32_t query_plane(uint32_t x, int j) {
    const uint32_t e = (x >> j) & 0x01010101u;           // bytes of 0/1
    const uint32_t w = j < 7 ? (1u << (6 - j)) : 1u;     // weight; -w as a byte = 256 - w
    return e * w | (e ^ 0x01010101u) * (256u - w);
}
__device__ __forceinline__ void expand32_query(uint32_t x, uint4& c0, uint4& c1) {
    c0 = make_uint4(query_plane(x, 0), query_plane(x, 1), query_plane(x, 2), query_plane(x, 3));
    c1 = make_uint4(query_plane(x, 4), query_plane(x, 5), query_plane(x, 6), query_plane(x, 7));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// same with a back-off between polls: the MMA warp shares its scheduler with four worker warps and must not eat
// their issue slots while it waits
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// drain 64 accumulator columns of one query row, loaded as packed int16 pairs: v[r] = columns 2r (low half) and
// 2r + 1 (high half), each 64 * g.  Columns >= valid come from zero rows and are skipped.
template <bool FULL>
__device__ __forceinline__ void tc_drain64(const uint32_t (&v)[32], int pa, int idx0, int valid, int glim, int* gptr, Top2& best) {
    uint32_t s[8];
#pragma unroll
    for (int g = 0; g < 8; g++) s[g] = __vmaxs2(__vimax3_s16x2(v[4 * g], v[4 * g + 1], v[4 * g + 2]), v[4 * g + 3]);
    uint32_t mx = __vimax3_s16x2(s[0], s[1], s[2]);
    mx = __vimax3_s16x2(mx, s[3], s[4]);
    mx = __vimax3_s16x2(mx, s[5], s[6]);
    mx = __vmaxs2(mx, s[7]);
    // a column enters this CTA's top-2 iff g > pa - d2, and it can only matter for the merged result if its distance
    // does not exceed the smallest second-best any CTA has published for this query (g > glim = pa - G - 1: two real
    // columns are at distance <= G, so nothing farther than G survives the merge; ties are kept).  In the packed
    // domain the test is max(x, bound) != bound for some half.
    const int d2_in = best.d2;
    int lim = max(pa - best.d2, glim);
    int bound = lim << kTcScaleLog2;
    uint32_t bound2 = __byte_perm((uint32_t)bound, 0u, 0x1010);
    if (__vmaxs2(mx, bound2) != bound2) {
#pragma unroll
        for (int g = 0; g < 8; g++) {
            if (__vmaxs2(s[g], bound2) != bound2) {
                // visit the 8 columns of this group by (g descending, column ascending) through packed keys
                // key = 8 * g_value + (7 - column): a later column never displaces an equal earlier one, so
                // the visit order keeps the first-index rule
                int key[8];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t x = v[4 * g + r];
                    const int lo = ((int)(x << 16) >> (16 + kTcScaleLog2 - 3)) | (7 - 2 * r);
                    const int hi = ((int)x >> (16 + kTcScaleLog2 - 3)) | (6 - 2 * r);
                    key[2 * r] = FULL || 8 * g + 2 * r < valid ? lo : (int)0x80000000;
                    key[2 * r + 1] = FULL || 8 * g + 2 * r + 1 < valid ? hi : (int)0x80000000;
                }
                // the largest key certainly beats the bound when FULL (that is why we are here)
                int k = max(__vimax3_s32(key[0], key[1], key[2]), __vimax3_s32(key[3], key[4], key[5]));
                k = __vimax3_s32(k, key[6], key[7]);
                while ((k >> 3) > lim) {
                    top2_push(best, pa - (k >> 3), idx0 + 8 * g + 7 - (k & 7));
                    lim = max(pa - best.d2, glim);
                    const int last = k;
                    k = (int)0x80000000;
#pragma unroll
                    for (int j = 0; j < 8; j++) k = max(k, key[j] < last ? key[j] : (int)0x80000000);
                }
                bound = lim << kTcScaleLog2;
                bound2 = __byte_perm((uint32_t)bound, 0u, 0x1010);
            }
        }
        if (best.d2 < d2_in) atomicMin(gptr, best.d2);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) knn2_tc_kernel(const uint32_t* __restrict__ q, int nq,
                                                               const uint32_t* __restrict__ m, long long nm,
                                                               long long per_split, long long index_base,
                                                               int4* __restrict__ partial, int* __restrict__ shared_d2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* s_a = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 2 x [128 rows][256 B] (64 KB)
    unsigned char* s_b = s_a + kTcQ * 256;                                               // 2 stages x [kTcN rows][256 B]
    __shared__ __align__(8) unsigned long long s_full[kTcBStages], s_done[2], s_empty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kTcQ;
    constexpr int kWorkThreads = kTcWorkers * 32;

    if (warp == kTcWorkers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
#pragma unroll
        for (int sb = 0; sb < kTcBStages; sb++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_full[sb])), "r"(kTcWorkers));
#pragma unroll
        for (int st = 0; st < 2; st++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_done[st])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_empty[st])), "r"(kTcWorkers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    // ---- A tiles (+1 / -1): element (r, c) of a tile at (r/8)*2048 + (c/16)*128 + (r%8)*16 + c%16
    for (int it = tid; it < kTcQ * 8; it += kTcThreads) {
        const int row = it >> 3, w = it & 7, r = row & 127;
        const uint32_t bits = q0 + row < nq ? q[(size_t)(q0 + row) * 8 + w] : 0u;
        uint4 c0, c1;
        expand32_query(bits, c0, c1);
        unsigned char* dst = s_a + (row >> 7) * (kTcM * 256) + (r >> 3) * 2048 + (r & 7) * 16 + (2 * w) * 128;
        *reinterpret_cast<uint4*>(dst) = c0;
        *reinterpret_cast<uint4*>(dst + 128) = c1;
    }
    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    const int ntiles = hi > lo ? (int)((hi - lo + kTcN - 1) / kTcN) : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::);
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t tmem = s_tmem;
    // shared-window addresses pinned in registers (otherwise they are re-derived from the CTA id at every use)
    uint32_t full0 = smem_u32(&s_full[0]), done0 = smem_u32(&s_done[0]), empty0 = smem_u32(&s_empty[0]), b0 = smem_u32(s_b);
    asm volatile("" : "+r"(full0), "+r"(done0), "+r"(empty0), "+r"(b0));

    Top2 best = {256, -1, 256, -1};
    if (warp == kTcWorkers) {
        // =========================== MMA warp ===========================
        // instruction descriptor: D = s32, A = signed 8 bit, B = unsigned 8 bit, both K-major, N = kTcN, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        const uint64_t a_desc = umma_desc(smem_u32(s_a)), b_desc = umma_desc(b0);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        for (int i = 0; i < ntiles; i++) {
            const int st = i & 1, sb = i & (kTcBStages - 1);
            mbar_wait_parked(full0 + 8 * sb, (uint32_t)((i / kTcBStages) & 1), 32);   // B tile i expanded by all workers
            mbar_wait_parked(empty0 + 8 * st, (uint32_t)(((i >> 1) & 1) ^ 1), 32);    // accumulators of tile i-2 drained (passes at once for i < 2)
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            if (leader) {
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        // descriptors address in 16-byte units: A tile t at +t*32 KB, K step at +256 B, stage at +kTcN*256 B
                        const uint64_t da = a_desc + (uint64_t)((t * (kTcM * 256) + ks * 256) >> 4);
                        const uint64_t db = b_desc + (uint64_t)((sb * (kTcN * 256) + ks * 256) >> 4);
                        const uint32_t accumulate = ks ? 1u : 0u;
                        asm volatile(
                            "{\n\t"
                            ".reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
                            "}\n"
                            :: "r"(tmem + (uint32_t)((2 * st + t) * kTcN)), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done0 + 8 * st));
            }
            __syncwarp();
        }
    } else {
        // =========================== worker warps ===========================
        // my query row (TMEM lane) and column half
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        const int chalf = warp >> 3;
        constexpr int kColsPerWarp = kTcN / (kTcWorkers / 8);
        int pa = 0;
        if (q0 + row < nq) {
            const uint4 x = *reinterpret_cast<const uint4*>(q + (size_t)(q0 + row) * 8);
            const uint4 y = *reinterpret_cast<const uint4*>(q + (size_t)(q0 + row) * 8 + 4);
            pa = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
        }
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * kTcN + chalf * kColsPerWarp);
        // raw words of a tile: thread t takes word w of rows r0 + 64*i (consecutive threads: the 8 rows of a group, then w)
        constexpr int kWordsPerThread = kTcN * 8 / kWorkThreads;   // 2
        constexpr int kRowStep = kWorkThreads / 8;                  // 64
        const int r0 = (tid & 7) | ((tid >> 6) << 3), w0 = (tid >> 3) & 7;
        const uint32_t* src = m + ((size_t)lo + r0) * 8 + w0;
        const uint32_t b_off = (uint32_t)((r0 >> 3) * 2048 + (r0 & 7) * 16 + (2 * w0) * 128);
        uint32_t raw_a[kWordsPerThread], raw_b[kWordsPerThread];     // tiles i and i + 1: global latency > one iteration
        const int ib0 = (int)(index_base + lo) + chalf * kColsPerWarp;
        const int n_full = (int)((hi - lo) / kTcN);                 // tiles whose kTcN rows all exist
        const int last_cnt = (int)(hi - lo) - n_full * kTcN;        // rows of the ragged last tile
        auto fetch = [&](int tile, uint32_t (&raw)[kWordsPerThread]) {      // called with tile = 0, 1, 2, ...
            if (tile < n_full) {
#pragma unroll
                for (int i = 0; i < kWordsPerThread; i++) raw[i] = __ldg(src + i * (kRowStep * 8));
            } else {
#pragma unroll
                for (int i = 0; i < kWordsPerThread; i++) raw[i] = r0 + i * kRowStep < last_cnt ? __ldg(src + i * (kRowStep * 8)) : 0u;
            }
            src += kTcN * 8;
        };
        static_assert(kColsPerWarp == 64, "one packed 32-register load covers the warp's 64 columns");
        // drain, part 1: wait for the tile's MMAs and start the TMEM load (64 columns as int16 pairs)
        int* const gptr = shared_d2 + q0 + row;     // padded to whole query blocks
        int gval = 0x7f7f7f7f;
        auto drain_load = [&](int tile, uint32_t (&v)[32]) {
            const int st = tile & 1;
            mbar_wait(done0 + 8 * st, (uint32_t)((tile >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            const int cnt = (tile < n_full ? kTcN : last_cnt) - chalf * kColsPerWarp;   // valid columns of my half
            if (cnt > 0) {      // warp-uniform
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(trow + (uint32_t)(2 * st * kTcN)));
            }
        };
        // drain, part 2: the load has landed; update the top-2 and hand the accumulators back
        auto drain_finish = [&](int tile, uint32_t (&v)[32], int glim) {
            const int st = tile & 1;
            const int cnt = (tile < n_full ? kTcN : last_cnt) - chalf * kColsPerWarp;
            const int ib = ib0 + tile * kTcN;
            if (cnt > 0) {
                // the wait makes the registers of the load above valid: tie them to it for the compiler
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                    : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                      "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                      "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                    :: "memory");
            }
            // the accumulators are in registers: hand the TMEM stage back before looking at them
            asm volatile("tcgen05.fence::before_thread_sync;" ::);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
            if (cnt >= kColsPerWarp) tc_drain64<true>(v, pa, ib, kColsPerWarp, glim, gptr, best);
            else if (cnt > 0) tc_drain64<false>(v, pa, ib, cnt, glim, gptr, best);
        };
        if (ntiles > 0) fetch(0, raw_a);
        if (ntiles > 1) fetch(1, raw_b);
        for (int i = 0; i < ntiles; i++) {
            const int sb = i & (kTcBStages - 1);
            uint32_t v[32];
            if (i >= kTcLag) drain_load(i - kTcLag, v);      // TMEM latency hides behind the expansion below
            // buffer sb was last read by the MMAs of tile i - kTcBStages, whose completion this warp saw before draining it
            const uint32_t dst = b0 + sb * (kTcN * 256) + b_off;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) {
                uint4 c0, c1;
                expand32_map(raw_a[k], c0, c1);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst + k * (kRowStep / 8) * 2048), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst + k * (kRowStep / 8) * 2048 + 128), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * sb);
            // tile i + 1 was requested a whole iteration ago; tile i + 2 goes out now
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) raw_a[k] = raw_b[k];
            if (i + 2 < ntiles) fetch(i + 2, raw_b);
            const int glim = pa - gval - 1;        // the bound read one iteration ago
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(gval) : "l"(gptr) : "memory");
            if (i >= kTcLag) drain_finish(i - kTcLag, v, glim);
        }
        for (int t = max(0, ntiles - kTcLag); t < ntiles; t++) {
            uint32_t v[32];
            drain_load(t, v);
            int gnow;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(gnow) : "l"(gptr) : "memory");
            drain_finish(t, v, pa - gnow - 1);
        }
    }
    // ---- merge the column halves (index ranges interleave: lexicographic merge) and store
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    int4* s_rec = reinterpret_cast<int4*>(s_b);     // all MMAs retired: the B stages are free
    if (warp >= 8 && warp < kTcWorkers) {
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        s_rec[row] = make_int4(best.d1, best.i1, best.d2, best.i2);
    }
    __syncthreads();
    if (warp < 8) {
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        if (kTcWorkers > 8) {
            const int4 o = s_rec[row];
            top2_merge(best.d1, best.i1, best.d2, best.i2, o.x, o.y, o.z, o.w);
        }
        if (q0 + row < nq) partial[(size_t)blockIdx.y * nq + q0 + row] = make_int4(best.d1, best.i1, best.d2, best.i2);
    }
    if (warp == kTcWorkers) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// K7d: K7c with the query operand held in tensor memory.  With both operands in shared memory an M128 x N128 x K32
// MMA reads 8 KB per 64 cycles -- the whole shared-memory bandwidth -- and every expansion store of the workers
// steals from it.  The queries never change, so they are written ONCE into TMEM (row = lane, 4 K-bytes per 32-bit
// column: 64 columns per 128-query tile, 128 columns for both) and each MMA only streams its 4 KB B slab.
//   TMEM: columns [0,128) = A0 | A1; three accumulator slots of 128 columns behind them.
//   Unit u = (tile u/2, A tile u%2) -> slot u%3: 8 MMAs, one commit.  Workers: 4 lane quarters x 4 column quarters;
//   a thread follows the two query rows (one per A tile) of its lane and 32 columns of every unit.
// ---------------------------------------------------------------------------------------------
constexpr int kTsSlots = 3;
constexpr int kTsLag = 2;         // tiles between expansion and drain: the three slots hold 1.5 tiles of accumulators
constexpr int kTsAccCol0 = 128;      // first accumulator column

// drain 32 accumulator columns (16 packed int16 pairs) of one query row; see tc_drain64
template <bool FULL>
__device__ __forceinline__ void tc_drain32p(const uint32_t (&v)[16], int pa, int idx0, int valid, int glim, int* gptr, Top2& best) {
    uint32_t s[4];
#pragma unroll
    for (int g = 0; g < 4; g++) s[g] = __vmaxs2(__vimax3_s16x2(v[4 * g], v[4 * g + 1], v[4 * g + 2]), v[4 * g + 3]);
    const uint32_t mx = __vmaxs2(__vimax3_s16x2(s[0], s[1], s[2]), s[3]);
    const int d2_in = best.d2;
    int lim = max(pa - best.d2, glim);
    uint32_t bound2 = __byte_perm((uint32_t)(lim << kTcScaleLog2), 0u, 0x1010);
    if (__vmaxs2(mx, bound2) != bound2) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if (__vmaxs2(s[g], bound2) != bound2) {
                int key[8];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint32_t x = v[4 * g + r];
                    const int lo = ((int)(x << 16) >> (16 + kTcScaleLog2 - 3)) | (7 - 2 * r);
                    const int hi = ((int)x >> (16 + kTcScaleLog2 - 3)) | (6 - 2 * r);
                    key[2 * r] = FULL || 8 * g + 2 * r < valid ? lo : (int)0x80000000;
                    key[2 * r + 1] = FULL || 8 * g + 2 * r + 1 < valid ? hi : (int)0x80000000;
                }
                int k = max(__vimax3_s32(key[0], key[1], key[2]), __vimax3_s32(key[3], key[4], key[5]));
                k = __vimax3_s32(k, key[6], key[7]);
                while ((k >> 3) > lim) {
                    top2_push(best, pa - (k >> 3), idx0 + 8 * g + 7 - (k & 7));
                    lim = max(pa - best.d2, glim);
                    const int last = k;
                    k = (int)0x80000000;
#pragma unroll
                    for (int j = 0; j < 8; j++) k = max(k, key[j] < last ? key[j] : (int)0x80000000);
                }
                bound2 = __byte_perm((uint32_t)(lim << kTcScaleLog2), 0u, 0x1010);
            }
        }
        if (best.d2 < d2_in) atomicMin(gptr, best.d2);
    }
}

#define TMEM_LD16P(v, addr)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "                                                 \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"                  \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(addr))
#define TMEM_WAIT_LD16(v)                                                                                            \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                    \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),   \
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) \
                 :: "memory")

__global__ void __launch_bounds__(kTcThreads, 1) knn2_ts_kernel(const uint32_t* __restrict__ q, int nq,
                                                               const uint32_t* __restrict__ m, long long nm,
                                                               long long per_split, long long index_base,
                                                               int4* __restrict__ partial, int* __restrict__ shared_d2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* s_b = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // kTcBStages x [kTcN rows][256 B]
    __shared__ __align__(8) unsigned long long s_full[kTcBStages], s_done[kTsSlots], s_empty[kTsSlots];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kTcQ;
    constexpr int kWorkThreads = kTcWorkers * 32;

    if (warp == kTcWorkers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
#pragma unroll
        for (int sb = 0; sb < kTcBStages; sb++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_full[sb])), "r"(kTcWorkers));
#pragma unroll
        for (int st = 0; st < kTsSlots; st++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_done[st])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_empty[st])), "r"(kTcWorkers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t tmem = s_tmem;
    const int quarter = warp & 3;                               // TMEM lanes 32*quarter .. +31
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
    // ---- query operand -> TMEM: warps 0-3 write A0, warps 4-7 A1; a thread writes the 64 columns of its row
    if (warp < 8) {
        const int t = warp >> 2;
        const int row = q0 + t * kTcM + quarter * 32 + lane;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t bits = row < nq ? q[(size_t)row * 8 + w] : 0u;
            uint4 c0, c1;
            expand32_query(bits, c0, c1);       // plane j -> column 8*w + j, byte k -> K index 4*j + k of the word's slab
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "r"(tlane + (uint32_t)(t * 64 + w * 8)), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w), "r"(c1.x), "r"(c1.y),
                            "r"(c1.z), "r"(c1.w));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    const int ntiles = hi > lo ? (int)((hi - lo + kTcN - 1) / kTcN) : 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    uint32_t full0 = smem_u32(&s_full[0]), done0 = smem_u32(&s_done[0]), empty0 = smem_u32(&s_empty[0]), b0 = smem_u32(s_b);
    asm volatile("" : "+r"(full0), "+r"(done0), "+r"(empty0), "+r"(b0));

    Top2 best[2] = {{256, -1, 256, -1}, {256, -1, 256, -1}};
    const int cq = warp >> 2;                                   // column quarter (workers)
    if (warp == kTcWorkers) {
        // =========================== MMA warp ===========================
        const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        const uint64_t b_desc = umma_desc(b0);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        int slot = 0;
        uint32_t use_par = 0;       // parity of the slot's use count
        for (int i = 0; i < ntiles; i++) {
            const int sb = i & (kTcBStages - 1);
            mbar_wait_parked(full0 + 8 * sb, (uint32_t)((i / kTcBStages) & 1), 32);   // B tile i expanded by all workers
#pragma unroll
            for (int t = 0; t < 2; t++) {
                mbar_wait_parked(empty0 + 8 * slot, use_par ^ 1u, 32);   // previous use of the slot loaded out (passes at once the first time)
                asm volatile("tcgen05.fence::after_thread_sync;" ::);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        const uint64_t db = b_desc + (uint64_t)((sb * (kTcN * 256) + ks * 256) >> 4);
                        const uint32_t accumulate = ks ? 1u : 0u;
                        asm volatile(
                            "{\n\t"
                            ".reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
                            "}\n"
                            :: "r"(tmem + (uint32_t)(kTsAccCol0 + slot * kTcN)), "r"(tmem + (uint32_t)(t * 64 + ks * 8)), "l"(db), "r"(idesc),
                               "r"(accumulate), "r"(0u));
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(done0 + 8 * slot));
                }
                __syncwarp();
                if (++slot == kTsSlots) { slot = 0; use_par ^= 1u; }
            }
        }
    } else {
        // =========================== worker warps ===========================
        int pa[2];
        int* gptr[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int row = q0 + t * kTcM + quarter * 32 + lane;
            gptr[t] = shared_d2 + row;      // padded to whole query blocks
            pa[t] = 0;
            if (row < nq) {
                const uint4 x = *reinterpret_cast<const uint4*>(q + (size_t)row * 8);
                const uint4 y = *reinterpret_cast<const uint4*>(q + (size_t)row * 8 + 4);
                pa[t] = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
            }
        }
        constexpr int kColsPerWarp = 32;
        const uint32_t tacc = tlane + (uint32_t)(kTsAccCol0 + cq * kColsPerWarp);
        constexpr int kWordsPerThread = kTcN * 8 / kWorkThreads;   // 2
        constexpr int kRowStep = kWorkThreads / 8;                  // 64
        const int r0 = (tid & 7) | ((tid >> 6) << 3), w0 = (tid >> 3) & 7;
        const uint32_t* src = m + ((size_t)lo + r0) * 8 + w0;
        const uint32_t b_off = (uint32_t)((r0 >> 3) * 2048 + (r0 & 7) * 16 + (2 * w0) * 128);
        uint32_t raw[kWordsPerThread], raw_next[kWordsPerThread];     // tiles i and i + 1
        const int ib0 = (int)(index_base + lo) + cq * kColsPerWarp;
        const int n_full = (int)((hi - lo) / kTcN);
        const int last_cnt = (int)(hi - lo) - n_full * kTcN;
        auto fetch = [&](int tile, uint32_t (&dstw)[kWordsPerThread]) {      // called with tile = 0, 1, 2, ...
            if (tile < n_full) {
#pragma unroll
                for (int i = 0; i < kWordsPerThread; i++) dstw[i] = __ldg(src + i * (kRowStep * 8));
            } else {
#pragma unroll
                for (int i = 0; i < kWordsPerThread; i++) dstw[i] = r0 + i * kRowStep < last_cnt ? __ldg(src + i * (kRowStep * 8)) : 0u;
            }
            src += kTcN * 8;
        };
        int dslot = 0;              // slot / use parity of the next unit this warp drains
        uint32_t dpar = 0;
        int gval[2] = {0x7f7f7f7f, 0x7f7f7f7f};
        // wait for a unit's MMAs and start loading my 32 columns of it
        auto unit_load = [&](uint32_t (&v)[16]) {
            mbar_wait(done0 + 8 * dslot, dpar);
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            TMEM_LD16P(v, tacc + (uint32_t)(dslot * kTcN));
        };
        // the load has landed: give the slot back
        auto unit_release = [&](uint32_t (&v)[16]) {
            TMEM_WAIT_LD16(v);
            asm volatile("tcgen05.fence::before_thread_sync;" ::);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * dslot);
            if (++dslot == kTsSlots) { dslot = 0; dpar ^= 1u; }
        };
        auto unit_update = [&](const uint32_t (&v)[16], int tile, int t, int glim) {
            const int cnt = (tile < n_full ? kTcN : last_cnt) - cq * kColsPerWarp;     // valid columns of my quarter
            const int ib = ib0 + tile * kTcN;
            if (cnt >= kColsPerWarp) tc_drain32p<true>(v, pa[t], ib, kColsPerWarp, glim, gptr[t], best[t]);
            else if (cnt > 0) tc_drain32p<false>(v, pa[t], ib, cnt, glim, gptr[t], best[t]);
        };
        auto drain_tile = [&](int tile, uint32_t (&v0)[16], uint32_t (&v1)[16], int glim0, int glim1) {    // after unit_load(v0)
            unit_release(v0);
            unit_load(v1);
            unit_update(v0, tile, 0, glim0);
            unit_release(v1);
            unit_update(v1, tile, 1, glim1);
        };
        if (ntiles > 0) fetch(0, raw);
        if (ntiles > 1) fetch(1, raw_next);
        for (int i = 0; i < ntiles; i++) {
            const int sb = i & (kTcBStages - 1);
            uint32_t v0[16], v1[16];
            if (i >= kTsLag) unit_load(v0);      // TMEM latency hides behind the expansion below
            // buffer sb was last read by the MMAs of tile i - kTcBStages, whose completion this warp saw before draining it
            const uint32_t dst = b0 + sb * (kTcN * 256) + b_off;
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) {
                uint4 c0, c1;
                expand32_map(raw[k], c0, c1);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst + k * (kRowStep / 8) * 2048), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst + k * (kRowStep / 8) * 2048 + 128), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * sb);
            // tile i + 1 was requested a whole iteration ago; tile i + 2 goes out now
#pragma unroll
            for (int k = 0; k < kWordsPerThread; k++) raw[k] = raw_next[k];
            if (i + 2 < ntiles) fetch(i + 2, raw_next);
            const int glim0 = pa[0] - gval[0] - 1, glim1 = pa[1] - gval[1] - 1;
            if ((i & 3) == 0) {     // a stale bound is still a bound: one L2 round trip every fourth tile
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "+r"(gval[0]) : "l"(gptr[0]) : "memory");
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "+r"(gval[1]) : "l"(gptr[1]) : "memory");
            }
            if (i >= kTsLag) drain_tile(i - kTsLag, v0, v1, glim0, glim1);
        }
        for (int t = max(0, ntiles - kTsLag); t < ntiles; t++) {
            uint32_t v0[16], v1[16];
            unit_load(v0);
            drain_tile(t, v0, v1, pa[0] - gval[0] - 1, pa[1] - gval[1] - 1);
        }
    }
    // ---- merge the column quarters (index ranges interleave: lexicographic merge) and store
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    int4* s_rec = reinterpret_cast<int4*>(s_b);     // all MMAs retired: the B stages are free
    if (warp >= 4 && warp < kTcWorkers) {
#pragma unroll
        for (int t = 0; t < 2; t++)
            s_rec[(cq - 1) * kTcQ + t * kTcM + quarter * 32 + lane] = make_int4(best[t].d1, best[t].i1, best[t].d2, best[t].i2);
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int row = t * kTcM + quarter * 32 + lane;
            Top2 b = best[t];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const int4 o = s_rec[c * kTcQ + row];
                top2_merge(b.d1, b.i1, b.d2, b.i2, o.x, o.y, o.z, o.w);
            }
            if (q0 + row < nq) partial[(size_t)blockIdx.y * nq + q0 + row] = make_int4(b.d1, b.i1, b.d2, b.i2);
        }
    }
    if (warp == kTcWorkers) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// K7e: K7c on a CTA pair (tcgen05 cta_group::2).  Two CTAs of a cluster (two SMs of a TPC) take 512 queries of one map
// split: each keeps its own 256 queries (two A tiles) and expands only HALF of every 128-descriptor B tile (64 rows) into
// its own shared memory; one tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA, N = 128, K = 32) multiplies both A
// tiles against the whole B tile, each tensor core taking the other half from the peer.  Per CTA that is half the
// expansion instructions and stores per comparison and 6 KB instead of 8 KB of operand reads per MMA.
//   * both CTAs allocate TMEM with .cta_group::2; accumulators as in K7c (2 stages x 2 A tiles x 128 columns per CTA);
//   * only the leader (cluster rank 0) runs the MMA warp; its full/empty barriers count the worker warps of BOTH CTAs
//     (the peer arrives through mapa + mbarrier.arrive.release.cluster); tcgen05.commit multicasts `done` to both;
//   * everything else -- drain, shared pruning bound, merge -- is K7c's, per CTA.
// Protocol established with tools/probe/cta_pair.cu.
// ---------------------------------------------------------------------------------------------
constexpr int kPairHalfN = kTcN / 2;      // B rows expanded per CTA per tile
#ifndef PAIR_BOUND_EVERY
#define PAIR_BOUND_EVERY 7
#endif
constexpr int kPairBoundEvery = PAIR_BOUND_EVERY;   // refresh the shared bound when (tile & this) == 0
constexpr int kPairBStages = 8;           // expanded half tiles in flight per CTA (16 KB each)

__device__ __forceinline__ uint32_t cluster_cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
// arrive on the barrier at the same shared-memory offset in CTA `target` of the cluster.
// .relaxed on purpose: .release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive (measured: 55 % of
// all stall samples, 2.3x the run time).  What the consumer needs is already ordered by the instructions in front of
// the arrive: `fence.proxy.async` has made the expansion stores visible to the tensor cores' proxy (full barrier), and
// `tcgen05.wait::ld` + `tcgen05.fence::before_thread_sync` have completed the TMEM reads (empty barrier); the arrive
// itself is issued after them in program order by the same thread.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t target) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(target));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}

template <bool kShareBound>      // true: the sharded search of several ranks that forward their pruning bounds to each other (bp)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
knn2_pair_kernel(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ m, long long nm, long long per_split,
                 long long index_base, int4* __restrict__ partial, int* __restrict__ shared_d2, BoundPeers bp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* s_a = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 2 x [128 rows][256 B] (64 KB)
    unsigned char* s_b = s_a + kTcQ * 256;                                               // kPairBStages x [64 rows][256 B]
    __shared__ __align__(8) unsigned long long s_full[kPairBStages], s_done[2], s_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_workers_done;          // kShareBound: worker warps of this CTA that have left their tile loop
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_cta_rank();
    const int q0 = blockIdx.x * kTcQ;

    if (tid == 0) s_workers_done = 0;
    if (warp == kTcWorkers) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    if (tid == 0) {
#pragma unroll
        for (int sb = 0; sb < kPairBStages; sb++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_full[sb])), "r"(2 * kTcWorkers));
#pragma unroll
        for (int st = 0; st < 2; st++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_done[st])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_empty[st])), "r"(2 * kTcWorkers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    // ---- my A tiles (weighted +/-): element (r, c) of a tile at (r/8)*2048 + (c/16)*128 + (r%8)*16 + c%16
    for (int it = tid; it < kTcQ * 8; it += kTcThreads) {
        const int row = it >> 3, w = it & 7, r = row & 127;
        const uint32_t bits = q0 + row < nq ? q[(size_t)(q0 + row) * 8 + w] : 0u;
        uint4 c0, c1;
        expand32_query(bits, c0, c1);
        unsigned char* dst = s_a + (row >> 7) * (kTcM * 256) + (r >> 3) * 2048 + (r & 7) * 16 + (2 * w) * 128;
        *reinterpret_cast<uint4*>(dst) = c0;
        *reinterpret_cast<uint4*>(dst + 128) = c1;
    }
    const long long lo = (long long)blockIdx.y * per_split;
    const long long hi = min(nm, lo + per_split);
    const int ntiles = hi > lo ? (int)((hi - lo + kTcN - 1) / kTcN) : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::);
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    cluster_sync_all();             // barriers initialised, TMEM allocated and A tiles written in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::);
    const uint32_t tmem = s_tmem;
    uint32_t full0 = smem_u32(&s_full[0]), done0 = smem_u32(&s_done[0]), empty0 = smem_u32(&s_empty[0]), b0 = smem_u32(s_b);
    asm volatile("" : "+r"(full0), "+r"(done0), "+r"(empty0), "+r"(b0));

    Top2 best = {256, -1, 256, -1};
    if (warp == kTcWorkers) {
        if (rank == 0) {
            // =========================== MMA warp of the leader ===========================
            // D = s32, A = signed 8 bit, B = unsigned 8 bit, K-major, N = 128, M = 256 (128 rows in each CTA)
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)((2 * kTcM) >> 4) << 24);
            const uint64_t a_desc = umma_desc(smem_u32(s_a)), b_desc = umma_desc(b0);
            uint32_t leader;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
            for (int i = 0; i < ntiles; i++) {
                const int st = i & 1, sb = i & (kPairBStages - 1);
                // CTA-scope waits on purpose: an acquire.cluster wait adds an L1 invalidate (CCTL.IVALL) per success, and what
                // the MMAs read afterwards goes through the async proxy (shared memory, TMEM), not through L1
                mbar_wait(full0 + 8 * sb, (uint32_t)((i / kPairBStages) & 1));      // both halves of B tile i expanded
                mbar_wait(empty0 + 8 * st, (uint32_t)(((i >> 1) & 1) ^ 1));          // both CTAs loaded tile i-2 out
                asm volatile("tcgen05.fence::after_thread_sync;" ::);
                if (leader) {
#pragma unroll
                    for (int t = 0; t < 2; t++)
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint64_t da = a_desc + (uint64_t)((t * (kTcM * 256) + ks * 256) >> 4);
                            const uint64_t db = b_desc + (uint64_t)((sb * (kPairHalfN * 256) + ks * 256) >> 4);
                            const uint32_t accumulate = ks ? 1u : 0u;
                            asm volatile(
                                "{\n\t"
                                ".reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
                                "}\n"
                                :: "r"(tmem + (uint32_t)((2 * st + t) * kTcN)), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
                        }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 :: "r"(done0 + 8 * st), "h"((unsigned short)3));
                }
                __syncwarp();
            }
        } else if (kShareBound && bp.n > 0 && blockIdx.y == 0) {
            // =========================== bound forwarder (the otherwise idle 17th warp of the second CTA) ===========================
            // Sharded map: the other ranks prune with this rank's bounds too.  Publishing every improvement of every worker
            // thread to every peer was measured SLOWER than not sharing (millions of 4-byte NVLink transactions per search), and
            // a forwarding test inside the worker loop cost 3 % of the kernel.  This warp has nothing else to do: at ~1, 3, 7, 15,
            // ... us into the search (the bounds fall fastest at the start) it reads the
            // rank's bounds of the pair's 512 queries (the minimum over all the rank's CTAs, in the local array) and sends those
            // that fell to every peer as fire-and-forget system-scope minima.  A bound of one shard is a bound on the merged result.
            const int qb = (blockIdx.x & ~1) * kTcQ;
            int last[2 * kTcQ / 32];
#pragma unroll
            for (int k = 0; k < 2 * kTcQ / 32; k++) last[k] = 0x7f7f7f7f;
            // (the workers' "done" count is polled every ~0.1 us: __nanosleep(1000) was measured to hold the CTA -- and with it
            // the kernel -- for ~6 us after the last worker had left)
            unsigned long long t_next = 0, t_step = 1000;
            {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                t_next = now + t_step;
            }
            while (true) {
                __nanosleep(100);
                if (*reinterpret_cast<volatile int*>(&s_workers_done) == kTcWorkers) break;
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now < t_next) continue;
                t_step *= 2; t_next = now + t_step;      // forwarding passes ~1, 3, 7, 15, 31, ... us into the search
#pragma unroll
                for (int k = 0; k < 2 * kTcQ / 32; k++) {
                    const int qi = qb + k * 32 + lane;
                    if (qi < nq) {
                        int g;
                        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(g) : "l"(shared_d2 + qi) : "memory");
                        if (g < last[k]) {
                            last[k] = g;
                            for (int p = 0; p < bp.n; p++)
                                asm volatile("red.relaxed.sys.global.min.s32 [%0], %1;" :: "l"(bp.remote[p] + qi), "r"(g) : "memory");
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // =========================== worker warps (both CTAs) ===========================
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        const int chalf = warp >> 3;
        constexpr int kColsPerWarp = kTcN / (kTcWorkers / 8);
        static_assert(kColsPerWarp == 64, "one packed 32-register load covers the warp's 64 columns");
        int pa = 0;
        if (q0 + row < nq) {
            const uint4 x = *reinterpret_cast<const uint4*>(q + (size_t)(q0 + row) * 8);
            const uint4 y = *reinterpret_cast<const uint4*>(q + (size_t)(q0 + row) * 8 + 4);
            pa = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
        }
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * kTcN + chalf * kColsPerWarp);
        // my word of a tile: row 64*rank + r0 of the tile, word w0 (512 threads x 1 word = 64 rows x 8 words)
        const int r0 = (tid & 7) | ((tid >> 6) << 3), w0 = (tid >> 3) & 7;
        const int tile_row = (int)rank * kPairHalfN + r0;
        const uint32_t* src = m + ((size_t)lo + tile_row) * 8 + w0;
        const uint32_t b_off = (uint32_t)((r0 >> 3) * 2048 + (r0 & 7) * 16 + (2 * w0) * 128);
        const int ib0 = (int)(index_base + lo) + chalf * kColsPerWarp;
        const int n_full = (int)((hi - lo) / kTcN);
        const int last_cnt = (int)(hi - lo) - n_full * kTcN;
        uint32_t raw_a = 0, raw_b = 0;        // tiles i and i + 1
        auto fetch = [&](int tile) {          // called with tile = 0, 1, 2, ...
            const uint32_t v = (tile < n_full || tile_row < last_cnt) ? __ldg(src) : 0u;
            src += kTcN * 8;
            return v;
        };
        int* const gptr = shared_d2 + q0 + row;
        int gval = 0x7f7f7f7f;
        auto drain_load = [&](int tile, uint32_t (&v)[32]) {
            const int st = tile & 1;
            mbar_wait(done0 + 8 * st, (uint32_t)((tile >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::);
            const int cnt = (tile < n_full ? kTcN : last_cnt) - chalf * kColsPerWarp;
            if (cnt > 0) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(trow + (uint32_t)(2 * st * kTcN)));
            }
        };
        // the load has landed: hand the TMEM stage back at once (the recycle latency of the two stages, not the worker
        // instruction count, bounds the pair kernel), the top-2 update follows after the expansion
        auto drain_release = [&](int tile, uint32_t (&v)[32]) {
            const int st = tile & 1;
            const int cnt = (tile < n_full ? kTcN : last_cnt) - chalf * kColsPerWarp;
            if (cnt > 0) {
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                    : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                      "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                      "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                    :: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::);
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(empty0 + 8 * st, 0);
        };
        auto drain_update = [&](int tile, uint32_t (&v)[32], int glim) {
            const int cnt = (tile < n_full ? kTcN : last_cnt) - chalf * kColsPerWarp;
            const int ib = ib0 + tile * kTcN;
            if (cnt >= kColsPerWarp) tc_drain64<true>(v, pa, ib, kColsPerWarp, glim, gptr, best);
            else if (cnt > 0) tc_drain64<false>(v, pa, ib, cnt, glim, gptr, best);
        };
        if (ntiles > 0) raw_a = fetch(0);
        if (ntiles > 1) raw_b = fetch(1);
        for (int i = 0; i < ntiles; i++) {
            const int sb = i & (kPairBStages - 1);
            uint32_t v[32];
            if (i >= kTcLag) { drain_load(i - kTcLag, v); drain_release(i - kTcLag, v); }
            // buffer sb was last read by the MMAs of tile i - kPairBStages, whose completion this warp saw (multicast commit)
            const uint32_t dst = b0 + sb * (kPairHalfN * 256) + b_off;
            uint4 c0, c1;
            expand32_map(raw_a, c0, c1);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(dst + 128), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(full0 + 8 * sb, 0);
            raw_a = raw_b;
            if (i + 2 < ntiles) raw_b = fetch(i + 2);
            const int glim = pa - gval - 1;
            // a stale bound is still a bound: one L2 round trip every eighth tile, issued where nothing waits for it soon
            // (a load per tile stalled the loop top -- i.e. the TMEM hand-back -- for 23 % of the samples)
            // ... except over the first tiles, where the bound falls fastest: a search over a small map (one rank's shard of a
            // sharded map is ~50 tiles per CTA) would otherwise run a sixth of its tiles against "no bound yet"
            if ((i & kPairBoundEvery) == 0 || i < 8) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "+r"(gval) : "l"(gptr) : "memory");
            if (i >= kTcLag) drain_update(i - kTcLag, v, glim);
        }
        for (int t = max(0, ntiles - kTcLag); t < ntiles; t++) {
            uint32_t v[32];
            drain_load(t, v);
            drain_release(t, v);
            int gnow;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(gnow) : "l"(gptr) : "memory");
            drain_update(t, v, pa - gnow - 1);
        }
        if (kShareBound) {
            __syncwarp();
            if (lane == 0) atomicAdd(&s_workers_done, 1);
        }
    }
    // ---- merge the column halves and store (per CTA, as K7c)
    asm volatile("tcgen05.fence::before_thread_sync;" ::);
    __syncthreads();
    int4* s_rec = reinterpret_cast<int4*>(s_a);     // my A tiles are no longer read once every MMA has retired: see the cluster sync
    cluster_sync_all();                             // both CTAs are past their last drain: all MMAs retired, nobody reads my smem
    if (warp >= 8 && warp < kTcWorkers) {
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        s_rec[row] = make_int4(best.d1, best.i1, best.d2, best.i2);
    }
    __syncthreads();
    if (warp < 8) {
        const int row = ((warp >> 2) & 1) * kTcM + (warp & 3) * 32 + lane;
        const int4 o = s_rec[row];
        top2_merge(best.d1, best.i1, best.d2, best.i2, o.x, o.y, o.z, o.w);
        if (q0 + row < nq) partial[(size_t)blockIdx.y * nq + q0 + row] = make_int4(best.d1, best.i1, best.d2, best.i2);
    }
    if (warp == kTcWorkers) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::);
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
    }
}

// Merge of the per-split (or per-rank) records of every query: ONE WARP PER QUERY, lane = split, a shuffle butterfly of
// the associative lexicographic (distance, index) top-2 merge.  (A thread per query walking the splits one after the other
// is a chain of ~18 dependent L2 round trips: 14 us for 2000 queries x 18 splits, a third of a 125k-descriptor search.)
// bound / n_bound: the cross-CTA pruning bounds of the search that produced `parts` (tcgen05 variants); they are put back
// to "no bound yet" here, behind the search, so that the next search on this stream finds them ready (no memset per call).
// match != NULL: the reference's acceptance test on the merged record (R21/src/ORBmatcher.cc:228-230 / :598-600):
// match[q] = i1 if d1 <= th (strict: d1 < th) and (float)d1 < ratio * (float)d2, else -1 -- the 2-NN search and its ratio test
// leave the device as one result.
__global__ void __launch_bounds__(256) merge_top2_kernel(const int4* __restrict__ parts, int nparts, int nq, int4* __restrict__ out,
                                                         int* __restrict__ bound, int n_bound, int* __restrict__ match, float ratio, int th,
                                                         int strict) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = gtid; i < n_bound; i += gridDim.x * blockDim.x) bound[i] = 0x7f7f7f7f;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int qi = gtid >> 5; qi < nq; qi += nwarps) {
        int d1 = 256, i1 = -1, d2 = 256, i2 = -1;
        for (int p = lane; p < nparts; p += 32) {
            const int4 r = parts[(size_t)p * nq + qi];
            top2_merge(d1, i1, d2, i2, r.x, r.y, r.z, r.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int e1 = __shfl_xor_sync(0xffffffffu, d1, o), j1 = __shfl_xor_sync(0xffffffffu, i1, o);
            const int e2 = __shfl_xor_sync(0xffffffffu, d2, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
            top2_merge(d1, i1, d2, i2, e1, j1, e2, j2);
        }
        if (lane == 0) {
            if (out) out[qi] = make_int4(d1, i1, d2, i2);
            if (match) match[qi] = ((strict ? d1 < th : d1 <= th) && __int2float_rn(d1) < __fmul_rn(ratio, __int2float_rn(d2))) ? i1 : -1;
        }
    }
}

namespace {
inline int merge_grid(int nq) { return (std::max(1, std::min((nq + 7) / 8, 1184)) + 1) & ~1; }      // a warp per query, 8 warps per CTA, <= 8 CTAs per SM

constexpr size_t kKnnCounters = 4096;      // query blocks of one search (256 queries each)
struct KnnScratch { int* bound; int* counters; int4* partial; };
struct KnnScratchSlot { char* base = nullptr; size_t bound_cap = 0, partial_cap = 0; };
std::mutex g_scratch_mu;
std::map<std::pair<int, cudaStream_t>, KnnScratchSlot> g_scratch;

// n_bound ints of pruning bounds + partial_bytes of records for a search on (dev, s).  Growing (rare: first call, larger
// problem) drains the stream, replaces the buffer and initialises the whole bound region.
bool knn_scratch(int dev, cudaStream_t s, size_t n_bound, size_t partial_bytes, KnnScratch* out) {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    KnnScratchSlot& slot = g_scratch[std::make_pair(dev, s)];
    if (!slot.base || n_bound > slot.bound_cap || partial_bytes > slot.partial_cap) {
        const size_t bcap = std::max<size_t>(std::max(n_bound, slot.bound_cap), 4096);
        const size_t pcap = std::max<size_t>(std::max(partial_bytes, slot.partial_cap) * 3 / 2, (size_t)1 << 20);
        if (cudaStreamSynchronize(s) != cudaSuccess) return false;
        if (slot.base) cudaFree(slot.base);
        slot = KnnScratchSlot();
        char* p = nullptr;
        if (cudaMalloc((void**)&p, (bcap + kKnnCounters) * sizeof(int) + pcap) != cudaSuccess) return false;
        if (cudaMemsetAsync(p, 0x7f, bcap * sizeof(int), s) != cudaSuccess ||
            cudaMemsetAsync(p + bcap * sizeof(int), 0, kKnnCounters * sizeof(int), s) != cudaSuccess) { cudaFree(p); return false; }
        slot.base = p; slot.bound_cap = bcap; slot.partial_cap = pcap;
    }
    out->bound = reinterpret_cast<int*>(slot.base);
    out->counters = out->bound + slot.bound_cap;
    out->partial = reinterpret_cast<int4*>(slot.base + (slot.bound_cap + kKnnCounters) * sizeof(int));
    return true;
}
}  // namespace

int launch_knn2(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, int32_t* d_out, int variant,
                cudaStream_t s, PeerExchange* peer, const RatioTest* rt) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int qper = variant >= 3 ? kTcQ : (variant >= 1 ? kMmaQPerCta : kKnnThreads);
    const int tile = variant >= 3 ? kTcN : (variant == 1 ? kMmaTile : (variant == 2 ? 8 : kKnnTile));
    const int per_sm = variant >= 3 ? 1 : 2;
    int qblocks = (nq + qper - 1) / qper;
    if (variant == 5) qblocks = (qblocks + 1) & ~1;      // CTA pairs: an even number of query blocks (a padding block stores nothing)
    // enough map splits to fill the SMs, each at least one tile
    // never more CTAs than fit at once (a partial second wave would double the run time)
    int splits = std::max(1, (per_sm * sms) / qblocks);
    const int64_t max_splits = (nm + tile - 1) / tile;
    if (splits > max_splits) splits = (int)std::max<int64_t>(max_splits, 1);
    int64_t per_split = (nm + splits - 1) / splits;
    per_split = (per_split + tile - 1) / tile * tile;
    splits = nm > 0 ? (int)((nm + per_split - 1) / per_split) : 1;
    // scratch: one grow-only buffer per (device, stream) -- calls on one stream are ordered, so they can share it, and calls
    // on different streams (the reference's three matcher threads) never do.  Layout: [pruning bounds][per-split records].
    // The bounds are "no bound yet" (0x7f7f7f7f) whenever no search is in flight: the merge kernel behind every search
    // restores the entries that search used.  Nothing is allocated, freed or cleared per call.
    const int n_bound = variant >= 3 ? qblocks * kTcQ : 0;
    KnnScratch sc;
    if (!knn_scratch(dev, s, (size_t)n_bound, (size_t)splits * nq * sizeof(int4), &sc)) return -1;
    int4* partial = sc.partial;
    int* const bound = sc.bound;
    int* shared_bound = nullptr;       // set when the bounds of this search live in the peer buffer
    if (variant == 1) {
        const size_t smem = (size_t)kMmaTile * 256 + kMmaTile * sizeof(int);
        static DeviceOnce once_configured;
        if (!once_configured.run([&] { return cudaFuncSetAttribute(knn2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
        knn2_mma_kernel<<<dim3(qblocks, splits), kMmaThreads, smem, s>>>((const uint32_t*)d_q, nq, (const uint4*)d_m, nm, per_split,
                                                                       index_base, partial);
    } else if (variant == 3) {
        const size_t smem = (size_t)kTcQ * 256 + kTcBStages * (size_t)kTcN * 256 + 1024;
        static DeviceOnce once_configured3;
        if (!once_configured3.run([&] { return cudaFuncSetAttribute(knn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
        int* shared_d2 = bound;
        knn2_tc_kernel<<<dim3(qblocks, splits), kTcThreads, smem, s>>>((const uint32_t*)d_q, nq, (const uint32_t*)d_m, nm, per_split,
                                                                      index_base, partial, shared_d2);
    } else if (variant == 5) {
        const size_t smem = (size_t)kTcQ * 256 + kPairBStages * (size_t)kPairHalfN * 256 + 1024;
        static DeviceOnce once_configured5;
        if (!once_configured5.run([&] {
                return cudaFuncSetAttribute(knn2_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                       cudaFuncSetAttribute(knn2_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
            })) return -1;
        // sharded search over peer buffers: the bounds live in the peer buffer and every rank's kernel publishes into all of them
        // (ORBCUDA_SHARE_BOUND=0: each rank keeps its bounds to itself -- A/B switch, results are identical)
        // OFF unless ORBCUDA_SHARE_BOUND=1: measured on 2 / 4 / 8 B200s (profiles/r2_shard_search_*.txt) a search with shared
        // bounds takes ~5 us LONGER than one without (0.1107 vs 0.1059 ms at 4 GPUs, 0.0800 vs 0.0742 at 8), with one or two
        // searches in flight; only the first form of it -- the forwarding test inside the worker loop, which cost the single-GPU
        // kernel 3 % -- was ahead at 8 GPUs (0.0729 vs 0.0755).  Kept as a measured alternative; results are identical.
        static const bool share = [] { const char* e = getenv("ORBCUDA_SHARE_BOUND"); return e ? atoi(e) > 0 : false; }();
        BoundPeers bp{};
        int* shared_d2 = bound;
        if (peer && peer->connected && peer->world > 1 && share && n_bound <= peer->layout.bound_ints()) {
            shared_d2 = reinterpret_cast<int*>(peer->local + peer->layout.bound_offset());
            for (int r = 0; r < peer->world; r++)
                if (r != peer->rank) bp.remote[bp.n++] = reinterpret_cast<int*>(peer->peers.base[r] + peer->layout.bound_offset());
            shared_bound = shared_d2;
        }
        if (bp.n > 0)
            knn2_pair_kernel<true><<<dim3(qblocks, splits), kTcThreads, smem, s>>>((const uint32_t*)d_q, nq, (const uint32_t*)d_m, nm, per_split,
                                                                                  index_base, partial, shared_d2, bp);
        else
            knn2_pair_kernel<false><<<dim3(qblocks, splits), kTcThreads, smem, s>>>((const uint32_t*)d_q, nq, (const uint32_t*)d_m, nm, per_split,
                                                                                   index_base, partial, shared_d2, bp);
    } else if (variant == 4) {
        const size_t smem = kTcBStages * (size_t)kTcN * 256 + 1024;
        static DeviceOnce once_configured4;
        if (!once_configured4.run([&] { return cudaFuncSetAttribute(knn2_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
        int* shared_d2 = bound;
        knn2_ts_kernel<<<dim3(qblocks, splits), kTcThreads, smem, s>>>((const uint32_t*)d_q, nq, (const uint32_t*)d_m, nm, per_split,
                                                                      index_base, partial, shared_d2);
    } else if (variant == 2) {
        knn2_mma_stream_kernel<<<dim3(qblocks, splits), kMmaThreads, 0, s>>>((const uint32_t*)d_q, nq, (const uint4*)d_m, nm, per_split,
                                                                              index_base, partial);
    } else {
        knn2_popc_kernel<<<dim3(qblocks, splits), kKnnThreads, 0, s>>>((const uint4*)d_q, nq, (const uint4*)d_m, nm, per_split,
                                                                       index_base, partial);
    }
    if (peer) {
        // multi-GPU: merge my splits, push my records into every rank's buffer over NVLink, wait for theirs, merge: one kernel
        if (launch_merge_exchange(peer, partial, splits, nq, d_out, shared_bound ? shared_bound : bound, n_bound, s, shared_bound != nullptr) < 0) return -1;
    } else {
        merge_top2_kernel<<<merge_grid(nq), 256, 0, s>>>(partial, splits, nq, (int4*)d_out, bound, n_bound, rt ? rt->d_match : nullptr,
                                                           rt ? rt->ratio : 0.f, rt ? rt->th : 0, rt ? rt->strict : 0);
    }
    return 2;
}

int launch_merge_top2(const int32_t* d_parts, int parts, int nq, int32_t* d_out, cudaStream_t s) {
    merge_top2_kernel<<<merge_grid(nq), 256, 0, s>>>((const int4*)d_parts, parts, nq, (int4*)d_out, nullptr, 0, nullptr, 0.f, 0, 0);
    return 1;
}

// the ratio test alone on records already on the device (one "split")
int launch_ratio_test(const int32_t* d_rec, int nq, const RatioTest& rt, cudaStream_t s) {
    merge_top2_kernel<<<merge_grid(nq), 256, 0, s>>>((const int4*)d_rec, 1, nq, nullptr, nullptr, 0, rt.d_match, rt.ratio, rt.th, rt.strict);
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
