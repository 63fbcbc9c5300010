// peer.cu -- the one exchange step of the sharded map search, fused with the merges around it.
//
// Each rank searches its shard of the map and ends up with per-split {d1, i1, d2, i2} records for every query.  The
// NCCL formulation is: merge my splits (kernel) -> ncclAllGather (32 KB per rank) -> merge the ranks' records (kernel).
// Here it is ONE kernel over peer memory: every rank owns a symmetric buffer (cudaMalloc + CUDA IPC, so it works with one
// process per GPU); a thread merges the splits of its query and stores the record straight into slot [my rank] of EVERY
// rank's buffer (16-byte stores over NVLink / NVSwitch), the last block of the grid publishes "rank r, epoch e is
// complete" with a system-scope release store into every buffer, then every block waits (system-scope acquire loads,
// bounded) for all ranks' flags in its OWN buffer and merges the world's records of its queries.  The merge is the
// associative lexicographic (distance, index) reduction, so the result equals the single-GPU search bit for bit.
//   * no deadlock: the grid is sized to be fully resident (grid-stride loops), a rank's scatter depends on nothing remote,
//     and the wait is bounded by a timer (it raises an error flag instead of hanging if a peer died);
//   * two parities of slots/flags: a fast rank's scatter of call k+1 cannot overwrite what a slow rank still reads in
//     call k, and nobody can be two calls ahead (call k+2 needs everybody's flags of call k+1).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "internal.h"
#include "orbcuda.h"

namespace orbcuda {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ bool lex_less_rec(int da, int ia, int db, int ib) { return da < db || (da == db && (unsigned)ia < (unsigned)ib); }
__device__ __forceinline__ void merge_rec(int4& a, const int4 b) {
    // a and b each hold (best, second) in lexicographic (distance, index) order
    if (lex_less_rec(b.x, b.y, a.x, a.y)) {
        if (lex_less_rec(b.z, b.w, a.x, a.y)) { a.z = b.z; a.w = b.w; } else { a.z = a.x; a.w = a.y; }
        a.x = b.x; a.y = b.y;
    } else if (lex_less_rec(b.x, b.y, a.z, a.w)) { a.z = b.x; a.w = b.y; }
}

__global__ void __launch_bounds__(256) merge_exchange_kernel(const int4* __restrict__ parts, int nparts, int nq, int rank, PeerLayout L, PeerPtrs peers,
                                                             unsigned epoch, unsigned* __restrict__ counter, int* __restrict__ error,
                                                             unsigned long long timeout_ns, int4* __restrict__ out, int* __restrict__ bound,
                                                             int n_bound, int bound_is_shared) {
    const int parity = (int)(epoch & 1u);
    // the pruning bounds of the search in front of this kernel go back to "no bound yet" for the next one (match.cu).  Bounds that
    // the other ranks write into (shared bounds) are reset only behind the wait below: every rank's search of this call is over
    // then, and nobody starts the next one before this kernel ends, so no late value of this call can leak into the next.
    if (!bound_is_shared)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bound; i += gridDim.x * blockDim.x) bound[i] = 0x7f7f7f7f;
    // ---- phase 1: merge my splits (a warp per query, lane = split: one round of loads instead of a chain of them) and store
    // the record into every rank's buffer (lane r stores to rank r)
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = gwarp; q < nq; q += nwarps) {
        int4 rec = make_int4(256, -1, 256, -1);
        for (int p = lane; p < nparts; p += 32) merge_rec(rec, parts[(size_t)p * nq + q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int4 v;
            v.x = __shfl_xor_sync(0xffffffffu, rec.x, o); v.y = __shfl_xor_sync(0xffffffffu, rec.y, o);
            v.z = __shfl_xor_sync(0xffffffffu, rec.z, o); v.w = __shfl_xor_sync(0xffffffffu, rec.w, o);
            merge_rec(rec, v);
        }
        const size_t at = L.record_index(parity, rank, q);
        for (int r = lane; r < L.world; r += 32) reinterpret_cast<int4*>(peers.base[r])[at] = rec;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        // every block's records are out (fence + counter): publish "rank, epoch" everywhere, reset the counter for the next call
        if ((int)threadIdx.x < L.world)
            st_release_sys(reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + L.flags_offset()) + parity * kMaxPeers + rank, epoch);
        if (threadIdx.x == 0) *counter = 0;
    }
    // ---- phase 2: wait for everybody's flag in MY buffer, then merge the world's records of my queries
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(peers.base[rank] + L.flags_offset()) + parity * kMaxPeers;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < L.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(my_flags + threadIdx.x) != epoch) {
            if (global_timer_ns() - t0 > timeout_ns) { s_ok = 0; atomicExch(error, 1 + (int)threadIdx.x); break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
    // the bound array in MY peer buffer: other ranks may have published into it during this search whether or not my own
    // search read it (the ranks decide per shard size) -- it goes back to "no bound yet" here in every case
    {
        int* pb = reinterpret_cast<int*>(peers.base[rank] + L.bound_offset());
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.bound_ints(); i += gridDim.x * blockDim.x) pb[i] = 0x7f7f7f7f;
    }
    if (!s_ok) {
        // a peer did not arrive in time: the call's output is the "no match" record for every query, never stale data
        for (int q = gwarp; q < nq; q += nwarps) if (lane == 0) out[q] = make_int4(256, -1, 256, -1);
        return;
    }
    const int4* mine = reinterpret_cast<const int4*>(peers.base[rank]);
    for (int q = gwarp; q < nq; q += nwarps) {
        int4 rec = make_int4(256, -1, 256, -1);
        for (int r = lane; r < L.world; r += 32) {
            int4 v;      // written by another GPU: read it past the L1
            asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "l"(mine + L.record_index(parity, r, q)) : "memory");
            merge_rec(rec, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int4 v;
            v.x = __shfl_xor_sync(0xffffffffu, rec.x, o); v.y = __shfl_xor_sync(0xffffffffu, rec.y, o);
            v.z = __shfl_xor_sync(0xffffffffu, rec.z, o); v.w = __shfl_xor_sync(0xffffffffu, rec.w, o);
            merge_rec(rec, v);
        }
        if (lane == 0) out[q] = rec;
    }
}

// The same exchange as a SMALL grid (a thread per query, 128 threads per block: 16 blocks for 2000 queries).  A grid this small
// always finds room next to the tensor-core kernel of ANOTHER search (144 of the 148 SMs, one CTA each, most of the registers):
// with two or three searches in flight on their own streams and peer buffers, the cross-GPU wait of one search hides behind the
// kernel of the next.  (The warp-per-query grid above is 250 blocks; next to a running search kernel only a few of them are
// resident, they spin on the flags, and the blocks that still have to send their records wait for an SM: the exchange then ends
// only when the other search's kernel does.)  The loads of a thread are issued in batches of six before the first merge.
__global__ void __launch_bounds__(128) merge_exchange_small_kernel(const int4* __restrict__ parts, int nparts, int nq, int rank, PeerLayout L,
                                                                   PeerPtrs peers, unsigned epoch, unsigned* __restrict__ counter,
                                                                   int* __restrict__ error, unsigned long long timeout_ns, int4* __restrict__ out,
                                                                   int* __restrict__ bound, int n_bound, int bound_is_shared) {
    const int parity = (int)(epoch & 1u);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    if (!bound_is_shared)
        for (int i = gtid; i < n_bound; i += gsize) bound[i] = 0x7f7f7f7f;
    for (int q = gtid; q < nq; q += gsize) {
        int4 rec = make_int4(256, -1, 256, -1);
        for (int p0 = 0; p0 < nparts; p0 += 6) {
            int4 v[6];
#pragma unroll
            for (int j = 0; j < 6; j++) v[j] = p0 + j < nparts ? parts[(size_t)(p0 + j) * nq + q] : make_int4(256, -1, 256, -1);
#pragma unroll
            for (int j = 0; j < 6; j++) merge_rec(rec, v[j]);
        }
        const size_t at = L.record_index(parity, rank, q);
        for (int r = 0; r < L.world; r++) reinterpret_cast<int4*>(peers.base[r])[at] = rec;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        if ((int)threadIdx.x < L.world)
            st_release_sys(reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + L.flags_offset()) + parity * kMaxPeers + rank, epoch);
        if (threadIdx.x == 0) *counter = 0;
    }
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(peers.base[rank] + L.flags_offset()) + parity * kMaxPeers;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < L.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(my_flags + threadIdx.x) != epoch) {
            if (global_timer_ns() - t0 > timeout_ns) { s_ok = 0; atomicExch(error, 1 + (int)threadIdx.x); break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
    {   // the bound array in MY peer buffer (see merge_exchange_kernel)
        int* pb = reinterpret_cast<int*>(peers.base[rank] + L.bound_offset());
        for (int i = gtid; i < L.bound_ints(); i += gsize) pb[i] = 0x7f7f7f7f;
    }
    if (!s_ok) {
        for (int q = gtid; q < nq; q += gsize) out[q] = make_int4(256, -1, 256, -1);
        return;
    }
    const int4* mine = reinterpret_cast<const int4*>(peers.base[rank]);
    for (int q = gtid; q < nq; q += gsize) {
        int4 rec = make_int4(256, -1, 256, -1);
        for (int r0 = 0; r0 < L.world; r0 += 8) {
            int4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                v[j] = make_int4(256, -1, 256, -1);
                if (r0 + j < L.world)      // written by another GPU: read it past the L1
                    asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                                 : "l"(mine + L.record_index(parity, r0 + j, q)) : "memory");
            }
#pragma unroll
            for (int j = 0; j < 8; j++) merge_rec(rec, v[j]);
        }
        out[q] = rec;
    }
}

// The same protocol for an opaque message (the key-frame message of wire.cu): every rank stores its `units` 16-byte units into
// slot [rank] of every rank's buffer, the last block publishes the flag, every block waits for all flags and copies the
// world's slots out of its own buffer into dst ([world][dst_units]).
__global__ void __launch_bounds__(256) blob_exchange_kernel(const int4* __restrict__ src, int units, int rank, PeerLayout L, PeerPtrs peers, unsigned epoch,
                                                            unsigned* __restrict__ counter, int* __restrict__ error, unsigned long long timeout_ns,
                                                            int4* __restrict__ dst, size_t dst_units) {
    const int parity = (int)(epoch & 1u);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    for (int r = 0; r < L.world; r++) {
        int4* slot = reinterpret_cast<int4*>(peers.base[r]) + L.record_index(parity, rank, 0);
        for (int i = gtid; i < units; i += gsize) slot[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        if ((int)threadIdx.x < L.world)
            st_release_sys(reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + L.flags_offset()) + parity * kMaxPeers + rank, epoch);
        if (threadIdx.x == 0) *counter = 0;
    }
    const unsigned* my_flags = reinterpret_cast<const unsigned*>(peers.base[rank] + L.flags_offset()) + parity * kMaxPeers;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < L.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(my_flags + threadIdx.x) != epoch) {
            if (global_timer_ns() - t0 > timeout_ns) { s_ok = 0; atomicExch(error, 1 + (int)threadIdx.x); break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
    if (!s_ok) return;
    for (int r = 0; r < L.world; r++) {
        const int4* slot = reinterpret_cast<const int4*>(peers.base[rank]) + L.record_index(parity, r, 0);
        for (int i = gtid; i < units; i += gsize) {
            int4 v;      // written by another GPU: read it past the L1
            asm volatile("ld.relaxed.sys.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(slot + i) : "memory");
            dst[(size_t)r * dst_units + i] = v;
        }
    }
}

int launch_merge_exchange(PeerExchange* pe, const void* d_partial, int parts, int nq, int32_t* d_out, int* d_bound, int n_bound,
                          cudaStream_t s, bool bound_is_shared) {
    if (!pe || !pe->connected || nq > pe->nq_cap) { set_error("merge-exchange: peer buffers not connected or nq > capacity"); return -1; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pe->device);
    // ORBCUDA_MEX=0: the warp-per-query grid (up to 4 blocks per SM), 1 (default): the small thread-per-query grid (see its header)
    static const int small = [] { const char* e = getenv("ORBCUDA_MEX"); return e ? atoi(e) : 1; }();
    if (small) {
        const int grid = std::max(1, std::min((nq + 127) / 128, sms / 4));
        merge_exchange_small_kernel<<<grid, 128, 0, s>>>((const int4*)d_partial, parts, nq, pe->rank, pe->layout, pe->peers, pe->epoch + 1,
                                                        pe->d_counter, pe->d_error, 5ull * 1000 * 1000 * 1000, (int4*)d_out, d_bound, n_bound,
                                                        bound_is_shared ? 1 : 0);
    } else {
        const int grid = std::max(1, std::min((nq + 7) / 8, 4 * sms));      // a warp per query; resident for sure: 256 threads, a few registers
        merge_exchange_kernel<<<grid, 256, 0, s>>>((const int4*)d_partial, parts, nq, pe->rank, pe->layout, pe->peers, pe->epoch + 1, pe->d_counter,
                                                  pe->d_error, 5ull * 1000 * 1000 * 1000, (int4*)d_out, d_bound, n_bound, bound_is_shared ? 1 : 0);
    }
    if (cudaPeekAtLastError() != cudaSuccess) return -1;      // the epoch advances only with a launch that went out: ranks stay in step
    pe->epoch++;
    return 1;
}

}  // namespace orbcuda

using namespace orbcuda;

extern "C" {

int orbm_peer_create(int nq_cap, int rank, int world, int device, orbm_peer_t* out, void* ipc_handle64) {
    if (nq_cap < 1 || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !out || !ipc_handle64) { set_error("orbm_peer_create: bad arguments (world <= %d)", kMaxPeers); return ORB_ERR_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the IPC handle travels as 64 opaque bytes");
    ORB_CUDA_TRY(cudaSetDevice(device));
    PeerExchange* pe = new PeerExchange;
    pe->device = device; pe->rank = rank; pe->world = world; pe->nq_cap = nq_cap;
    pe->layout.world = world; pe->layout.nq_cap = nq_cap;
    if (!cuda_ok(cudaMalloc((void**)&pe->local, pe->layout.bytes()), "cudaMalloc") || !cuda_ok(cudaMemset(pe->local, 0, pe->layout.bytes()), "cudaMemset") ||
        !cuda_ok(cudaMemset(pe->local + pe->layout.bound_offset(), 0x7f, (size_t)pe->layout.bound_ints() * sizeof(int)), "cudaMemset") ||      // "no bound yet"
        !cuda_ok(cudaMalloc((void**)&pe->d_counter, 64), "cudaMalloc") || !cuda_ok(cudaMemset(pe->d_counter, 0, 64), "cudaMemset")) {
        orbm_peer_destroy(reinterpret_cast<orbm_peer_t>(pe));
        return ORB_ERR_CUDA;
    }
    pe->d_error = reinterpret_cast<int*>(pe->d_counter) + 8;
    cudaIpcMemHandle_t hnd;
    if (!cuda_ok(cudaIpcGetMemHandle(&hnd, pe->local), "cudaIpcGetMemHandle")) { orbm_peer_destroy(reinterpret_cast<orbm_peer_t>(pe)); return ORB_ERR_CUDA; }
    memcpy(ipc_handle64, &hnd, 64);
    pe->peers.base[rank] = pe->local;
    if (world == 1) pe->connected = true;
    *out = reinterpret_cast<orbm_peer_t>(pe);
    return ORB_OK;
}

int orbm_peer_connect(orbm_peer_t p, const void* handles) {
    PeerExchange* pe = reinterpret_cast<PeerExchange*>(p);
    if (!pe || !handles) { set_error("orbm_peer_connect: bad arguments"); return ORB_ERR_ARG; }
    ORB_CUDA_TRY(cudaSetDevice(pe->device));
    for (int r = 0; r < pe->world; r++) {
        if (r == pe->rank || pe->opened[r]) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, static_cast<const unsigned char*>(handles) + (size_t)r * 64, 64);
        void* ptr = nullptr;
        if (!cuda_ok(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return ORB_ERR_CUDA;
        pe->peers.base[r] = static_cast<unsigned char*>(ptr);
        pe->opened[r] = true;
    }
    pe->connected = true;
    return ORB_OK;
}

int orbm_peer_destroy(orbm_peer_t p) {
    PeerExchange* pe = reinterpret_cast<PeerExchange*>(p);
    if (!pe) return ORB_OK;
    cudaSetDevice(pe->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < pe->world; r++)
        if (pe->opened[r]) cudaIpcCloseMemHandle(pe->peers.base[r]);
    if (pe->local) cudaFree(pe->local);
    if (pe->d_counter) cudaFree(pe->d_counter);
    cudaGetLastError();
    delete pe;
    return ORB_OK;
}

int orbm_knn2_exchange_device(orbm_peer_t p, const uint8_t* d_q, int nq, const uint8_t* d_m_shard, int64_t nm, int64_t index_base,
                              int32_t* d_out, int variant, void* stream) {
    PeerExchange* pe = reinterpret_cast<PeerExchange*>(p);
    if (!pe || !d_q || nq < 1 || !d_out || nm < 0 || (nm > 0 && !d_m_shard) || variant < 0 || variant > 5) { set_error("orbm_knn2_exchange_device: bad arguments"); return ORB_ERR_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_q) & 15) || (reinterpret_cast<uintptr_t>(d_m_shard) & 15)) { set_error("orbm_knn2_exchange_device: descriptor arrays must be 16-byte aligned"); return ORB_ERR_ARG; }
    ORB_CUDA_TRY(cudaSetDevice(pe->device));
    if (launch_knn2(d_q, nq, d_m_shard, nm, index_base, d_out, variant, (cudaStream_t)stream, pe) < 0) {
        cuda_ok(cudaGetLastError(), "knn2 + exchange launch");
        return ORB_ERR_CUDA;
    }
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

// All-gather of one opaque message per rank over the peer buffers (orbm_peer_create with nq_cap >= bytes / 16): d_all receives
// [world][slot_bytes].  Same calling rules as orbm_knn2_exchange_device (every rank makes the same sequence of calls on this
// peer object; use a separate object for messages and for searches).
int orbw_exchange_messages_device(orbm_peer_t p, const void* d_msg, size_t bytes, void* d_all, size_t slot_bytes, void* stream) {
    PeerExchange* pe = reinterpret_cast<PeerExchange*>(p);
    if (!pe || !pe->connected || !d_msg || !d_all || (bytes & 15) || (slot_bytes & 15) || bytes > slot_bytes || bytes / 16 > (size_t)pe->nq_cap ||
        (reinterpret_cast<uintptr_t>(d_msg) & 15) || (reinterpret_cast<uintptr_t>(d_all) & 15)) {
        set_error("orbw_exchange_messages_device: bad arguments (16-byte aligned buffers, bytes <= slot_bytes, bytes / 16 <= the peer's capacity)");
        return ORB_ERR_ARG;
    }
    ORB_CUDA_TRY(cudaSetDevice(pe->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pe->device);
    const int units = (int)(bytes / 16);
    const int grid = std::max(1, std::min((units + 255) / 256, 2 * sms));
    blob_exchange_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4*)d_msg, units, pe->rank, pe->layout, pe->peers, pe->epoch + 1, pe->d_counter, pe->d_error,
                                                                 5ull * 1000 * 1000 * 1000, (int4*)d_all, slot_bytes / 16);
    ORB_CUDA_TRY(cudaGetLastError());
    pe->epoch++;
    return ORB_OK;
}

int orbm_peer_error(orbm_peer_t p, int* error) {
    PeerExchange* pe = reinterpret_cast<PeerExchange*>(p);
    if (!pe || !error) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(pe->device));
    ORB_CUDA_TRY(cudaMemcpy(error, pe->d_error, 4, cudaMemcpyDeviceToHost));
    if (*error) ORB_CUDA_TRY(cudaMemset(pe->d_error, 0, 4));      // reported once: the next exchange starts clean
    return ORB_OK;
}

}  // extern "C"
