// internal.h -- shared host/device definitions of liborbcuda (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>

#include "orbcuda.h"

namespace orbcuda {

constexpr int kEdge = 19;          // EDGE_THRESHOLD (R21 ORBextractor.cc:74)
constexpr int kXPad = 32;          // interior column origin inside a padded plane (16-B aligned)
constexpr int kMinBorder = 16;     // EDGE_THRESHOLD-3 (R21 :773)
constexpr int kMaxLevels = ORB_MAX_LEVELS;
constexpr int kPyrTileW = 128;     // output tile of the resize kernel
constexpr int kPyrTileH = 64;       // 8 warps x 8 rows

// Geometry of one pyramid level for the current image size.  Lives in device constant-like global
// memory (one array per handle) and on the host.
struct LevelGeom {
    int w, h;               // level size (R21 :1112)
    int pitch;              // padded plane pitch in bytes; interior pixel (x,y) at (y+19)*pitch + 32 + x
    int plane_rows;         // h + 38
    int64_t plane_off;      // byte offset of the padded plane inside one frame's pyramid block
    int spitch;             // pitch of the un-padded planes (blurred level, score map); origin (0,0)
    int64_t splane_off;     // byte offset inside one frame's blurred / score block
    // FAST cell grid (R21 :773-807)
    int n_cols, n_rows, w_cell, h_cell;
    int cell_base;          // index of this level's first cell in the per-frame cell table
    int64_t cand_off;       // offset (in uint32 entries) of this level's candidate slots in a frame
    // quadtree (R21 :539-545)
    int n_feat;             // mnFeaturesPerLevel[level]
    int n_ini;              // number of root nodes
    float h_x;              // root width
    int kp_slot;            // first keypoint slot of this level inside a frame (slot = max(N+3, 4*nIni))
    int kp_cap;             // slot size
    // resize tables for producing this level from level-1 (offsets into the table buffers)
    int xtab_off, ytab_off;
    int rs_rows, rs_cols;   // source rows / columns (word aligned) one kPyrTileH x kPyrTileW output tile of this level needs
    // tiling of the two-phase resize kernel (pyr_resize2_kernel): output tiles of t2_w x t2_h, the largest source
    // footprint of a tile (t2_rows x t2_cols bytes, t2_cols a multiple of 4 incl. 4 bytes of slack) and its shared memory
    int t2_w, t2_h, t2_nx, t2_ny, t2_rows, t2_cols, t2_smem;
    // tiling of the pair-staged resize kernel (pyr_resize3_kernel, the default): output tiles of kPyrTileW x t3_h (8 warps x
    // t3_h / 8 rows), t3_rows staged source rows of t3_row_bytes (two bytes per source pixel), t3_smem bytes of shared memory
    int t3_h, t3_rows, t3_row_bytes, t3_smem;
    float scale;            // mvScaleFactor[level]
    float patch_size;       // (float)(int)(31*scale)
};

struct CellInfo {           // one FAST cell; interior = [x0,x1) x [y0,y1) in level coordinates
    int16_t x0, y0, x1, y1;
    int32_t slot_off;       // offset (uint32 entries) of the cell's candidate slot inside the level
};

struct ResizeTap {          // one destination column/row of cv::resize INTER_LINEAR
    int16_t ofs;            // source index
    int16_t c0, c1;         // 11-bit fixed-point weights
    int16_t pad;
};

struct FrameLayout {        // sizes of the per-frame device blocks
    int nlevels;
    int width, height;
    int in_pitch;           // device input image pitch
    int64_t pyr_bytes;      // padded pyramid block
    int64_t splane_bytes;   // blurred block == score block
    int n_cells;            // cells per frame
    int64_t cand_entries;   // candidate slot entries (uint32) per frame
    int kp_cap;             // keypoint slots per frame
    int node_cap;           // quadtree node capacity
    int ini_th;             // iniThFAST (per-cell threshold fallback, applied in the quadtree gather)
};

// flattened (level, block) grids: blocks [start[l], start[l+1]) belong to level l
struct LevelBlocks {
    int start[kMaxLevels + 1];
};

// ---- kernel launchers (each returns the number of kernels it launched) ----
struct DevPtrs {
    const uint8_t* in;          // [B][height][in_pitch]: pyramid level 0 is read from here directly (no copy)
    size_t in_frame_stride;     // bytes between frames of `in`
    uint8_t* pyr;               // [B][pyr_bytes]
    uint8_t* blur;              // [B][splane_bytes]
    uint8_t* score;             // [B][splane_bytes]  FAST score at minThFAST
    uint32_t* cand;             // [B][cand_entries]  per level: NMS survivors, packed x | y<<12 | score<<24 (relative to 16,16)
    int32_t* cell_count;        // [B][n_cells]  flag: the cell holds a survivor with score >= iniThFAST
    int32_t* level_raw;         // [B][kMaxLevels] NMS survivors appended to each level's list
    uint32_t* oct_scratch;      // [B][cand_entries] packed candidates in list order (quadtree input)
    uint16_t* oct_node;         // [B][cand_entries] node id per candidate
    uint32_t* sel;              // [B][kp_cap] selected candidate (packed) per slot
    int32_t* level_count;       // [B][kMaxLevels] keypoints per level
    const LevelGeom* geom;      // [nlevels]
    const CellInfo* cells;      // [n_cells]
    const ResizeTap* xtab;
    const ResizeTap* ytab;
};

int launch_pyramid(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hgeom, int n_frames, cudaStream_t s);
int launch_repack(const uint8_t* src, size_t src_row, size_t src_frame, uint8_t* dst, int dst_pitch, size_t dst_frame, int width,
                  int height, int n_frames, cudaStream_t s);
int launch_fast_score(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hgeom, int n_frames, int min_th,
                      cudaStream_t s);
int launch_blur(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hgeom, int n_frames, cudaStream_t s);
int launch_fast_cells(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hgeom, int n_frames, int ini_th,
                      cudaStream_t s);
int launch_octree(const DevPtrs& d, const FrameLayout& fl, int n_frames, cudaStream_t s);
// tensor maps (TMA descriptors) of the describe kernel's two staged neighbourhoods, cached per extractor handle
struct DescribeMaps;
DescribeMaps* describe_maps_create();
void describe_maps_destroy(DescribeMaps* m);
int launch_describe(const DevPtrs& d, const FrameLayout& fl, const LevelGeom* hgeom, int n_frames, orb_keypoint_t* d_kps, uint8_t* d_desc,
                    int32_t* d_counts, int cap, DescribeMaps* maps, cudaStream_t s);

// stand-alone quadtree on packed candidates already in device memory
int launch_octree_single(const uint32_t* d_cand, int n, int width, int height, int n_feat, int n_ini, float h_x,
                         int32_t* d_sel_idx, uint16_t* d_node, uint32_t* d_sel, int32_t* d_count, int kp_cap,
                         int node_cap, cudaStream_t s);
size_t octree_smem_bytes(int node_cap);

// matcher
constexpr int kMaxPeers = 16;
// symmetric record buffers of the sharded map search (peer.cu): every rank owns one, mapped into every other rank through CUDA IPC
struct PeerLayout {
    // [2 parities][world][nq_cap] int4 records, then [2][kMaxPeers] flags, then an error word
    int world, nq_cap;
    __host__ __device__ size_t record_index(int parity, int rank, int q) const { return ((size_t)parity * world + rank) * nq_cap + q; }
    __host__ __device__ size_t flags_offset() const { return (size_t)2 * world * nq_cap * sizeof(int4); }
    // pruning bounds of the sharded search, one int per query (rounded up to whole query blocks): every rank's tcgen05 kernel
    // publishes its second-best distances into EVERY rank's array, so a shard does not have to warm its bound up alone
    __host__ __device__ size_t bound_offset() const { return (flags_offset() + 2 * kMaxPeers * sizeof(unsigned) + 64 + 255) & ~(size_t)255; }
    __host__ __device__ int bound_ints() const { return (nq_cap + 511) / 512 * 512; }
    __host__ __device__ size_t bytes() const { return bound_offset() + (size_t)bound_ints() * sizeof(int); }
};
struct BoundPeers { int* remote[kMaxPeers - 1]; int n; };      // the other ranks' bound arrays (n = 0: not shared)
struct PeerPtrs { unsigned char* base[kMaxPeers]; };
struct PeerExchange {
    int device = 0, rank = 0, world = 0, nq_cap = 0;
    PeerLayout layout{};
    unsigned char* local = nullptr;
    PeerPtrs peers{};
    bool opened[kMaxPeers] = {false};
    unsigned epoch = 0;
    unsigned* d_counter = nullptr;      // blocks that finished the scatter phase
    int* d_error = nullptr;             // set by the kernel when the bounded wait ran out
    bool connected = false;
};
int launch_merge_exchange(PeerExchange* peer, const void* d_partial, int parts, int nq, int32_t* d_out, int* d_bound, int n_bound,
                          cudaStream_t s, bool bound_is_shared = false);
struct RatioTest { int32_t* d_match; float ratio; int th; int strict; };   // R21/src/ORBmatcher.cc:228-230 / :598-600 on the merged records
int launch_knn2(const uint8_t* d_q, int nq, const uint8_t* d_m, int64_t nm, int64_t index_base, int32_t* d_out,
                int variant, cudaStream_t s, PeerExchange* peer = nullptr, const RatioTest* rt = nullptr);
int launch_ratio_test(const int32_t* d_rec, int nq, const RatioTest& rt, cudaStream_t s);
int launch_merge_top2(const int32_t* d_parts, int parts, int nq, int32_t* d_out, cudaStream_t s);
void host_merge_top2(const int32_t* parts, int nparts, int nq, int32_t* out);

#ifdef __CUDACC__
// level of a block of a flattened (level, block) grid: start[l] of the unused levels equals the grid size, so a straight count
// of the starts at or below the block index is the level (7 uniform compares instead of a data-dependent loop per thread)
__device__ __forceinline__ int level_of_block(const LevelBlocks& lb, int bx, int& first_block) {
    int level = 0;
    first_block = lb.start[0];
#pragma unroll
    for (int l = 1; l < kMaxLevels; l++)
        if (bx >= lb.start[l]) { level++; first_block = lb.start[l]; }      // constant indices only: the struct stays in the parameter bank
    return level;
}
// Pointer to pixel (0,0) of a pyramid level and its row pitch.  Level 0 is the input image itself; levels >= 1
// live in the padded planes of the pyramid block (only the interior is ever written or read on the device:
// every consumer either stays inside the level or applies REFLECT_101 itself).
__device__ __forceinline__ const uint8_t* level_roi(const DevPtrs& d, const FrameLayout& fl, const LevelGeom& g, int level,
                                                    int frame, int& pitch) {
    if (level == 0) { pitch = fl.in_pitch; return d.in + (size_t)frame * d.in_frame_stride; }
    pitch = g.pitch;
    return d.pyr + (size_t)frame * fl.pyr_bytes + g.plane_off + (size_t)kEdge * g.pitch + kXPad;
}
#endif

// Per-thread workspace of the matcher entry points: one stream, one grow-only device arena and one pinned
// host arena per calling thread, so a call costs no cudaMalloc / stream creation and all its copies are
// truly asynchronous.  The reference constructs ORBmatcher objects on the stack from three threads at once
// (SURVEY 3.3); a thread-local context keeps the entry points re-entrant without locks.
struct MatchCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    char* dbase = nullptr; size_t dcap = 0, doff = 0;
    char* hbase = nullptr; size_t hcap = 0, hoff = 0;
    struct Pending { void* dst; const void* staged; size_t bytes; };
    Pending pend[8]; int npend = 0;
    // Uploads that follow each other are contiguous in the device arena AND in the staging arena (both advance by the same
    // padded size), so they leave as ONE copy when the first kernel is launched; downloads of neighbouring device blocks
    // likewise come back as one copy in finish().  (A call like the window search made 8 + 3 copies of a few KB each: ~5 us
    // apiece on the stream.)  Kernels and copies of the entry points therefore take their stream from s().
    char* up_d = nullptr; char* up_h = nullptr; size_t up_bytes = 0;
    const char* dl_d = nullptr; char* dl_h = nullptr; size_t dl_bytes = 0;
    bool copy_failed = false;
    bool begin(int dev, size_t dev_bytes, size_t host_bytes);   // select device, size the arenas, reset them
    void* dalloc(size_t bytes);                                   // device bump allocation (256-B aligned)
    void* upload(const void* src, size_t bytes);                  // staged H2D (sent by s() / finish()), returns the device copy
    bool download(void* dst, const void* dsrc, size_t bytes);     // D2H into staging (sent by finish()); copied out by finish()
    cudaStream_t s();                                             // the stream, after the pending uploads have been queued on it
    bool finish();                                                // stream sync + deliver the downloads
    ~MatchCtx();
};
MatchCtx& match_ctx();

// Function attributes (dynamic shared memory opt-in, memory-pool thresholds) are per device.  The matcher entry points
// are called concurrently from the reference's Tracking / LocalMapping / LoopClosing threads, so the per-device
// configuration runs under a lock and a device is marked configured only AFTER its configuration succeeded: a second
// thread arriving during the first call waits instead of launching with the attribute still unset.
struct DeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    // runs f() (returns true on success) once per device; returns false if f failed (it is retried on the next call)
    template <class F> bool run(F&& f) {
        int dev = 0;
        const bool indexed = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64;
        std::lock_guard<std::mutex> lock(mu);
        if (indexed && done[dev]) return true;
        if (!f()) return false;
        if (indexed) done[dev] = true;
        return true;
    }
};

void set_error(const char* fmt, ...);
bool cuda_ok(cudaError_t e, const char* what);

}  // namespace orbcuda

#define ORB_CUDA_TRY(expr)                                            \
    do {                                                              \
        if (!::orbcuda::cuda_ok((expr), #expr)) return ORB_ERR_CUDA;  \
    } while (0)
