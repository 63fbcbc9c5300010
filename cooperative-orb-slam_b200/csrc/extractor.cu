// extractor.cu -- host side of the extractor: handle, level geometry, staging and the orbx_* C ABI.
// Mirrors ORB_SLAM2::ORBextractor (R21/include/ORBextractor.h:45-111, R21/src/ORBextractor.cc:410-470,
// :1043-1132): same constructor arithmetic for the scale tables / features per level / cell grid,
// kernels K1..K5 replace the per-frame work.
#include "internal.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

namespace orbcuda {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_error("CUDA error %s (%d) in %s", cudaGetErrorString(e), (int)e, what);
    return false;
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound: round-half-even
static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace orbcuda

using namespace orbcuda;

struct orbx_handle_s {
    orbx_params_t prm;
    int device = 0;
    double scale_factor_d = 1.2;   // R21/include/ORBextractor.h:98: a double member holding the float argument
    std::vector<float> sf, isf, s2, is2;
    std::vector<int> nfeat;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[10] = {};
    cudaEvent_t ev_done = nullptr;   // "last extraction enqueued on `stream`" for consumers on other streams
    float* d_tables = nullptr;       // [2][kMaxLevels]: mvScaleFactor, mvInvScaleFactor (device copy for the batched stereo search)
    bool profiling = false;
    float stage_ms[9] = {};
    bool timing_pending = false;
    int64_t launches = 0;

    // geometry of the current image size
    int cur_w = 0, cur_h = 0;
    FrameLayout fl = {};
    std::vector<LevelGeom> geom;
    std::vector<CellInfo> cells;
    std::vector<ResizeTap> xtab, ytab;
    LevelGeom* d_geom = nullptr;
    CellInfo* d_cells = nullptr;
    ResizeTap* d_xtab = nullptr;
    ResizeTap* d_ytab = nullptr;
    size_t cap_geom = 0, cap_cells = 0, cap_xtab = 0, cap_ytab = 0;

    // per-batch device buffers (grown on demand)
    int batch_cap = 0;
    uint8_t* d_in = nullptr; size_t cap_in = 0;
    uint8_t* d_raw = nullptr; size_t cap_raw = 0;     // caller-strided upload, repacked into d_in on the device
    uint8_t* d_pyr = nullptr; size_t cap_pyr = 0;
    uint8_t* d_blur = nullptr; size_t cap_blur = 0;
    uint8_t* d_score = nullptr; size_t cap_score = 0;
    uint32_t* d_cand = nullptr; size_t cap_cand = 0;
    uint32_t* d_scratch = nullptr; size_t cap_scratch = 0;
    uint16_t* d_node = nullptr; size_t cap_node = 0;
    int32_t* d_cell_count = nullptr; size_t cap_cell_count = 0;
    int32_t* d_level_raw = nullptr; size_t cap_level_raw = 0;
    uint32_t* d_sel = nullptr; size_t cap_sel = 0;
    int32_t* d_level_count = nullptr; size_t cap_level_count = 0;
    orb_keypoint_t* d_kps = nullptr; size_t cap_kps = 0;
    uint8_t* d_desc = nullptr; size_t cap_desc = 0;
    int32_t* d_counts = nullptr; size_t cap_counts = 0;
    // pinned staging
    uint8_t* h_in = nullptr; size_t cap_h_in = 0;
    orb_keypoint_t* h_kps = nullptr; size_t cap_h_kps = 0;
    uint8_t* h_desc = nullptr; size_t cap_h_desc = 0;
    int32_t* h_counts = nullptr; size_t cap_h_counts = 0;

    // CUDA graph of the kernel chain for small (latency-bound) calls: captured on the second call with the same signature,
    // replayed afterwards -- one graph launch instead of 12 kernel launches (ORBCUDA_GRAPH=0 disables it)
    struct GraphKey {
        int n_frames, width, height, cap, in_pitch; const void *in, *kps, *desc, *counts; size_t in_stride;
        bool operator==(const GraphKey& o) const {
            return n_frames == o.n_frames && width == o.width && height == o.height && cap == o.cap && in_pitch == o.in_pitch && in == o.in &&
                   kps == o.kps && desc == o.desc && counts == o.counts && in_stride == o.in_stride;
        }
    };
    cudaGraphExec_t graph_exec = nullptr;
    GraphKey graph_key = {}, graph_seen = {};
    int graph_kernels = 0;
    int64_t graph_replays = 0;
    DescribeMaps* dmaps = nullptr;      // TMA descriptors of the describe kernel (describe.cu)
    ~orbx_handle_s() { if (dmaps) describe_maps_destroy(dmaps); }

    // pending async call
    bool pending = false;
    int p_frames = 0, p_cap = 0;
    orb_keypoint_t* p_kps = nullptr; uint8_t* p_desc = nullptr; int32_t* p_counts = nullptr;
    bool p_direct = false;
    int last_frames = 0;
    // level 0 of the last extraction is the input itself (the handle's own d_in, or the caller's device buffer)
    const uint8_t* last_in = nullptr; size_t last_in_stride = 0; int last_in_pitch = 0;
};

namespace {

template <class T> int grow_dev(T*& p, size_t& cap, size_t need_bytes) {
    if (need_bytes <= cap) return ORB_OK;
    if (p) ORB_CUDA_TRY(cudaFree(p));
    p = nullptr; cap = 0;
    ORB_CUDA_TRY(cudaMalloc((void**)&p, need_bytes));
    cap = need_bytes;
    return ORB_OK;
}
template <class T> int grow_host(T*& p, size_t& cap, size_t need_bytes) {
    if (need_bytes <= cap) return ORB_OK;
    if (p) ORB_CUDA_TRY(cudaFreeHost(p));
    p = nullptr; cap = 0;
    ORB_CUDA_TRY(cudaHostAlloc((void**)&p, need_bytes, cudaHostAllocDefault));
    cap = need_bytes;
    return ORB_OK;
}

// cv::resize INTER_LINEAR tap table (OpenCV imgproc resize.cpp)
void make_taps(int ssize, int dsize, bool clamp_x, std::vector<ResizeTap>& out) {
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        ResizeTap t;
        if (clamp_x) {   // horizontal taps clamp the position and zero the fraction
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
            t.ofs = (int16_t)s;
            t.pad = (int16_t)std::min(s + 1, ssize - 1);
        } else {         // vertical taps keep the fraction and clip the row indices
            t.ofs = (int16_t)std::min(std::max(s, 0), ssize - 1);
            t.pad = (int16_t)std::min(std::max(s + 1, 0), ssize - 1);
        }
        t.c0 = (int16_t)cv_round_f((1.f - f) * 2048.f);
        t.c1 = (int16_t)cv_round_f(f * 2048.f);
        out.push_back(t);
    }
}

// Level geometry for an image size.  Returns ORB_ERR_ARG for shapes the reference itself cannot
// process (a level narrower than one FAST cell divides by zero at R21 :781-784).
int build_geometry(orbx_handle_s* h, int width, int height) {
    const int L = h->prm.nlevels;
    if (width > 4096 || height > 4096) { set_error("image larger than 4096 px is not supported"); return ORB_ERR_ARG; }
    std::vector<LevelGeom> geom(L);
    std::vector<CellInfo> cells;
    std::vector<ResizeTap> xtab, ytab;
    int64_t pyr_off = 0, sp_off = 0, cand_off = 0;
    int kp_slot = 0, node_cap = 0;
    for (int l = 0; l < L; l++) {
        LevelGeom& g = geom[l];
        memset(&g, 0, sizeof(g));
        const float scale = h->isf[l];
        g.w = cv_round_f((float)width * scale);    // R21 :1112
        g.h = cv_round_f((float)height * scale);
        const int minB = kMinBorder, maxBX = g.w - kEdge + 3, maxBY = g.h - kEdge + 3;
        const float fw = (float)(maxBX - minB), fh = (float)(maxBY - minB);
        const float W = 30;
        g.n_cols = (int)(fw / W);
        g.n_rows = (int)(fh / W);
        if (g.n_cols < 1 || g.n_rows < 1) {
            set_error("level %d (%dx%d) is smaller than one 30-px FAST cell; the reference divides by zero here", l, g.w, g.h);
            return ORB_ERR_ARG;
        }
        g.w_cell = (int)ceilf(fw / g.n_cols);
        g.h_cell = (int)ceilf(fh / g.n_rows);
        g.pitch = (int)align_up(kXPad + g.w + kEdge, 64);
        g.plane_rows = g.h + 2 * kEdge;
        g.plane_off = pyr_off;
        pyr_off += align_up((int64_t)g.pitch * g.plane_rows, 256);
        g.spitch = (int)align_up(g.w, 16);
        g.splane_off = sp_off;
        sp_off += align_up((int64_t)g.spitch * g.h, 256);
        g.cell_base = (int)cells.size();
        g.cand_off = cand_off;
        int slot = 0;
        for (int i = 0; i < g.n_rows; i++) {
            const int iniY = minB + i * g.h_cell;
            int maxY = iniY + g.h_cell + 6;
            const bool skip_row = iniY >= maxBY - 3;
            if (maxY > maxBY) maxY = maxBY;
            for (int j = 0; j < g.n_cols; j++) {
                const int iniX = minB + j * g.w_cell;
                int maxX = iniX + g.w_cell + 6;
                const bool skip = skip_row || iniX >= maxBX - 6;
                if (maxX > maxBX) maxX = maxBX;
                CellInfo c;
                c.x0 = (int16_t)(iniX + 3); c.x1 = (int16_t)(maxX - 3);
                c.y0 = (int16_t)(iniY + 3); c.y1 = (int16_t)(maxY - 3);
                if (skip || c.x1 <= c.x0 || c.y1 <= c.y0) { c.x1 = c.x0; c.y1 = c.y0; }
                const int iw = c.x1 - c.x0, ih = c.y1 - c.y0;
                if (iw > 64) { set_error("FAST cell wider than 64 px"); return ORB_ERR_ARG; }
                c.slot_off = slot;
                slot += ((iw + 1) / 2) * ((ih + 1) / 2);   // strict 3x3 NMS keeps at most one pixel per 2x2 block
                cells.push_back(c);
            }
        }
        cand_off += align_up(std::max(slot, 1), 4);
        g.n_feat = h->nfeat[l];
        g.n_ini = (int)roundf((float)(maxBX - minB) / (maxBY - minB));   // R21 :543
        if (g.n_ini < 1) {
            set_error("level %d: height > 2x width gives zero quadtree roots (the reference divides by zero)", l);
            return ORB_ERR_ARG;
        }
        g.h_x = (float)(maxBX - minB) / g.n_ini;                            // R21 :545
        g.kp_cap = std::max(g.n_feat + 3, 4 * g.n_ini);
        g.kp_slot = kp_slot;
        kp_slot += g.kp_cap;
        node_cap = std::max(node_cap, g.kp_cap);
        g.scale = h->sf[l];
        g.patch_size = (float)(int)(31 * h->sf[l]);                         // R21 :836
        if (l > 0) {
            g.xtab_off = (int)xtab.size();
            g.ytab_off = (int)ytab.size();
            make_taps(geom[l - 1].w, g.w, true, xtab);
            make_taps(geom[l - 1].h, g.h, false, ytab);
            // source footprint of one output tile (for the resize kernel's shared-memory staging)
            for (int x0 = 0; x0 < g.w; x0 += kPyrTileW) {
                const int x1 = std::min(x0 + kPyrTileW, g.w) - 1;
                const int lo = xtab[g.xtab_off + x0].ofs & ~3, hi = xtab[g.xtab_off + x1].pad;
                g.rs_cols = std::max(g.rs_cols, (int)align_up(hi - lo + 1, 4));
            }
            for (int y0 = 0; y0 < g.h; y0 += kPyrTileH) {
                const int y1 = std::min(y0 + kPyrTileH, g.h) - 1;
                int lo = 1 << 30, hi = 0;   // vertical taps are monotone but take min/max anyway
                for (int y = y0; y <= y1; y++) {
                    lo = std::min(lo, (int)ytab[g.ytab_off + y].ofs);
                    hi = std::max(hi, (int)ytab[g.ytab_off + y].pad);
                }
                g.rs_rows = std::max(g.rs_rows, hi - lo + 1);
            }
            // pair-staged kernel (pyramid.cu, the default): tiles as tall as the level allows (<= 128 rows, a multiple of 8, the
            // level height split evenly) within 96 KB of staged source pairs
            {
                int ny = (g.h + 127) / 128;
                int th = std::min(128, (int)align_up((g.h + ny - 1) / ny, 8));
                g.t3_row_bytes = 2 * (g.rs_cols + 4);
                for (;; th -= 8) {
                    g.t3_rows = 0;
                    for (int y0 = 0; y0 < g.h; y0 += th) {
                        const int y1 = std::min(y0 + th, g.h) - 1;
                        int lo = 1 << 30, hi = 0;
                        for (int y = y0; y <= y1; y++) {
                            lo = std::min(lo, (int)ytab[g.ytab_off + y].ofs);
                            hi = std::max(hi, (int)ytab[g.ytab_off + y].pad);
                        }
                        g.t3_rows = std::max(g.t3_rows, hi - lo + 1);
                    }
                    g.t3_smem = g.t3_rows * g.t3_row_bytes;
                    if (g.t3_smem <= 96 * 1024 || th <= 8) break;
                }
                g.t3_h = th;
                if (g.t3_smem > 96 * 1024 || g.rs_cols > 256) g.t3_smem = -1;     // scale factors this large go to the first kernel
            }
            // two-phase kernel (pyramid.cu): tiles span the level width, split so that a tile has at most 160 four-column groups
            // and its source footprint at most 64 sixteen-byte vectors; as tall as a 40 KB block of horizontal sums allows
            {
                const int w4 = (g.w + 3) / 4;
                int nx = (w4 + 159) / 160;
                for (;; nx++) {
                    g.t2_w = (int)align_up((g.w + nx - 1) / nx, 4);
                    g.t2_nx = (g.w + g.t2_w - 1) / g.t2_w;
                    g.t2_cols = 0;
                    for (int x0 = 0; x0 < g.w; x0 += g.t2_w) {
                        const int x1 = std::min(x0 + g.t2_w, g.w) - 1;
                        const int lo = xtab[g.xtab_off + x0].ofs & ~15, hi = xtab[g.xtab_off + x1].pad;
                        g.t2_cols = std::max(g.t2_cols, (int)align_up(hi - lo + 2, 16));     // + the (zero-weight) byte behind the last pixel
                    }
                    if (g.t2_cols <= 64 * 16 || g.t2_w <= 4) break;
                }
                const int cgx = g.t2_w / 4;
                const int max_rows = std::max(6, 40960 / (cgx * 16));
                const double ratio = (double)geom[l - 1].h / g.h;
                int th = std::max(2, std::min(96, (int)((max_rows - 3) / ratio)));
                g.t2_ny = (g.h + th - 1) / th;
                g.t2_h = (g.h + g.t2_ny - 1) / g.t2_ny;
                g.t2_ny = (g.h + g.t2_h - 1) / g.t2_h;
                g.t2_rows = 0;
                for (int y0 = 0; y0 < g.h; y0 += g.t2_h) {
                    const int y1 = std::min(y0 + g.t2_h, g.h) - 1;
                    int lo = 1 << 30, hi = 0;
                    for (int y = y0; y <= y1; y++) {
                        lo = std::min(lo, (int)ytab[g.ytab_off + y].ofs);
                        hi = std::max(hi, (int)ytab[g.ytab_off + y].pad);
                    }
                    g.t2_rows = std::max(g.t2_rows, hi - lo + 1);
                }
                g.t2_smem = g.t2_rows * g.t2_cols + g.t2_rows * cgx * 16 + g.t2_w * 8 + g.t2_h * 16 + 64;
                if (g.t2_cols > 64 * 16 || cgx > 176) g.t2_smem = -1;       // scale factors this large go to the first kernel
            }
        }
    }
    if (node_cap > 60000) { set_error("nfeatures per level too large"); return ORB_ERR_ARG; }
    FrameLayout fl;
    fl.nlevels = L; fl.width = width; fl.height = height;
    fl.in_pitch = (int)align_up(width, 16);
    fl.pyr_bytes = pyr_off; fl.splane_bytes = sp_off;
    fl.n_cells = (int)cells.size();
    fl.cand_entries = cand_off;
    fl.kp_cap = kp_slot;
    fl.node_cap = (int)align_up(node_cap + 4, 32);
    fl.ini_th = h->prm.ini_th_fast;
    h->geom.swap(geom); h->cells.swap(cells); h->xtab.swap(xtab); h->ytab.swap(ytab);
    if (h->xtab.empty()) { h->xtab.push_back(ResizeTap()); h->ytab.push_back(ResizeTap()); }
    h->fl = fl;
    return ORB_OK;
}

int upload_geometry(orbx_handle_s* h) {
    int rc;
    if ((rc = grow_dev(h->d_geom, h->cap_geom, h->geom.size() * sizeof(LevelGeom)))) return rc;
    if ((rc = grow_dev(h->d_cells, h->cap_cells, h->cells.size() * sizeof(CellInfo)))) return rc;
    if ((rc = grow_dev(h->d_xtab, h->cap_xtab, h->xtab.size() * sizeof(ResizeTap)))) return rc;
    if ((rc = grow_dev(h->d_ytab, h->cap_ytab, h->ytab.size() * sizeof(ResizeTap)))) return rc;
    // synchronous copies from pageable vectors: only when the image size changes
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    ORB_CUDA_TRY(cudaMemcpy(h->d_geom, h->geom.data(), h->geom.size() * sizeof(LevelGeom), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_cells, h->cells.data(), h->cells.size() * sizeof(CellInfo), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_xtab, h->xtab.data(), h->xtab.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    ORB_CUDA_TRY(cudaMemcpy(h->d_ytab, h->ytab.data(), h->ytab.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
    return ORB_OK;
}

int ensure_size(orbx_handle_s* h, int width, int height, int n_frames) {
    int rc;
    if (width != h->cur_w || height != h->cur_h) {
        if ((rc = build_geometry(h, width, height))) return rc;
        if ((rc = upload_geometry(h))) return rc;
        h->cur_w = width; h->cur_h = height;
        if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
        h->graph_seen = orbx_handle_s::GraphKey{};
    }
    const FrameLayout& fl = h->fl;
    const size_t B = (size_t)n_frames;
    if ((rc = grow_dev(h->d_pyr, h->cap_pyr, B * fl.pyr_bytes))) return rc;
    if ((rc = grow_dev(h->d_blur, h->cap_blur, B * fl.splane_bytes))) return rc;
    if ((rc = grow_dev(h->d_score, h->cap_score, B * fl.splane_bytes + 256))) return rc;
    if ((rc = grow_dev(h->d_cand, h->cap_cand, B * fl.cand_entries * 4))) return rc;
    if ((rc = grow_dev(h->d_scratch, h->cap_scratch, B * fl.cand_entries * 4))) return rc;
    if ((rc = grow_dev(h->d_node, h->cap_node, B * fl.cand_entries * 2))) return rc;
    if ((rc = grow_dev(h->d_cell_count, h->cap_cell_count, B * fl.n_cells * 4))) return rc;
    if ((rc = grow_dev(h->d_level_raw, h->cap_level_raw, B * kMaxLevels * 4))) return rc;
    if ((rc = grow_dev(h->d_sel, h->cap_sel, B * fl.kp_cap * 4))) return rc;
    if ((rc = grow_dev(h->d_level_count, h->cap_level_count, B * kMaxLevels * 4))) return rc;
    return ORB_OK;
}

void fill_ptrs(orbx_handle_s* h, DevPtrs& d, const uint8_t* d_in, size_t in_frame_stride) {
    d.in = d_in; d.in_frame_stride = in_frame_stride; d.pyr = h->d_pyr; d.blur = h->d_blur; d.score = h->d_score; d.cand = h->d_cand;
    d.cell_count = h->d_cell_count; d.level_raw = h->d_level_raw; d.oct_scratch = h->d_scratch; d.oct_node = h->d_node; d.sel = h->d_sel;
    d.level_count = h->d_level_count; d.geom = h->d_geom; d.cells = h->d_cells; d.xtab = h->d_xtab; d.ytab = h->d_ytab;
}

// Enqueue K1..K5 for n_frames images resident at d_in (row pitch fl.in_pitch).
int enqueue_kernels(orbx_handle_s* h, const uint8_t* d_in, size_t in_frame_stride, int n_frames, orb_keypoint_t* d_kps,
                    uint8_t* d_desc, int32_t* d_counts, int cap) {
    DevPtrs d;
    fill_ptrs(h, d, d_in, in_frame_stride);
    h->last_in = d_in; h->last_in_stride = in_frame_stride; h->last_in_pitch = h->fl.in_pitch;
    cudaStream_t s = h->stream;
    const bool prof = h->profiling;
    int n;
    if (prof) cudaEventRecord(h->ev[1], s);
    if ((n = launch_pyramid(d, h->fl, h->geom.data(), n_frames, s)) < 0) {
        set_error("scale factor too large for the resize kernel's shared-memory tile");
        return ORB_ERR_ARG;
    }
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[2], s);
    if ((n = launch_fast_score(d, h->fl, h->geom.data(), n_frames, h->prm.min_th_fast, s)) < 0) return ORB_ERR_CUDA;
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[3], s);
    if ((n = launch_blur(d, h->fl, h->geom.data(), n_frames, s)) < 0) return ORB_ERR_CUDA;
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[4], s);
    if ((n = launch_fast_cells(d, h->fl, h->geom.data(), n_frames, h->prm.ini_th_fast, s)) < 0) return ORB_ERR_CUDA;
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[5], s);
    if ((n = launch_octree(d, h->fl, n_frames, s)) < 0) { set_error("quadtree kernel configuration failed"); return ORB_ERR_CUDA; }
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[6], s);
    if ((n = launch_describe(d, h->fl, h->geom.data(), n_frames, d_kps, d_desc, d_counts, cap, h->dmaps, s)) < 0) return ORB_ERR_CUDA;
    h->launches += n;
    if (prof) cudaEventRecord(h->ev[7], s);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

constexpr int kGraphMaxFrames = 4;

// enqueue_kernels, through a CUDA graph when the call is small and repeats an earlier call's signature
int enqueue_kernels_maybe_graph(orbx_handle_s* h, const uint8_t* d_in, size_t in_frame_stride, int n_frames, int width, int height,
                                orb_keypoint_t* d_kps, uint8_t* d_desc, int32_t* d_counts, int cap) {
    static const bool enabled = [] { const char* e = getenv("ORBCUDA_GRAPH"); return e ? atoi(e) != 0 : true; }();
    if (!enabled || h->profiling || n_frames > kGraphMaxFrames) return enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap);
    const orbx_handle_s::GraphKey key = {n_frames, width, height, cap, h->fl.in_pitch, d_in, d_kps, d_desc, d_counts, in_frame_stride};
    if (h->graph_exec && key == h->graph_key) {
        ORB_CUDA_TRY(cudaGraphLaunch(h->graph_exec, h->stream));
        h->launches += h->graph_kernels;
        h->graph_replays++;
        h->last_in = d_in; h->last_in_stride = in_frame_stride; h->last_in_pitch = h->fl.in_pitch;
        return ORB_OK;
    }
    if (!(key == h->graph_seen)) {       // first call with this signature: plain launches (also warms lazily loaded kernels)
        h->graph_seen = key;
        return enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap);
    }
    // second call: capture the chain, instantiate, replay
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    const int64_t before = h->launches;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap);
    }
    const int rc = enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc != ORB_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->graph_seen = orbx_handle_s::GraphKey{};
        h->launches = before;
        return rc != ORB_OK ? rc : enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap);
    }
    h->graph_kernels = (int)(h->launches - before);
    h->launches = before;
    const cudaError_t ie = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); h->graph_exec = nullptr; return enqueue_kernels(h, d_in, in_frame_stride, n_frames, d_kps, d_desc, d_counts, cap); }
    h->graph_key = key;
    ORB_CUDA_TRY(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->graph_kernels;
    h->graph_replays++;
    return ORB_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

const char* orb_last_error(void) { return g_err; }

int orb_device_count(int* count) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count) *count = n;
    return ORB_OK;
}

int orb_host_alloc(void** ptr, size_t bytes) {
    ORB_CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return ORB_OK;
}
int orb_host_free(void* ptr) {
    ORB_CUDA_TRY(cudaFreeHost(ptr));
    return ORB_OK;
}

int orbx_create(const orbx_params_t* p, int max_width, int max_height, int max_batch, int device, orbx_handle_t* out) {
    if (!p || !out || p->nlevels < 1 || p->nlevels > kMaxLevels || p->nfeatures < 1 || !(p->scale_factor > 1.0f)) {
        set_error("orbx_create: bad parameters");
        return ORB_ERR_ARG;
    }
    orbx_handle_s* h = new orbx_handle_s;
    h->dmaps = describe_maps_create();
    h->prm = *p;
    h->device = device;
    // R21 ORBextractor.cc:415-446 -- same float/double mix
    const int L = p->nlevels;
    h->scale_factor_d = p->scale_factor;
    h->sf.resize(L); h->s2.resize(L); h->isf.resize(L); h->is2.resize(L); h->nfeat.resize(L);
    h->sf[0] = 1.0f; h->s2[0] = 1.0f;
    for (int i = 1; i < L; i++) {
        h->sf[i] = (float)(h->sf[i - 1] * h->scale_factor_d);
        h->s2[i] = h->sf[i] * h->sf[i];
    }
    for (int i = 0; i < L; i++) { h->isf[i] = 1.0f / h->sf[i]; h->is2[i] = 1.0f / h->s2[i]; }
    float factor = (float)(1.0f / h->scale_factor_d);
    float ndesired = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; l++) {
        h->nfeat[l] = cv_round_f(ndesired);
        sum += h->nfeat[l];
        ndesired *= factor;
    }
    h->nfeat[L - 1] = std::max(p->nfeatures - sum, 0);
    *out = h;
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("orbx_create: no usable CUDA device %d (this library has no CPU fallback)", device);
        cudaGetLastError();
        delete h; *out = nullptr;
        return ORB_ERR_CUDA;
    }
    if (!cuda_ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete h; *out = nullptr; return ORB_ERR_CUDA; }
    for (auto& e : h->ev) if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) { orbx_destroy(h); *out = nullptr; return ORB_ERR_CUDA; }
    {
        float tab[2 * kMaxLevels] = {};
        for (int i = 0; i < L; i++) { tab[i] = h->sf[i]; tab[kMaxLevels + i] = h->isf[i]; }
        if (!cuda_ok(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming), "cudaEventCreate") ||
            !cuda_ok(cudaMalloc((void**)&h->d_tables, sizeof(tab)), "cudaMalloc") ||
            !cuda_ok(cudaMemcpy(h->d_tables, tab, sizeof(tab), cudaMemcpyHostToDevice), "cudaMemcpy")) { orbx_destroy(h); *out = nullptr; return ORB_ERR_CUDA; }
    }
    if (max_width > 0 && max_height > 0 && max_batch > 0) {
        int rc = ensure_size(h, max_width, max_height, max_batch);
        if (rc) { orbx_destroy(h); *out = nullptr; return rc; }
        h->batch_cap = max_batch;
    }
    return ORB_OK;
}

int orbx_destroy(orbx_handle_t h) {
    if (!h) return ORB_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* dev[] = {h->d_geom, h->d_cells, h->d_xtab, h->d_ytab, h->d_in, h->d_pyr, h->d_blur, h->d_score, h->d_cand,
                   h->d_scratch, h->d_node, h->d_cell_count, h->d_level_raw, h->d_sel, h->d_level_count, h->d_kps, h->d_desc, h->d_counts,
                   h->d_tables, h->d_raw};
    for (void* p : dev) if (p) cudaFree(p);
    void* host[] = {h->h_in, h->h_kps, h->h_desc, h->h_counts};
    for (void* p : host) if (p) cudaFreeHost(p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return ORB_OK;
}

int orbx_tables(orbx_handle_t h, float* sf, float* isf, float* s2, float* is2, int32_t* nfeat) {
    if (!h) return ORB_ERR_ARG;
    for (int i = 0; i < h->prm.nlevels; i++) {
        if (sf) sf[i] = h->sf[i];
        if (isf) isf[i] = h->isf[i];
        if (s2) s2[i] = h->s2[i];
        if (is2) is2[i] = h->is2[i];
        if (nfeat) nfeat[i] = h->nfeat[i];
    }
    return ORB_OK;
}

int orbx_max_keypoints(orbx_handle_t h, int width, int height, int* cap) {
    if (!h || !cap) return ORB_ERR_ARG;
    orbx_handle_s tmp;
    tmp.prm = h->prm; tmp.sf = h->sf; tmp.isf = h->isf; tmp.nfeat = h->nfeat;
    const int rc = build_geometry(&tmp, width, height);
    if (rc) return rc;
    *cap = tmp.fl.kp_cap;
    return ORB_OK;
}

int orbx_set_profiling(orbx_handle_t h, int enable) {
    if (!h) return ORB_ERR_ARG;
    h->profiling = enable != 0;
    return ORB_OK;
}

int orbx_stage_times(orbx_handle_t h, float* ms) {
    if (!h || !ms) return ORB_ERR_ARG;
    if (h->timing_pending) {
        cudaSetDevice(h->device);
        ORB_CUDA_TRY(cudaEventSynchronize(h->ev[9]));
        for (int i = 0; i < 8; i++) {
            float t = 0;
            if (cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); t = 0; }
            h->stage_ms[i] = t;
        }
        float t = 0;
        if (cudaEventElapsedTime(&t, h->ev[0], h->ev[8]) != cudaSuccess) { cudaGetLastError(); t = 0; }
        h->stage_ms[8] = t;
        h->timing_pending = false;
    }
    memcpy(ms, h->stage_ms, sizeof(h->stage_ms));
    return ORB_OK;
}

int orbx_launch_count(orbx_handle_t h, int64_t* n) {
    if (!h || !n) return ORB_ERR_ARG;
    *n = h->launches;
    return ORB_OK;
}

int orbx_stream(orbx_handle_t h, void** stream) {
    if (!h || !stream) return ORB_ERR_ARG;
    *stream = (void*)h->stream;
    return ORB_OK;
}

int orbx_extract_batch_device(orbx_handle_t h, const uint8_t* d_images, int n_frames, int width, int height,
                              size_t row_stride, size_t frame_stride, orb_keypoint_t* d_kps, uint8_t* d_desc, int cap,
                              int32_t* d_counts) {
    if (!h || !d_images || n_frames < 1 || width < 1 || height < 1 || !d_kps || !d_desc || !d_counts || cap < 1) {
        set_error("orbx_extract_batch_device: bad arguments");
        return ORB_ERR_ARG;
    }
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_size(h, width, height, n_frames))) return rc;
    const uint8_t* src = d_images;
    size_t fstride = frame_stride;
    int pitch = (int)row_stride;
    if ((row_stride & 15) || (frame_stride & 15) || (reinterpret_cast<uintptr_t>(d_images) & 15)) {
        // user layout without 16-byte aligned rows (the resize kernel stages level 0 with 16-byte loads): re-pitch into the handle's own buffer
        if ((rc = grow_dev(h->d_in, h->cap_in, (size_t)n_frames * h->fl.in_pitch * height))) return rc;
        h->launches += launch_repack(d_images, row_stride, frame_stride, h->d_in, h->fl.in_pitch, (size_t)h->fl.in_pitch * height, width, height,
                                     n_frames, h->stream);
        src = h->d_in; fstride = (size_t)h->fl.in_pitch * height; pitch = h->fl.in_pitch;
    }
    const int saved = h->fl.in_pitch;
    h->fl.in_pitch = pitch;
    if (h->profiling) { cudaEventRecord(h->ev[0], h->stream); }
    rc = enqueue_kernels(h, src, fstride, n_frames, d_kps, d_desc, d_counts, cap);
    if (h->profiling) { cudaEventRecord(h->ev[8], h->stream); cudaEventRecord(h->ev[9], h->stream); h->timing_pending = true; }
    h->fl.in_pitch = saved;
    h->last_frames = n_frames;
    return rc;
}

int orbx_extract_batch_async(orbx_handle_t h, const uint8_t* images, int n_frames, int width, int height,
                             size_t row_stride, size_t frame_stride, orb_keypoint_t* kps, uint8_t* desc, int cap,
                             int32_t* counts) {
    if (!h || n_frames < 1 || !counts || cap < 1 || !kps || !desc) { set_error("orbx_extract_batch: bad arguments"); return ORB_ERR_ARG; }
    if (h->pending) { set_error("orbx_extract_batch_async: previous call not waited for"); return ORB_ERR_ARG; }
    if (!images || width <= 0 || height <= 0) {   // R21 :1046-1047: silent return on an empty image
        for (int f = 0; f < n_frames; f++) counts[f] = 0;
        return ORB_OK;
    }
    if (row_stride < (size_t)width) { set_error("row stride smaller than width"); return ORB_ERR_ARG; }
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_size(h, width, height, n_frames))) return rc;
    const FrameLayout& fl = h->fl;
    const size_t dev_frame = (size_t)fl.in_pitch * height;
    if ((rc = grow_dev(h->d_in, h->cap_in, (size_t)n_frames * dev_frame))) return rc;
    if ((rc = grow_dev(h->d_kps, h->cap_kps, (size_t)n_frames * cap * sizeof(orb_keypoint_t)))) return rc;
    if ((rc = grow_dev(h->d_desc, h->cap_desc, (size_t)n_frames * cap * 32))) return rc;
    if ((rc = grow_dev(h->d_counts, h->cap_counts, (size_t)n_frames * 4))) return rc;
    cudaStream_t s = h->stream;
    if (h->profiling) cudaEventRecord(h->ev[0], s);
    // ---- upload: straight from the caller's buffer when it is pinned, else through pinned staging
    const bool in_pinned = is_pinned(images);
    if (in_pinned && row_stride == (size_t)fl.in_pitch && frame_stride == dev_frame) {
        ORB_CUDA_TRY(cudaMemcpyAsync(h->d_in, images, (size_t)n_frames * dev_frame, cudaMemcpyHostToDevice, s));
    } else if (in_pinned) {
        // caller's own strides (e.g. a tight 1241-byte row): ONE copy of the spanned block, rows re-pitched on the device
        // (a cudaMemcpy2DAsync per frame from an unaligned pitch ran at a fifth of the link rate)
        const size_t span = (size_t)(n_frames - 1) * frame_stride + (size_t)(height - 1) * row_stride + width;
        if (frame_stride >= row_stride * (size_t)(height - 1) + width && span <= 2 * (size_t)n_frames * width * height + 4096) {
            if ((rc = grow_dev(h->d_raw, h->cap_raw, span))) return rc;
            ORB_CUDA_TRY(cudaMemcpyAsync(h->d_raw, images, span, cudaMemcpyHostToDevice, s));
            h->launches += launch_repack(h->d_raw, row_stride, frame_stride, h->d_in, fl.in_pitch, dev_frame, width, height, n_frames, s);
        } else {
            for (int f = 0; f < n_frames; f++)
                ORB_CUDA_TRY(cudaMemcpy2DAsync(h->d_in + f * dev_frame, fl.in_pitch, images + f * frame_stride, row_stride, width,
                                               height, cudaMemcpyHostToDevice, s));
        }
    } else {
        if ((rc = grow_host(h->h_in, h->cap_h_in, (size_t)n_frames * dev_frame))) return rc;
        for (int f = 0; f < n_frames; f++)
            for (int y = 0; y < height; y++)
                memcpy(h->h_in + f * dev_frame + (size_t)y * fl.in_pitch, images + f * frame_stride + (size_t)y * row_stride, width);
        ORB_CUDA_TRY(cudaMemcpyAsync(h->d_in, h->h_in, (size_t)n_frames * dev_frame, cudaMemcpyHostToDevice, s));
    }
    if ((rc = enqueue_kernels_maybe_graph(h, h->d_in, dev_frame, n_frames, width, height, h->d_kps, h->d_desc, h->d_counts, cap))) {
        cudaStreamSynchronize(h->stream);   // the upload may still be reading the pinned staging buffer
        return rc;
    }
    // ---- download
    const size_t kb = (size_t)n_frames * cap * sizeof(orb_keypoint_t), db = (size_t)n_frames * cap * 32;
    const bool direct = is_pinned(kps) && is_pinned(desc) && is_pinned(counts);
    if (direct) {
        ORB_CUDA_TRY(cudaMemcpyAsync(counts, h->d_counts, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA_TRY(cudaMemcpyAsync(kps, h->d_kps, kb, cudaMemcpyDeviceToHost, s));
        ORB_CUDA_TRY(cudaMemcpyAsync(desc, h->d_desc, db, cudaMemcpyDeviceToHost, s));
    } else {
        if ((rc = grow_host(h->h_kps, h->cap_h_kps, kb))) return rc;
        if ((rc = grow_host(h->h_desc, h->cap_h_desc, db))) return rc;
        if ((rc = grow_host(h->h_counts, h->cap_h_counts, (size_t)n_frames * 4))) return rc;
        ORB_CUDA_TRY(cudaMemcpyAsync(h->h_counts, h->d_counts, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA_TRY(cudaMemcpyAsync(h->h_kps, h->d_kps, kb, cudaMemcpyDeviceToHost, s));
        ORB_CUDA_TRY(cudaMemcpyAsync(h->h_desc, h->d_desc, db, cudaMemcpyDeviceToHost, s));
    }
    if (h->profiling) { cudaEventRecord(h->ev[8], s); cudaEventRecord(h->ev[9], s); h->timing_pending = true; }
    h->pending = true; h->p_direct = direct; h->p_frames = n_frames; h->p_cap = cap;
    h->p_kps = kps; h->p_desc = desc; h->p_counts = counts;
    h->last_frames = n_frames;
    return ORB_OK;
}

int orbx_wait(orbx_handle_t h) {
    if (!h) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    const bool was_pending = h->pending;
    h->pending = false;     // also when the synchronisation fails: the handle must not stay wedged ("previous call not waited for")
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!was_pending) return ORB_OK;
    int rc = ORB_OK;
    const int cap = h->p_cap;
    const int32_t* cnt = h->p_direct ? h->p_counts : h->h_counts;
    for (int f = 0; f < h->p_frames; f++) {
        int n = cnt[f];
        if (n > cap) { set_error("frame %d produced %d key points but cap is %d", f, n, cap); rc = ORB_ERR_CAPACITY; n = cap; }
        if (!h->p_direct) {
            memcpy(h->p_kps + (size_t)f * cap, h->h_kps + (size_t)f * cap, (size_t)n * sizeof(orb_keypoint_t));
            memcpy(h->p_desc + (size_t)f * cap * 32, h->h_desc + (size_t)f * cap * 32, (size_t)n * 32);
            h->p_counts[f] = cnt[f];
        }
    }
    return rc;
}

int orbx_extract_batch(orbx_handle_t h, const uint8_t* images, int n_frames, int width, int height, size_t row_stride,
                       size_t frame_stride, orb_keypoint_t* kps, uint8_t* desc, int cap, int32_t* counts) {
    const int rc = orbx_extract_batch_async(h, images, n_frames, width, height, row_stride, frame_stride, kps, desc, cap, counts);
    if (rc) return rc;
    return orbx_wait(h);
}

int orbx_extract(orbx_handle_t h, const uint8_t* image, int width, int height, size_t stride, orb_keypoint_t* kps,
                 uint8_t* desc, int cap, int* n) {
    int32_t cnt = 0;
    const int rc = orbx_extract_batch(h, image, 1, width, height, stride, stride * (size_t)std::max(height, 0), kps, desc, cap, &cnt);
    if (n) *n = cnt;
    return rc;
}

// ---- views of intermediate results --------------------------------------------------------------
int orbx_level_size(orbx_handle_t h, int level, int* w, int* hh) {
    if (!h || level < 0 || level >= (int)h->geom.size()) return ORB_ERR_ARG;
    *w = h->geom[level].w; *hh = h->geom[level].h;
    return ORB_OK;
}

int orbx_download_level(orbx_handle_t h, int frame, int level, int with_border, uint8_t* dst, size_t dst_stride) {
    if (!h || level < 0 || level >= (int)h->geom.size() || frame < 0 || frame >= h->last_frames || !dst) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const LevelGeom& g = h->geom[level];
    const int b = with_border ? kEdge : 0;
    uint8_t* roi = dst + (size_t)b * dst_stride + b;
    if (level == 0) {
        ORB_CUDA_TRY(cudaMemcpy2D(roi, dst_stride, h->last_in + (size_t)frame * h->last_in_stride, h->last_in_pitch, g.w, g.h,
                                  cudaMemcpyDeviceToHost));
    } else {
        const uint8_t* src = h->d_pyr + (size_t)frame * h->fl.pyr_bytes + g.plane_off + (size_t)kEdge * g.pitch + kXPad;
        ORB_CUDA_TRY(cudaMemcpy2D(roi, dst_stride, src, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    }
    if (b) {
        // the REFLECT_101 border of cv::copyMakeBorder (R21 :1122-1128) is rebuilt here: the device never needs it
        auto refl = [](int p, int len) { if (p < 0) p = -p; if (p >= len) p = 2 * len - 2 - p; return p; };
        for (int y = 0; y < g.h; y++) {
            uint8_t* row = roi + (size_t)y * dst_stride;
            for (int k = 1; k <= b; k++) { row[-k] = row[refl(-k, g.w)]; row[g.w - 1 + k] = row[refl(g.w - 1 + k, g.w)]; }
        }
        for (int k = 1; k <= b; k++) {
            memcpy(roi + (ptrdiff_t)(-k) * (ptrdiff_t)dst_stride - b, roi + (size_t)refl(-k, g.h) * dst_stride - b, g.w + 2 * b);
            memcpy(roi + (size_t)(g.h - 1 + k) * dst_stride - b, roi + (size_t)refl(g.h - 1 + k, g.h) * dst_stride - b, g.w + 2 * b);
        }
    }
    return ORB_OK;
}

// All levels of one frame in one go (the mvImagePyramid mirror of the C++ class): the copies are queued back to back on the
// handle's stream -- truly asynchronous when the destination planes are page-locked (orb_host_alloc) -- and waited for once;
// the REFLECT_101 borders are rebuilt on the host as in orbx_download_level.
int orbx_download_pyramid(orbx_handle_t h, int frame, int with_border, uint8_t* const* dst, const size_t* dst_stride) {
    if (!h || !dst || !dst_stride || frame < 0 || frame >= h->last_frames) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    const int b = with_border ? kEdge : 0;
    const int L = (int)h->geom.size();
    for (int level = 0; level < L; level++) {
        const LevelGeom& g = h->geom[level];
        if (!dst[level]) return ORB_ERR_ARG;
        uint8_t* roi = dst[level] + (size_t)b * dst_stride[level] + b;
        if (level == 0)
            ORB_CUDA_TRY(cudaMemcpy2DAsync(roi, dst_stride[level], h->last_in + (size_t)frame * h->last_in_stride, h->last_in_pitch, g.w, g.h,
                                           cudaMemcpyDeviceToHost, h->stream));
        else
            ORB_CUDA_TRY(cudaMemcpy2DAsync(roi, dst_stride[level], h->d_pyr + (size_t)frame * h->fl.pyr_bytes + g.plane_off + (size_t)kEdge * g.pitch + kXPad,
                                           g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, h->stream));
    }
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (b) {
        auto refl = [](int p, int len) { if (p < 0) p = -p; if (p >= len) p = 2 * len - 2 - p; return p; };
        for (int level = 0; level < L; level++) {
            const LevelGeom& g = h->geom[level];
            const size_t st = dst_stride[level];
            uint8_t* roi = dst[level] + (size_t)b * st + b;
            for (int y = 0; y < g.h; y++) {
                uint8_t* row = roi + (size_t)y * st;
                for (int k = 1; k <= b; k++) { row[-k] = row[refl(-k, g.w)]; row[g.w - 1 + k] = row[refl(g.w - 1 + k, g.w)]; }
            }
            for (int k = 1; k <= b; k++) {
                memcpy(roi + (ptrdiff_t)(-k) * (ptrdiff_t)st - b, roi + (size_t)refl(-k, g.h) * st - b, g.w + 2 * b);
                memcpy(roi + (size_t)(g.h - 1 + k) * st - b, roi + (size_t)refl(g.h - 1 + k, g.h) * st - b, g.w + 2 * b);
            }
        }
    }
    return ORB_OK;
}

static int download_splane(orbx_handle_t h, const uint8_t* base, int frame, int level, uint8_t* dst, size_t dst_stride) {
    if (!h || level < 0 || level >= (int)h->geom.size() || frame < 0 || frame >= h->last_frames || !dst) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const LevelGeom& g = h->geom[level];
    ORB_CUDA_TRY(cudaMemcpy2D(dst, dst_stride, base + (size_t)frame * h->fl.splane_bytes + g.splane_off, g.spitch, g.w, g.h,
                              cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbx_download_blurred(orbx_handle_t h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    return download_splane(h, h ? h->d_blur : nullptr, frame, level, dst, dst_stride);
}
int orbx_download_scores(orbx_handle_t h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    return download_splane(h, h ? h->d_score : nullptr, frame, level, dst, dst_stride);
}

int orbx_download_candidates(orbx_handle_t h, int frame, int level, int16_t* x, int16_t* y, uint8_t* score, int cap, int* n) {
    if (!h || level < 0 || level >= (int)h->geom.size() || frame < 0 || frame >= h->last_frames || !n) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(h->device));
    ORB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const LevelGeom& g = h->geom[level];
    const int ncell = g.n_cols * g.n_rows;
    std::vector<int32_t> flag(ncell);
    ORB_CUDA_TRY(cudaMemcpy(flag.data(), h->d_cell_count + (size_t)frame * h->fl.n_cells + g.cell_base, ncell * 4, cudaMemcpyDeviceToHost));
    int32_t nraw = 0;
    ORB_CUDA_TRY(cudaMemcpy(&nraw, h->d_level_raw + (size_t)frame * kMaxLevels + level, 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> buf(std::max(nraw, 1));
    ORB_CUDA_TRY(cudaMemcpy(buf.data(), h->d_cand + (size_t)frame * h->fl.cand_entries + g.cand_off, (size_t)nraw * 4, cudaMemcpyDeviceToHost));
    // same rule as the quadtree gather (iniTh survivors where the cell has any, else all), then the
    // reference's order: cells row-major, raster inside a cell
    std::vector<std::pair<uint32_t, uint32_t>> keyed;
    for (int i = 0; i < nraw; i++) {
        const uint32_t v = buf[i];
        const int xr = (int)(v & 0xfff) - 3, yr = (int)((v >> 12) & 0xfff) - 3;
        const int cx = xr / g.w_cell, cy = yr / g.h_cell;
        const int cell = cy * g.n_cols + cx;
        if (flag[cell] && (int)(v >> 24) < h->prm.ini_th_fast) continue;
        keyed.push_back(std::make_pair(((uint32_t)cell << 12) | (uint32_t)((yr - cy * g.h_cell) * g.w_cell + (xr - cx * g.w_cell)), v));
    }
    std::sort(keyed.begin(), keyed.end());
    int k = 0;
    for (size_t i = 0; i < keyed.size(); i++, k++) {
        const uint32_t v = keyed[i].second;
        if (k < cap) { x[k] = (int16_t)(v & 0xfff); y[k] = (int16_t)((v >> 12) & 0xfff); score[k] = (uint8_t)(v >> 24); }
    }
    *n = k;
    return ORB_OK;
}

int orbx_distribute_octtree(const int16_t* x, const int16_t* y, const uint8_t* score, int n, int min_x, int max_x, int min_y,
                            int max_y, int n_features, int32_t* out_index, int cap, int* n_out, int device) {
    if (!n_out || n < 0 || (n > 0 && (!x || !y || !score))) return ORB_ERR_ARG;
    *n_out = 0;
    if (n == 0) return ORB_OK;
    const int width = max_x - min_x, height = max_y - min_y;
    if (width <= 0 || height <= 0 || width > 4095 || height > 4095 || n >= (1 << 24)) return ORB_ERR_ARG;
    const int n_ini = (int)roundf((float)width / height);
    if (n_ini < 1) { set_error("zero quadtree roots"); return ORB_ERR_ARG; }
    const float h_x = (float)width / n_ini;
    const int kp_cap = std::max(n_features + 3, 4 * n_ini);
    const int node_cap = (int)align_up(kp_cap + 4, 32);
    if (node_cap > 60000) return ORB_ERR_ARG;
    ORB_CUDA_TRY(cudaSetDevice(device));
    std::vector<uint32_t> packed(n);
    for (int i = 0; i < n; i++) packed[i] = (uint32_t)(x[i] & 0xfff) | ((uint32_t)(y[i] & 0xfff) << 12) | ((uint32_t)score[i] << 24);
    uint32_t *d_kp = nullptr, *d_sel = nullptr; uint16_t* d_node = nullptr; int32_t *d_cnt = nullptr, *d_idx = nullptr;
    int rc = ORB_OK;
    std::vector<int32_t> sel(kp_cap);
    int32_t cnt = 0;
    do {
        if (!cuda_ok(cudaMalloc((void**)&d_kp, (size_t)n * 4), "cudaMalloc") || !cuda_ok(cudaMalloc((void**)&d_sel, (size_t)kp_cap * 4), "cudaMalloc") ||
            !cuda_ok(cudaMalloc((void**)&d_node, (size_t)n * 2), "cudaMalloc") || !cuda_ok(cudaMalloc((void**)&d_cnt, 4), "cudaMalloc") ||
            !cuda_ok(cudaMalloc((void**)&d_idx, (size_t)kp_cap * 4), "cudaMalloc")) { rc = ORB_ERR_CUDA; break; }
        if (!cuda_ok(cudaMemcpy(d_kp, packed.data(), (size_t)n * 4, cudaMemcpyHostToDevice), "cudaMemcpy")) { rc = ORB_ERR_CUDA; break; }
        if (launch_octree_single(d_kp, n, width, height, n_features, n_ini, h_x, d_idx, d_node, d_sel, d_cnt, kp_cap, node_cap, 0) < 0) { rc = ORB_ERR_CUDA; break; }
        if (!cuda_ok(cudaDeviceSynchronize(), "octree kernel")) { rc = ORB_ERR_CUDA; break; }
        if (!cuda_ok(cudaMemcpy(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost), "cudaMemcpy") ||
            !cuda_ok(cudaMemcpy(sel.data(), d_idx, (size_t)kp_cap * 4, cudaMemcpyDeviceToHost), "cudaMemcpy")) { rc = ORB_ERR_CUDA; break; }
    } while (0);
    cudaFree(d_kp); cudaFree(d_sel); cudaFree(d_node); cudaFree(d_cnt); cudaFree(d_idx);
    if (rc) return rc;
    *n_out = cnt;
    for (int i = 0; i < cnt && i < cap; i++) out_index[i] = sel[i];
    return cnt > cap ? ORB_ERR_CAPACITY : ORB_OK;
}

}  // extern "C"

// Internal (not in orbcuda.h): device views of a handle's last extraction for orbm_stereo_matches.
extern "C" int orbx_internal_view(orbx_handle_t h, const uint8_t** d_pyr, const orbcuda::LevelGeom** d_geom,
                                  const orbcuda::LevelGeom** h_geom, orbcuda::FrameLayout* fl, int* device,
                                  const float** sf, const float** isf, const uint8_t** d_level0, int* level0_pitch) {
    if (!h || h->last_frames < 1) return ORB_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) return ORB_ERR_CUDA;
    *d_pyr = h->d_pyr; *d_geom = h->d_geom; *h_geom = h->geom.data(); *fl = h->fl; *device = h->device;
    *sf = h->sf.data(); *isf = h->isf.data();
    *d_level0 = h->last_in; *level0_pitch = h->last_in_pitch;
    return ORB_OK;
}

// Internal (not in orbcuda.h): the same for a batch, stream-ordered instead of synchronising -- `consumer` is made to
// wait for everything enqueued on the handle's stream so far.  d_tables = [mvScaleFactor | mvInvScaleFactor] on the device.
extern "C" int orbx_internal_view_batch(orbx_handle_t h, const uint8_t** d_pyr, const orbcuda::LevelGeom** d_geom,
                                        const orbcuda::LevelGeom** h_geom, orbcuda::FrameLayout* fl, int* device,
                                        const float** d_tables, const uint8_t** d_level0, int* level0_pitch,
                                        size_t* level0_stride, int* n_frames, void* consumer) {
    if (!h || h->last_frames < 1) return ORB_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return ORB_ERR_CUDA;
    if ((cudaStream_t)consumer != h->stream) {
        if (cudaEventRecord(h->ev_done, h->stream) != cudaSuccess || cudaStreamWaitEvent((cudaStream_t)consumer, h->ev_done, 0) != cudaSuccess)
            return ORB_ERR_CUDA;
    }
    *d_pyr = h->d_pyr; *d_geom = h->d_geom; *h_geom = h->geom.data(); *fl = h->fl; *device = h->device;
    *d_tables = h->d_tables; *d_level0 = h->last_in; *level0_pitch = h->last_in_pitch; *level0_stride = h->last_in_stride;
    *n_frames = h->last_frames;
    return ORB_OK;
}
