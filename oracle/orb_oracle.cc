// orb_oracle.cc -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see orb_oracle.h).
//
// Scalar restatement of the reference hot path.  R21 = /root/reference/ORB_SLAM2.1.
// Every function cites the reference lines (or the OpenCV routine the reference calls)
// that it follows.  Build with -ffp-contract=off.
#include "orb_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace {

const int kEdge = 19;        // EDGE_THRESHOLD   R21/src/ORBextractor.cc:74
const int kHalfPatch = 15;   // HALF_PATCH_SIZE  :73
const int kPatch = 31;       // PATCH_SIZE       :72

const int8_t kPattern[1024] = {
#include "../cooperative-orb-slam_b200/csrc/orb_pattern.inc"
};

// cvRound(float): SSE cvtss2si == round-half-to-even in the default rounding mode.
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(float v) { return (int)floorf(v); }

inline int reflect101(int p, int len) {
    // cv::borderInterpolate(p, len, BORDER_REFLECT_101)
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

}  // namespace

extern "C" int orc_cv_round_f(float v) { return cv_round(v); }

// ----------------------------------------------------------------------------------------------
// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 -- called at R21 ORBextractor.cc:1120.
// OpenCV imgproc resize.cpp: coefficient tables in float -> 11-bit fixed point (INTER_RESIZE_COEF_BITS),
// HResizeLinear (int32 rows), VResizeLinear<uchar> fixed-point pack.  SURVEY.md App. A.1.
// ----------------------------------------------------------------------------------------------
extern "C" void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst,
                                     int dw, int dh, size_t dstride) {
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = (short)cv_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
        beta[2 * dy + 1] = (short)cv_round(fy * 2048.f);
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; dy++) {
        const int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        const int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            const int sx = xofs[dx];
            const int sx1 = std::min(sx + 1, sw - 1);
            row0[dx] = S0[sx] * alpha[2 * dx] + S0[sx1] * alpha[2 * dx + 1];
            row1[dx] = S1[sx] * alpha[2 * dx] + S1[sx1] * alpha[2 * dx + 1];
        }
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101 [| BORDER_ISOLATED]) -- R21 ORBextractor.cc:1122-1128.
extern "C" void orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstride,
                                                uint8_t* dst, size_t dstride, int top, int bottom,
                                                int left, int right) {
    const int dw = w + left + right, dh = h + top + bottom;
    std::vector<int> xmap(dw);
    for (int x = 0; x < dw; x++) xmap[x] = reflect101(x - left, w);
    std::vector<uint8_t> line(dw);
    for (int y = 0; y < dh; y++) {
        const uint8_t* S = src + (size_t)reflect101(y - top, h) * sstride;
        for (int x = 0; x < dw; x++) line[x] = S[xmap[x]];  // via a line buffer: src may alias dst's ROI
        memcpy(dst + (size_t)y * dstride, line.data(), dw);
    }
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 -- R21 :1085-1086.
// OpenCV's fixed-point smoothing path: 8.8 kernel, single rounding.  SURVEY.md App. A.2.
extern "C" void orc_gaussian_blur7_sigma2(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst,
                                          size_t dstride) {
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<int> xm(w + 6);
    for (int x = 0; x < w + 6; x++) xm[x] = reflect101(x - 3, w);
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        uint16_t* H = hbuf.data() + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int k = 0; k < 7; k++) s += K[k] * S[xm[x + k]];
            H[x] = (uint16_t)s;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; k++) R[k] = hbuf.data() + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int k = 0; k < 7; k++) s += (uint32_t)K[k] * R[k][x];
            D[x] = (uint8_t)((s + 32768u) >> 16);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression) == FAST_t<16> (TYPE_9_16) -- R21 :809-815.
// SURVEY.md App. A.3.
// ----------------------------------------------------------------------------------------------
namespace {
const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                            {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// max over the 16 arcs of 9 contiguous circle pixels of max(min d, -max d), d = I(p) - I(circle).
// Arcs are visited in pairs sharing their 8 middle pixels, with the running best as an early-out bound --
// the evaluation order of cv::cornerScore<16>.  floor: results below it are reported as `floor`.
inline int fast_best_from(const uint8_t* c, size_t stride, int floor_) {
    int d[25];
    const int v = c[0];
    for (int k = 0; k < 16; k++) d[k] = v - c[(ptrdiff_t)kCircle[k][1] * (ptrdiff_t)stride + kCircle[k][0]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int a0 = floor_;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], std::min(d[k + 2], d[k + 3]));
        if (a <= a0) continue;
        a = std::min(a, std::min(std::min(d[k + 4], d[k + 5]), std::min(d[k + 6], std::min(d[k + 7], d[k + 8]))));
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], std::max(d[k + 2], d[k + 3]));
        if (b >= b0) continue;
        b = std::max(b, std::max(std::max(d[k + 4], d[k + 5]), std::max(d[k + 6], std::max(d[k + 7], d[k + 8]))));
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0;
}
inline int fast_best(const uint8_t* c, size_t stride) { return fast_best_from(c, stride, -256); }
}  // namespace

// cornerScore<16>() of a pixel that is a corner at `threshold`; 0 otherwise.
extern "C" int orc_fast_score(const uint8_t* center, size_t stride, int threshold) {
    const int best = fast_best(center, stride);
    return best > threshold ? best - 1 : 0;
}

extern "C" int orc_fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold, int nms,
                            orc_keypoint* out, int cap) {
    if (w < 7 || h < 7) return 0;
    threshold = std::min(std::max(threshold, 0), 255);
    // Same pruning as cv::FAST_t<16>: a 9-arc must contain one pixel of every opposite pair (k, k+8), so a
    // pixel is dropped as soon as some pair has neither member brighter (resp. darker) than the centre by
    // more than the threshold.  Only survivors get the exact arc test / cornerScore.  (Keeps the CPU
    // baseline honest: the full arc scan on every pixel would be ~5x slower than OpenCV's own detector.)
    ptrdiff_t off[16];
    for (int k = 0; k < 16; k++) off[k] = (ptrdiff_t)kCircle[k][1] * (ptrdiff_t)stride + kCircle[k][0];
    std::vector<uint8_t> score((size_t)w * h, 0);
    std::vector<uint8_t> corner((size_t)w * h, 0);
    const int iw = w - 6;
    std::vector<uint8_t> lo(iw), hi(iw), dk(iw), br(iw);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* c = img + (size_t)y * stride + 3;
        // byte-wise loops written so that the compiler vectorises them (32 pixels per AVX2 op)
        for (int x = 0; x < iw; x++) {
            const int v = c[x];
            lo[x] = (uint8_t)std::max(v - threshold, 0);     // a < lo  <=>  a < v - th
            hi[x] = (uint8_t)std::min(v + threshold, 255);   // a > hi  <=>  a > v + th
            dk[x] = 0xff; br[x] = 0xff;
        }
        for (int k = 0; k < 8; k++) {
            const uint8_t* a = c + off[k];
            const uint8_t* b = c + off[k + 8];
            for (int x = 0; x < iw; x++) {
                dk[x] &= (uint8_t)(-(int)((a[x] < lo[x]) | (b[x] < lo[x])));
                br[x] &= (uint8_t)(-(int)((a[x] > hi[x]) | (b[x] > hi[x])));
            }
        }
        for (int x = 0; x < iw; x++) {
            if (!(dk[x] | br[x])) continue;
            const int best = fast_best_from(c + x, stride, threshold);
            if (best > threshold) {
                corner[(size_t)y * w + x + 3] = 1;
                // without NMS OpenCV never fills the score buffer -> response 0
                score[(size_t)y * w + x + 3] = nms ? (uint8_t)(best - 1) : 0;
            }
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            const int s = score[(size_t)y * w + x];
            bool keep = true;
            if (nms) {
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (s <= score[(size_t)(y + dy) * w + (x + dx)]) { keep = false; break; }
                    }
            }
            if (!keep) continue;
            if (n < cap) {
                orc_keypoint kp;
                kp.x = (float)x; kp.y = (float)y; kp.size = 7.f; kp.angle = -1.f;
                kp.response = (float)s; kp.octave = 0; kp.class_id = -1;
                out[n] = kp;
            }
            n++;
        }
    return n;
}

// cv::fastAtan2(y, x) scalar path (OpenCV core mathfuncs_core: atan_f32) -- R21 :103.  App. A.5.
extern "C" float orc_fast_atan2(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ----------------------------------------------------------------------------------------------
// Extractor
// ----------------------------------------------------------------------------------------------
namespace {

struct Cand { int16_t x, y; uint8_t score; };

struct Level {
    int w = 0, h = 0;
    size_t pstride = 0;                 // padded stride = w + 38
    std::vector<uint8_t> padded;        // (w+38) x (h+38); mvImagePyramid[l] is the ROI at (+19,+19)
    std::vector<uint8_t> blurred;       // w x h
    bool has_blur = false;
    std::vector<Cand> cands;            // vToDistributeKeys
    std::vector<orc_keypoint> kps;      // allKeypoints[level] after orientation
    const uint8_t* roi() const { return padded.data() + (size_t)kEdge * pstride + kEdge; }
};

}  // namespace

struct orc_extractor {
    int nfeatures, nlevels, ini_th, min_th, trig_mode, fma_mode;
    double scale_factor;  // R21/include/ORBextractor.h:98 -- a double member set from a float
    std::vector<float> sf, isf, s2, is2;
    std::vector<int> nfeat;
    int umax[16];
    std::vector<Level> lv;
};

// R21 ORBextractor.cc:410-470
extern "C" orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th,
                                               int min_th, int trig_mode, int fma_mode) {
    if (nlevels < 1 || nfeatures < 1) return nullptr;
    orc_extractor* e = new orc_extractor;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = ini_th; e->min_th = min_th;
    e->trig_mode = trig_mode; e->fma_mode = fma_mode;
    e->scale_factor = scale_factor;
    e->sf.resize(nlevels); e->s2.resize(nlevels); e->isf.resize(nlevels); e->is2.resize(nlevels);
    e->sf[0] = 1.0f; e->s2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->sf[i] = (float)(e->sf[i - 1] * e->scale_factor);   // float*double -> double -> float  (:421)
        e->s2[i] = e->sf[i] * e->sf[i];
    }
    for (int i = 0; i < nlevels; i++) { e->isf[i] = 1.0f / e->sf[i]; e->is2[i] = 1.0f / e->s2[i]; }
    e->nfeat.resize(nlevels);
    float factor = (float)(1.0f / e->scale_factor);           // :436
    float ndesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->nfeat[l] = cv_round(ndesired);
        sum += e->nfeat[l];
        ndesired *= factor;
    }
    e->nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax (:452-469)
    int v, v0;
    const int vmax = cv_floor((float)(kHalfPatch * sqrtf(2.f) / 2 + 1));
    const int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round(sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->lv.resize(nlevels);
    return e;
}

extern "C" void orc_extractor_destroy(orc_extractor* e) { delete e; }

extern "C" void orc_extractor_tables(const orc_extractor* e, float* sf, float* isf, float* s2, float* is2,
                                     int* nfeat, int* umax16) {
    for (int i = 0; i < e->nlevels; i++) {
        if (sf) sf[i] = e->sf[i];
        if (isf) isf[i] = e->isf[i];
        if (s2) s2[i] = e->s2[i];
        if (is2) is2[i] = e->is2[i];
        if (nfeat) nfeat[i] = e->nfeat[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

// IC_Angle R21 :77-104 (umax is a function of HALF_PATCH_SIZE only; recomputed tables are identical).
static float ic_angle_umax(const uint8_t* center, size_t stride, const int* umax) {
    int m_01 = 0, m_10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m_10 += u * center[u];
    const ptrdiff_t step = (ptrdiff_t)stride;
    for (int v = 1; v <= kHalfPatch; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

extern "C" float orc_ic_angle(const uint8_t* center, size_t stride) {
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    return ic_angle_umax(center, stride, umax);
}

// computeOrbDescriptor R21 :108-147.  App. A.6.
extern "C" void orc_orb_descriptor(const uint8_t* center, size_t stride, float angle_deg, int trig_mode,
                                   int fma_mode, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    const float angle = angle_deg * factorPI;
    float a, b;
    if (trig_mode == 0) { a = cosf(angle); b = sinf(angle); }
    else { a = (float)cos((double)angle); b = (float)sin((double)angle); }
    const int step = (int)stride;
    const int8_t* p = kPattern;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int j = 0; j < 8; ++j, p += 4) {
            int t[2];
            for (int s = 0; s < 2; ++s) {
                const float px = (float)p[2 * s], py = (float)p[2 * s + 1];
                float fr, fc;
                if (fma_mode == 0) {
                    const float xb = px * b, ya = py * a, xa = px * a, yb = py * b;
                    fr = xb + ya;
                    fc = xa - yb;
                } else {
                    fr = fmaf(px, b, py * a);
                    fc = fmaf(px, a, -(py * b));
                }
                t[s] = center[cv_round(fr) * step + cv_round(fc)];
            }
            val |= (t[0] < t[1]) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

// ----------------------------------------------------------------------------------------------
// DistributeOctTree (R21 :539-763) + ExtractorNode::DivideNode (:481-537), restated
// level-synchronously.  Node ids are creation order.  The reference keeps nodes in a std::list
// where every new node is push_front'ed, so list order == descending creation id, with the
// initial (push_back'ed) root nodes last in ascending order.  The phase-2 sort (:684) orders by
// (size, node pointer); the reference's pointer order is allocator dependent, the
// canonical rule here is "later created == larger pointer".  SURVEY.md App. A.4.
// ----------------------------------------------------------------------------------------------
namespace {

struct QNode {
    int ulx, uly, brx, bry;
    std::vector<int> keys;  // candidate indices, input order
    bool alive = true;
};

struct QTree {
    const int16_t* x;
    const int16_t* y;
    std::vector<QNode> nodes;
    int n_root = 0;
    int alive = 0;

    // DivideNode: returns ids of the non-empty children in creation order n1..n4
    void divide(int id, std::vector<int>& to_expand) {
        const QNode P = nodes[id];
        const int halfX = (int)ceilf((float)(P.brx - P.ulx) / 2);
        const int halfY = (int)ceilf((float)(P.bry - P.uly) / 2);
        const int mx = P.ulx + halfX, my = P.uly + halfY;
        QNode c[4];
        c[0].ulx = P.ulx; c[0].uly = P.uly; c[0].brx = mx;    c[0].bry = my;      // n1 UL
        c[1].ulx = mx;    c[1].uly = P.uly; c[1].brx = P.brx; c[1].bry = my;      // n2 UR
        c[2].ulx = P.ulx; c[2].uly = my;    c[2].brx = mx;    c[2].bry = P.bry;   // n3 BL
        c[3].ulx = mx;    c[3].uly = my;    c[3].brx = P.brx; c[3].bry = P.bry;   // n4 BR
        for (int k : P.keys) {
            const int q = (x[k] < mx ? 0 : 1) + (y[k] < my ? 0 : 2);
            c[q].keys.push_back(k);
        }
        nodes[id].alive = false;
        nodes[id].keys.clear();
        alive--;
        for (int q = 0; q < 4; q++) {
            if (c[q].keys.empty()) continue;
            nodes.push_back(c[q]);
            alive++;
            if (c[q].keys.size() > 1) to_expand.push_back((int)nodes.size() - 1);
        }
    }
    // list order: non-root nodes descending id, then roots ascending id
    template <class F> void for_each_in_list_order(F f) const {
        for (int i = (int)nodes.size() - 1; i >= n_root; i--) if (nodes[i].alive) f(i);
        for (int i = 0; i < n_root; i++) if (nodes[i].alive) f(i);
    }
};

int distribute_octtree(const int16_t* x, const int16_t* y, const uint8_t* score, int n, int min_x, int max_x,
                       int min_y, int max_y, int N, std::vector<int>& result) {
    result.clear();
    const int nIni = (int)roundf((float)(max_x - min_x) / (max_y - min_y));   // :543
    if (nIni < 1) return -1;   // the reference divides by zero here
    const float hX = (float)(max_x - min_x) / nIni;                           // :545
    QTree t;
    t.x = x; t.y = y;
    t.nodes.resize(nIni);
    t.n_root = nIni;
    for (int i = 0; i < nIni; i++) {
        QNode& r = t.nodes[i];
        r.ulx = (int)(hX * (float)i);
        r.brx = (int)(hX * (float)(i + 1));
        r.uly = 0;
        r.bry = max_y - min_y;
    }
    for (int k = 0; k < n; k++) {
        const int r = (int)((float)x[k] / hX);                                // :569
        if (r < 0 || r >= nIni) return -2;
        t.nodes[r].keys.push_back(k);
    }
    t.alive = 0;
    for (int i = 0; i < nIni; i++) {
        if (t.nodes[i].keys.empty()) t.nodes[i].alive = false;                // :581-582
        else t.alive++;
    }
    std::vector<int> expand;       // vSizeAndPointerToNode (ids of nodes with >1 key)
    bool finish = false;
    while (!finish) {
        const int prev = t.alive;
        std::vector<int> todo;
        t.for_each_in_list_order([&](int i) { if (t.nodes[i].keys.size() > 1) todo.push_back(i); });
        expand.clear();
        for (int id : todo) t.divide(id, expand);                             // :598-665
        const int nToExpand = (int)expand.size();
        if (t.alive >= N || t.alive == prev) {
            finish = true;
        } else if (t.alive + nToExpand * 3 > N) {                             // :673
            while (!finish) {
                const int prev2 = t.alive;
                std::vector<int> prevExpand = expand;
                expand.clear();
                // sort by (size asc, pointer asc) and walk from the back (:684-685)
                std::sort(prevExpand.begin(), prevExpand.end(), [&](int a, int b) {
                    const size_t sa = t.nodes[a].keys.size(), sb = t.nodes[b].keys.size();
                    if (sa != sb) return sa < sb;
                    return a < b;
                });
                for (int j = (int)prevExpand.size() - 1; j >= 0; j--) {
                    t.divide(prevExpand[j], expand);
                    if (t.alive >= N) break;                                  // :731-732
                }
                if (t.alive >= N || t.alive == prev2) finish = true;
            }
        }
    }
    // retain the best point per node (:741-760): first max wins
    t.for_each_in_list_order([&](int i) {
        const std::vector<int>& keys = t.nodes[i].keys;
        int best = keys[0];
        for (size_t k = 1; k < keys.size(); k++)
            if (score[keys[k]] > score[best]) best = keys[k];
        result.push_back(best);
    });
    return 0;
}

}  // namespace

extern "C" int orc_distribute_octtree(const int16_t* x, const int16_t* y, const uint8_t* score, int n, int min_x,
                                      int max_x, int min_y, int max_y, int n_features, int32_t* out, int cap) {
    std::vector<int> res;
    if (n == 0) return 0;
    const int rc = distribute_octtree(x, y, score, n, min_x, max_x, min_y, max_y, n_features, res);
    if (rc) return rc;
    for (size_t i = 0; i < res.size() && (int)i < cap; i++) out[i] = res[i];
    return (int)res.size();
}

// operator() R21 :1043-1105
extern "C" int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride, orc_keypoint* kps,
                           uint8_t* desc, int cap, int* n_out) {
    if (n_out) *n_out = 0;
    if (!img || w <= 0 || h <= 0) return 0;  // :1046-1047 silent return on empty image
    const int L = e->nlevels;
    // ---- ComputePyramid :1107-1132
    for (int l = 0; l < L; l++) {
        Level& lv = e->lv[l];
        const float scale = e->isf[l];
        lv.w = cv_round((float)w * scale);
        lv.h = cv_round((float)h * scale);
        if (lv.w < 2 * kEdge + 1 || lv.h < 2 * kEdge + 1) return -3;  // reflect101 of 19 px needs >19 px
        lv.pstride = lv.w + 2 * kEdge;
        lv.padded.assign(lv.pstride * (lv.h + 2 * kEdge), 0);
        lv.has_blur = false;
        lv.cands.clear();
        lv.kps.clear();
        uint8_t* roi = lv.padded.data() + (size_t)kEdge * lv.pstride + kEdge;
        if (l != 0) {
            const Level& pv = e->lv[l - 1];
            orc_resize_linear_u8(pv.roi(), pv.w, pv.h, pv.pstride, roi, lv.w, lv.h, lv.pstride);
            orc_copy_make_border_reflect101(roi, lv.w, lv.h, lv.pstride, lv.padded.data(), lv.pstride, kEdge,
                                            kEdge, kEdge, kEdge);
        } else {
            orc_copy_make_border_reflect101(img, w, h, stride, lv.padded.data(), lv.pstride, kEdge, kEdge, kEdge,
                                            kEdge);
        }
    }
    // ---- ComputeKeyPointsOctTree :765-853
    const float W = 30;
    std::vector<orc_keypoint> cell(4096);
    for (int l = 0; l < L; l++) {
        Level& lv = e->lv[l];
        const int minBX = kEdge - 3, minBY = minBX;
        const int maxBX = lv.w - kEdge + 3, maxBY = lv.h - kEdge + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) return -4;  // the reference divides by zero
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        const uint8_t* roi = lv.roi();
        for (int i = 0; i < nRows; i++) {
            const int iniY = minBY + i * hCell;
            int maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = maxBY;
            for (int j = 0; j < nCols; j++) {
                const int iniX = minBX + j * wCell;
                int maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = maxBX;
                const uint8_t* sub = roi + (ptrdiff_t)iniY * (ptrdiff_t)lv.pstride + iniX;
                const int cw = maxX - iniX, ch = maxY - iniY;
                if ((size_t)cw * ch > cell.size()) cell.resize((size_t)cw * ch);
                int nk = orc_fast9_16(sub, cw, ch, lv.pstride, e->ini_th, 1, cell.data(), (int)cell.size());
                if (nk == 0) nk = orc_fast9_16(sub, cw, ch, lv.pstride, e->min_th, 1, cell.data(), (int)cell.size());
                for (int k = 0; k < nk; k++) {
                    Cand c;
                    c.x = (int16_t)((int)cell[k].x + j * wCell);
                    c.y = (int16_t)((int)cell[k].y + i * hCell);
                    c.score = (uint8_t)cell[k].response;
                    lv.cands.push_back(c);
                }
            }
        }
        const int nc = (int)lv.cands.size();
        std::vector<int16_t> cx(nc), cy(nc);
        std::vector<uint8_t> cs(nc);
        for (int k = 0; k < nc; k++) { cx[k] = lv.cands[k].x; cy[k] = lv.cands[k].y; cs[k] = lv.cands[k].score; }
        std::vector<int> sel;
        if (nc > 0) {
            const int rc = distribute_octtree(cx.data(), cy.data(), cs.data(), nc, minBX, maxBX, minBY, maxBY,
                                              e->nfeat[l], sel);
            if (rc) return rc;
        }
        const int scaledPatchSize = (int)(kPatch * e->sf[l]);  // :836
        for (int k : sel) {
            orc_keypoint kp;
            kp.x = (float)(cx[k] + minBX);
            kp.y = (float)(cy[k] + minBY);
            kp.size = (float)scaledPatchSize;
            kp.angle = -1.f;
            kp.response = (float)cs[k];
            kp.octave = l;
            kp.class_id = -1;
            lv.kps.push_back(kp);
        }
    }
    for (int l = 0; l < L; l++) {  // computeOrientation :851-852
        Level& lv = e->lv[l];
        for (orc_keypoint& kp : lv.kps)
            kp.angle = ic_angle_umax(lv.roi() + (ptrdiff_t)cv_round(kp.y) * (ptrdiff_t)lv.pstride + cv_round(kp.x),
                                     lv.pstride, e->umax);
    }
    // ---- descriptors :1058-1104
    int total = 0;
    for (int l = 0; l < L; l++) {
        Level& lv = e->lv[l];
        const int nk = (int)lv.kps.size();
        if (nk == 0) continue;
        lv.blurred.resize((size_t)lv.w * lv.h);
        orc_gaussian_blur7_sigma2(lv.roi(), lv.w, lv.h, lv.pstride, lv.blurred.data(), lv.w);
        lv.has_blur = true;
        for (int k = 0; k < nk; k++) {
            const orc_keypoint& kp = lv.kps[k];
            const int idx = total + k;
            if (idx >= cap) continue;
            if (desc)
                orc_orb_descriptor(lv.blurred.data() + (size_t)cv_round(kp.y) * lv.w + cv_round(kp.x), lv.w, kp.angle,
                                   e->trig_mode, e->fma_mode, desc + (size_t)idx * 32);
            if (kps) {
                orc_keypoint o = kp;
                if (l != 0) { const float s = e->sf[l]; o.x = o.x * s; o.y = o.y * s; }   // :1095-1101
                kps[idx] = o;
            }
        }
        total += nk;
    }
    if (n_out) *n_out = total;
    return 0;
}

extern "C" int orc_level_size(const orc_extractor* e, int l, int* w, int* h) {
    if (l < 0 || l >= e->nlevels) return -1;
    *w = e->lv[l].w; *h = e->lv[l].h;
    return 0;
}

extern "C" int orc_get_pyramid(const orc_extractor* e, int l, int with_border, uint8_t* dst, size_t dstride) {
    if (l < 0 || l >= e->nlevels) return -1;
    const Level& lv = e->lv[l];
    if (with_border) {
        for (int y = 0; y < lv.h + 2 * kEdge; y++) memcpy(dst + y * dstride, lv.padded.data() + y * lv.pstride, lv.pstride);
    } else {
        for (int y = 0; y < lv.h; y++) memcpy(dst + y * dstride, lv.roi() + y * lv.pstride, lv.w);
    }
    return 0;
}

extern "C" int orc_get_blurred(const orc_extractor* e, int l, uint8_t* dst, size_t dstride) {
    if (l < 0 || l >= e->nlevels) return -1;
    const Level& lv = e->lv[l];
    if (!lv.has_blur) return 1;
    for (int y = 0; y < lv.h; y++) memcpy(dst + y * dstride, lv.blurred.data() + (size_t)y * lv.w, lv.w);
    return 0;
}

extern "C" int orc_get_candidates(const orc_extractor* e, int l, int16_t* x, int16_t* y, uint8_t* s, int cap) {
    if (l < 0 || l >= e->nlevels) return -1;
    const Level& lv = e->lv[l];
    const int n = (int)lv.cands.size();
    for (int k = 0; k < n && k < cap; k++) { x[k] = lv.cands[k].x; y[k] = lv.cands[k].y; s[k] = lv.cands[k].score; }
    return n;
}

extern "C" int orc_get_level_keypoints(const orc_extractor* e, int l, orc_keypoint* out, int cap) {
    if (l < 0 || l >= e->nlevels) return -1;
    const Level& lv = e->lv[l];
    const int n = (int)lv.kps.size();
    for (int k = 0; k < n && k < cap; k++) out[k] = lv.kps[k];
    return n;
}

// ----------------------------------------------------------------------------------------------
// Matcher
// ----------------------------------------------------------------------------------------------
// ORBmatcher::DescriptorDistance R21/src/ORBmatcher.cc:1647-1663 (SWAR popcount over 8 x int32)
extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4);
        memcpy(&pb, b + 4 * i, 4);
        unsigned int v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

namespace {
template <class F> void parallel_for(int n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) { f(0, n); return; }
    nthreads = std::min(nthreads, n);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        const int lo = (int)((int64_t)n * t / nthreads), hi = (int)((int64_t)n * (t + 1) / nthreads);
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}
}  // namespace

// best / second-best update rule of R21 ORBmatcher.cc:216-225
extern "C" void orc_knn2(const uint8_t* q, int nq, const uint8_t* m, int64_t nm, int64_t base, int32_t* bi,
                         int32_t* bd1, int32_t* bd2, int nthreads) {
    parallel_for(nq, nthreads, [=](int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            int best1 = 256, best2 = 256, idx = -1;
            const uint8_t* dq = q + (size_t)i * 32;
            for (int64_t j = 0; j < nm; j++) {
                const int dist = orc_descriptor_distance(dq, m + (size_t)j * 32);
                if (dist < best1) { best2 = best1; best1 = dist; idx = (int)(base + j); }
                else if (dist < best2) { best2 = dist; }
            }
            bi[i] = idx; bd1[i] = best1; bd2[i] = best2;
        }
    });
}

extern "C" void orc_knn2_full(const uint8_t* q, int nq, const uint8_t* m, int64_t nm, int64_t base, int32_t* i1,
                              int32_t* d1, int32_t* i2, int32_t* d2, int nthreads) {
    parallel_for(nq, nthreads, [=](int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            int b1 = 256, b2 = 256, x1 = -1, x2 = -1;
            const uint8_t* dq = q + (size_t)i * 32;
            for (int64_t j = 0; j < nm; j++) {
                const int dist = orc_descriptor_distance(dq, m + (size_t)j * 32);
                if (dist < b1) { b2 = b1; x2 = x1; b1 = dist; x1 = (int)(base + j); }
                else if (dist < b2) { b2 = dist; x2 = (int)(base + j); }
            }
            i1[i] = x1; d1[i] = b1; i2[i] = x2; d2[i] = b2;
        }
    });
}

// MapPoint::ComputeDistinctiveDescriptors R21/src/MapPoint.cc:242-307
extern "C" void orc_distinctive_descriptors(const uint8_t* desc, const int32_t* ptr, int n_points, int32_t* best) {
    for (int p = 0; p < n_points; p++) {
        const int N = ptr[p + 1] - ptr[p];
        best[p] = -1;
        if (N <= 0) continue;
        const uint8_t* D = desc + (size_t)ptr[p] * 32;
        std::vector<float> dist((size_t)N * N);
        for (int i = 0; i < N; i++) {
            dist[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int dij = orc_descriptor_distance(D + (size_t)i * 32, D + (size_t)j * 32);
                dist[(size_t)i * N + j] = (float)dij;
                dist[(size_t)j * N + i] = (float)dij;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> v(dist.begin() + (size_t)i * N, dist.begin() + (size_t)(i + 1) * N);
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best[p] = BestIdx;
    }
}

namespace {
const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;  // R21 ORBmatcher.cc:37-39

// ComputeThreeMaxima R21 :1601-1642
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {   // :236-243
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// walk two sorted CSR feature vectors, calling f(node_pos1, node_pos2) on shared node ids (:175-263)
template <class F> void merge_walk(const orc_featvec* a, const orc_featvec* b, F f) {
    int i = 0, j = 0;
    while (i < a->n_nodes && j < b->n_nodes) {
        if (a->node_ids[i] == b->node_ids[j]) { f(i, j); i++; j++; }
        else if (a->node_ids[i] < b->node_ids[j]) i++;   // lower_bound on a sorted map == advance
        else j++;
    }
}
}  // namespace

// SearchByBoW(KeyFrame*, Frame&, ...) R21 :159-288
extern "C" int orc_search_by_bow_kf_f(const uint8_t* dkf, const float* akf, const uint8_t* kf_valid, int n_kf,
                                      const orc_featvec* fvk, const uint8_t* df, const float* af, int n_f,
                                      const orc_featvec* fvf, float nnratio, int check_ori, int32_t* match_f) {
    (void)n_kf;
    for (int i = 0; i < n_f; i++) match_f[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    merge_walk(fvk, fvf, [&](int nk, int nf) {
        for (int p = fvk->ptr[nk]; p < fvk->ptr[nk + 1]; p++) {
            const int realIdxKF = fvk->idx[p];
            if (!kf_valid[realIdxKF]) continue;
            int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
            for (int r = fvf->ptr[nf]; r < fvf->ptr[nf + 1]; r++) {
                const int realIdxF = fvf->idx[r];
                if (match_f[realIdxF] >= 0) continue;
                const int dist = orc_descriptor_distance(dkf + (size_t)realIdxKF * 32, df + (size_t)realIdxF * 32);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                else if (dist < bestDist2) { bestDist2 = dist; }
            }
            if (bestDist1 <= TH_LOW) {
                if ((float)bestDist1 < nnratio * (float)bestDist2) {
                    match_f[bestIdxF] = realIdxKF;
                    if (check_ori) rotHist[rot_bin(akf[realIdxKF], af[bestIdxF])].push_back(bestIdxF);
                    nmatches++;
                }
            }
        }
    });
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match_f[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// SearchByBoW(KeyFrame*, KeyFrame*, ...) R21 :522-655
extern "C" int orc_search_by_bow_kf_kf(const uint8_t* d1, const float* a1, const uint8_t* valid1, int n1,
                                       const orc_featvec* fv1, const uint8_t* d2, const float* a2,
                                       const uint8_t* valid2, int n2, const orc_featvec* fv2, float nnratio,
                                       int check_ori, int32_t* match12) {
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<uint8_t> matched2(n2, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    merge_walk(fv1, fv2, [&](int na, int nb) {
        for (int p = fv1->ptr[na]; p < fv1->ptr[na + 1]; p++) {
            const int idx1 = fv1->idx[p];
            if (!valid1[idx1]) continue;
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int r = fv2->ptr[nb]; r < fv2->ptr[nb + 1]; r++) {
                const int idx2 = fv2->idx[r];
                if (matched2[idx2] || !valid2[idx2]) continue;
                const int dist = orc_descriptor_distance(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                else if (dist < bestDist2) { bestDist2 = dist; }
            }
            if (bestDist1 < TH_LOW) {
                if ((float)bestDist1 < nnratio * (float)bestDist2) {
                    match12[idx1] = bestIdx2;
                    matched2[bestIdx2] = 1;
                    if (check_ori) rotHist[rot_bin(a1[idx1], a2[bestIdx2])].push_back(idx1);
                    nmatches++;
                }
            }
        }
    });
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// CheckDistEpipolarLine R21 :140-157
static bool check_dist_epipolar(float x1, float y1, float x2, float y2, const float* F, float sigma2) {
    const float a = x1 * F[0] + y1 * F[3] + F[6];
    const float b = x1 * F[1] + y1 * F[4] + F[7];
    const float c = x1 * F[2] + y1 * F[5] + F[8];
    const float num = a * x2 + b * y2 + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2;   // float < double*float -> double compare, as in the reference
}

// SearchForTriangulation R21 :657-823
extern "C" int orc_search_for_triangulation(const uint8_t* d1, const orc_tri_feature* f1, int n1,
                                            const orc_featvec* fv1, const uint8_t* d2, const orc_tri_feature* f2,
                                            int n2, const orc_featvec* fv2, const float* F12, float ex, float ey,
                                            const float* sf2, const float* sigma2_2, int only_stereo, int check_ori,
                                            int32_t* out_pairs, int cap_pairs) {
    int nmatches = 0;
    std::vector<uint8_t> matched2(n2, 0);   // never set in the reference (:677,:725)
    std::vector<int> matches12(n1, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    merge_walk(fv1, fv2, [&](int na, int nb) {
        for (int p = fv1->ptr[na]; p < fv1->ptr[na + 1]; p++) {
            const int idx1 = fv1->idx[p];
            if (f1[idx1].has_mp) continue;
            const bool bStereo1 = f1[idx1].u_right >= 0;
            if (only_stereo && !bStereo1) continue;
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int r = fv2->ptr[nb]; r < fv2->ptr[nb + 1]; r++) {
                const int idx2 = fv2->idx[r];
                if (matched2[idx2] || f2[idx2].has_mp) continue;
                const bool bStereo2 = f2[idx2].u_right >= 0;
                if (only_stereo && !bStereo2) continue;
                const int dist = orc_descriptor_distance(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                if (dist > TH_LOW || dist > bestDist) continue;
                if (!bStereo1 && !bStereo2) {
                    const float distex = ex - f2[idx2].x, distey = ey - f2[idx2].y;
                    if (distex * distex + distey * distey < 100 * sf2[f2[idx2].octave]) continue;
                }
                if (check_dist_epipolar(f1[idx1].x, f1[idx1].y, f2[idx2].x, f2[idx2].y, F12,
                                        sigma2_2[f2[idx2].octave])) {
                    bestIdx2 = idx2;
                    bestDist = dist;
                }
            }
            if (bestIdx2 >= 0) {
                matches12[idx1] = bestIdx2;
                nmatches++;
                if (check_ori) rotHist[rot_bin(f1[idx1].angle, f2[bestIdx2].angle)].push_back(idx1);
            }
        }
    });
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { matches12[j] = -1; nmatches--; }
        }
    }
    int np = 0;
    for (int i = 0; i < n1; i++) {
        if (matches12[i] < 0) continue;
        if (np < cap_pairs) { out_pairs[2 * np] = i; out_pairs[2 * np + 1] = matches12[i]; }
        np++;
    }
    return nmatches;
}

// Frame::ComputeStereoMatches R21/src/Frame.cc:471-645
extern "C" int orc_stereo_matches(const orc_keypoint* kl, const uint8_t* dl, int N, const orc_keypoint* kr,
                                  const uint8_t* dr, int Nr, int nlevels, const float* sfs, const float* isfs,
                                  const uint8_t* const* pyr_l, const uint8_t* const* pyr_r, const int* lvl_w,
                                  const int* lvl_h, const size_t* strides, float mbf, float mb, float* u_right,
                                  float* depth) {
    (void)nlevels;
    for (int i = 0; i < N; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = lvl_h[0];
    std::vector<std::vector<int>> rowIdx(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * sfs[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // guard: the reference indexes unchecked
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < N; iL++) {
        const orc_keypoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& cand = rowIdx[row];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        for (size_t iC = 0; iC < cand.size(); iC++) {
            const int iR = cand[iC];
            const orc_keypoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_descriptor_distance(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = isfs[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const int oct = kpL.octave;
            const size_t st = strides[oct];
            const uint8_t* PL = pyr_l[oct] + (size_t)kEdge * st + kEdge;   // ROI origin
            const uint8_t* PR = pyr_r[oct] + (size_t)kEdge * st + kEdge;
            const int cu = (int)scaleduL, cv = (int)scaledvL, cr0 = (int)scaleduR0;
            int bestSad = INT_MAX, bestincR = 0;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= lvl_w[oct]) continue;
            const int cL = PL[(ptrdiff_t)cv * (ptrdiff_t)st + cu];
            for (int incR = -L; incR <= +L; incR++) {
                const int cR = PR[(ptrdiff_t)cv * (ptrdiff_t)st + (cr0 + incR)];
                // cv::norm(IL, IR, NORM_L1) on centre-normalised float patches; all values are integers
                float dist = 0;
                for (int dy = -w; dy <= w; dy++)
                    for (int dx = -w; dx <= w; dx++) {
                        const float a = (float)PL[(ptrdiff_t)(cv + dy) * (ptrdiff_t)st + (cu + dx)] - (float)cL;
                        const float b = (float)PR[(ptrdiff_t)(cv + dy) * (ptrdiff_t)st + (cr0 + incR + dx)] - (float)cR;
                        dist += fabsf(a - b);
                    }
                if (dist < bestSad) { bestSad = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = sfs[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                u_right[iL] = bestuR;
                distIdx.push_back(std::pair<int, int>(bestSad, iL));
            }
        }
    }
    if (distIdx.empty()) return 0;   // guard: the reference reads vDistIdx[0] of an empty vector (:632)
    std::sort(distIdx.begin(), distIdx.end());
    const float median = distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)distIdx.size();
    for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
        if (distIdx[i].first < thDist) break;
        u_right[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
        kept--;
    }
    return kept;
}
