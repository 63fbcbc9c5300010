// ref_wrap.cc -- C entry points around the REFERENCE's own ORBextractor class
// (R21/include/ORBextractor.h, R21/src/ORBextractor.cc compiled verbatim from /root/reference
// over oracle/cvshim by oracle/Makefile into oracle/_ref/).  TEST INFRASTRUCTURE ONLY.
#include "ORBextractor.h"
#include "orb_oracle.h"

using ORB_SLAM2::ORBextractor;

extern "C" {

void* orbref_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    return new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void orbref_destroy(void* h) { delete (ORBextractor*)h; }

void orbref_tables(void* h, float* sf, float* isf, float* s2, float* is2) {
    ORBextractor* e = (ORBextractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                       d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->GetLevels(); i++) { sf[i] = a[i]; isf[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; }
}

int orbref_extract(void* h, const uint8_t* img, int w, int hgt, size_t stride, orc_keypoint* kps, uint8_t* desc,
                   int cap, int* n) {
    ORBextractor* e = (ORBextractor*)h;
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    (*e)(image, cv::Mat(), keys, descriptors);
    const int nk = (int)keys.size();
    *n = nk;
    for (int i = 0; i < nk && i < cap; i++) {
        memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
        memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return 0;
}

int orbref_level_size(void* h, int level, int* w, int* hgt) {
    ORBextractor* e = (ORBextractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size()) return -1;
    *w = e->mvImagePyramid[level].cols;
    *hgt = e->mvImagePyramid[level].rows;
    return 0;
}

// mvImagePyramid[level] is the ROI at (+19,+19) of the padded plane (R21 ORBextractor.cc:1113-1116)
int orbref_get_pyramid(void* h, int level, int with_border, uint8_t* dst, size_t dstride) {
    ORBextractor* e = (ORBextractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    const int b = with_border ? 19 : 0;
    for (int y = -b; y < m.rows + b; y++)
        memcpy(dst + (size_t)(y + b) * dstride, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step.p - b, m.cols + 2 * b);
    return 0;
}

// seconds the calling thread has spent in the shim's resize / copyMakeBorder / GaussianBlur / FAST since the last reset
void orbref_stage_times(double* out4, int reset) {
    double* acc = cvshim_stage_acc();
    for (int i = 0; i < 4; i++) { out4[i] = acc[i]; if (reset) acc[i] = 0; }
}

}  // extern "C"
