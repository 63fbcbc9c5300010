// cvshim features2d (TEST INFRASTRUCTURE ONLY)
#ifndef CVSHIM_FEATURES2D_HPP
#define CVSHIM_FEATURES2D_HPP
#include "opencv2/core/core.hpp"
namespace cv {
// cv::FAST(image, keypoints, threshold, nonmaxSuppression) -- TYPE_9_16 (R21 ORBextractor.cc:809-815)
inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nms = true) {
    CvshimStageTimer timer(3);
    Mat img = _img.getMat();
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    std::vector<orc_keypoint> buf((size_t)img.cols * img.rows / 2 + 16);
    const int n = orc_fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms ? 1 : 0, buf.data(), (int)buf.size());
    keypoints.resize(n);
    if (n) memcpy(&keypoints[0], buf.data(), (size_t)n * sizeof(KeyPoint));
}
struct KeyPointsFilter {
    // only referenced by the dead ComputeKeyPointsOld path (R21 :855-1031)
    static void retainBest(std::vector<KeyPoint>& kps, int n) {
        if (n >= 0 && (int)kps.size() > n) {
            std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            kps.resize(n);
        }
    }
};
}  // namespace cv
#endif
