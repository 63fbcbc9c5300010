// cvshim imgproc (TEST INFRASTRUCTURE ONLY) -- forwards to the cv2-pinned integer models.
#ifndef CVSHIM_IMGPROC_HPP
#define CVSHIM_IMGPROC_HPP
#include "opencv2/core/core.hpp"
namespace cv {
// cv::resize, INTER_LINEAR, CV_8UC1 only (R21 ORBextractor.cc:1120)
inline void resize(InputArray _src, OutputArray _dst, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR) {
    assert(interp == INTER_LINEAR);
    CvshimStageTimer timer(0);
    (void)interp;
    Mat src = _src.getMat();
    _dst.create(dsize, CV_8UC1);
    Mat dst = _dst.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
// cv::copyMakeBorder, BORDER_REFLECT_101 (+BORDER_ISOLATED) only (R21 :1122-1128).  A non-isolated
// call on a sub-view would read the parent's pixels in OpenCV; the shim always isolates.
inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int type) {
    assert((type & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    CvshimStageTimer timer(1);
    (void)type;
    Mat src = _src.getMat();
    _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = _dst.getMat();
    orc_copy_make_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom, left, right);
}
// cv::GaussianBlur, 7x7 sigma 2 BORDER_REFLECT_101 only (R21 :1086)
inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy = 0, int border = BORDER_DEFAULT) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    CvshimStageTimer timer(2);
    (void)ksize; (void)sx; (void)sy; (void)border;
    Mat src = _src.getMat();
    _dst.create(src.rows, src.cols, CV_8UC1);
    Mat dst = _dst.getMat();
    orc_gaussian_blur7_sigma2(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}
}  // namespace cv
#endif
