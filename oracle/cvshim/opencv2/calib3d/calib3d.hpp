// cvshim calib3d (TEST INFRASTRUCTURE ONLY): cv::undistortPoints as Frame.cc:428,:453 call it (R empty, P = K),
// forwarded to the cv2-pinned model of oracle/frame_oracle.cc.
#ifndef CVSHIM_CALIB3D_HPP
#define CVSHIM_CALIB3D_HPP
#include "opencv2/core/core.hpp"
namespace cv {
inline void undistortPoints(InputArray _src, OutputArray _dst, InputArray _K, InputArray _D, InputArray _R, InputArray _P) {
    Mat src = _src.getMat(), K = _K.getMat(), D = _D.getMat(), P = _P.getMat();
    assert(_R.empty());
    assert(src.type() == CV_32FC2 && src.cols == 1 && K.type() == CV_32F && D.type() == CV_32F);
    (void)_R;
    const float k[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
    assert(P.at<float>(0, 0) == k[0] && P.at<float>(1, 1) == k[1] && P.at<float>(0, 2) == k[2] && P.at<float>(1, 2) == k[3]);
    const int nd = (int)D.total();
    std::vector<float> d(nd), in((size_t)src.rows * 2), out((size_t)src.rows * 2);
    for (int i = 0; i < nd; i++) d[i] = D.at<float>(i);
    for (int i = 0; i < src.rows; i++) { in[2 * i] = src.ptr<float>(i)[0]; in[2 * i + 1] = src.ptr<float>(i)[1]; }
    orc_undistort_points(in.data(), src.rows, k, d.data(), nd, out.data());
    _dst.create(src.rows, 1, CV_32FC2);
    Mat dst = _dst.getMat();
    for (int i = 0; i < src.rows; i++) { dst.ptr<float>(i)[0] = out[2 * i]; dst.ptr<float>(i)[1] = out[2 * i + 1]; }
}
}  // namespace cv
#endif
