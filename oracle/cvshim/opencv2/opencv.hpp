// cvshim umbrella header (TEST INFRASTRUCTURE ONLY)
#ifndef CVSHIM_OPENCV_HPP
#define CVSHIM_OPENCV_HPP
#include "opencv2/core/core.hpp"
#include "opencv2/imgproc/imgproc.hpp"
#include "opencv2/features2d/features2d.hpp"
#include "opencv2/calib3d/calib3d.hpp"
#endif
