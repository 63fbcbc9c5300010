// cvshim <opencv/cv.h> (TEST INFRASTRUCTURE ONLY) -- included by R21/include/ORBextractor.h:26
#ifndef CVSHIM_CV_H
#define CVSHIM_CV_H
#include "opencv2/core/core.hpp"
#include "opencv2/imgproc/imgproc.hpp"
#include "opencv2/features2d/features2d.hpp"
#endif
