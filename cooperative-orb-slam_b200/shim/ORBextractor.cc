// ORBextractor.cc -- shim over the C ABI (include/orbcuda.h).  Replaces R21/src/ORBextractor.cc.
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orbcuda.h"

namespace ORB_SLAM2
{

int ORBextractor::msDevice = 0;

static void orb_throw(int rc, const char* what)
{
    // the reference aborts through CV_Assert / assert on misuse; errors never pass silently here either
    throw std::runtime_error(std::string(what) + " failed: " + orb_last_error() + " (status " + std::to_string(rc) + ")");
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST):
    mpHandle(0), mbMirrorPyramid(true),
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST)
{
    orbx_params_t p;
    p.nfeatures = _nfeatures; p.scale_factor = _scaleFactor; p.nlevels = _nlevels;
    p.ini_th_fast = _iniThFAST; p.min_th_fast = _minThFAST;
    int rc = orbx_create(&p, 0, 0, 0, msDevice, &mpHandle);
    if(rc != ORB_OK)
        orb_throw(rc, "orbx_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    orbx_tables(mpHandle, &mvScaleFactor[0], &mvInvScaleFactor[0], &mvLevelSigma2[0], &mvInvLevelSigma2[0], 0);
    mvImagePyramid.resize(nlevels);
    mvPadded.resize(nlevels);
}

ORBextractor::~ORBextractor()
{
    orbx_destroy(mpHandle);
}

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                      cv::OutputArray _descriptors)
{
    (void)_mask;
    if(_image.empty())
        return;

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );

    int cap = 0;
    int rc = orbx_max_keypoints(mpHandle, image.cols, image.rows, &cap);
    if(rc != ORB_OK)
        orb_throw(rc, "orbx_max_keypoints");

    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint_t), "cv::KeyPoint layout");
    std::vector<cv::KeyPoint> keys(cap);
    std::vector<unsigned char> desc((size_t)cap*32);
    int n = 0;
    rc = orbx_extract(mpHandle, image.data, image.cols, image.rows, (size_t)image.step,
                      reinterpret_cast<orb_keypoint_t*>(&keys[0]), &desc[0], cap, &n);
    if(rc != ORB_OK)
        orb_throw(rc, "orbx_extract");

    if(n == 0)
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        for(int i=0; i<n; i++)
            memcpy(descriptors.ptr(i), &desc[(size_t)i*32], 32);
    }
    keys.resize(n);
    _keypoints.swap(keys);

    if(mbMirrorPyramid)
    {
        const int EDGE_THRESHOLD = 19;
        for(int level = 0; level < nlevels; ++level)
        {
            int w = 0, h = 0;
            orbx_level_size(mpHandle, level, &w, &h);
            if(mvPadded[level].cols != w + 2*EDGE_THRESHOLD || mvPadded[level].rows != h + 2*EDGE_THRESHOLD)
                mvPadded[level] = cv::Mat(cv::Size(w + 2*EDGE_THRESHOLD, h + 2*EDGE_THRESHOLD), CV_8UC1);
            rc = orbx_download_level(mpHandle, 0, level, 1, mvPadded[level].data, (size_t)mvPadded[level].step);
            if(rc != ORB_OK)
                orb_throw(rc, "orbx_download_level");
            mvImagePyramid[level] = mvPadded[level](cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }
    }
}

} //namespace ORB_SLAM
