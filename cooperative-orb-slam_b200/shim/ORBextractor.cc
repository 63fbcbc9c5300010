// ORBextractor.cc -- shim over the C ABI (include/orbcuda.h).  Replaces R21/src/ORBextractor.cc.
#include "ORBextractor.h"

#include <cassert>
#include <cstdlib>
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
    mpHandle(0), mbMirrorPyramid(true), mnCapWidth(0), mnCapHeight(0), mnCap(0),
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST)
{
    // ORBCUDA_PYRAMID_MIRROR=0: never fill mvImagePyramid (monocular tracking does not read it; the stereo search of the
    // drop-in runs on the device pyramids).  The default keeps the reference's contract: the public member is valid after
    // every call.
    const char* e = getenv("ORBCUDA_PYRAMID_MIRROR");
    if(e && atoi(e) == 0) mbMirrorPyramid = false;
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
    mvPinned.assign(nlevels, static_cast<void*>(0));
}

ORBextractor::~ORBextractor()
{
    mvImagePyramid.clear();
    mvPadded.clear();
    for(size_t i = 0; i < mvPinned.size(); i++)
        if(mvPinned[i]) orb_host_free(mvPinned[i]);
    if(mpStage) orb_host_free(mpStage);
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

    int rc;
    if(image.cols != mnCapWidth || image.rows != mnCapHeight)
    {
        // per image size: key point capacity, page-locked result staging, page-locked pyramid planes
        rc = orbx_max_keypoints(mpHandle, image.cols, image.rows, &mnCap);
        if(rc != ORB_OK)
            orb_throw(rc, "orbx_max_keypoints");
        mnCapWidth = image.cols; mnCapHeight = image.rows;
        if(mpStage) { orb_host_free(mpStage); mpStage = 0; }
        rc = orb_host_alloc(&mpStage, (size_t)mnCap * (sizeof(orb_keypoint_t) + 32) + 64);
        if(rc != ORB_OK)
            orb_throw(rc, "orb_host_alloc");
    }

    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint_t), "cv::KeyPoint layout");
    orb_keypoint_t* keys = static_cast<orb_keypoint_t*>(mpStage);
    unsigned char* desc = static_cast<unsigned char*>(mpStage) + (size_t)mnCap * sizeof(orb_keypoint_t);
    int n = 0;
    rc = orbx_extract(mpHandle, image.data, image.cols, image.rows, (size_t)image.step, keys, desc, mnCap, &n);
    if(rc != ORB_OK)
        orb_throw(rc, "orbx_extract");

    if(n == 0)
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        if(descriptors.isContinuous())
            memcpy(descriptors.ptr(0), desc, (size_t)n*32);
        else
            for(int i=0; i<n; i++)
                memcpy(descriptors.ptr(i), desc + (size_t)i*32, 32);
    }
    _keypoints.resize(n);
    if(n)
        memcpy(reinterpret_cast<void*>(&_keypoints[0]), keys, (size_t)n*sizeof(cv::KeyPoint));

    if(mbMirrorPyramid)
    {
        const int EDGE_THRESHOLD = 19;
        std::vector<unsigned char*> planes(nlevels);
        std::vector<size_t> strides(nlevels);
        for(int level = 0; level < nlevels; ++level)
        {
            int w = 0, h = 0;
            orbx_level_size(mpHandle, level, &w, &h);
            const int pw = w + 2*EDGE_THRESHOLD, ph = h + 2*EDGE_THRESHOLD;
            if(mvPadded[level].cols != pw || mvPadded[level].rows != ph)
            {
                // page-locked planes: the eight downloads are queued back to back and waited for once
                mvImagePyramid[level] = cv::Mat();
                mvPadded[level] = cv::Mat();
                if(mvPinned[level]) { orb_host_free(mvPinned[level]); mvPinned[level] = 0; }
                rc = orb_host_alloc(&mvPinned[level], (size_t)pw*ph);
                if(rc != ORB_OK)
                    orb_throw(rc, "orb_host_alloc");
                mvPadded[level] = cv::Mat(ph, pw, CV_8UC1, mvPinned[level], (size_t)pw);
            }
            planes[level] = mvPadded[level].data;
            strides[level] = (size_t)mvPadded[level].step;
        }
        rc = orbx_download_pyramid(mpHandle, 0, 1, &planes[0], &strides[0]);
        if(rc != ORB_OK)
            orb_throw(rc, "orbx_download_pyramid");
        for(int level = 0; level < nlevels; ++level)
        {
            const int w = mvPadded[level].cols - 2*EDGE_THRESHOLD, h = mvPadded[level].rows - 2*EDGE_THRESHOLD;
            mvImagePyramid[level] = mvPadded[level](cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }
    }
}

} //namespace ORB_SLAM
