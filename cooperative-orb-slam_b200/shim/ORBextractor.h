// ORBextractor.h -- drop-in replacement of the reference's include/ORBextractor.h
// (R21/include/ORBextractor.h:45-111 of 530300865/Cooperative-ORB-SLAM): same namespace, class name,
// constructor, operator(), getters and public mvImagePyramid, backed by liborbcuda.so (B200).
// Link-time substitution: compile this file's .cc instead of the reference's src/ORBextractor.cc and
// link -lorbcuda.  Frame.cc / Tracking.cc compile against it unchanged.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include <opencv/cv.h>

struct orbx_handle_s;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int inline GetLevels(){ return nlevels; }
    float inline GetScaleFactor(){ return scaleFactor; }
    std::vector<float> inline GetScaleFactors(){ return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors(){ return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares(){ return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares(){ return mvInvLevelSigma2; }

    // Host mirror of the device pyramid: every element is the ROI (+19,+19) of a padded plane, exactly
    // like the reference (R21/src/ORBextractor.cc:1113-1116), so Frame::ComputeStereoMatches may read
    // it (R21/src/Frame.cc:478,568,585).  Filled after each call (eight queued downloads into page-locked planes, one
    // synchronisation) unless SetPyramidMirror(false) or ORBCUDA_PYRAMID_MIRROR=0: monocular tracking never reads it and
    // the drop-in's stereo search (orbaccel::ComputeStereoMatches) uses the device pyramids.
    std::vector<cv::Mat> mvImagePyramid;

    // ---- additions (not in the reference) ----
    void SetPyramidMirror(bool on) { mbMirrorPyramid = on; }   // skip the 1.2 MB/frame download
    orbx_handle_s* Handle() const { return mpHandle; }           // for orbm_stereo_matches
    static void SetDevice(int device) { msDevice = device; }     // CUDA device of extractors created next

protected:
    orbx_handle_s* mpHandle;
    bool mbMirrorPyramid;
    static int msDevice;
    std::vector<cv::Mat> mvPadded;       // headers over the page-locked planes below
    std::vector<void*> mvPinned;
    void* mpStage = 0;                   // page-locked key point / descriptor staging of the current image size
    int mnCapWidth, mnCapHeight, mnCap;

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
};

} //namespace ORB_SLAM

#endif
