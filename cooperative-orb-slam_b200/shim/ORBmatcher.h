// ORBmatcher.h -- drop-in replacement of R21/include/ORBmatcher.h:37-102 backed by liborbcuda.
//
// Same class name, namespace, constructor and public methods as the reference's ORB_SLAM2::ORBmatcher, so
// Tracking / LocalMapping / LoopClosing / MapPoint / Frame compile and link against it unchanged.  Frame, KeyFrame
// and MapPoint are the reference's own classes (their headers are included, not replaced).  Every candidate loop
// runs on the GPU through the C ABI of include/orbcuda.h; what stays on the host is the reference's geometry in front
// of each window search (a handful of 3x3 products per point) and the map bookkeeping behind Fuse.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <set>
#include <utility>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"

namespace ORB_SLAM2
{

class ORBmatcher
{
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    // 256-bit Hamming distance of two descriptor rows (R21/src/ORBmatcher.cc:1647-1663) -> orb_hamming256
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

    // Tracking::SearchLocalPoints (:45-130) -> orbm_search_by_projection_frame
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    // Tracking::TrackWithMotionModel (:1328-1470) -> orbm_search_by_projection_last_frame
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    // Tracking::Relocalization (:1472-1599) -> orbm_search_by_projection_keyframe
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const int ORBdist);
    // LoopClosing::ComputeSim3 (:290-403) -> orbm_search_by_projection_sim3
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints,
                           std::vector<MapPoint*>& vpMatched, int th);

    // Relocalisation / loop detection (:159-288, :522-655) -> orbm_search_by_bow_kf_f / _kf_kf
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);

    // Monocular initialisation (:405-520) -> orbm_search_for_initialization
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);

    // LocalMapping::CreateNewMapPoints (:657-823) -> orbm_search_for_triangulation
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);

    // LoopClosing::ComputeSim3 (:1102-1326) -> orbm_window_best_match, twice + the mutual-consistency test
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);

    // LocalMapping::SearchInNeighbors (:825-975) and LoopClosing::SearchAndFuse (:977-1100) -> orbm_window_best_match
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);

public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

    // CUDA device the searches of this process run on (default 0); not part of the reference interface
    static int msDevice;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2

#endif  // ORBMATCHER_H
