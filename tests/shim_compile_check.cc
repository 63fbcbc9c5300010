// Compile-time check (CPU, tests/test_boundary.py): the drop-in shim classes build against an
// OpenCV-compatible header set and the adapter templates instantiate on types shaped like the
// reference's Frame / KeyFrame / MapPoint.  Uses oracle/cvshim only as a stand-in for OpenCV headers.
#include <map>
#include <type_traits>
#include "ORBextractor.h"
#include "ORBmatcher_accel.h"

struct MapPoint {
    bool isBad() { return false; }
    bool mbTrackInView; float mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos; int mnTrackScaleLevel;
    int Observations() { return 1; }
    cv::Mat GetDescriptor() { return cv::Mat(); }
};
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
struct KeyFrame {
    int N;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2;
    cv::Mat mDescriptors;
    FeatureVector mFeatVec;
    std::vector<MapPoint*> mps;
    std::vector<MapPoint*> GetMapPointMatches() { return mps; }
    MapPoint* GetMapPoint(size_t i) { return mps[i]; }
};
struct Frame {
    int N;
    float mbf, fx, fy, cx, cy;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors;
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat mDistCoef;
    cv::Mat mDescriptors, mDescriptorsRight;
    FeatureVector mFeatVec;
    std::map<unsigned int, double> mBowVec;
    ORB_SLAM2::ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
};

int shim_instantiate(KeyFrame* a, KeyFrame* b, Frame& f) {
    std::vector<MapPoint*> out;
    std::vector<std::pair<size_t, size_t> > pairs;
    float F12[9] = {0};
    int n = orbaccel::SearchByBoW_KF_F(a, f, out, 0.7f, true);
    n += orbaccel::SearchByBoW_KF_KF(a, b, out, 0.75f, true);
    n += orbaccel::SearchForTriangulation(a, b, F12, 0.f, 0.f, pairs, false, false);
    orbaccel::ComputeStereoMatches(f);
    orbaccel::UndistortKeyPoints(f);
    orbaccel::FrameGrid grid = orbaccel::AssignFeaturesToGrid(f, 640, 480);
    n += orbaccel::SearchByProjection(f, grid, out, 1.0f, 0.8f);
    orbaccel::FrameGrid grid2 = orbaccel::BuildFrame(f, grid.bounds);
    n += (int)grid2.cell_idx.size();
    std::vector<orbm_proj_point_t> pts(out.size());
    n += orbaccel::SearchByProjectionLastFrame(f, grid, pts, out, 15.0f, 0, true);
    orbaccel::ComputeBoW(f, (orbv_handle_t)0);
    ORB_SLAM2::ORBextractor e(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> k; cv::Mat d, img;
    e(img, cv::Mat(), k, d);
    return n + e.GetLevels() + (int)e.GetScaleFactors().size() + (int)e.mvImagePyramid.size();
}
