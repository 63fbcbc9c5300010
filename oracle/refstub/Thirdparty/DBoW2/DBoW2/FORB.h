// refstub (TEST INFRASTRUCTURE ONLY): DBoW2::FORB names only (descriptor = one cv::Mat row of 32 bytes).
#ifndef REFSTUB_DBOW2_FORB_H
#define REFSTUB_DBOW2_FORB_H
#include <opencv2/core/core.hpp>
#include <vector>
namespace DBoW2 {
class FORB {
public:
    typedef cv::Mat TDescriptor;
    typedef const TDescriptor* pDescriptor;
    static const int L = 32;
};
}  // namespace DBoW2
#endif
