// refstub (TEST INFRASTRUCTURE ONLY): the container types of DBoW2 that the reference's Frame / KeyFrame / ORBmatcher
// name.  DBoW2 is a third-party dependency the reference does not vendor (R21/src/ORBmatcher.cc:28,
// R21/include/Frame.h:27-28); only the std::map-shaped containers are needed to compile those files from
// /root/reference.  Written from DBoW2's published interface (BowVector = map<WordId, WordValue>).
#ifndef REFSTUB_DBOW2_BOWVECTOR_H
#define REFSTUB_DBOW2_BOWVECTOR_H
#include <cmath>
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
enum LNorm { L1, L2 };
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };
class BowVector : public std::map<WordId, WordValue> {
public:
    void addWeight(WordId id, WordValue v) {
        iterator it = lower_bound(id);
        if (it != end() && !(key_comp()(id, it->first))) it->second += v;
        else insert(it, value_type(id, v));
    }
    void addIfNotExist(WordId id, WordValue v) {
        iterator it = lower_bound(id);
        if (it == end() || key_comp()(id, it->first)) insert(it, value_type(id, v));
    }
    void normalize(LNorm norm_type) {
        double norm = 0.0;
        for (iterator it = begin(); it != end(); ++it) norm += norm_type == L1 ? std::fabs(it->second) : it->second * it->second;
        if (norm_type == L2) norm = std::sqrt(norm);
        if (norm > 0.0) for (iterator it = begin(); it != end(); ++it) it->second /= norm;
    }
};
}  // namespace DBoW2
#endif
