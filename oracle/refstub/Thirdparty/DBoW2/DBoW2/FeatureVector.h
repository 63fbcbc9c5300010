// refstub (TEST INFRASTRUCTURE ONLY): DBoW2::FeatureVector = map<NodeId, vector<unsigned int>> (see BowVector.h here).
#ifndef REFSTUB_DBOW2_FEATUREVECTOR_H
#define REFSTUB_DBOW2_FEATUREVECTOR_H
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {
public:
    void addFeature(NodeId id, unsigned int i_feature) {
        iterator it = lower_bound(id);
        if (it != end() && it->first == id) it->second.push_back(i_feature);
        else { it = insert(it, value_type(id, std::vector<unsigned int>())); it->second.push_back(i_feature); }
    }
};
}  // namespace DBoW2
#endif
