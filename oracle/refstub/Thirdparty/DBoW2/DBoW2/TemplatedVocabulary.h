// refstub (TEST INFRASTRUCTURE ONLY): DBoW2::TemplatedVocabulary is only named by the reference headers compiled for
// the parity harness (ORBVocabulary.h:31); no vocabulary file ships with the reference, the harness passes a null
// vocabulary and sets Frame::mFeatVec / mBowVec itself.  transform() aborts if it is ever reached.
#ifndef REFSTUB_DBOW2_TEMPLATEDVOCABULARY_H
#define REFSTUB_DBOW2_TEMPLATEDVOCABULARY_H
#include <cstdlib>
#include <string>
#include <vector>
#include "BowVector.h"
#include "FeatureVector.h"
#include <list>
#include <set>
#include <map>
// DBoW2's own TemplatedVocabulary.h has a file-scope using-directive; the reference's headers depend on it
// (KeyFrameDatabase.h:66 writes list<KeyFrame*>, Map.h:64 vector<KeyFrame*>).
using namespace std;
namespace DBoW2 {
template <class TDescriptor, class F> class TemplatedVocabulary {
public:
    TemplatedVocabulary() {}
    virtual ~TemplatedVocabulary() {}
    virtual void transform(const std::vector<TDescriptor>&, BowVector&, FeatureVector&, int) const { abort(); }
    virtual unsigned int size() const { return 0; }
    double score(const BowVector&, const BowVector&) const { abort(); return 0; }
    bool loadFromTextFile(const std::string&) { return false; }
};
}  // namespace DBoW2
#endif
