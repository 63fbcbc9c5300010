/*
 * bow_oracle.cc -- CPU ORACLE (TEST INFRASTRUCTURE ONLY), part of liborb_oracle.so.
 *
 * Frame::ComputeBoW / KeyFrame::ComputeBoW (R21/src/Frame.cc:400-407, R21/src/KeyFrame.cc:60-69) call
 *   mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
 * ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (R21/include/ORBVocabulary.h:31-32).
 * DBoW2 is a third-party dependency that the reference does NOT vendor (it links
 * Thirdparty/DBoW2/lib/libDBoW2.so of the stock ORB_SLAM2 tree, R21/CMakeLists.txt:55: the ORB-SLAM2 fork of
 * DBoW2 by D. Galvez-Lopez, unversioned, as shipped with raulmur/ORB_SLAM2 v1.0), and no vocabulary file is shipped.
 * PARITY UNPINNED: this file restates the published algorithm of that version,
 *   TemplatedVocabulary::transform(const vector<TDescriptor>&, BowVector&, FeatureVector&, int levelsup)
 *   TemplatedVocabulary::transform(const TDescriptor&, WordId&, WordValue&, NodeId*, int levelsup)
 *   FORB::distance (256-bit Hamming), BowVector::addWeight / normalize(L1), FeatureVector::addFeature
 * and is anchored on the reference's call sites (levelsup = 4; BowVector/FeatureVector consumed by SearchByBoW
 * R21/src/ORBmatcher.cc:159-288 and KeyFrameDatabase).  The vocabulary is given as flat arrays:
 *   node i: children child_idx[child_ptr[i] .. child_ptr[i+1]) (empty: leaf), descriptor node_desc[i][32],
 *   word_id[i] and weight[i] (leaves).  Node 0 is the root.
 */
#include "orb_oracle.h"

#include <math.h>
#include <map>
#include <vector>

extern "C" {

/* transform(feature, id, weight, nid, levelsup): greedy descent, first child wins ties (strict <). */
void orc_bow_transform(const uint8_t* desc, int n, const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc,
                       const int32_t* word_id, const double* weight, int depth_L, int levelsup, int32_t* out_word,
                       int32_t* out_node, double* out_weight) {
    const int nid_level = depth_L - levelsup;
    for (int f = 0; f < n; f++) {
        int nid = 0;                       /* nid_level <= 0: the root */
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int b = child_ptr[final_id], e = child_ptr[final_id + 1];
            final_id = child_idx[b];
            double best_d = (double)orc_descriptor_distance(desc + (size_t)f * 32, node_desc + (size_t)final_id * 32);
            for (int c = b + 1; c < e; c++) {
                const int id = child_idx[c];
                const double d = (double)orc_descriptor_distance(desc + (size_t)f * 32, node_desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_ptr[final_id + 1] > child_ptr[final_id]);
        out_word[f] = word_id[final_id];
        out_weight[f] = weight[final_id];
        out_node[f] = nid;
    }
}

/* The rest of transform(features, v, fv, levelsup): v.addWeight(id, w) / fv.addFeature(nid, i) for w > 0 in feature
 * order, then v.normalize(L1).  BowVector -> (bow_words ascending, bow_values); FeatureVector -> CSR (fv_nodes
 * ascending, fv_ptr, fv_idx).  Returns the number of words; *n_nodes = number of feature-vector nodes.
 * Buffers: bow_* [n], fv_nodes [n], fv_ptr [n+1], fv_idx [n]. */
int orc_bow_vectors(const int32_t* word, const int32_t* node, const double* weight, int n, int normalize_l1, int32_t* bow_words,
                    double* bow_values, int32_t* fv_nodes, int32_t* fv_ptr, int32_t* fv_idx, int* n_nodes) {
    std::map<int32_t, double> v;
    std::map<int32_t, std::vector<int32_t> > fv;
    for (int i = 0; i < n; i++) {
        if (weight[i] > 0) {
            std::map<int32_t, double>::iterator it = v.lower_bound(word[i]);
            if (it != v.end() && !(v.key_comp()(word[i], it->first))) it->second += weight[i];
            else v.insert(it, std::map<int32_t, double>::value_type(word[i], weight[i]));
            fv[node[i]].push_back(i);
        }
    }
    if (!v.empty() && normalize_l1) {
        double norm = 0.0;
        for (std::map<int32_t, double>::iterator it = v.begin(); it != v.end(); ++it) norm += fabs(it->second);
        if (norm > 0.0)
            for (std::map<int32_t, double>::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }
    int k = 0;
    for (std::map<int32_t, double>::iterator it = v.begin(); it != v.end(); ++it, ++k) { bow_words[k] = it->first; bow_values[k] = it->second; }
    int m = 0, at = 0;
    for (std::map<int32_t, std::vector<int32_t> >::iterator it = fv.begin(); it != fv.end(); ++it, ++m) {
        fv_nodes[m] = it->first;
        fv_ptr[m] = at;
        for (size_t j = 0; j < it->second.size(); j++) fv_idx[at++] = it->second[j];
    }
    fv_ptr[m] = at;
    *n_nodes = m;
    return k;
}

}  // extern "C"
