// cvshim: stand-in for DBoW2/FeatureVector.h -- published layout: std::map<NodeId, std::vector<unsigned int>>, filled in
// ascending feature order by addFeature (what FeatureMatcher.cc:289-335 walks).
#pragma once
#include <map>
#include <vector>
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
public:
    void addFeature(NodeId id, unsigned int i_feature)
    {
        auto it = this->lower_bound(id);
        if (it != this->end() && it->first == id) it->second.push_back(i_feature);
        else this->insert(it, value_type(id, std::vector<unsigned int>(1, i_feature)));
    }
};
}
