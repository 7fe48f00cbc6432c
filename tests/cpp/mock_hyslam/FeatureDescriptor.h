// TEST DOUBLE mirroring hySLAM src/features/low_level/FeatureDescriptor.h:26-38.
#pragma once
#include <DescriptorDistance.h>
#include <memory>
namespace HYSLAM {
class FeatureDescriptor {
public:
    FeatureDescriptor() {}
    FeatureDescriptor(cv::Mat desc, std::shared_ptr<DescriptorDistance> distfunc_) : descriptor(desc.clone()), distfunc(distfunc_), is_empty(false) {}
    float distance(const FeatureDescriptor &d2) const { return distfunc->distance(descriptor, d2.rawDescriptor()); }
    cv::Mat rawDescriptor() const { return descriptor.clone(); }
    bool isEmpty() { return is_empty; }
private:
    cv::Mat descriptor; std::shared_ptr<DescriptorDistance> distfunc; bool is_empty = true;
};
}
