// instantiation of the dim-2 flow kernel for hidden width 16, 8 spline bins
#include "flow_fast.cuh"
namespace mnf {
MNF_FLOW_FAST_DEFINE(16, 8)
}  // namespace mnf
