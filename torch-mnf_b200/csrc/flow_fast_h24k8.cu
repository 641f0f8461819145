// instantiation of the dim-2 flow kernel for hidden width 24, 8 spline bins
#include "flow_fast.cuh"
namespace mnf {
MNF_FLOW_FAST_DEFINE(24, 8)
}  // namespace mnf
