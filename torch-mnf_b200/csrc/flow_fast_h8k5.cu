// instantiation of the dim-2 flow kernel for hidden width 8, 5 spline bins
#include "flow_fast.cuh"
namespace mnf {
MNF_FLOW_FAST_DEFINE(8, 5)
}  // namespace mnf
