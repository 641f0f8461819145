// flow_fast_h8k8.cu -- instantiation of the dim-2 register-resident flow kernel for hidden width 8, 8 spline bins.
#include "flow_fast.cuh"
namespace mnf {
MNF_FLOW_FAST_DEFINE(8, 8)
}
