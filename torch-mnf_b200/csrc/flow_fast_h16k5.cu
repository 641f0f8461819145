// flow_fast_h16k5.cu -- instantiation of the dim-2 register-resident flow kernel for hidden width 16, 5 spline bins.
#include "flow_fast.cuh"
namespace mnf {
MNF_FLOW_FAST_DEFINE(16, 5)
}
