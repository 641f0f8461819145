// flow_fast.cu -- host-side planning and dispatch for the dim-2 register-resident flow kernel
// (templates in flow_fast.cuh; one translation unit per (H, K) instantiation so they build
// in parallel).
#include "flow_fast.cuh"

namespace mnf {

int launch_fast_16_8(MNF_FLOW_FAST_ARGS);
int launch_fast_24_8(MNF_FLOW_FAST_ARGS);
int launch_fast_8_5(MNF_FLOW_FAST_ARGS);

// ---- host side: plan + launch ---------------------------------------------------------

struct FastPlan {
    int H = 0, K = 0;
    bool ok = false;
    FastLayout lay;
};

static FastPlan plan_fast(const mnf_flow_op *ops, int n_ops, int dim, int variant) {
    FastPlan p;
    if (dim != 2 || n_ops < 1) return p;
    int H = 0, K = 0, slots = 0;
    for (int k = 0; k < n_ops; ++k) {
        const mnf_flow_op &op = ops[k];
        p.lay.net_slot[k][0] = p.lay.net_slot[k][1] = 0;
        if (op.type == MNF_OP_AFFINE_CONST || op.type == MNF_OP_GLOW) continue;
        if (op.type != MNF_OP_AFFINE_HALF && op.type != MNF_OP_NSF_CL) return p;
        if (op.n_lin != 4 || op.sizes[0] != 1) return p;
        const int h = op.sizes[1];
        if (op.sizes[2] != h || op.sizes[3] != h) return p;
        if (H && h != H) return p;
        H = h;
        int n_out = 1;
        if (op.type == MNF_OP_NSF_CL) {
            if (K && op.K != K) return p;
            K = op.K;
            n_out = 3 * K - 1;
        }
        if (op.sizes[4] != n_out) return p;
        const int per_net = fast_net_slots(h, n_out, variant);
        for (int which = 0; which < 2; ++which) {
            p.lay.net_slot[k][which] = slots;
            slots += per_net;
        }
    }
    if (H == 0) return p;  // nothing for this kernel to accelerate -> generic handles it
    if (K == 0) K = 8;
    const bool have = (H == 16 && K == 8) || (H == 24 && K == 8) || (H == 8 && K == 5);
    if (!have) return p;
    p.H = H;
    p.K = K;
    p.lay.total_slots = slots;
    p.ok = true;
    return p;
}

// default: constant-bank weights for large batches (3 segment launches + copies amortise), the
// single-launch shared-memory variant for small ones
constexpr long long kCbankMinRows = 1 << 16;
constexpr int kNumVariants = 4;

// returns 1 if the program is not eligible (caller falls back to the generic kernel)
int launch_flow_fast(const mnf_flow_op *ops, int n_ops, const float *params, int64_t n_params, const float *x,
                     float *y, float *log_det, float *base_lp, float *inter, int64_t n_rows, int dim, int inverse,
                     int variant, float *workspace, cudaStream_t stream, bool plan_only) {
    bool has_spline = false;
    for (int k = 0; k < n_ops; ++k) has_spline |= ops[k].type == MNF_OP_NSF_CL;
    // measured (r01): the constant-bank variant wins on spline stacks (4.10 vs 5.52 ms per 2^24 points) and loses
    // slightly on pure AffineHalfFlow stacks (16.1 vs 15.0 ms), where a segment still holds two conditioners
    int mode = (variant >= 0 && variant < kNumVariants) ? variant : (n_rows >= kCbankMinRows && has_spline ? 3 : 2);
    FastPlan p = plan_fast(ops, n_ops, dim, mode);
    if (!p.ok) return 1;
    if (mode == 3 && inter && (n_rows % 2)) return 1;
    const DeviceProps *dp = plan_only ? nullptr : device_props();
    const size_t smem_bytes = (size_t)p.lay.total_slots * sizeof(float);
    if (plan_only) return (mode == 3 || smem_bytes <= 227 * 1024) ? 0 : 1;
    MNF_REQUIRE(dp != nullptr, MNF_E_DEVICE, "no CUDA device");
    if (mode != 3 && smem_bytes > (size_t)dp->smem_optin) return 1;
    MNF_REQUIRE(((uintptr_t)x % 16) == 0 && (!y || ((uintptr_t)y % 16) == 0), MNF_E_ALIGN,
                "x and y must be 16-byte aligned for the dim-2 kernel");
    MNF_REQUIRE(!log_det || ((uintptr_t)log_det % 8) == 0, MNF_E_ALIGN, "log_det must be 8-byte aligned");
    MNF_REQUIRE(!base_lp || ((uintptr_t)base_lp % 8) == 0, MNF_E_ALIGN, "base_log_prob must be 8-byte aligned");
    if (inter && (n_rows % 2)) return 1;  // per-flow outputs of an odd batch are not 16 B aligned: generic path
    MNF_REQUIRE(!inter || ((uintptr_t)inter % 16) == 0, MNF_E_ALIGN, "intermediates must be 16-byte aligned");
    FlowProgram prog;
    prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) prog.ops[k] = ops[k];
    (void)n_params;
    if (p.H == 16 && p.K == 8)
        return launch_fast_16_8(mode, prog, p.lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, inverse,
                                workspace, dp, stream);
    if (p.H == 24 && p.K == 8)
        return launch_fast_24_8(mode, prog, p.lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, inverse,
                                workspace, dp, stream);
    if (p.H == 8 && p.K == 5)
        return launch_fast_8_5(mode, prog, p.lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, inverse,
                               workspace, dp, stream);
    return 1;
}

}  // namespace mnf
