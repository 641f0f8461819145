// flow_lanes.cu -- small batches of dim-2 AffineHalfFlow stacks (BASELINE config 1: RNVP x9, h = 24 x 3, 4096 points per
// call): a call is bound by the LATENCY of one point's dependent chain (18 conditioner evaluations of ~1 200 FMAs each),
// not by throughput -- with one point per thread the kernel occupies 32 SMs for 30 us.  Here EIGHT lanes share a point:
// a lane owns H / 8 hidden units of every layer, the activations of a layer are exchanged through a per-warp
// shared-memory strip (st.shared -> __syncwarp -> 16-byte broadcast loads), every lane keeps the point's state
// redundantly, so nothing but the activations ever moves between lanes.  The chain per point is 8x shorter and 4096 points
// fill every SM (256 CTAs of 16 points).  The conditioner weights are read in the blob's own layout (weight[out][in]
// rows are what a lane needs); the next conditioner's 5 KB arrive by cp.async while the current one is evaluated.
#include <stdlib.h>

#include "flow_math.cuh"

namespace mnf {
namespace lanes {

constexpr int THREADS = 128, RING = 4;  // conditioners in flight in shared memory (cp.async ring)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int H>
struct Net {  // float offsets inside one conditioner 1 -> H -> H -> H -> 1 in the blob: per Linear weight[out][in], bias[out]
    static constexpr int W0 = 0, B0 = H, W1 = 2 * H, B1 = W1 + H * H, W2 = B1 + H, B2 = W2 + H * H, W3 = B2 + H, B3 = W3 + H;
    static constexpr int FLOATS = B3 + 1, CHUNKS = (FLOATS + 3) / 4;  // 16-byte chunks copied per net
};

// one conditioner on the scalar c for the point this lane belongs to; every lane of the point returns the output
template <int H, int LANES>
__device__ __forceinline__ float eval_net(const float *__restrict__ w, float c, float *strip, int sub) {
    constexpr int U = H / LANES;
    using N = Net<H>;
    float h[U], a[H];
#pragma unroll
    for (int u = 0; u < U; ++u) h[u] = leaky02(fmaf(w[N::W0 + U * sub + u], c, w[N::B0 + U * sub + u]));
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
        // all-gather of the layer input over the point's 8 lanes through the strip
#pragma unroll
        for (int u = 0; u < U; ++u) strip[U * sub + u] = h[u];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(strip + i);
            a[i] = v.x, a[i + 1] = v.y, a[i + 2] = v.z, a[i + 3] = v.w;
        }
        __syncwarp();  // everyone has read the strip before the next layer overwrites it
        const float *W = w + (layer == 0 ? N::W1 : N::W2), *B = w + (layer == 0 ? N::B1 : N::B2);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float *row = W + (U * sub + u) * H;
            float acc0 = B[U * sub + u], acc1 = 0.f;
#pragma unroll
            for (int i = 0; i < H; i += 4) {
                const float4 r = *reinterpret_cast<const float4 *>(row + i);
                acc0 = fmaf(r.x, a[i], acc0), acc1 = fmaf(r.y, a[i + 1], acc1);
                acc0 = fmaf(r.z, a[i + 2], acc0), acc1 = fmaf(r.w, a[i + 3], acc1);
            }
            h[u] = leaky02(acc0 + acc1);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) strip[U * sub + u] = h[u];
    __syncwarp();
    float o0 = w[N::B3], o1 = 0.f;
#pragma unroll
    for (int i = 0; i < H; i += 4) {  // the single output: every lane evaluates it (no broadcast needed afterwards)
        const float4 v = *reinterpret_cast<const float4 *>(strip + i), r = *reinterpret_cast<const float4 *>(w + N::W3 + i);
        o0 = fmaf(r.x, v.x, o0), o1 = fmaf(r.y, v.y, o1);
        o0 = fmaf(r.z, v.z, o0), o1 = fmaf(r.w, v.w, o1);
    }
    __syncwarp();
    return o0 + o1;
}

struct NetPlan {  // conditioners in execution order
    int n;
    int off[2 * MNF_MAX_OPS];
};

template <int H, int LANES>
__global__ void __launch_bounds__(THREADS) flow_lanes_kernel(const __grid_constant__ FlowProgram prog, const __grid_constant__ NetPlan nets,
                                                             const float *__restrict__ params, const float *__restrict__ x,
                                                             float *__restrict__ y, float *__restrict__ log_det,
                                                             float *__restrict__ base_lp, float *__restrict__ inter, long long n_rows,
                                                             int dir_flags) {
    using N = Net<H>;
    constexpr int NETF = N::CHUNKS * 4, PTS = THREADS / LANES;
    __shared__ __align__(16) float wbuf[RING][NETF];
    __shared__ __align__(16) float strips[PTS][H];
    const int tid = threadIdx.x, sub = tid & (LANES - 1), pl = tid / LANES;
    const int inverse = dir_flags & 1;
    const bool sum_lp = dir_flags & 2;
    const long long pt = (long long)blockIdx.x * PTS + pl;
    const bool live = pt < n_rows;
    float *strip = strips[pl];
    auto fetch = [&](int n) {  // conditioner n of the execution order into ring slot n % RING
        if (n < nets.n) {
            const float *src = params + nets.off[n];
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(wbuf[n % RING]);
            for (int i = tid; i < N::CHUNKS; i += THREADS) cp_async16(dst + 16u * i, src + 4 * i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int n = 0; n < RING - 1; ++n) fetch(n);
    float v0, v1, ld = 0.f;
    {
        const float2 q = ld_stream2(reinterpret_cast<const float2 *>(x) + (live ? pt : n_rows - 1));
        v0 = q.x, v1 = q.y;
    }
    int net = 0;
    auto next_net = [&]() -> const float * {  // weights of conditioner `net` are in place; RING - 1 more are on their way
        // (slot (net + RING - 1) % RING held conditioner net - 1: the barrier after its evaluation released it)
        fetch(net + RING - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
        __syncthreads();
        return wbuf[net % RING];
    };
#pragma unroll 1
    for (int kk = 0; kk < prog.n_ops; ++kk) {
        const mnf_flow_op &op = prog.ops[inverse ? prog.n_ops - 1 - kk : kk];
        if (op.type == MNF_OP_AFFINE_CONST) {
            const float s0 = params[op.aux_off], s1 = params[op.aux_off + 1], t0 = params[op.aux_off + 2], t1 = params[op.aux_off + 3];
            if (inverse) {  // affine_constant_flow.py:24
                v0 = (v0 - t0) * expf(-s0), v1 = (v1 - t1) * expf(-s1), ld -= s0 + s1;
            } else {  // affine_constant_flow.py:19
                v0 = v0 * expf(s0) + t0, v1 = v1 * expf(s1) + t1, ld += s0 + s1;
            }
        } else if (op.type == MNF_OP_GLOW) {
            const float *W = params + op.aux_off + (inverse ? 4 : 0);  // glow.py:28,36: v @ W
            const float n0 = fmaf(v1, W[2], v0 * W[0]), n1 = fmaf(v1, W[3], v0 * W[1]);
            v0 = n0, v1 = n1;
            ld += inverse ? -params[op.aux_off + 8] : params[op.aux_off + 8];
        } else {  // AffineHalfFlow (affine_half_flow.py:44-62)
            const bool parity = op.flags & MNF_FLAG_PARITY;
            const float cond = parity ? v1 : v0;
            float tr = parity ? v0 : v1, s = 0.f, t = 0.f;
            if (op.flags & MNF_FLAG_SCALE) {
                const float *w = next_net();
                s = eval_net<H, LANES>(w, cond, strip, sub);
                __syncthreads();  // every warp is done with this buffer before the fetch after next reuses it
                ++net;
            }
            if (op.flags & MNF_FLAG_SHIFT) {
                const float *w = next_net();
                t = eval_net<H, LANES>(w, cond, strip, sub);
                __syncthreads();
                ++net;
            }
            if (inverse) {  // affine_half_flow.py:54-56
                tr = (tr - t) / expf(s), ld -= s;
            } else {  // affine_half_flow.py:58
                tr = expf(s) * tr + t, ld += s;
            }
            if (parity) v0 = tr; else v1 = tr;
        }
        if (inter && live && sub == 0)
            st_stream2(reinterpret_cast<float2 *>(inter + ((size_t)kk * n_rows + pt) * 2), make_float2(v0, v1));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!live || sub != 0) return;
    float lp = fmaf(-0.5f, fmaf(v0, v0, v1 * v1), -1.8378770664093453f);  // -(D/2) log(2 pi), D = 2
    if (sum_lp) lp += ld;
    if (y) st_stream2(reinterpret_cast<float2 *>(y) + pt, make_float2(v0, v1));
    if (log_det) log_det[pt] = ld;
    if (base_lp) base_lp[pt] = lp;
}

template <int H, int LANES>
static int launch_hl(const FlowProgram &prog, const NetPlan &nets, const float *params, const float *x, float *y, float *log_det,
                    float *base_lp, float *inter, long long n_rows, int dir_flags, cudaStream_t st) {
    constexpr int PTS = THREADS / LANES;
    const unsigned blocks = (unsigned)((n_rows + PTS - 1) / PTS);
    flow_lanes_kernel<H, LANES><<<blocks, THREADS, 0, st>>>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags);
    return launch_status("flow_lanes_kernel");
}

// lanes per point by batch size: enough CTAs to cover the SMs, no more redundancy than that needs
template <int H>
static int launch_h(const FlowProgram &prog, const NetPlan &nets, const float *params, const float *x, float *y, float *log_det,
                    float *base_lp, float *inter, long long n_rows, int dir_flags, cudaStream_t st) {
    // measured on config 1 (us per bound call, kernel-bound): 8 lanes 22.6 up to 1024 rows, 36.9 at 4096; 4 lanes 34.9 up to
    // 4096; 2 lanes 41.0 up to 8192; the one-point-per-thread kernel 32.8 at 4096 (profiles/r02_cfg1_lanes.md)
    static const char *force = getenv("MNF_LANES");
    const int lanes = force ? atoi(force) : 8;
    if constexpr (H % 8 == 0) {
        if (lanes == 8) return launch_hl<H, 8>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags, st);
    }
    if (lanes >= 4) return launch_hl<H, 4>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags, st);
    return launch_hl<H, 2>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags, st);
}

}  // namespace lanes

// rows up to which the lane-split kernel is the default (beyond it the one-point-per-thread kernel is as fast: the SMs are
// covered and eight lanes do ~3x the instructions of one); an explicit request (variant 5) may go up to kLanesMaxRows
constexpr long long kLanesDefaultRows = 2048, kLanesMaxRows = 32768;

// returns 1 if the program / batch is not of this kernel's class (caller falls back)
int launch_flow_lanes(const mnf_flow_op *ops, int n_ops, const float *params, const float *x, float *y, float *log_det,
                      float *base_lp, float *inter, int64_t n_rows, int dim, int dir_flags, cudaStream_t stream, bool plan_only,
                      bool forced) {
    using namespace lanes;
    if (dim != 2 || n_ops < 1 || n_ops > MNF_MAX_OPS) return 1;
    if (!plan_only && (n_rows < 1 || n_rows > (forced ? kLanesMaxRows : kLanesDefaultRows))) return 1;
    const int inverse = dir_flags & 1;
    NetPlan nets;
    nets.n = 0;
    int H = 0;
    for (int kk = 0; kk < n_ops; ++kk) {
        const mnf_flow_op &op = ops[inverse ? n_ops - 1 - kk : kk];
        if (op.type == MNF_OP_AFFINE_CONST || op.type == MNF_OP_GLOW) continue;
        if (op.type != MNF_OP_AFFINE_HALF || op.n_lin != 4 || op.sizes[0] != 1 || op.sizes[4] != 1) return 1;
        const int h = op.sizes[1];
        if (op.sizes[2] != h || op.sizes[3] != h || (H && h != H)) return 1;
        H = h;
        for (int which = 0; which < 2; ++which) {
            if (!(op.flags & (which ? MNF_FLAG_SHIFT : MNF_FLAG_SCALE))) continue;
            if (op.net_off[which] % 4 != 0) return 1;  // 16-byte copies of the net
            nets.off[nets.n++] = op.net_off[which];
        }
    }
    if (nets.n == 0 || !(H == 8 || H == 16 || H == 24 || H == 32)) return 1;
    if (plan_only) return 0;
    if (((uintptr_t)params % 16) != 0 || ((uintptr_t)x % 8) != 0 || (y && ((uintptr_t)y % 8) != 0) || (inter && ((uintptr_t)inter % 8) != 0))
        return 1;
    FlowProgram prog;
    prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) prog.ops[k] = ops[k];
    switch (H) {
        case 8: return lanes::launch_h<8>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags & 3, stream);
        case 16: return lanes::launch_h<16>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags & 3, stream);
        case 24: return lanes::launch_h<24>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags & 3, stream);
        default: return lanes::launch_h<32>(prog, nets, params, x, y, log_det, base_lp, inter, n_rows, dir_flags & 3, stream);
    }
}

}  // namespace mnf
