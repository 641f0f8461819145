// flow_pl.cu -- dim-2 coupling stacks (BASELINE configs 1 and 2) with every conditioner evaluated as what it is: a
// PIECEWISE-LINEAR function of ONE scalar.
//
// In two dimensions the conditioner of AffineHalfFlow (affine_half_flow.py:46-50) and of NSF_CL (spline_flow.py:249-257)
// sees a single coordinate c, and an MLP of Linear / LeakyReLU layers (mlp.py:4-12) is piecewise linear in its input:
// along the real line it has a finite list of breakpoints (the zero crossings of every hidden pre-activation; about one
// per hidden unit in practice: ~50 for 1-16-16-16-23, ~70 for 1-24-24-24-1) and between two breakpoints EVERY output is
// A c + B.  So the whole net -- not only its first two layers (flow_fast.cuh, flow_tc.cu) -- is a table:
//
//   flow_pl_build_kernel   one CTA per conditioner (an AffineHalfFlow's s- and t-net share their input and one table): finds
//                          the breakpoints layer by layer in fp64 (on every current piece each pre-activation is affine, so
//                          it has at most one zero there), sorts them, and writes per piece the slope and the value at the
//                          piece's origin of every output.  Exact up to fp64 rounding -- closer to the real-number net than
//                          an fp32 evaluation of its layers.
//   flow_pl_kernel<K>      one thread per point, the whole stack in one launch: a conditioner is a branch-free binary
//                          search over <= 511 sorted breakpoints in shared memory and one FMA per output
//                          (out = V_i + A_i (c - origin_i)); the spline (flow_math.cuh) then runs on the raw outputs, and only
//                          the two knot derivatives of the bin the point falls into are ever formed.
//
// Per point and conditioner: ~60 instructions instead of 5 376 multiply-adds (or three MMA round trips).  No tensor core
// and no FMA chain is left to feed: what remains is the spline arithmetic and 12 B/pt of HBM traffic.
// A table with more than 511 breakpoints (e.g. a pair of 1-64-64-64-64-64-1 nets; the bound is pieces x width per layer) is flagged by the
// builder and evaluated layer by layer in fp32 by the threads that need it.  The library owns no device memory: the
// tables live in the caller's workspace (built per call) or in the image of mnf_flow_stack_stage (built per parameter
// version).
#include "flow_math.cuh"
#include "tc_common.cuh"

namespace mnf {
namespace fpl {
using tc::smem_u32; using tc::mbar_init; using tc::mbar_arrive; using tc::mbar_wait; using tc::bulk_load;

constexpr int PMAX = 511;       // breakpoints per table
constexpr int BP_FLOATS = 512;  // sorted breakpoints, padded with +inf (search array)
constexpr int MAX_GROUPS = 2 * MNF_MAX_OPS;
constexpr int HDR_INTS = 4;     // per table: breakpoints, overflow flag, 2 spare
constexpr int HDR_FLOATS = MAX_GROUPS * HDR_INTS;
constexpr int MAX_H = 64;       // hidden width the in-kernel fp32 fallback holds

__host__ __device__ constexpr int round4i(int n) { return (n + 3) / 4 * 4; }
// NSF_CL row: K float4 (A_2j, A_2j+1, V_2j, V_2j+1) for the 2K width / height outputs | K - 1 float2 (A, V) of the knot
// derivatives | origin.  AffineHalfFlow row: (A_s, A_t, V_s, V_t) | origin.  Strides are an odd number of float4 so that
// rows of different pieces spread over the banks.
__host__ __device__ constexpr int nsf_doff(int K) { return 4 * K; }
__host__ __device__ constexpr int nsf_org(int K) { return 4 * K + round4i(2 * (K - 1)); }
__host__ __device__ constexpr int nsf_stride4(int K) { return (nsf_org(K) / 4 + 1) | 1; }
constexpr int AFF_ORG = 4, AFF_STRIDE4 = 3;
__host__ __device__ constexpr int region_floats(int stride4) { return BP_FLOATS + (PMAX + 1) * 4 * stride4; }

struct GroupList {
    int n;
    int region[MAX_GROUPS];            // float offset of the table in the image
    unsigned char op[MAX_GROUPS];      // flow it belongs to
    unsigned char which[MAX_GROUPS];   // NSF_CL: 0 = f1, 1 = f2
    signed char group_of[MNF_MAX_OPS][2];
};

// ---------------------------------------------------------------------------------------------------------
// builder
// ---------------------------------------------------------------------------------------------------------
constexpr int BW = 8;  // warps per builder CTA

// a point strictly inside piece i of the n sorted breakpoints (fixes the sign of every pre-activation on the piece)
__device__ __forceinline__ double piece_point(const double *bp, int n, int i, double &lo, double &hi) {
    lo = i > 0 ? bp[i - 1] : -INFINITY;
    hi = i < n ? bp[i] : INFINITY;
    if (n == 0) return 0.0;
    if (i == 0) return hi - 1.0 - fabs(hi);
    if (i == n) return lo + 1.0 + fabs(lo);
    return 0.5 * lo + 0.5 * hi;
}

// One warp: affine maps (a c + b) of the pre-activations of Linear layer `upto` (0-based) of a net on the piece that
// contains m.  Two ping-pong vectors per warp; returns the index of the one holding the result.
__device__ int net_pre_maps(const float *__restrict__ net, const int *sizes, int upto, double m, double (*va)[MNF_MAX_HIDDEN],
                            double (*vb)[MNF_MAX_HIDDEN], int lane) {
    const double slope_neg = (double)0.2f;  // LeakyReLU(0.2) on fp32 tensors (mlp.py:9)
    int cur = 0;
    if (lane == 0) va[0][0] = 1.0, vb[0][0] = 0.0;  // layer input: c itself
    __syncwarp();
    const float *p = net;
    for (int t = 0; t <= upto; ++t) {
        const int n_in = sizes[t], n_out = sizes[t + 1];
        const float *bias = p + n_in * n_out;
        for (int k = lane; k < n_out; k += 32) {
            double A = 0.0, B = (double)bias[k];
            const float *w = p + (size_t)k * n_in;
            for (int j = 0; j < n_in; ++j) {
                const double wj = (double)w[j];
                A = fma(wj, va[cur][j], A);
                B = fma(wj, vb[cur][j], B);
            }
            if (t < upto) {  // LeakyReLU with the sign the pre-activation has on this piece
                const double s = fma(A, m, B) > 0.0 ? 1.0 : slope_neg;
                A *= s, B *= s;
            }
            va[cur ^ 1][k] = A, vb[cur ^ 1][k] = B;
        }
        __syncwarp();
        cur ^= 1;
        p += n_in * n_out + n_out;
    }
    return cur;
}

__global__ void __launch_bounds__(32 * BW) flow_pl_build_kernel(const __grid_constant__ FlowProgram prog, const __grid_constant__ GroupList gl,
                                                                const float *__restrict__ params, float *__restrict__ image) {
    __shared__ double s_bp[512], s_tmp[512];
    __shared__ double s_va[BW][2][MNF_MAX_HIDDEN], s_vb[BW][2][MNF_MAX_HIDDEN];
    __shared__ int s_n, s_ncand;
    const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nt = blockDim.x;
    const mnf_flow_op &op = prog.ops[gl.op[g]];
    const bool nsf = op.type == MNF_OP_NSF_CL;
    const int K = op.K, L = op.n_lin, n_out = op.sizes[L];
    int n_nets = 0, net_off[2], net_slot[2];
    if (nsf) {
        net_off[0] = op.net_off[gl.which[g]], net_slot[0] = 0, n_nets = 1;
    } else {  // AffineHalfFlow: s_net and t_net read the same coordinate (affine_half_flow.py:49-50)
        if (op.flags & MNF_FLAG_SCALE) net_off[n_nets] = op.net_off[0], net_slot[n_nets++] = 0;
        if (op.flags & MNF_FLAG_SHIFT) net_off[n_nets] = op.net_off[1], net_slot[n_nets++] = 1;
    }
    if (tid == 0) s_n = 0, s_ncand = 0;
    __syncthreads();

    bool over = false;
    for (int l = 1; l < L; ++l) {  // new breakpoints: zeros of the pre-activations of hidden layer l inside the current pieces
        const int n = s_n;
        for (int i = warp; i <= n; i += BW) {
            double lo, hi;
            const double m = piece_point(s_bp, n, i, lo, hi);
            for (int q = 0; q < n_nets; ++q) {
                const int cur = net_pre_maps(params + net_off[q], op.sizes, l - 1, m, s_va[warp], s_vb[warp], lane);
                for (int k = lane; k < op.sizes[l]; k += 32) {
                    const double a = s_va[warp][cur][k], b = s_vb[warp][cur][k];
                    if (a != 0.0) {
                        const double tz = -b / a;
                        // (a kink beyond the fp32 range -- a first-layer weight of ~1e-39 -- is no kink for any fp32 input,
                        // and would give its piece an infinite origin)
                        if (tz > lo && tz < hi && fabs(tz) < 3.0e38) {
                            const int pos = atomicAdd(&s_ncand, 1);
                            if (pos < 512) s_tmp[pos] = tz;
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        const int nc = s_ncand, tot = n + nc;
        if (tot > PMAX) {
            over = true;
            break;
        }
        for (int e = tid; e < nc; e += nt) s_bp[n + e] = s_tmp[e];
        __syncthreads();
        for (int e = tid; e < tot; e += nt) {  // rank sort (ties by index)
            const double v = s_bp[e];
            int r = 0;
            for (int u = 0; u < tot; ++u) r += (s_bp[u] < v || (s_bp[u] == v && u < e)) ? 1 : 0;
            s_tmp[r] = v;
        }
        __syncthreads();
        for (int e = tid; e < tot; e += nt) s_bp[e] = s_tmp[e];
        if (tid == 0) s_n = tot, s_ncand = 0;
        __syncthreads();
    }

    float *region = image + gl.region[g];
    int *hdr = reinterpret_cast<int *>(image) + g * HDR_INTS;
    const int n = over ? 0 : s_n;
    if (tid == 0) hdr[0] = n, hdr[1] = over ? 1 : 0, hdr[2] = 0, hdr[3] = 0;
    for (int e = tid; e < BP_FLOATS; e += nt) region[e] = e < n ? (float)s_bp[e] : INFINITY;
    if (over) return;
    const int stride = 4 * (nsf ? nsf_stride4(K) : AFF_STRIDE4);
    float *rows = region + BP_FLOATS;
    for (int i = warp; i <= n; i += BW) {
        double lo, hi;
        const double m = piece_point(s_bp, n, i, lo, hi);
        // origin: the (fp32) breakpoint the piece starts at -- the kernel forms c - origin exactly there
        const float org = n == 0 ? 0.f : (float)(i == 0 ? s_bp[0] : s_bp[i - 1]);
        float *row = rows + (size_t)i * stride;
        for (int e = lane; e < stride; e += 32) row[e] = 0.f;
        __syncwarp();
        for (int q = 0; q < n_nets; ++q) {
            const int cur = net_pre_maps(params + net_off[q], op.sizes, L - 1, m, s_va[warp], s_vb[warp], lane);
            for (int o = lane; o < n_out; o += 32) {
                const double A = s_va[warp][cur][o], V = fma(A, (double)org, s_vb[warp][cur][o]);
                int ia, iv;
                if (!nsf) ia = net_slot[q], iv = 2 + net_slot[q];
                else if (o < 2 * K) ia = 4 * (o >> 1) + (o & 1), iv = ia + 2;
                else ia = nsf_doff(K) + 2 * (o - 2 * K), iv = ia + 1;
                row[ia] = (float)A, row[iv] = (float)V;
            }
            __syncwarp();
        }
        if (lane == 0) row[nsf ? nsf_org(K) : AFF_ORG] = org;
    }
}

// ---------------------------------------------------------------------------------------------------------
// evaluation
// ---------------------------------------------------------------------------------------------------------
struct Params {
    FlowProgram prog;  // module order
    GroupList gl;
    const float *params, *x, *image;
    float *y, *log_det, *base_lp, *inter;
    long long n_rows;
    int dir_flags, smem_floats;
    mnf_gather_out gather;
};

__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// exact-fp32 evaluation of a net layer by layer (tables that overflowed): parameter blob, per Linear weight[out][in], bias[out]
__device__ __noinline__ void mlp_fp32(const float *__restrict__ net, int n_lin, const int *sizes, float c, float *out) {
    float h[2][MAX_H];
    h[0][0] = c;
    int cur = 0;
    const float *p = net;
    for (int t = 0; t < n_lin; ++t) {
        const int n_in = sizes[t], n_o = sizes[t + 1];
        for (int k = 0; k < n_o; ++k) {
            float acc = p[n_in * n_o + k];
            for (int j = 0; j < n_in; ++j) acc = fmaf(p[k * n_in + j], h[cur][j], acc);
            if (t + 1 < n_lin) h[cur ^ 1][k] = leaky02(acc);
            else out[k] = acc;
        }
        cur ^= 1;
        p += n_in * n_o + n_o;
    }
}

// (values in and out by value: a reference parameter of a non-inlined function would pin the caller's accumulators to
// local memory on the table path as well)
template <int K>
__device__ __noinline__ float2 nsf_fallback(const float *__restrict__ net, int n_lin, const int *sizes, float B, float edge_deriv, float c,
                                            bool inverse, float tr) {
    float raw[3 * K - 1], ld = 0.f;
    mlp_fp32(net, n_lin, sizes, c, raw);
    rq_spline<K, true>(raw, K, B, edge_deriv, inverse, tr, ld);
    return make_float2(tr, ld);
}
__device__ __noinline__ float mlp_fp32_scalar(const float *__restrict__ net, int n_lin, const int *sizes, float c) {
    float out[1];
    mlp_fp32(net, n_lin, sizes, c, out);
    return out[0];
}

template <int S>
__device__ __forceinline__ void search_steps(uint32_t &a, float c) {
    float t;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(t) : "r"(a), "n"(4 * (S - 1)));
    if (c > t) a += 4u * S;
    if constexpr (S > 1) search_steps<S / 2>(a, c);
}

// piece of c: count of breakpoints below it, branch-free over the +inf padded array.  SM: the table is in shared memory and
// the search walks a byte address (load, compare, predicated add per step).
template <bool SM>
__device__ __forceinline__ int find_piece(const float *tb, float c) {
    if constexpr (SM) {
        const uint32_t base = smem_u32(tb);
        uint32_t a = base;
        search_steps<(PMAX + 1) / 2>(a, c);
        return (int)((a - base) >> 2);
    } else {
        int pos = 0;
#pragma unroll
        for (int s = (PMAX + 1) / 2; s >= 1; s >>= 1) pos += (c > tb[pos + s - 1]) ? s : 0;
        return pos;
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Knots of one spline axis in units of B: t[0] = -1, t[K] = 1.  The formulas of flow_math.cuh::spline_knots (both
// softmaxes of spline_flow.py:253-255 and :95, floor :96-97, cumsum :99, pinned ends :100-101) with the cumulative sum
// taken over the second softmax's terms while they are summed: x_{k+1} / B = -1 + 2 (k + 1) min_bin + (2 span / sum f) * (f_0 + .. + f_k),
// one FFMA per knot.
template <int K>
__device__ __forceinline__ void unit_knots(const float *raw, float twoB, float *t) {
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float span = 1.f - kMinBin * (float)K;
    float m = raw[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, raw[k]);
    const float mneg = -m * LOG2E;
    float e[K], sum = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        e[k] = ex2_approx(fmaf(raw[k], LOG2E, mneg));
        sum += e[k];
    }
    const float sc = twoB * frcp<true>(sum) * LOG2E;  // first softmax scaled by 2B, in log2 units
    const float off = twoB > 80.f ? -sc : 0.f;        // arguments of the second softmax lie in [0, 2B]
    float pre = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pre += ex2_approx(fmaf(e[k], sc, off));
        e[k] = pre;
    }
    const float c1 = 2.f * span * frcp<true>(pre);
    t[0] = -1.f;
#pragma unroll
    for (int k = 0; k + 1 < K; ++k) t[k + 1] = fmaf(e[k], c1, -1.f + 2.f * (float)(k + 1) * kMinBin);
    t[K] = 1.f;
}

// The rational-quadratic spline of flow_math.cuh::rq_spline (FAST flavour) for a point already known to lie in [-B, B],
// evaluated in units of B (the bin ratios, the knot derivatives and the log-det do not depend on the unit).  The raw
// derivatives of the two knots around the point's bin are fetched through `deriv(j)` once the bin is known -- the
// conditioner's other K - 3 derivative outputs are never formed.
template <int K, class DerivFn>
__device__ __forceinline__ void rq_spline_lazy(const float *raw, float B, float inv_B, float edge_deriv, bool inverse, float &v, float &ld,
                                               DerivFn deriv) {
    float cw[K + 1], ch[K + 1];
    unit_knots<K>(raw, 2.f * B, cw);
    unit_knots<K>(raw + K, 2.f * B, ch);
    const float u = v * inv_B;
    const float *sk = inverse ? ch : cw;
    // bin = number of interior knots <= u (spline_flow.py:22-24,115-118; u lies inside [t_0, t_K], so the count needs no
    // clamp); the knots of both axes around it are picked with the same predicates
    int idx = 0;
    float xk = cw[0], xk1 = cw[1], yk = ch[0], yk1 = ch[1];
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const bool ge = u >= sk[k];
        idx += ge ? 1 : 0;
        xk = ge ? cw[k] : xk, xk1 = ge ? cw[k + 1] : xk1;
        yk = ge ? ch[k] : yk, yk1 = ge ? ch[k + 1] : yk1;
    }
    const float wk = xk1 - xk, hk = yk1 - yk;  // spline_flow.py:102,113
    // interior knot derivatives (spline_flow.py:256,104), the padded constant at the two ends (:46-49); the loads are
    // unconditional on a clamped index so that lanes in different bins do not diverge
    const float r0 = deriv(max(idx - 1, 0)), r1 = deriv(min(idx, K - 2));
    const float dk = (idx > 0) ? knot_derivative<true>(r0) : edge_deriv;
    const float dk1 = (idx < K - 1) ? knot_derivative<true>(r1) : edge_deriv;
    const float sk_ = __fdividef(hk, wk);  // spline_flow.py:123
    const float dsum = dk + dk1 - 2.f * sk_;
    if (inverse) {  // spline_flow.py:133-162
        const float dy = u - yk;
        const float a = dy * dsum + hk * (sk_ - dk);
        const float b = hk * dk - dy * dsum;
        const float c = -sk_ * dy;
        const float disc = fmaxf(b * b - 4.f * a * c, 0.f);
        const float root = __fdividef(2.f * c, -b - sqrt_approx(disc));
        v = (root * wk + xk) * B;
        const float tt = root * (1.f - root);
        const float den = sk_ + dsum * tt;
        const float omr = 1.f - root;
        const float num = (sk_ * sk_) * (dk1 * (root * root) + 2.f * sk_ * tt + dk * (omr * omr));
        const float rd = frcp<true>(den);
        ld -= __logf(num * rd * rd);
    } else {  // spline_flow.py:163-179
        const float th = __fdividef(u - xk, wk);
        const float tt = th * (1.f - th);
        const float numer = hk * (sk_ * (th * th) + dk * tt);
        const float den = sk_ + dsum * tt;
        const float omt = 1.f - th;
        const float num = (sk_ * sk_) * (dk1 * (th * th) + 2.f * sk_ * tt + dk * (omt * omt));
        const float rd = frcp<true>(den);
        v = fmaf(numer, rd, yk) * B;
        ld += __logf(num * rd * rd);
    }
}

constexpr int THREADS = 1024;

// The stack as the point loop walks it (execution order, built once per CTA): runs of AffineConstantFlow / ActNormFlow / Glow
// are composed into ONE affine map v -> v M + b with a constant log-det (fp64; not when per-flow outputs are requested),
// the coupling flows carry their constants and table indices -- the loop never touches the descriptors again.
struct Exec {
    int type;   // 0 affine map, 1 AffineHalfFlow, 2 NSF_CL
    int k;      // flow index (parameters of the fp32 fallback)
    int g0, g1; // tables: AffineHalfFlow g0; NSF_CL g0 = first step's, g1 = second step's
    int flag;   // AffineHalfFlow: parity; NSF_CL: the first step is f1 (forward direction)
    int pad[3];
    float c[8]; // affine: m00 m01 m10 m11 | b0 b1 dld - ; NSF_CL: B 1/B edge_deriv - ; AffineHalfFlow: -
};
static_assert(sizeof(Exec) == 64, "two 16-byte rows of ints, two of floats");

__device__ __forceinline__ void build_exec(const Params &p, Exec *ex, int *n_exec, int lane) {
    const int n = p.prog.n_ops, inverse = p.dir_flags & 1;
    if (lane < n) {
        const int k = inverse ? n - 1 - lane : lane;
        const mnf_flow_op &op = p.prog.ops[k];
        Exec e{};
        e.k = k;
        if (op.type == MNF_OP_AFFINE_CONST) {
            const double s0 = p.params[op.aux_off], s1 = p.params[op.aux_off + 1];
            const double t0 = p.params[op.aux_off + 2], t1 = p.params[op.aux_off + 3];
            if (inverse) {  // (v - t) exp(-s), affine_constant_flow.py:24
                const double e0 = exp(-s0), e1 = exp(-s1);
                e.c[0] = (float)e0, e.c[3] = (float)e1, e.c[4] = (float)(-t0 * e0), e.c[5] = (float)(-t1 * e1), e.c[6] = (float)(-(s0 + s1));
            } else {  // v exp(s) + t, affine_constant_flow.py:19
                e.c[0] = (float)exp(s0), e.c[3] = (float)exp(s1), e.c[4] = (float)t0, e.c[5] = (float)t1, e.c[6] = (float)(s0 + s1);
            }
        } else if (op.type == MNF_OP_GLOW) {  // v @ W (glow.py:28) or v @ W^-1 (glow.py:36)
            const float *W = p.params + op.aux_off + (inverse ? 4 : 0);
            e.c[0] = W[0], e.c[1] = W[1], e.c[2] = W[2], e.c[3] = W[3];
            e.c[6] = inverse ? -p.params[op.aux_off + 8] : p.params[op.aux_off + 8];
        } else if (op.type == MNF_OP_AFFINE_HALF) {
            e.type = 1, e.g0 = p.gl.group_of[k][0], e.flag = (op.flags & MNF_FLAG_PARITY) ? 1 : 0;
        } else {  // forward: f1 then f2 (spline_flow.py:249-266); inverse: f2 then f1 (:268-285)
            e.type = 2, e.flag = inverse ? 0 : 1;
            e.g0 = p.gl.group_of[k][inverse ? 1 : 0], e.g1 = p.gl.group_of[k][inverse ? 0 : 1];
            e.c[0] = op.bound, e.c[1] = 1.f / op.bound, e.c[2] = op.edge_deriv;
        }
        ex[lane] = e;
    }
    __syncwarp();
    if (lane == 0) {
        int j = 0;
        bool open = false;  // ex[j - 1] is an affine map that may still absorb the next one
        double M[4], b[2], ld;
        for (int i = 0; i < n; ++i) {
            const Exec e = ex[i];
            if (e.type != 0 || p.inter) {
                ex[j++] = e, open = false;
                continue;
            }
            if (!open) {
                for (int q = 0; q < 4; ++q) M[q] = e.c[q];
                b[0] = e.c[4], b[1] = e.c[5], ld = e.c[6];
                ++j, open = true;
            } else {  // (v M + b) W + t = v (M W) + (b W + t)
                const double W0 = e.c[0], W1 = e.c[1], W2 = e.c[2], W3 = e.c[3];
                const double n00 = M[0] * W0 + M[1] * W2, n01 = M[0] * W1 + M[1] * W3;
                const double n10 = M[2] * W0 + M[3] * W2, n11 = M[2] * W1 + M[3] * W3;
                const double nb0 = b[0] * W0 + b[1] * W2 + e.c[4], nb1 = b[0] * W1 + b[1] * W3 + e.c[5];
                M[0] = n00, M[1] = n01, M[2] = n10, M[3] = n11, b[0] = nb0, b[1] = nb1, ld += e.c[6];
            }
            Exec f{};
            for (int q = 0; q < 4; ++q) f.c[q] = (float)M[q];
            f.c[4] = (float)b[0], f.c[5] = (float)b[1], f.c[6] = (float)ld;
            ex[j - 1] = f;
        }
        *n_exec = j;
    }
}

// SM: every table of the program sits in shared memory (plain LDS); otherwise a table is read through a generic pointer
// (shared memory if it fitted, the image in global memory if not).
template <int K, bool SM>
__device__ __forceinline__ void run_points(const Params &p, const float *smem, const int *s_off, const int *s_over, const Exec *ex,
                                           int n_exec) {
    const bool sum_lp = p.dir_flags & 2;
    auto table = [&](int g) -> const float * {
        if constexpr (SM) return smem + s_off[g];
        const int o = s_off[g];
        return o >= 0 ? smem + o : p.image + p.gl.region[g];
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.n_rows; r += stride) {
        float v0, v1, ld = 0.f;
        {
            const float2 xin = ld_stream2(reinterpret_cast<const float2 *>(p.x) + r);
            v0 = xin.x, v1 = xin.y;
        }
#pragma unroll 1
        for (int kk = 0; kk < n_exec; ++kk) {
            const int4 ei = *reinterpret_cast<const int4 *>(&ex[kk]);  // type, k, g0, g1
            if (ei.x == 0) {
                const float4 m = *reinterpret_cast<const float4 *>(ex[kk].c), b = *reinterpret_cast<const float4 *>(ex[kk].c + 4);
                const float n0 = fmaf(v1, m.z, fmaf(v0, m.x, b.x)), n1 = fmaf(v1, m.w, fmaf(v0, m.y, b.y));
                v0 = n0, v1 = n1, ld += b.z;
            } else if (ei.x == 1) {
                const mnf_flow_op &op = p.prog.ops[ei.y];
                const bool parity = ex[kk].flag != 0;
                const float c = parity ? v1 : v0;  // affine_half_flow.py:46-50
                float tr = parity ? v0 : v1;
                const int g = ei.z;
                float s = 0.f, t = 0.f;
                if (g < 0) {  // neither net: s = t = 0
                } else if (!s_over[g]) {
                    const float *tb = table(g);
                    const float4 *row = reinterpret_cast<const float4 *>(tb + BP_FLOATS) + find_piece<SM>(tb, c) * AFF_STRIDE4;
                    const float4 q = row[0];
                    const float d = c - row[1].x;
                    s = fmaf(q.x, d, q.z), t = fmaf(q.y, d, q.w);
                } else {
                    if (op.flags & MNF_FLAG_SCALE) s = mlp_fp32_scalar(p.params + op.net_off[0], op.n_lin, op.sizes, c);
                    if (op.flags & MNF_FLAG_SHIFT) t = mlp_fp32_scalar(p.params + op.net_off[1], op.n_lin, op.sizes, c);
                }
                if (p.dir_flags & 1) {  // affine_half_flow.py:54-56
                    tr = (tr - t) / expf(s), ld -= s;
                } else {  // affine_half_flow.py:58
                    tr = expf(s) * tr + t, ld += s;
                }
                if (parity) v0 = tr; else v1 = tr;
            } else {
                const float4 cs = *reinterpret_cast<const float4 *>(ex[kk].c);  // B, 1/B, edge_deriv
                const bool first_f1 = ex[kk].flag != 0;
#pragma unroll 1
                for (int step = 0; step < 2; ++step) {
                    const bool use_f1 = (step == 0) == first_f1;
                    const float c = use_f1 ? v0 : v1;
                    float tr = use_f1 ? v1 : v0;
                    const float B = cs.x;
                    if (tr >= -B && tr <= B) {  // identity tails (also NaN), spline_flow.py:40,51-52: the conditioner is not needed
                        const int g = step ? ei.w : ei.z;
                        if (!s_over[g]) {
                            const float *tb = table(g);
                            const float4 *row = reinterpret_cast<const float4 *>(tb + BP_FLOATS) + find_piece<SM>(tb, c) * nsf_stride4(K);
                            const float d = c - row[nsf_org(K) / 4].x;
                            const float2 dd = make_float2(d, d);
                            float raw[2 * K];
#pragma unroll
                            for (int j = 0; j < K; ++j) {
                                const float4 q = row[j];
                                const float2 o = f2_fma(make_float2(q.x, q.y), dd, make_float2(q.z, q.w));
                                raw[2 * j] = o.x, raw[2 * j + 1] = o.y;
                            }
                            const float2 *dv = reinterpret_cast<const float2 *>(row) + nsf_doff(K) / 2;
                            rq_spline_lazy<K>(raw, B, cs.y, cs.z, (p.dir_flags & 1) != 0, tr, ld, [&](int j) {
                                const float2 q = dv[j];
                                return fmaf(q.x, d, q.y);
                            });
                        } else {
                            const mnf_flow_op &op = p.prog.ops[ei.y];
                            const float2 o = nsf_fallback<K>(p.params + op.net_off[use_f1 ? 0 : 1], op.n_lin, op.sizes, B, cs.z, c,
                                                             (p.dir_flags & 1) != 0, tr);
                            tr = o.x, ld += o.y;
                        }
                        if (use_f1) v1 = tr; else v0 = tr;
                    }
                }
            }
            if (p.inter) st_stream2(reinterpret_cast<float2 *>(p.inter + ((size_t)kk * p.n_rows + r) * 2), make_float2(v0, v1));
        }
        float lp = fmaf(-0.5f, fmaf(v0, v0, v1 * v1), -1.8378770664093453f);  // -(D/2) log(2 pi), D = 2
        if (sum_lp) lp += ld;
        if (p.y) st_stream2(reinterpret_cast<float2 *>(p.y) + r, make_float2(v0, v1));
        if (p.log_det) p.log_det[r] = ld;
        if (p.base_lp) p.base_lp[r] = lp;
        // fused gather: the result also goes straight to the other ranks over NVLink
        if (p.gather.multicast_ptr) {
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.gather.multicast_ptr + p.gather.row_offset + r), "f"(lp)
                         : "memory");
        } else {
            for (int g = 0; g < p.gather.n_peers; ++g) p.gather.peer_ptrs[g][p.gather.row_offset + r] = lp;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(THREADS, 1) flow_pl_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(16) Exec s_exec[MNF_MAX_OPS];
    __shared__ int s_off[MAX_GROUPS], s_over[MAX_GROUPS];
    __shared__ int s_allfit, s_nexec;
    __shared__ __align__(8) unsigned long long s_bar;
    const int lane = threadIdx.x & 31, ng = p.gl.n;
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) build_exec(p, s_exec, &s_nexec, lane);
    if (threadIdx.x < 32) {
        // table sizes from the builder's header -> shared-memory offsets, one bulk copy per table
        const int *hdr = reinterpret_cast<const int *>(p.image);
        int total = 0, fit = 1;
        for (int base = 0; base < ng; base += 32) {
            const int g = base + lane;
            int sz = 0, ov = 0;
            if (g < ng) {
                const int2 h = *reinterpret_cast<const int2 *>(hdr + g * HDR_INTS);
                const int s4 = p.prog.ops[p.gl.op[g]].type == MNF_OP_NSF_CL ? nsf_stride4(K) : AFF_STRIDE4;
                ov = h.y;
                sz = ov ? 0 : BP_FLOATS + (h.x + 1) * 4 * s4;
            }
            int inc = sz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int off = total + inc - sz;
            const bool ok = off + sz <= p.smem_floats;
            if (g < ng) {
                s_off[g] = (ov || !ok) ? -1 : off;
                s_over[g] = ov;
                if (!ov && ok) {
                    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"((uint32_t)sz * 4u) : "memory");
                    bulk_load(smem_u32(smem + off), p.image + p.gl.region[g], (uint32_t)sz * 4u, bar);
                }
            }
            fit &= __all_sync(0xffffffffu, g >= ng || ov || ok) ? 1 : 0;
            total += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        if (lane == 0) {
            s_allfit = fit;
            mbar_arrive(bar);
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);
    if (s_allfit) run_points<K, true>(p, smem, s_off, s_over, s_exec, s_nexec);
    else run_points<K, false>(p, smem, s_off, s_over, s_exec, s_nexec);
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
// bins K the point kernel is instantiated for (the spline lives in registers: K is a template parameter)
#define MNF_PL_BINS(X) X(4) X(5) X(6) X(8) X(10) X(12) X(16)
constexpr int K_MAX = 16;
static bool k_instantiated(int K) {
#define MNF_PL_HAVE(KK) if (K == KK) return true;
    MNF_PL_BINS(MNF_PL_HAVE)
#undef MNF_PL_HAVE
    return false;
}

struct Plan {
    bool ok = false;
    int K = 8;
    GroupList gl;
    int64_t image_floats = 0;
};

static Plan make_plan(const mnf_flow_op *ops, int n_ops, int dim) {
    Plan pl;
    if (dim != 2 || n_ops < 1 || n_ops > MNF_MAX_OPS) return pl;
    int K = 0, n = 0;
    int64_t off = HDR_FLOATS;
    for (int k = 0; k < n_ops; ++k) {
        const mnf_flow_op &op = ops[k];
        pl.gl.group_of[k][0] = pl.gl.group_of[k][1] = -1;
        if (op.type == MNF_OP_AFFINE_CONST || op.type == MNF_OP_GLOW) continue;
        if (op.type != MNF_OP_AFFINE_HALF && op.type != MNF_OP_NSF_CL) return pl;
        if (op.n_lin < 2 || op.n_lin > MNF_MAX_LIN || op.sizes[0] != 1) return pl;
        for (int l = 1; l < op.n_lin; ++l)
            if (op.sizes[l] < 1 || op.sizes[l] > MAX_H) return pl;
        if (op.type == MNF_OP_NSF_CL) {
            if (!k_instantiated(op.K)) return pl;
            if (K && op.K != K) return pl;
            K = op.K;
            if (op.sizes[op.n_lin] != 3 * K - 1) return pl;
            for (int which = 0; which < 2; ++which) {
                pl.gl.group_of[k][which] = (signed char)n;
                pl.gl.op[n] = (unsigned char)k, pl.gl.which[n] = (unsigned char)which, pl.gl.region[n] = (int)off;
                off += region_floats(nsf_stride4(K));
                ++n;
            }
        } else {
            if (op.sizes[op.n_lin] != 1) return pl;
            if (!(op.flags & (MNF_FLAG_SCALE | MNF_FLAG_SHIFT))) continue;  // no net at all: s = t = 0 ... handled as identity below
            pl.gl.group_of[k][0] = (signed char)n;
            pl.gl.op[n] = (unsigned char)k, pl.gl.which[n] = 0, pl.gl.region[n] = (int)off;
            off += region_floats(AFF_STRIDE4);
            ++n;
        }
    }
    if (n == 0) return pl;
    pl.gl.n = n;
    pl.K = K ? K : 8;
    pl.image_floats = off;
    pl.ok = true;
    return pl;
}

template <int K>
static int launch_k(const Params &p, const DeviceProps *dp, size_t smem_bytes, cudaStream_t st) {
    static thread_local size_t attr_set[64] = {};  // per device: the attribute belongs to the device's copy of the function
    int dev = 0;
    MNF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || attr_set[dev] < smem_bytes) {
        MNF_CUDA(cudaFuncSetAttribute(flow_pl_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = smem_bytes;
    }
    // small batches are bound by the latency of one point's chain: spread the points over the SMs
    const int threads = p.n_rows <= (long long)dp->sm_count * 64 ? 64 : p.n_rows <= (long long)dp->sm_count * 256 ? 256 : THREADS;
    long long blocks = (p.n_rows + threads - 1) / threads;
    if (blocks > dp->sm_count) blocks = dp->sm_count;
    flow_pl_kernel<K><<<(unsigned)blocks, threads, smem_bytes, st>>>(p);
    return launch_status("flow_pl_kernel");
}

}  // namespace fpl

// floats of workspace / staged image the piecewise-linear kernel needs for a program of n_ops flows (upper bound)
int64_t flow_pl_workspace_floats(int n_ops) {
    if (n_ops < 1) n_ops = 1;
    if (n_ops > MNF_MAX_OPS) n_ops = MNF_MAX_OPS;
    return fpl::HDR_FLOATS + (int64_t)2 * n_ops * fpl::region_floats(fpl::nsf_stride4(fpl::K_MAX));
}

// floats of the table image of this program, 0 if it has no piecewise-linear form
int64_t flow_pl_image_floats(const mnf_flow_op *ops, int n_ops, int dim) {
    const fpl::Plan pl = fpl::make_plan(ops, n_ops, dim);
    return pl.ok ? pl.image_floats : 0;
}

int flow_pl_build(const mnf_flow_op *ops, int n_ops, const float *params, int dim, float *image, cudaStream_t stream) {
    using namespace fpl;
    const Plan pl = make_plan(ops, n_ops, dim);
    MNF_REQUIRE(pl.ok, MNF_E_SHAPE, "program has no piecewise-linear form");
    MNF_REQUIRE(params && image && ((uintptr_t)image % 16) == 0, MNF_E_ARG, "NULL or misaligned pointer");
    FlowProgram prog;
    prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) prog.ops[k] = ops[k];
    flow_pl_build_kernel<<<pl.gl.n, 32 * BW, 0, stream>>>(prog, pl.gl, params, image);
    return launch_status("flow_pl_build_kernel");
}

// returns 1 if the program is not eligible (caller falls back to the other dim-2 kernels).  dir_flags: bit0 inverse,
// bit1 log-det summed into base_lp, bit2 `workspace` already holds the image of flow_pl_build for these parameters.
int launch_flow_pl(const mnf_flow_op *ops, int n_ops, const float *params, const float *x, float *y, float *log_det,
                   float *base_lp, float *inter, int64_t n_rows, int dim, int dir_flags, float *workspace,
                   const mnf_gather_out *gather, cudaStream_t stream, bool plan_only) {
    using namespace fpl;
    const Plan pl = make_plan(ops, n_ops, dim);
    if (!pl.ok) return 1;
    if (plan_only) return 0;
    if (!workspace) return 1;  // no room for the tables: the shared-memory kernels need none
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr, MNF_E_DEVICE, "no CUDA device");
    MNF_REQUIRE(((uintptr_t)workspace % 16) == 0, MNF_E_ALIGN, "workspace must be 16-byte aligned");
    MNF_REQUIRE(((uintptr_t)x % 8) == 0 && (!y || ((uintptr_t)y % 8) == 0) && (!inter || ((uintptr_t)inter % 8) == 0), MNF_E_ALIGN,
                "x, y and intermediates must be 8-byte aligned");
    MNF_REQUIRE(n_rows > 0, MNF_E_ARG, "bad row count");
    if (!(dir_flags & 4)) {
        const int rc = flow_pl_build(ops, n_ops, params, dim, workspace, stream);
        if (rc) return rc;
    }
    Params p{};
    p.prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) p.prog.ops[k] = ops[k];
    p.gl = pl.gl;
    p.params = params, p.x = x, p.image = workspace, p.y = y, p.log_det = log_det, p.base_lp = base_lp, p.inter = inter;
    p.n_rows = n_rows, p.dir_flags = dir_flags & 3;
    if (gather) {
        MNF_REQUIRE(gather->n_peers >= 0 && gather->n_peers <= MNF_MAX_PEERS, MNF_E_ARG, "bad n_peers");
        p.gather = *gather;
    }
    // shared memory: every table at its largest, capped by what a CTA can have
    size_t want = (size_t)(pl.image_floats - HDR_FLOATS) * sizeof(float);
    const size_t cap = (size_t)dp->smem_optin - 4096;  // static shared memory of the kernel: execution list, offsets, barrier
    if (want > cap) want = cap;
    want &= ~(size_t)15;
    p.smem_floats = (int)(want / sizeof(float));
#define MNF_PL_LAUNCH(KK) if (pl.K == KK) return launch_k<KK>(p, dp, want, stream);
    MNF_PL_BINS(MNF_PL_LAUNCH)
#undef MNF_PL_LAUNCH
    return 1;
}

}  // namespace mnf
