// flow_fast.cu -- register-resident flow-stack kernel for 2-D points (BASELINE configs 1, 2).
//
// Whole stack in one launch: a thread owns TWO points as an fp32x2 pair, keeps both points and
// their log-dets in registers across every flow, and evaluates the coupling conditioners
// (MLP 1 -> H -> H -> H -> n_out, LeakyReLU 0.2) fully unrolled with the weights broadcast
// from shared memory (staged once per CTA; the grid is persistent, a multiple of the SM
// count).  The two points share every weight load and are multiplied with Blackwell's packed
// FFMA2 (fma.rn.f32x2).  HBM traffic is the algorithmic minimum: 8 B in, 8 B out, 4 B log-det
// per point (+4 B when the fused base log-density is requested).
//
// Supported ops (all with dim == 2): AffineConstantFlow/ActNormFlow, Glow, AffineHalfFlow with
// h_sizes (H,H,H), NSF_CL with n_h = H and K bins.  Anything else -> flow_generic.cu.
//
// MODE selects how the packed multiply gets its (w,w) operand:
//   0: scalar FFMA (two per weight), plain smem      1: FFMA2, pair built with a MOV
//   2: FFMA2, weights stored duplicated in smem (one LDS.128 = two ready pairs)
#pragma once
#include "flow_math.cuh"

namespace mnf {

struct FlowProgram {
    int n_ops;
    mnf_flow_op ops[MNF_MAX_OPS];
};

// shared-memory offsets (in weight slots) of each op's nets, computed on the host
struct FastLayout {
    int net_slot[MNF_MAX_OPS][2];
    int total_slots;
};

__device__ __forceinline__ float2 fma2_packed(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

template <int MODE>
__device__ __forceinline__ float2 fma2(float2 w, float2 x, float2 c) {
    if constexpr (MODE == 0)
        return make_float2(fmaf(w.x, x.x, c.x), fmaf(w.y, x.y, c.y));
    else
        return fma2_packed(w, x, c);
}

__device__ __forceinline__ float2 leaky2(float2 v) { return make_float2(leaky02(v.x), leaky02(v.y)); }

// Weight slot accessors.  Slot i of a net holds weight i; in MODE 2 a slot is a float2 (w,w).
template <int MODE>
struct Wts {
    const float *base;
    // four consecutive slots starting at compile-time-constant i (multiple of 4 relative to an
    // aligned net start), as four (w,w) pairs
    __device__ __forceinline__ void load4(int i, float2 (&w)[4]) const {
        if constexpr (MODE == 2) {
            const float4 *p = reinterpret_cast<const float4 *>(base) + (i >> 1);
            const float4 a = p[0], b = p[1];
            w[0] = make_float2(a.x, a.y);
            w[1] = make_float2(a.z, a.w);
            w[2] = make_float2(b.x, b.y);
            w[3] = make_float2(b.z, b.w);
        } else {
            const float4 a = reinterpret_cast<const float4 *>(base)[i >> 2];
            w[0] = make_float2(a.x, a.x);
            w[1] = make_float2(a.y, a.y);
            w[2] = make_float2(a.z, a.z);
            w[3] = make_float2(a.w, a.w);
        }
    }
    __device__ __forceinline__ float2 load1(int i) const {
        if constexpr (MODE == 2) return reinterpret_cast<const float2 *>(base)[i];
        const float w = base[i];
        return make_float2(w, w);
    }
};

// Hidden part of the conditioner for a scalar input: 1 -> H -> H -> H.  Result in h.
template <int H, int MODE>
__device__ __forceinline__ void mlp_hidden3(const Wts<MODE> W, float2 x, float2 (&h)[H]) {
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4");
    float2 g[H];
    // layer 0: weight[H][1], bias[H]
#pragma unroll
    for (int j = 0; j < H; j += 4) {
        float2 w[4], b[4];
        W.load4(j, w);
        W.load4(H + j, b);
#pragma unroll
        for (int u = 0; u < 4; ++u) g[j + u] = leaky2(fma2<MODE>(w[u], x, b[u]));
    }
    // layers 1, 2: weight[H][H], bias[H]
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
        const int wo = 2 * H + layer * (H * H + H);
        float2(&src)[H] = layer == 0 ? g : h;
        float2(&dst)[H] = layer == 0 ? h : g;
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            float2 acc[4];
            W.load4(wo + H * H + j, acc);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < H; i += 4) {
                    float2 w[4];
                    W.load4(wo + (j + u) * H + i, w);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[u] = fma2<MODE>(w[q], src[i + q], acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[j + u] = leaky2(acc[u]);
        }
    }
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = g[j];
}

// Last Linear layer: n_out outputs from the H hidden units.
template <int H, int NO, int MODE>
__device__ __forceinline__ void mlp_last(const Wts<MODE> W, const float2 (&h)[H], float2 (&out)[NO]) {
    const int wo = 2 * H + 2 * (H * H + H);
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        float2 acc = W.load1(wo + NO * H + j);
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            float2 w[4];
            W.load4(wo + j * H + i, w);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc = fma2<MODE>(w[q], h[i + q], acc);
        }
        out[j] = acc;
    }
}

template <int H>
constexpr int net_slots(int n_out) {
    return 2 * H + 2 * (H * H + H) + n_out * H + ((n_out + 3) / 4) * 4;
}

// conditioner on `cond`, spline on `trans` (both points of the pair)
template <int H, int K, int MODE>
__device__ __forceinline__ void spline_half(const Wts<MODE> W, const mnf_flow_op &op, float2 cond, float2 &trans,
                                            bool rqs_inverse, float2 &ld) {
    constexpr int NB = 3 * K - 1;
    float2 raw2[NB];
    {
        float2 h[H];
        mlp_hidden3<H, MODE>(W, cond, h);
        mlp_last<H, NB, MODE>(W, h, raw2);
    }
#pragma unroll 1
    for (int pt = 0; pt < 2; ++pt) {
        float raw[NB];
#pragma unroll
        for (int o = 0; o < NB; ++o) raw[o] = pt ? raw2[o].y : raw2[o].x;
        float v = pt ? trans.y : trans.x;
        float l = 0.f;
        rq_spline<K>(raw, K, op.bound, op.edge_deriv, rqs_inverse, v, l);
        if (pt) { trans.y = v; ld.y += l; } else { trans.x = v; ld.x += l; }
    }
}

template <int H, int K, int MODE>
__global__ void __launch_bounds__(128, 4)
flow_fast_kernel(const __grid_constant__ FlowProgram prog, const __grid_constant__ FastLayout lay,
                 const float *__restrict__ params, const float *__restrict__ x, float *__restrict__ y,
                 float *__restrict__ log_det, float *__restrict__ base_lp, float *__restrict__ inter,
                 long long n_rows, int inverse) {
    extern __shared__ __align__(16) float smem[];
    // ---- stage every net of the program into shared memory (once per CTA) ----
    for (int k = 0; k < prog.n_ops; ++k) {
        const mnf_flow_op &op = prog.ops[k];
        if (op.type != MNF_OP_AFFINE_HALF && op.type != MNF_OP_NSF_CL) continue;
        const int n_out = op.sizes[op.n_lin];
        const int n = 2 * H + 2 * (H * H + H) + n_out * H + n_out;
        for (int which = 0; which < 2; ++which) {
            if (op.type == MNF_OP_AFFINE_HALF && !(op.flags & (which ? MNF_FLAG_SHIFT : MNF_FLAG_SCALE))) continue;
            const float *src = params + op.net_off[which];
            const int slot0 = lay.net_slot[k][which];
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const float w = src[i];
                if constexpr (MODE == 2)
                    reinterpret_cast<float2 *>(smem)[slot0 + i] = make_float2(w, w);
                else
                    smem[slot0 + i] = w;
            }
        }
    }
    __syncthreads();

    const long long n_pairs = (n_rows + 1) >> 1;
    for (long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x; pair < n_pairs;
         pair += (long long)gridDim.x * blockDim.x) {
        const bool has_b = 2 * pair + 1 < n_rows;
        float2 v0, v1;  // v0 = (x_A, x_B) first coordinate of both points, v1 = second coordinate
        if (has_b) {
            const float4 q = ld_stream4(reinterpret_cast<const float4 *>(x) + pair);
            v0 = make_float2(q.x, q.z);
            v1 = make_float2(q.y, q.w);
        } else {
            const float2 q = ld_stream2(reinterpret_cast<const float2 *>(x) + 2 * pair);
            v0 = make_float2(q.x, q.x);
            v1 = make_float2(q.y, q.y);
        }
        float2 ld = make_float2(0.f, 0.f);

        for (int kk = 0; kk < prog.n_ops; ++kk) {
            const int k = inverse ? prog.n_ops - 1 - kk : kk;
            const mnf_flow_op &op = prog.ops[k];
            if (op.type == MNF_OP_AFFINE_CONST) {
                const float s0 = params[op.aux_off], s1 = params[op.aux_off + 1];
                const float t0 = params[op.aux_off + 2], t1 = params[op.aux_off + 3];
                if (inverse) {  // affine_constant_flow.py:24
                    const float e0 = expf(-s0), e1 = expf(-s1);
                    v0 = make_float2((v0.x - t0) * e0, (v0.y - t0) * e0);
                    v1 = make_float2((v1.x - t1) * e1, (v1.y - t1) * e1);
                    ld.x -= s0 + s1;
                    ld.y -= s0 + s1;
                } else {  // affine_constant_flow.py:19
                    const float e0 = expf(s0), e1 = expf(s1);
                    v0 = make_float2(v0.x * e0 + t0, v0.y * e0 + t0);
                    v1 = make_float2(v1.x * e1 + t1, v1.y * e1 + t1);
                    ld.x += s0 + s1;
                    ld.y += s0 + s1;
                }
            } else if (op.type == MNF_OP_GLOW) {
                const float *W = params + op.aux_off + (inverse ? 4 : 0);  // glow.py:28,36: v @ W
                const float w00 = W[0], w01 = W[1], w10 = W[2], w11 = W[3];
                const float lg = params[op.aux_off + 8];
                const float2 n0 = make_float2(fmaf(v1.x, w10, v0.x * w00), fmaf(v1.y, w10, v0.y * w00));
                const float2 n1 = make_float2(fmaf(v1.x, w11, v0.x * w01), fmaf(v1.y, w11, v0.y * w01));
                v0 = n0;
                v1 = n1;
                ld.x += inverse ? -lg : lg;
                ld.y += inverse ? -lg : lg;
            } else if (op.type == MNF_OP_AFFINE_HALF) {
                const bool parity = op.flags & MNF_FLAG_PARITY;
                const float2 cond = parity ? v1 : v0;  // affine_half_flow.py:46-50
                float2 tr = parity ? v0 : v1;
                float2 st[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 1
                for (int which = 0; which < 2; ++which) {
                    if (!(op.flags & (which ? MNF_FLAG_SHIFT : MNF_FLAG_SCALE))) continue;
                    Wts<MODE> W{smem + (MODE == 2 ? 2 : 1) * lay.net_slot[k][which]};
                    float2 h[H], o[1];
                    mlp_hidden3<H, MODE>(W, cond, h);
                    mlp_last<H, 1, MODE>(W, h, o);
                    if (which) st[1] = o[0]; else st[0] = o[0];
                }
                const float2 s = st[0], t = st[1];
                if (inverse) {  // affine_half_flow.py:54-56
                    tr = make_float2((tr.x - t.x) / expf(s.x), (tr.y - t.y) / expf(s.y));
                    ld.x -= s.x;
                    ld.y -= s.y;
                } else {  // affine_half_flow.py:58
                    tr = make_float2(expf(s.x) * tr.x + t.x, expf(s.y) * tr.y + t.y);
                    ld.x += s.x;
                    ld.y += s.y;
                }
                if (parity) v0 = tr; else v1 = tr;
            } else if (op.type == MNF_OP_NSF_CL) {
                // forward: f1 on (lower -> upper) then f2 on (upper -> lower) (spline_flow.py:249-266);
                // inverse: f2 first, then f1, both with the spline inverse (spline_flow.py:268-285).
                // One call site, two trips: keeps the unrolled body in the instruction cache.
#pragma unroll 1
                for (int step = 0; step < 2; ++step) {
                    const bool use_f1 = (step == 0) != (inverse != 0);
                    Wts<MODE> W{smem + (MODE == 2 ? 2 : 1) * lay.net_slot[k][use_f1 ? 0 : 1]};
                    const float2 cond = use_f1 ? v0 : v1;
                    float2 tr = use_f1 ? v1 : v0;
                    spline_half<H, K, MODE>(W, op, cond, tr, inverse != 0, ld);
                    if (use_f1) v1 = tr; else v0 = tr;
                }
            }
            if (inter) {
                float *dst = inter + ((size_t)kk * n_rows + 2 * pair) * 2;
                if (has_b)
                    st_stream4(reinterpret_cast<float4 *>(dst), make_float4(v0.x, v1.x, v0.y, v1.y));
                else
                    st_stream2(reinterpret_cast<float2 *>(dst), make_float2(v0.x, v1.x));
            }
        }

        const float c = -1.8378770664093453f;  // -(D/2) log(2 pi), D = 2
        const float2 lp = make_float2(fmaf(-0.5f, fmaf(v0.x, v0.x, v1.x * v1.x), c),
                                      fmaf(-0.5f, fmaf(v0.y, v0.y, v1.y * v1.y), c));
        if (has_b) {
            st_stream4(reinterpret_cast<float4 *>(y) + pair, make_float4(v0.x, v1.x, v0.y, v1.y));
            if (log_det) st_stream2(reinterpret_cast<float2 *>(log_det) + pair, ld);
            if (base_lp) st_stream2(reinterpret_cast<float2 *>(base_lp) + pair, lp);
        } else {
            st_stream2(reinterpret_cast<float2 *>(y) + 2 * pair, make_float2(v0.x, v1.x));
            if (log_det) log_det[2 * pair] = ld.x;
            if (base_lp) base_lp[2 * pair] = lp.x;
        }
    }
}

template <int H, int K, int MODE>
int launch_inst(const FlowProgram &prog, const FastLayout &lay, size_t smem_bytes, const float *params,
                       const float *x, float *y, float *log_det, float *base_lp, float *inter, int64_t n_rows,
                       int inverse, const DeviceProps *dp, cudaStream_t stream) {
    auto kern = flow_fast_kernel<H, K, MODE>;
    if (smem_bytes > 48 * 1024)
        MNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    int occ = 0;
    MNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem_bytes));
    if (occ < 1) return fail(MNF_E_SHAPE, "flow_fast_kernel does not fit on an SM (smem %zu B)", smem_bytes);
    const long long n_pairs = (n_rows + 1) / 2;
    long long blocks = (n_pairs + 127) / 128;
    const long long cap = (long long)dp->sm_count * occ;  // persistent: one wave, a multiple of the SM count
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 128, smem_bytes, stream>>>(prog, lay, params, x, y, log_det, base_lp, inter, n_rows,
                                                        inverse);
    return launch_status("flow_fast_kernel");
}


#define MNF_FLOW_FAST_ARGS                                                                                     \
    int mode, const FlowProgram &prog, const FastLayout &lay, size_t smem_bytes, const float *params,          \
        const float *x, float *y, float *log_det, float *base_lp, float *inter, int64_t n_rows, int inverse,   \
        const DeviceProps *dp, cudaStream_t stream

#define MNF_FLOW_FAST_DEFINE(HH, KK)                                                                            \
    int launch_fast_##HH##_##KK(MNF_FLOW_FAST_ARGS) {                                                           \
        if (mode == 0)                                                                                          \
            return launch_inst<HH, KK, 0>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, \
                                          inverse, dp, stream);                                                 \
        if (mode == 1)                                                                                          \
            return launch_inst<HH, KK, 1>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, \
                                          inverse, dp, stream);                                                 \
        return launch_inst<HH, KK, 2>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows,     \
                                      inverse, dp, stream);                                                     \
    }

}  // namespace mnf
