// flow_fast.cuh -- register-resident flow-stack kernel for 2-D points (BASELINE configs 1, 2).
//
// Whole stack in one launch: a thread owns TWO points, keeps both points and their log-dets in
// registers across every flow, and evaluates the coupling conditioners (MLP 1 -> H -> H -> H ->
// n_out, LeakyReLU 0.2) fully unrolled with the weights broadcast from shared memory (staged
// once per CTA; the grid is persistent, a multiple of the SM count).  The two points share every
// weight load; the multiplies use Blackwell's packed FFMA2 (fma.rn.f32x2).  HBM traffic is the
// algorithmic minimum: 8 B in, 8 B out, 4 B log-det per point (+4 B for the fused base
// log-density).
//
// Supported ops (all with dim == 2): AffineConstantFlow/ActNormFlow, Glow, AffineHalfFlow with
// h_sizes (H,H,H), NSF_CL with n_h = H and K bins.  Anything else -> flow_generic.cu.
//
// VARIANT selects how the MLP is mapped onto the packed FMA:
//   0: point-packed, scalar FFMA (two per weight)           -- reference variant
//   1: point-packed FFMA2, (w,w) operand built with a MOV per weight
//   2: output-packed FFMA2: weights stored transposed ([in][out]) so one LDS.128 delivers two
//      ready (w_j, w_j+1) operands; the activation is duplicated once per input instead
//   3: as 2, but the weights sit in the CONSTANT bank: they reach the FMA pipe through uniform
//      registers (LDCU.128 -> FFMA2 R, R.F32, UR.F32x2, R), no shared-memory -> register-file
//      traffic at all.  Needs the staged nets to fit in 64 KB (config 2: 23 KB).
#pragma once
#include <mutex>

#include "flow_math.cuh"

namespace mnf {

#ifndef MNF_SPLINE_FAST
#define MNF_SPLINE_FAST 1
#endif

// shared-memory offsets (in floats) of each op's nets, computed on the host
struct FastLayout {
    int net_slot[MNF_MAX_OPS][2];
    int total_slots;
};

constexpr int round4(int n) { return (n + 3) / 4 * 4; }

// floats one staged net occupies in shared memory
constexpr int fast_net_slots(int H, int n_out, int variant) {
    const int hidden = 2 * H + 2 * (H * H + H);
    if (variant >= 2) return hidden + (n_out == 1 ? H + 4 : H * round4(n_out) + round4(n_out));
    return hidden + n_out * H + round4(n_out);
}

__device__ __forceinline__ float2 fma2_packed(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 mul2_packed(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

template <int VARIANT>
__device__ __forceinline__ float2 fma2(float2 w, float2 x, float2 c) {
    if constexpr (VARIANT == 0)
        return make_float2(fmaf(w.x, x.x, c.x), fmaf(w.y, x.y, c.y));
    else
        return fma2_packed(w, x, c);
}

__device__ __forceinline__ float2 leaky2(float2 v) {
    const float2 s = mul2_packed(v, make_float2(0.2f, 0.2f));
    return make_float2(fmaxf(v.x, s.x), fmaxf(v.y, s.y));
}

__device__ __forceinline__ float4 lds4(const float *base, int i) {  // i: compile-time multiple of 4
    return reinterpret_cast<const float4 *>(base)[i >> 2];
}

// constant-bank staging area of this translation unit's kernels (variant 3) and the global
// scratch the nets are re-laid-out into before the device-to-device copy into the bank
constexpr int kConstFloats = 16384 - 64;
__constant__ float c_flow_w[kConstFloats];

// weight accessors: offsets are in floats from the start of the staged net
struct WShared {
    const float *p;
    __device__ __forceinline__ float4 ld4(int i) const { return reinterpret_cast<const float4 *>(p)[i >> 2]; }
    __device__ __forceinline__ float ld1(int i) const { return p[i]; }
};
// constant-bank accessor with a COMPILE-TIME base: every load becomes LDCU c[0x3][imm] into uniform
// registers (ptxas keeps a run-time base in vector registers and falls back to per-thread LDC)
template <int BASE>
struct WConstAt {
    __device__ __forceinline__ float4 ld4(int i) const {
        return reinterpret_cast<const float4 *>(c_flow_w)[(BASE + i) >> 2];
    }
    __device__ __forceinline__ float ld1(int i) const { return c_flow_w[BASE + i]; }
};

// =====================================================================================
// point-packed engine (variants 0, 1): float2 = (value for point A, value for point B)
// smem net layout = blob layout: per Linear weight[out][in], bias[out]
// =====================================================================================
template <int H, int VARIANT>
__device__ __forceinline__ void pp_hidden(const float *W, float2 x, float2 (&h)[H]) {
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4");
    float2 g[H];
#pragma unroll
    for (int j = 0; j < H; j += 4) {  // layer 0: weight[H][1], bias[H]
        const float4 w = lds4(W, j), b = lds4(W, H + j);
        g[j + 0] = leaky2(fma2<VARIANT>(make_float2(w.x, w.x), x, make_float2(b.x, b.x)));
        g[j + 1] = leaky2(fma2<VARIANT>(make_float2(w.y, w.y), x, make_float2(b.y, b.y)));
        g[j + 2] = leaky2(fma2<VARIANT>(make_float2(w.z, w.z), x, make_float2(b.z, b.z)));
        g[j + 3] = leaky2(fma2<VARIANT>(make_float2(w.w, w.w), x, make_float2(b.w, b.w)));
    }
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {  // layers 1, 2: weight[H][H], bias[H]
        const int wo = 2 * H + layer * (H * H + H);
        float2(&src)[H] = layer == 0 ? g : h;
        float2(&dst)[H] = layer == 0 ? h : g;
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 b = lds4(W, wo + H * H + j);
            float2 acc[4] = {make_float2(b.x, b.x), make_float2(b.y, b.y), make_float2(b.z, b.z),
                             make_float2(b.w, b.w)};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < H; i += 4) {
                    const float4 w = lds4(W, wo + (j + u) * H + i);
                    acc[u] = fma2<VARIANT>(make_float2(w.x, w.x), src[i + 0], acc[u]);
                    acc[u] = fma2<VARIANT>(make_float2(w.y, w.y), src[i + 1], acc[u]);
                    acc[u] = fma2<VARIANT>(make_float2(w.z, w.z), src[i + 2], acc[u]);
                    acc[u] = fma2<VARIANT>(make_float2(w.w, w.w), src[i + 3], acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[j + u] = leaky2(acc[u]);
        }
    }
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = g[j];
}

template <int H, int NO, int VARIANT>
__device__ __forceinline__ void pp_last(const float *W, const float2 (&h)[H], float2 (&out)[NO]) {
    const int wo = 2 * H + 2 * (H * H + H);
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        const float b = W[wo + NO * H + j];
        float2 acc = make_float2(b, b);
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 w = lds4(W, wo + j * H + i);
            acc = fma2<VARIANT>(make_float2(w.x, w.x), h[i + 0], acc);
            acc = fma2<VARIANT>(make_float2(w.y, w.y), h[i + 1], acc);
            acc = fma2<VARIANT>(make_float2(w.z, w.z), h[i + 2], acc);
            acc = fma2<VARIANT>(make_float2(w.w, w.w), h[i + 3], acc);
        }
        out[j] = acc;
    }
}

// =====================================================================================
// output-packed engine (variant 2): float2 = (unit j, unit j+1) of ONE point; two points A, B
// smem net layout: w0[H] b0[H] | Wt1[in][out] b1[H] | Wt2[in][out] b2[H] | Wt3[in][NOP] b3[NOP]
// (n_out == 1: w3[H] b3[4])
// =====================================================================================
template <int H, int NOUT2, class WT>
__device__ __forceinline__ void op_dense(const WT W, int wt_off, int bias_off, const float2 (&inA)[H / 2],
                                         const float2 (&inB)[H / 2], float2 (&outA)[NOUT2],
                                         float2 (&outB)[NOUT2]) {
    // out[j] = bias[j] + sum_i Wt[i][j] * in[i]; NOUT2 pairs of outputs, row stride 2*NOUT2
#pragma unroll
    for (int j = 0; j < NOUT2; j += 2) {
        const float4 b = W.ld4(bias_off + 2 * j);
        outA[j] = outB[j] = make_float2(b.x, b.y);
        if (j + 1 < NOUT2) outA[j + 1] = outB[j + 1] = make_float2(b.z, b.w);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float a = (i & 1) ? inA[i >> 1].y : inA[i >> 1].x;
        const float b = (i & 1) ? inB[i >> 1].y : inB[i >> 1].x;
        const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
        for (int j = 0; j < NOUT2; j += 2) {
            const float4 w = W.ld4(wt_off + i * 2 * NOUT2 + 2 * j);
            outA[j] = fma2_packed(make_float2(w.x, w.y), aa, outA[j]);
            outB[j] = fma2_packed(make_float2(w.x, w.y), bb, outB[j]);
            if (j + 1 < NOUT2) {
                outA[j + 1] = fma2_packed(make_float2(w.z, w.w), aa, outA[j + 1]);
                outB[j + 1] = fma2_packed(make_float2(w.z, w.w), bb, outB[j + 1]);
            }
        }
    }
}

template <int H, class WT>
__device__ __forceinline__ void op_hidden(const WT W, float xA, float xB, float2 (&hA)[H / 2],
                                          float2 (&hB)[H / 2]) {
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4");
    float2 gA[H / 2], gB[H / 2];
    const float2 xa = make_float2(xA, xA), xb = make_float2(xB, xB);
#pragma unroll
    for (int j = 0; j < H / 2; j += 2) {  // layer 0
        const float4 w = W.ld4(2 * j), b = W.ld4(H + 2 * j);
        gA[j] = leaky2(fma2_packed(make_float2(w.x, w.y), xa, make_float2(b.x, b.y)));
        gB[j] = leaky2(fma2_packed(make_float2(w.x, w.y), xb, make_float2(b.x, b.y)));
        gA[j + 1] = leaky2(fma2_packed(make_float2(w.z, w.w), xa, make_float2(b.z, b.w)));
        gB[j + 1] = leaky2(fma2_packed(make_float2(w.z, w.w), xb, make_float2(b.z, b.w)));
    }
    const int l1 = 2 * H, l2 = 2 * H + H * H + H;
    op_dense<H, H / 2>(W, l1, l1 + H * H, gA, gB, hA, hB);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        hA[j] = leaky2(hA[j]);
        hB[j] = leaky2(hB[j]);
    }
    op_dense<H, H / 2>(W, l2, l2 + H * H, hA, hB, gA, gB);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        hA[j] = leaky2(gA[j]);
        hB[j] = leaky2(gB[j]);
    }
}

// single output (AffineHalfFlow s / t): dot product packed over input pairs
template <int H, class WT>
__device__ __forceinline__ void op_last1(const WT W, const float2 (&hA)[H / 2], const float2 (&hB)[H / 2],
                                         float &oA, float &oB) {
    const int wo = 2 * H + 2 * (H * H + H);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < H / 2; i += 2) {
        const float4 w = W.ld4(wo + 2 * i);
        accA = fma2_packed(make_float2(w.x, w.y), hA[i], accA);
        accB = fma2_packed(make_float2(w.x, w.y), hB[i], accB);
        accA = fma2_packed(make_float2(w.z, w.w), hA[i + 1], accA);
        accB = fma2_packed(make_float2(w.z, w.w), hB[i + 1], accB);
    }
    const float b = W.ld1(wo + H);
    oA = (accA.x + accA.y) + b;
    oB = (accB.x + accB.y) + b;
}

// =====================================================================================
// conditioner -> spline for both points
// =====================================================================================
template <int VARIANT>
struct WSel {
    using type = WShared;
    static __device__ __forceinline__ WShared make(const float *smem, int slot) { return WShared{smem + slot}; }
};

template <int H, int K, int VARIANT, bool ONE = false>
__device__ __forceinline__ void spline_half(const float *smem, int slot, const mnf_flow_op &op, float2 cond,
                                            float2 &trans, bool rqs_inverse, float2 &ld) {
    constexpr int NB = 3 * K - 1;
    constexpr bool FAST = MNF_SPLINE_FAST != 0;
    const float *W = smem + slot;
    if constexpr (VARIANT >= 2) {
        constexpr int NP2 = round4(NB) / 2;
        float2 rA[NP2], rB[NP2];
        {
            const auto Wa = WSel<VARIANT>::make(smem, slot);
            float2 hA[H / 2], hB[H / 2];
            op_hidden<H>(Wa, cond.x, cond.y, hA, hB);
            const int wo = 2 * H + 2 * (H * H + H);
            op_dense<H, NP2>(Wa, wo, wo + H * 2 * NP2, hA, hB, rA, rB);
        }
#pragma unroll 1
        for (int pt = 0; pt < (ONE ? 1 : 2); ++pt) {
            float raw[NB];
#pragma unroll
            for (int o = 0; o < NB; ++o) {
                const float2 q = pt ? rB[o >> 1] : rA[o >> 1];
                raw[o] = (o & 1) ? q.y : q.x;
            }
            float v = pt ? trans.y : trans.x;
            float l = 0.f;
            rq_spline<K, FAST>(raw, K, op.bound, op.edge_deriv, rqs_inverse, v, l);
            if (pt) { trans.y = v; ld.y += l; } else { trans.x = v; ld.x += l; }
        }
    } else {
        float2 raw2[NB];
        {
            float2 h[H];
            pp_hidden<H, VARIANT>(W, cond, h);
            pp_last<H, NB, VARIANT>(W, h, raw2);
        }
#pragma unroll 1
        for (int pt = 0; pt < 2; ++pt) {
            float raw[NB];
#pragma unroll
            for (int o = 0; o < NB; ++o) raw[o] = pt ? raw2[o].y : raw2[o].x;
            float v = pt ? trans.y : trans.x;
            float l = 0.f;
            rq_spline<K, FAST>(raw, K, op.bound, op.edge_deriv, rqs_inverse, v, l);
            if (pt) { trans.y = v; ld.y += l; } else { trans.x = v; ld.x += l; }
        }
    }
}

// scalar-output conditioner (s or t of AffineHalfFlow) for both points
template <int H, int VARIANT>
__device__ __forceinline__ float2 affine_net(const float *smem, int slot, float2 cond) {
    const float *W = smem + slot;
    if constexpr (VARIANT >= 2) {
        const auto Wa = WSel<VARIANT>::make(smem, slot);
        float2 hA[H / 2], hB[H / 2];
        op_hidden<H>(Wa, cond.x, cond.y, hA, hB);
        float2 o;
        op_last1<H>(Wa, hA, hB, o.x, o.y);
        return o;
    } else {
        float2 h[H], o[1];
        pp_hidden<H, VARIANT>(W, cond, h);
        pp_last<H, 1, VARIANT>(W, h, o);
        return o[0];
    }
}

// stage one net from the parameter blob into shared memory in the variant's layout
template <int H, int VARIANT>
__device__ __forceinline__ void stage_net(const float *__restrict__ src, float *dst, int n_out) {
    const int hidden = 2 * H + 2 * (H * H + H);
    const int n = hidden + n_out * H + n_out;
    // eight independent loads in flight per thread: with one load per trip the copy runs at one L2 latency per
    // element and thread (measured: ~45 us to stage the 18 nets of BASELINE config 1, most of a small-batch call)
    for (int e0 = threadIdx.x; e0 < n; e0 += 8 * blockDim.x) {
        float w8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            w8[u] = e < n ? src[e] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e >= n) continue;
            int d = e;
            if constexpr (VARIANT >= 2) {
                if (e >= 2 * H && e < hidden) {  // the two H x H layers: transpose weight blocks
                    const int r = (e - 2 * H) % (H * H + H), base = e - r;
                    if (r < H * H) d = base + (r % H) * H + (r / H);
                } else if (e >= hidden && n_out > 1) {
                    const int nop = round4(n_out), r = e - hidden;
                    d = r < n_out * H ? hidden + (r % H) * nop + (r / H) : hidden + H * nop + (r - n_out * H);
                }
            }
            dst[d] = w8[u];
        }
    }
}

// all nets of a program into an image laid out by `lay` (shared memory in the kernel, global memory when pre-staging)
template <int H, int VARIANT>
__device__ __forceinline__ void stage_program(const FlowProgram &prog, const FastLayout &lay,
                                              const float *__restrict__ params, float *image) {
    if constexpr (VARIANT == 2) {  // padding lanes of the last layers must read zeros
        for (int e = threadIdx.x; e < lay.total_slots; e += blockDim.x) image[e] = 0.f;
        __syncthreads();
    }
    for (int k = 0; k < prog.n_ops; ++k) {
        const mnf_flow_op &op = prog.ops[k];
        if (op.type != MNF_OP_AFFINE_HALF && op.type != MNF_OP_NSF_CL) continue;
        for (int which = 0; which < 2; ++which) {
            if (op.type == MNF_OP_AFFINE_HALF && !(op.flags & (which ? MNF_FLAG_SHIFT : MNF_FLAG_SCALE))) continue;
            stage_net<H, VARIANT>(params + op.net_off[which], image + lay.net_slot[k][which], op.sizes[op.n_lin]);
        }
    }
}

template <int H>
__global__ void flow_stage_image_kernel(const __grid_constant__ FlowProgram prog, const __grid_constant__ FastLayout lay,
                                        const float *__restrict__ params, float *__restrict__ image) {
    stage_program<H, 2>(prog, lay, params, image);
}

// everything that happens to one pair of points: load, all flows, store.
// ONE: small-batch mode -- `pair` is a POINT index, the B lane of every float2 is a dead copy that the compiler removes
// (nothing is stored from it), so a thread walks half the dependent FFMA2 chain and twice as many warps share the work.
template <int H, int K, int VARIANT, bool ONE = false>
__device__ __forceinline__ void process_pair(const FlowProgram &prog, const FastLayout &lay,
                                             const float *__restrict__ params, const float *smem,
                                             const float *__restrict__ x, float *__restrict__ y,
                                             float *__restrict__ log_det, float *__restrict__ base_lp,
                                             float *__restrict__ inter, long long n_rows, int inverse, bool sum_lp,
                                             long long pair, bool live) {
    const bool has_b = ONE ? false : 2 * pair + 1 < n_rows;
    const long long first_pt = ONE ? pair : 2 * pair;  // index of point A
    float2 v0, v1;  // v0 = first coordinate of points (A, B), v1 = second coordinate
    if (has_b) {
        const float4 q = ld_stream4(reinterpret_cast<const float4 *>(x) + pair);
        v0 = make_float2(q.x, q.z);
        v1 = make_float2(q.y, q.w);
    } else {
        const float2 q = ld_stream2(reinterpret_cast<const float2 *>(x) + first_pt);
        v0 = make_float2(q.x, q.x);
        v1 = make_float2(q.y, q.y);
    }
    float2 ld = make_float2(0.f, 0.f);

#pragma unroll 1
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
                const float2 o = affine_net<H, VARIANT>(smem, lay.net_slot[k][which], cond);
                if (which) st[1] = o; else st[0] = o;
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
                const float2 cond = use_f1 ? v0 : v1;
                float2 tr = use_f1 ? v1 : v0;
                spline_half<H, K, VARIANT, ONE>(smem, lay.net_slot[k][use_f1 ? 0 : 1], op, cond, tr, inverse != 0, ld);
                if (use_f1) v1 = tr; else v0 = tr;
            }
        }
        if (inter && live) {
            float *dst = inter + ((size_t)kk * n_rows + first_pt) * 2;
            if (has_b)
                st_stream4(reinterpret_cast<float4 *>(dst), make_float4(v0.x, v1.x, v0.y, v1.y));
            else
                st_stream2(reinterpret_cast<float2 *>(dst), make_float2(v0.x, v1.x));
        }
    }

    const float c = -1.8378770664093453f;  // -(D/2) log(2 pi), D = 2
    float2 lp = make_float2(fmaf(-0.5f, fmaf(v0.x, v0.x, v1.x * v1.x), c),
                            fmaf(-0.5f, fmaf(v0.y, v0.y, v1.y * v1.y), c));
    if (sum_lp) lp = make_float2(lp.x + ld.x, lp.y + ld.y);
    if (!live) return;
    if (has_b) {
        if (y) st_stream4(reinterpret_cast<float4 *>(y) + pair, make_float4(v0.x, v1.x, v0.y, v1.y));
        if (log_det) st_stream2(reinterpret_cast<float2 *>(log_det) + pair, ld);
        if (base_lp) st_stream2(reinterpret_cast<float2 *>(base_lp) + pair, lp);
    } else {
        if (y) st_stream2(reinterpret_cast<float2 *>(y) + first_pt, make_float2(v0.x, v1.x));
        if (log_det) log_det[first_pt] = ld.x;
        if (base_lp) base_lp[first_pt] = lp.x;
    }
}

template <int H, int K, int VARIANT, bool ONE = false>
__global__ void __launch_bounds__(128, 4)
flow_fast_kernel(const __grid_constant__ FlowProgram prog, const __grid_constant__ FastLayout lay,
                 const float *__restrict__ params, const float *__restrict__ x, float *__restrict__ y,
                 float *__restrict__ log_det, float *__restrict__ base_lp, float *__restrict__ inter,
                 long long n_rows, int dir_flags, const float *__restrict__ staged) {
    extern __shared__ __align__(16) float smem[];
    const int inverse = dir_flags & 1;
    const bool sum_lp = dir_flags & 2;
    if (staged != nullptr) {
        // the shared-memory image was laid out once by flow_stage_image_kernel (mnf_flow_stack_stage): plain 16-byte
        // copy, eight loads in flight per thread (total_slots is a multiple of 4)
        const float4 *src = reinterpret_cast<const float4 *>(staged);
        float4 *dst = reinterpret_cast<float4 *>(smem);
        const int n4 = lay.total_slots >> 2;
        for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * blockDim.x < n4) v[u] = src[i0 + u * blockDim.x];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * blockDim.x < n4) dst[i0 + u * blockDim.x] = v[u];
        }
    } else {
        stage_program<H, VARIANT>(prog, lay, params, smem);
    }
    __syncthreads();

    const long long n_pairs = ONE ? n_rows : (n_rows + 1) >> 1;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // persistent grid-stride loop: the nets are staged once per CTA
    for (long long pair = first; pair < n_pairs; pair += (long long)gridDim.x * blockDim.x)
        process_pair<H, K, VARIANT, ONE>(prog, lay, params, smem, x, y, log_det, base_lp, inter, n_rows, inverse, sum_lp,
                                         pair, true);
}

// batches below this many points take the one-point-per-thread form of variant 2 (latency, not throughput, matters)
constexpr long long kOnePointMaxRows = 148LL * 128;

template <int H, int K, int VARIANT, bool ONE = false>
int launch_inst(const FlowProgram &prog, const FastLayout &lay, size_t smem_bytes, const float *params,
                const float *x, float *y, float *log_det, float *base_lp, float *inter, int64_t n_rows,
                int inverse, const DeviceProps *dp, cudaStream_t stream, const float *staged = nullptr) {
    if constexpr (VARIANT == 2 && !ONE) {
        if (n_rows <= kOnePointMaxRows)
            return launch_inst<H, K, 2, true>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, inverse, dp,
                                              stream, staged);
    }
    auto kern = flow_fast_kernel<H, K, VARIANT, ONE>;
    static thread_local int occ_cache = -1;
    static thread_local size_t occ_smem = 0;
    if (occ_cache < 0 || occ_smem != smem_bytes) {
        if (smem_bytes > 48 * 1024)
            MNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int occ = 0;
        MNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem_bytes));
        if (occ < 1) return fail(MNF_E_SHAPE, "flow_fast_kernel does not fit on an SM (smem %zu B)", smem_bytes);
        occ_cache = occ;
        occ_smem = smem_bytes;
    }
    const long long n_pairs = ONE ? n_rows : (n_rows + 1) / 2;
    const int threads = 128;  // measured on config 1: CTAs of 32 / 64 threads lose more to the per-CTA staging of the nets
                              // (128 / 80 us per call) than they gain from spreading the points over more SMs (63 us)
    long long blocks = (n_pairs + threads - 1) / threads;
    const long long cap = (long long)dp->sm_count * occ_cache;  // persistent: one wave, a multiple of the SM count
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, threads, smem_bytes, stream>>>(prog, lay, params, x, y, log_det, base_lp, inter, n_rows,
                                                            inverse & 3, VARIANT == 2 ? staged : nullptr);
    return launch_status("flow_fast_kernel");
}


// =====================================================================================
// VARIANT 3: constant-bank weights.  The stack is cut into segments with at most one
// net-bearing flow each; a segment's two nets are copied (device to device, stream ordered) to
// FIXED offsets of the constant bank, so every weight load is LDCU c[0x3][imm] -> uniform register
// -> FFMA2, with no shared-memory or register-file traffic for weights.  Between segments the points
// and the running log-det make one round trip through HBM (20 B/point, noise next to the FMA time).
// =====================================================================================
// launcher-internal op flags: an NSF_CL flow is cut into its two conditioner/spline halves, one per segment,
// so a segment kernel holds ONE unrolled conditioner (~30 KB of code) instead of two
constexpr uint32_t kFlagHalfA = 0x100u, kFlagHalfB = 0x200u;

template <int H, int K>
struct CbankLayout {
    static constexpr int kSpline = fast_net_slots(H, 3 * K - 1, 2);  // floats per staged spline conditioner
    static constexpr int kAffine = fast_net_slots(H, 1, 2);
    static constexpr int kSegStride = 2 * (kSpline > kAffine ? kSpline : kAffine);
};
static_assert(CbankLayout<24, 8>::kSegStride <= kConstFloats, "segment nets must fit the constant bank");

__device__ float g_flow_stage[MNF_MAX_OPS * 2 * 1900];  // staged nets of a whole program, execution order
static_assert(CbankLayout<24, 8>::kSegStride <= 2 * 1900 && CbankLayout<16, 8>::kSegStride <= 2 * 1900, "stage size");

// one CTA per (segment, slot): re-lay-out the segment's net(s) from the parameter blob into the stage.
// NSF_CL halves own one net (slot 0); an AffineHalfFlow segment owns s_net (slot 0) and t_net (slot 1).
template <int H>
__global__ void cbank_stage_kernel(const __grid_constant__ FlowProgram prog, const float *__restrict__ params,
                                   int seg_stride, int second_off_affine, int inverse) {
    // prog is in EXECUTION order; segment index = number of net-bearing ops before this one
    int target = blockIdx.x >> 1, slot = blockIdx.x & 1, seg = 0;
    for (int k = 0; k < prog.n_ops; ++k) {
        const mnf_flow_op &op = prog.ops[k];
        if (op.type != MNF_OP_AFFINE_HALF && op.type != MNF_OP_NSF_CL) continue;
        if (seg++ != target) continue;
        int which = slot;
        if (op.type == MNF_OP_NSF_CL) {
            if (slot) return;
            // forward: half A = f1, half B = f2 (spline_flow.py:249-266); inverse: half A = f2, half B = f1 (:268-285)
            const bool half_b = op.flags & kFlagHalfB;
            which = (half_b != (inverse != 0)) ? 1 : 0;
        } else if (!(op.flags & (which ? MNF_FLAG_SHIFT : MNF_FLAG_SCALE))) {
            return;
        }
        stage_net<H, 3>(params + op.net_off[which], g_flow_stage + target * seg_stride + (slot ? second_off_affine : 0),
                        op.sizes[op.n_lin]);
        return;
    }
}

template <int H, int K, int BASE>
__device__ __forceinline__ void cb_spline_half(const mnf_flow_op &op, float2 cond, float2 &trans, bool rqs_inverse,
                                               float2 &ld) {
    constexpr int NB = 3 * K - 1, NP2 = round4(NB) / 2;
    constexpr bool FAST = MNF_SPLINE_FAST != 0;
    float2 rA[NP2], rB[NP2];
    {
        const WConstAt<BASE> W;
        float2 hA[H / 2], hB[H / 2];
        op_hidden<H>(W, cond.x, cond.y, hA, hB);
        const int wo = 2 * H + 2 * (H * H + H);
        op_dense<H, NP2>(W, wo, wo + H * 2 * NP2, hA, hB, rA, rB);
    }
#pragma unroll 1
    for (int pt = 0; pt < 2; ++pt) {
        float raw[NB];
#pragma unroll
        for (int o = 0; o < NB; ++o) {
            const float2 q = pt ? rB[o >> 1] : rA[o >> 1];
            raw[o] = (o & 1) ? q.y : q.x;
        }
        float v = pt ? trans.y : trans.x;
        float l = 0.f;
        rq_spline<K, FAST>(raw, K, op.bound, op.edge_deriv, rqs_inverse, v, l);
        if (pt) { trans.y = v; ld.y += l; } else { trans.x = v; ld.x += l; }
    }
}

template <int H, int BASE>
__device__ __forceinline__ float2 cb_affine_net(float2 cond) {
    const WConstAt<BASE> W;
    float2 hA[H / 2], hB[H / 2];
    op_hidden<H>(W, cond.x, cond.y, hA, hB);
    float2 o;
    op_last1<H>(W, hA, hB, o.x, o.y);
    return o;
}

// one segment of the stack (ops in execution order, <= 1 net-bearing op); one pair of points per thread
template <int H, int K>
__global__ void __launch_bounds__(128, 4)
flow_cbank_kernel(const __grid_constant__ FlowProgram prog, const float *__restrict__ params,
                  const float *__restrict__ x, const float *__restrict__ ld_in, float *__restrict__ y,
                  float *__restrict__ log_det, float *__restrict__ base_lp, float *__restrict__ inter,
                  long long n_rows, int dir_flags, const __grid_constant__ mnf_gather_out gather) {
    using L = CbankLayout<H, K>;
    const int inverse = dir_flags & 1;
    const bool sum_lp = dir_flags & 2;
    const long long n_pairs = (n_rows + 1) >> 1;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = first < n_pairs;
    const long long pair = live ? first : n_pairs - 1;  // idle threads redo the last pair, stores predicated off
    const bool has_b = 2 * pair + 1 < n_rows;
    float2 v0, v1, ld = make_float2(0.f, 0.f);
    if (has_b) {
        const float4 q = ld_stream4(reinterpret_cast<const float4 *>(x) + pair);
        v0 = make_float2(q.x, q.z);
        v1 = make_float2(q.y, q.w);
        if (ld_in) ld = ld_stream2(reinterpret_cast<const float2 *>(ld_in) + pair);
    } else {
        const float2 q = ld_stream2(reinterpret_cast<const float2 *>(x) + 2 * pair);
        v0 = make_float2(q.x, q.x);
        v1 = make_float2(q.y, q.y);
        if (ld_in) ld.x = ld.y = ld_in[2 * pair];
    }
    int slot = 0;  // per-flow output slot within this segment
#pragma unroll 1
    for (int kk = 0; kk < prog.n_ops; ++kk) {
        const mnf_flow_op &op = prog.ops[kk];
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
            float2 s = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f);
            if (op.flags & MNF_FLAG_SCALE) s = cb_affine_net<H, 0>(cond);
            if (op.flags & MNF_FLAG_SHIFT) t = cb_affine_net<H, L::kAffine>(cond);
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
            // one half of the coupling layer per segment (its conditioner sits at constant-bank offset 0).
            // Half A conditions on v0 going forward (f1: lower -> upper) and on v1 going backward (f2);
            // half B the other way round.
            const bool cond_v0 = ((op.flags & kFlagHalfB) != 0) == (inverse != 0);
            const float2 cond = cond_v0 ? v0 : v1;
            float2 tr = cond_v0 ? v1 : v0;
            cb_spline_half<H, K, 0>(op, cond, tr, inverse != 0, ld);
            if (cond_v0) v1 = tr; else v0 = tr;
        }
        // per-flow outputs are defined after the flow's LAST half
        if ((op.flags & kFlagHalfA) != 0) continue;
        if (inter && live) {
            float *dst = inter + ((size_t)slot * n_rows + 2 * pair) * 2;
            if (has_b)
                st_stream4(reinterpret_cast<float4 *>(dst), make_float4(v0.x, v1.x, v0.y, v1.y));
            else
                st_stream2(reinterpret_cast<float2 *>(dst), make_float2(v0.x, v1.x));
        }
        ++slot;
    }
    if (!live) return;
    const float c = -1.8378770664093453f;  // -(D/2) log(2 pi), D = 2
    float2 lp = make_float2(fmaf(-0.5f, fmaf(v0.x, v0.x, v1.x * v1.x), c),
                            fmaf(-0.5f, fmaf(v0.y, v0.y, v1.y * v1.y), c));
    if (sum_lp) lp = make_float2(lp.x + ld.x, lp.y + ld.y);
    if (has_b) {
        if (y) st_stream4(reinterpret_cast<float4 *>(y) + pair, make_float4(v0.x, v1.x, v0.y, v1.y));
        if (log_det) st_stream2(reinterpret_cast<float2 *>(log_det) + pair, ld);
        if (base_lp) st_stream2(reinterpret_cast<float2 *>(base_lp) + pair, lp);
        // fused gather: the result also goes straight to the other ranks over NVLink (n_rows is even here)
        if (gather.multicast_ptr) {
            asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(gather.multicast_ptr + gather.row_offset + 2 * pair),
                         "f"(lp.x), "f"(lp.y)
                         : "memory");
        } else {
            for (int p = 0; p < gather.n_peers; ++p)
                st_stream2(reinterpret_cast<float2 *>(gather.peer_ptrs[p] + gather.row_offset) + pair, lp);
        }
    } else {
        if (y) st_stream2(reinterpret_cast<float2 *>(y) + 2 * pair, make_float2(v0.x, v1.x));
        if (log_det) log_det[2 * pair] = ld.x;
        if (base_lp) base_lp[2 * pair] = lp.x;
    }
}

struct ConstBankGuard {  // the bank and the stage are shared by every launch of this TU's kernels on a device
    std::mutex mu;
    cudaEvent_t done[64] = {};
};

// prog: ops in MODULE order.  workspace: 3 * n_rows floats (points + log-det between segments).
template <int H, int K>
int launch_cbank(const FlowProgram &prog, const float *params, const float *x, float *y, float *log_det,
                 float *base_lp, float *inter, int64_t n_rows, int dir_flags, float *workspace,
                 const mnf_gather_out *gather, cudaStream_t stream) {
    using L = CbankLayout<H, K>;
    mnf_gather_out no_gather{};
    if (gather) MNF_REQUIRE(gather->n_peers >= 0 && gather->n_peers <= MNF_MAX_PEERS, MNF_E_ARG, "bad n_peers");
    const int inverse = dir_flags & 1;
    // execution order, NSF_CL flows cut into halves; every net-bearing entry starts a new segment
    struct Exec {
        int n = 0;
        mnf_flow_op ops[2 * MNF_MAX_OPS];
        int flow_index[2 * MNF_MAX_OPS];  // execution index of the flow an entry belongs to (for intermediates)
    } ex;
    for (int k = 0; k < prog.n_ops; ++k) {
        const mnf_flow_op &op = prog.ops[inverse ? prog.n_ops - 1 - k : k];
        if (op.type == MNF_OP_NSF_CL) {
            ex.ops[ex.n] = op, ex.ops[ex.n].flags |= kFlagHalfA, ex.flow_index[ex.n++] = k;
            ex.ops[ex.n] = op, ex.ops[ex.n].flags |= kFlagHalfB, ex.flow_index[ex.n++] = k;
        } else {
            ex.ops[ex.n] = op, ex.flow_index[ex.n++] = k;
        }
    }
    int seg_begin[2 * MNF_MAX_OPS + 1], n_seg = 0, nets_seen = 0;
    seg_begin[n_seg++] = 0;
    for (int k = 0; k < ex.n; ++k) {
        const bool net = ex.ops[k].type == MNF_OP_AFFINE_HALF || ex.ops[k].type == MNF_OP_NSF_CL;
        if (net && nets_seen++ > 0) seg_begin[n_seg++] = k;
    }
    seg_begin[n_seg] = ex.n;
    MNF_REQUIRE(n_seg == 1 || workspace != nullptr, MNF_E_ARG,
                "multi-segment constant-bank run needs a workspace of 3*n_rows floats");
    MNF_REQUIRE(nets_seen <= MNF_MAX_OPS, MNF_E_SHAPE, "too many conditioner nets for the stage (%d)", nets_seen);

    static ConstBankGuard guard;
    int dev = 0;
    MNF_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(guard.mu);
    if (!guard.done[dev]) MNF_CUDA(cudaEventCreateWithFlags(&guard.done[dev], cudaEventDisableTiming));
    else MNF_CUDA(cudaStreamWaitEvent(stream, guard.done[dev], 0));  // other streams: wait, no host sync

    if (nets_seen > 0) {
        // the stage kernel walks net-bearing entries only; hand it those (<= MNF_MAX_OPS of them)
        FlowProgram nets;
        nets.n_ops = 0;
        for (int k = 0; k < ex.n; ++k)
            if (ex.ops[k].type == MNF_OP_AFFINE_HALF || ex.ops[k].type == MNF_OP_NSF_CL) nets.ops[nets.n_ops++] = ex.ops[k];
        cbank_stage_kernel<H><<<2 * nets_seen, 128, 0, stream>>>(nets, params, L::kSegStride, L::kAffine, inverse);
        int rc = launch_status("cbank_stage_kernel");
        if (rc) return rc;
    }
    void *stage_ptr = nullptr;
    MNF_CUDA(cudaGetSymbolAddress(&stage_ptr, g_flow_stage));
    const long long n_pairs = (n_rows + 1) / 2;
    const unsigned blocks = (unsigned)((n_pairs + 127) / 128);
    float *z_tmp = workspace, *ld_tmp = workspace ? workspace + 2 * n_rows : nullptr;
    int net_idx = 0, rc = 0;
    for (int sgm = 0; sgm < n_seg; ++sgm) {
        FlowProgram sp;
        sp.n_ops = seg_begin[sgm + 1] - seg_begin[sgm];
        bool has_net = false;
        for (int k = 0; k < sp.n_ops; ++k) {
            sp.ops[k] = ex.ops[seg_begin[sgm] + k];
            has_net |= sp.ops[k].type == MNF_OP_AFFINE_HALF || sp.ops[k].type == MNF_OP_NSF_CL;
        }
        if (has_net) {
            MNF_CUDA(cudaMemcpyToSymbolAsync(c_flow_w, (const float *)stage_ptr + (size_t)net_idx * L::kSegStride,
                                             sizeof(float) * L::kSegStride, 0, cudaMemcpyDeviceToDevice, stream));
            ++net_idx;
        }
        const bool first = sgm == 0, last = sgm == n_seg - 1;
        // per-flow outputs: entry k of the segment writes slot flow_index (half A entries are skipped in-kernel)
        float *inter_seg = inter ? inter + (size_t)ex.flow_index[seg_begin[sgm]] * n_rows * 2 : nullptr;
        flow_cbank_kernel<H, K><<<blocks, 128, 0, stream>>>(sp, params, first ? x : z_tmp, first ? nullptr : ld_tmp,
                                                            last ? y : z_tmp, last ? log_det : ld_tmp,
                                                            last ? base_lp : nullptr, inter_seg, n_rows, dir_flags,
                                                            (last && gather) ? *gather : no_gather);
        rc = launch_status("flow_cbank_kernel");
        if (rc) break;
    }
    MNF_CUDA(cudaEventRecord(guard.done[dev], stream));
    return rc;
}

#define MNF_FLOW_FAST_ARGS                                                                                    \
    int variant, const FlowProgram &prog, const FastLayout &lay, size_t smem_bytes, const float *params,      \
        const float *x, float *y, float *log_det, float *base_lp, float *inter, int64_t n_rows, int inverse,  \
        float *workspace, const mnf_gather_out *gather, const DeviceProps *dp, cudaStream_t stream

#define MNF_FLOW_FAST_DEFINE(HH, KK)                                                                            \
    int launch_fast_##HH##_##KK(MNF_FLOW_FAST_ARGS) {                                                           \
        if (variant == 0)                                                                                       \
            return launch_inst<HH, KK, 0>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, \
                                          inverse, dp, stream);                                                 \
        if (variant == 1)                                                                                       \
            return launch_inst<HH, KK, 1>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows, \
                                          inverse, dp, stream);                                                 \
        /* constant-bank variant: only the shapes it is validated on at >= 65536 rows (the BASELINE ones); the   \
           (8, 5) instantiation gave wrong results there in r02 testing and other widths are untested -- they take \
           the shared-memory variant at every batch size */                                                      \
        if (variant == 3 && KK == 8 && (HH == 16 || HH == 24))                                                  \
            return launch_cbank<HH, KK>(prog, params, x, y, log_det, base_lp, inter, n_rows, inverse, workspace, \
                                        gather, stream);                                                        \
        if (variant == 3 && gather && (gather->n_peers > 0 || gather->multicast_ptr)) return 1;                 \
        return launch_inst<HH, KK, 2>(prog, lay, smem_bytes, params, x, y, log_det, base_lp, inter, n_rows,     \
                                      inverse, dp, stream, (inverse & 4) ? workspace : nullptr);                \
    }                                                                                                           \
    int stage_image_##HH##_##KK(const FlowProgram &prog, const FastLayout &lay, const float *params,            \
                                float *image, cudaStream_t stream) {                                            \
        flow_stage_image_kernel<HH><<<1, 256, 0, stream>>>(prog, lay, params, image);                           \
        return launch_status("flow_stage_image_kernel");                                                        \
    }

}  // namespace mnf
