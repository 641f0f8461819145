// flow_backward.cu -- reverse-mode pass of a flow stack (the reference trains through
// torch autograd, tests/test_flows.py:14-31; there is no custom backward in the reference to restate, so this
// is the adjoint of the forward kernels in flow_generic.cu).
//
// One thread per point.  For every flow, last to first in execution order, the thread re-evaluates the flow
// from its saved input (the per-flow outputs of the forward pass), propagates (d loss / d output, d loss /
// d log_det) to the flow's input and adds the parameter gradients into a blob with the SAME layout as the
// packed parameter blob (atomicAdd; the host scatters it back into the nn.Parameters).
//   * conditioner MLPs: classic backprop over the recomputed activations
//   * rational-quadratic splines: the bin's rational map is differentiated with 7-wide forward-mode duals
//     (v, x_k, x_k+1, y_k, y_k+1, d_k, d_k+1); the knot construction (two softmaxes, cumsum, pinning) and the
//     double softplus are back-propagated by hand
//   * MAF/IAF sequential direction: D masked-MLP adjoint passes (dimension i only feeds dimensions > i)
// Exact-fp32 path (libdevice math), any dim <= 64 / width <= 128 / K <= 32, like flow_generic.cu.
#include "flow_math.cuh"

namespace mnf {

constexpr int kMaxNetOut = 1024;  // widest conditioner output the backward pass handles

// Parameter-gradient accumulation: the 32 lanes of a warp always add to the SAME parameter (control flow around every
// call is warp-uniform: op sequence, layer sizes and loop bounds come from the program, rows past the end run on
// zeros), so the lanes are summed with shuffles first and one lane issues the atomic.
__device__ __forceinline__ void red_add(float *addr, float v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(addr, v);
}

// ---------------------------------------------------------------------------------------
// 7-wide forward-mode dual numbers
// ---------------------------------------------------------------------------------------
struct D7 {
    float v;
    float d[7];
};
__device__ __forceinline__ D7 d7_const(float c) {
    D7 r;
    r.v = c;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = 0.f;
    return r;
}
__device__ __forceinline__ D7 d7_var(float c, int i) {
    D7 r = d7_const(c);
    r.d[i] = 1.f;
    return r;
}
__device__ __forceinline__ D7 operator+(const D7 &a, const D7 &b) {
    D7 r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ D7 operator-(const D7 &a, const D7 &b) {
    D7 r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ D7 operator*(const D7 &a, const D7 &b) {
    D7 r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
__device__ __forceinline__ D7 operator/(const D7 &a, const D7 &b) {
    D7 r;
    const float inv = 1.f / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
__device__ __forceinline__ D7 operator*(float c, const D7 &a) {
    D7 r;
    r.v = c * a.v;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = c * a.d[i];
    return r;
}
__device__ __forceinline__ D7 d7_sqrt(const D7 &a) {
    D7 r;
    r.v = sqrtf(a.v);
    const float k = a.v > 0.f ? 0.5f / r.v : 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = k * a.d[i];
    return r;
}
__device__ __forceinline__ D7 d7_log(const D7 &a) {
    D7 r;
    r.v = logf(a.v);
    const float k = 1.f / a.v;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.d[i] = k * a.d[i];
    return r;
}

// ---------------------------------------------------------------------------------------
// spline adjoint
// ---------------------------------------------------------------------------------------
// knots of one axis, keeping both softmaxes for the backward pass
__device__ __noinline__ void knots_with_tape(const float *raw, int K, float B, float *knots, float *p1, float *p2) {
    float m = raw[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, raw[k]);
    float sum = 0.f;
    for (int k = 0; k < K; ++k) {
        p1[k] = expf(raw[k] - m);
        sum += p1[k];
    }
    const float twoB = 2.f * B, inv = 1.f / sum;
    float m2 = 0.f;
    for (int k = 0; k < K; ++k) {
        p1[k] *= inv;
        m2 = fmaxf(m2, twoB * p1[k]);
    }
    float sum2 = 0.f;
    for (int k = 0; k < K; ++k) {
        p2[k] = expf(twoB * p1[k] - m2);
        sum2 += p2[k];
    }
    const float inv2 = 1.f / sum2, span = 1.f - kMinBin * (float)K;
    float cum = 0.f;
    knots[0] = -B;
    for (int k = 0; k < K; ++k) {
        p2[k] *= inv2;
        cum += kMinBin + span * p2[k];
        knots[k + 1] = twoB * cum + (-B);
    }
    knots[K] = B;
}

// d loss / d raw[0..K) += given d loss / d knots[j] for j in {k, k+1} (pinned end knots carry no gradient)
__device__ __noinline__ void knots_backward(const float *p1, const float *p2, int K, float B, int k, float g_k, float g_k1,
                               float *graw) {
    const float twoB = 2.f * B, span = 1.f - kMinBin * (float)K;
    const float gk = (k > 0 && k < K) ? g_k : 0.f, gk1 = (k + 1 < K) ? g_k1 : 0.f;
    if (gk == 0.f && gk1 == 0.f) return;
    // knots[j] = -B + 2B * sum_{i<j} (min + span * p2_i)
    float dot2 = 0.f;
    for (int i = 0; i < K; ++i) {
        const float gp2 = twoB * span * ((i < k ? gk : 0.f) + (i < k + 1 ? gk1 : 0.f));
        dot2 += gp2 * p2[i];
    }
    float dot1 = 0.f;
    float gp1_buf[MNF_MAX_BINS];
    for (int i = 0; i < K; ++i) {
        const float gp2 = twoB * span * ((i < k ? gk : 0.f) + (i < k + 1 ? gk1 : 0.f));
        const float gW = p2[i] * (gp2 - dot2);  // second softmax
        gp1_buf[i] = twoB * gW;                 // W = 2B * p1
        dot1 += gp1_buf[i] * p1[i];
    }
    for (int i = 0; i < K; ++i) graw[i] += p1[i] * (gp1_buf[i] - dot1);  // first softmax
}

// Adjoint of rq_spline<0,false> at input v_in: returns d loss / d v_in, adds d loss / d raw into graw[3K-1].
// g_out = d loss / d (spline output), g_ld = d loss / d (log-det accumulator).
__device__ __noinline__ float rq_spline_backward(const float *raw, int K, float B, float edge_deriv, bool inverse, float v_in,
                                    float g_out, float g_ld, float *graw) {
    if (!(v_in >= -B && v_in <= B)) return g_out;  // identity tails, no parameter dependence
    float cw[MNF_MAX_BINS + 1], ch[MNF_MAX_BINS + 1];
    float pw1[MNF_MAX_BINS], pw2[MNF_MAX_BINS], ph1[MNF_MAX_BINS], ph2[MNF_MAX_BINS];
    knots_with_tape(raw, K, B, cw, pw1, pw2);
    knots_with_tape(raw + K, K, B, ch, ph1, ph2);
    const float *sk = inverse ? ch : cw;
    int idx = -1;
    for (int k = 0; k < K; ++k) idx += (v_in >= sk[k]) ? 1 : 0;
    idx += (v_in >= sk[K] + 1e-6f) ? 1 : 0;
    idx = min(max(idx, 0), K - 1);
    const float r0 = idx > 0 ? raw[2 * K + idx - 1] : 0.f, r1 = idx < K - 1 ? raw[2 * K + idx] : 0.f;
    const float dk_v = idx > 0 ? kMinDeriv + softplus(softplus(r0)) : edge_deriv;
    const float dk1_v = idx < K - 1 ? kMinDeriv + softplus(softplus(r1)) : edge_deriv;

    // locals: 0 v, 1 x_k, 2 x_k+1, 3 y_k, 4 y_k+1, 5 d_k, 6 d_k+1
    const D7 v = d7_var(v_in, 0), xk = d7_var(cw[idx], 1), xk1 = d7_var(cw[idx + 1], 2), yk = d7_var(ch[idx], 3),
             yk1 = d7_var(ch[idx + 1], 4), dk = d7_var(dk_v, 5), dk1 = d7_var(dk1_v, 6);
    const D7 wk = xk1 - xk, hk = yk1 - yk, s = hk / wk, dsum = dk + dk1 - 2.f * s, one = d7_const(1.f);
    D7 out, ldc;
    if (inverse) {  // spline_flow.py:133-162
        const D7 dy = v - yk;
        const D7 a = dy * dsum + hk * (s - dk), b = hk * dk - dy * dsum, c = d7_const(0.f) - s * dy;
        D7 disc = b * b - 4.f * (a * c);
        if (disc.v < 0.f) disc = d7_const(0.f);
        const D7 root = (2.f * c) / (d7_const(0.f) - b - d7_sqrt(disc));
        out = root * wk + xk;
        const D7 tt = root * (one - root), den = s + dsum * tt, omr = one - root;
        const D7 num = (s * s) * (dk1 * (root * root) + 2.f * (s * tt) + dk * (omr * omr));
        ldc = d7_const(0.f) - (d7_log(num) - 2.f * d7_log(den));
    } else {  // spline_flow.py:163-179
        const D7 th = (v - xk) / wk, tt = th * (one - th);
        const D7 numer = hk * (s * (th * th) + dk * tt), den = s + dsum * tt, omt = one - th;
        out = yk + numer / den;
        const D7 num = (s * s) * (dk1 * (th * th) + 2.f * (s * tt) + dk * (omt * omt));
        ldc = d7_log(num) - 2.f * d7_log(den);
    }
    float a7[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) a7[i] = g_out * out.d[i] + g_ld * ldc.d[i];
    knots_backward(pw1, pw2, K, B, idx, a7[1], a7[2], graw);
    knots_backward(ph1, ph2, K, B, idx, a7[3], a7[4], graw + K);
    // d/dr [min + log(2 + e^r)] = e^r / (2 + e^r)   (softplus(softplus(r)), both below the threshold 20)
    if (idx > 0) graw[2 * K + idx - 1] += a7[5] * (r0 > 20.f ? 1.f : expf(r0) / (2.f + expf(r0)));
    if (idx < K - 1) graw[2 * K + idx] += a7[6] * (r1 > 20.f ? 1.f : expf(r1) / (2.f + expf(r1)));
    return a7[0];
}

// ---------------------------------------------------------------------------------------
// conditioner MLPs with a tape
// ---------------------------------------------------------------------------------------
struct NetRef {
    const float *w;  // packed parameters of the net
    float *gw;       // same offsets in the gradient blob
    int n_lin;
    const int *sizes;
    bool relu;
};

// forward over all layers; acts[l] = post-activation of hidden layer l; out = last layer (no activation)
__device__ __noinline__ void mlp_forward_tape(const NetRef &net, const float *in, float (*acts)[MNF_MAX_HIDDEN], float *out) {
    const float *w = net.w;
    const float *cur = in;
    for (int l = 0; l < net.n_lin; ++l) {
        const int n_i = net.sizes[l], n_o = net.sizes[l + 1];
        const float *bias = w + n_o * n_i;
        const bool last = l + 1 == net.n_lin;
        float *dst = last ? out : acts[l];
        for (int j = 0; j < n_o; ++j) {
            float acc = bias[j];
            for (int i = 0; i < n_i; ++i) acc = fmaf(w[j * n_i + i], cur[i], acc);
            dst[j] = last ? acc : (net.relu ? fmaxf(acc, 0.f) : leaky02(acc));
        }
        w = bias + n_o;
        cur = dst;
    }
}

// backprop gout (d loss / d net output) -> gin (accumulated into, n_in entries) and parameter gradients
__device__ __noinline__ void mlp_backward(const NetRef &net, const float *in, float (*acts)[MNF_MAX_HIDDEN], const float *gout,
                             float *gin) {
    int off[MNF_MAX_LIN + 1];
    off[0] = 0;
    for (int l = 0; l < net.n_lin; ++l) off[l + 1] = off[l] + net.sizes[l + 1] * net.sizes[l] + net.sizes[l + 1];
    float ga[MNF_MAX_HIDDEN], gb[MNF_MAX_HIDDEN];
    const float *g = gout;
    float *gprev = ga;
    for (int l = net.n_lin - 1; l >= 0; --l) {
        const int n_i = net.sizes[l], n_o = net.sizes[l + 1];
        const float *W = net.w + off[l];
        float *gW = net.gw + off[l], *gbias = gW + n_o * n_i;
        const float *in_l = l == 0 ? in : acts[l - 1];
        for (int i = 0; i < n_i; ++i) gprev[i] = 0.f;
        for (int j = 0; j < n_o; ++j) {
            const float gp = g[j];
            if (__all_sync(0xffffffffu, gp == 0.f)) continue;
            red_add(&gbias[j], gp);
            for (int i = 0; i < n_i; ++i) {
                red_add(&gW[j * n_i + i], gp * in_l[i]);
                gprev[i] = fmaf(W[j * n_i + i], gp, gprev[i]);
            }
        }
        if (l > 0) {
            const float slope = net.relu ? 0.f : 0.2f;
            for (int i = 0; i < n_i; ++i) gprev[i] *= acts[l - 1][i] > 0.f ? 1.f : slope;
            g = gprev;
            gprev = (gprev == ga) ? gb : ga;
        } else {
            for (int i = 0; i < n_i; ++i) gin[i] += gprev[i];
        }
    }
}

struct Scratch {
    float acts[MNF_MAX_LIN][MNF_MAX_HIDDEN];
    float out[kMaxNetOut];
    float gout[kMaxNetOut];
};

// ---------------------------------------------------------------------------------------
// per-op adjoints.  in: the op's input point; g: d loss / d output on entry, d loss / d input on exit.
// ---------------------------------------------------------------------------------------
__device__ void bw_affine_const(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl,
                                int D, bool inverse) {
    const float *s = P + op.aux_off, *t = s + D;
    float *gs = G + op.aux_off, *gt = gs + D;
    for (int d = 0; d < D; ++d) {
        if (inverse) {  // out = (in - t) e^{-s}; ld -= s
            const float e = expf(-s[d]), out = (in[d] - t[d]) * e;
            red_add(&gs[d], -g[d] * out - gl);
            red_add(&gt[d], -g[d] * e);
            g[d] *= e;
        } else {  // out = in e^{s} + t; ld += s
            const float e = expf(s[d]);
            red_add(&gs[d], g[d] * in[d] * e + gl);
            red_add(&gt[d], g[d]);
            g[d] *= e;
        }
    }
}

__device__ void bw_glow(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl, int D,
                        bool inverse) {
    // out = in @ M, M = W (forward) or W^-1 (inverse); the host chains dM and dlogdet back to L, S, U
    const int moff = op.aux_off + (inverse ? D * D : 0);
    const float *M = P + moff;
    float *gM = G + moff;
    float gin[MNF_MAX_DIM];
    for (int i = 0; i < D; ++i) {
        float acc = 0.f;
        for (int j = 0; j < D; ++j) {
            acc = fmaf(M[i * D + j], g[j], acc);
            red_add(&gM[i * D + j], in[i] * g[j]);
        }
        gin[i] = acc;
    }
    for (int i = 0; i < D; ++i) g[i] = gin[i];
    red_add(&G[op.aux_off + 2 * D * D], inverse ? -gl : gl);
}

__device__ void bw_affine_half(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl,
                               int D, bool inverse, Scratch &sc) {
    const int h = D / 2;
    const bool parity = op.flags & MNF_FLAG_PARITY;
    const int co = parity ? h : 0, to = parity ? 0 : h;  // conditioning / transformed halves
    const float *cond = in + co;
    float sv[MNF_MAX_DIM / 2], tv[MNF_MAX_DIM / 2], gs[MNF_MAX_DIM / 2], gt[MNF_MAX_DIM / 2];
    for (int j = 0; j < h; ++j) sv[j] = tv[j] = 0.f;
    NetRef ns{P + op.net_off[0], G + op.net_off[0], op.n_lin, op.sizes, false};
    NetRef nt{P + op.net_off[1], G + op.net_off[1], op.n_lin, op.sizes, false};
    if (op.flags & MNF_FLAG_SCALE) {
        mlp_forward_tape(ns, cond, sc.acts, sc.out);
        for (int j = 0; j < h; ++j) sv[j] = sc.out[j];
    }
    if (op.flags & MNF_FLAG_SHIFT) {
        mlp_forward_tape(nt, cond, sc.acts, sc.out);
        for (int j = 0; j < h; ++j) tv[j] = sc.out[j];
    }
    for (int j = 0; j < h; ++j) {
        const float go = g[to + j];
        if (inverse) {  // out = (x - t) e^{-s}; ld -= s
            const float e = expf(-sv[j]), out = (in[to + j] - tv[j]) * e;
            gs[j] = -go * out - gl;
            gt[j] = -go * e;
            g[to + j] = go * e;
        } else {  // out = e^{s} x + t; ld += s
            const float e = expf(sv[j]);
            gs[j] = go * in[to + j] * e + gl;
            gt[j] = go;
            g[to + j] = go * e;
        }
    }
    if (op.flags & MNF_FLAG_SCALE) {
        mlp_forward_tape(ns, cond, sc.acts, sc.out);
        mlp_backward(ns, cond, sc.acts, gs, g + co);
    }
    if (op.flags & MNF_FLAG_SHIFT) {
        mlp_forward_tape(nt, cond, sc.acts, sc.out);
        mlp_backward(nt, cond, sc.acts, gt, g + co);
    }
}

// one conditioner -> spline half: cond (values, fixed), t_in (inputs of the transformed dims);
// g_t: d loss / d outputs on entry, d loss / d inputs on exit; g_cond accumulated
__device__ void bw_spline_half(const mnf_flow_op &op, const NetRef &net, const float *cond, const float *t_in, int n_t,
                               bool rqs_inverse, float *g_t, float *g_cond, float gl, Scratch &sc) {
    const int nb = 3 * op.K - 1;
    mlp_forward_tape(net, cond, sc.acts, sc.out);
    for (int o = 0; o < n_t * nb; ++o) sc.gout[o] = 0.f;
    for (int j = 0; j < n_t; ++j)
        g_t[j] = rq_spline_backward(sc.out + j * nb, op.K, op.bound, op.edge_deriv, rqs_inverse, t_in[j], g_t[j], gl,
                                    sc.gout + j * nb);
    mlp_backward(net, cond, sc.acts, sc.gout, g_cond);
}

__device__ void bw_nsf_cl(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl, int D,
                          bool inverse, Scratch &sc) {
    const int h = D / 2, nb = 3 * op.K - 1;
    NetRef f1{P + op.net_off[0], G + op.net_off[0], op.n_lin, op.sizes, false};
    NetRef f2{P + op.net_off[1], G + op.net_off[1], op.n_lin, op.sizes, false};
    float mid[MNF_MAX_DIM / 2];  // the half transformed by the first step (input of the second step's conditioner)
    float ld_dummy = 0.f;
    if (!inverse) {
        // forward: A: upper' = RQS(upper; f1(lower)); B: lower' = RQS(lower; f2(upper'))   (spline_flow.py:249-266)
        mlp_forward_tape(f1, in, sc.acts, sc.out);
        for (int j = 0; j < h; ++j) {
            mid[j] = in[h + j];
            rq_spline<0, false>(sc.out + j * nb, op.K, op.bound, op.edge_deriv, false, mid[j], ld_dummy);
        }
        bw_spline_half(op, f2, mid, in, h, false, g, g + h, gl, sc);      // B: transforms lower, conditions on upper'
        bw_spline_half(op, f1, in, in + h, h, false, g + h, g, gl, sc);   // A: transforms upper, conditions on lower
    } else {
        // inverse: A: lower' = RQS^-1(lower; f2(upper)); B: upper' = RQS^-1(upper; f1(lower'))  (spline_flow.py:268-285)
        mlp_forward_tape(f2, in + h, sc.acts, sc.out);
        for (int j = 0; j < h; ++j) {
            mid[j] = in[j];
            rq_spline<0, false>(sc.out + j * nb, op.K, op.bound, op.edge_deriv, true, mid[j], ld_dummy);
        }
        bw_spline_half(op, f1, mid, in + h, h, true, g + h, g, gl, sc);   // B: transforms upper, conditions on lower'
        bw_spline_half(op, f2, in + h, in, h, true, g, g + h, gl, sc);    // A: transforms lower, conditions on upper
    }
}

__device__ void bw_nsf_ar_inverse(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl,
                                  int D, Scratch &sc) {
    // NSF_AR.inverse (spline_flow.py:218-235): dim i goes through the FORWARD spline parameterised by an MLP of the
    // INPUT dims < i (a learned constant for i = 0): no sequential dependency in the adjoint
    const int nb = 3 * op.K - 1;
    int sizes[MNF_MAX_LIN + 1];
    for (int l = 0; l <= op.n_lin; ++l) sizes[l] = op.sizes[l];
    float gin[MNF_MAX_DIM];
    for (int i = 0; i < D; ++i) gin[i] = 0.f;
    int woff = op.net_off[0];
    for (int i = 0; i < D; ++i) {
        if (i == 0) {
            for (int o = 0; o < nb; ++o) sc.gout[o] = 0.f;
            gin[0] += rq_spline_backward(P + op.aux_off, op.K, op.bound, op.edge_deriv, false, in[0], g[0], gl, sc.gout);
            for (int o = 0; o < nb; ++o)
                red_add(&G[op.aux_off + o], sc.gout[o]);
        } else {
            sizes[0] = i;
            NetRef net{P + woff, G + woff, op.n_lin, sizes, false};
            mlp_forward_tape(net, in, sc.acts, sc.out);
            for (int o = 0; o < nb; ++o) sc.gout[o] = 0.f;
            gin[i] += rq_spline_backward(sc.out, op.K, op.bound, op.edge_deriv, false, in[i], g[i], gl, sc.gout);
            mlp_backward(net, in, sc.acts, sc.gout, gin);
            for (int l = 0; l < op.n_lin; ++l) woff += sizes[l + 1] * sizes[l] + sizes[l + 1];
        }
    }
    for (int i = 0; i < D; ++i) g[i] = gin[i];
}

__device__ void bw_nsf_ar_forward(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl,
                                  int D, Scratch &sc) {
    // NSF_AR.forward (spline_flow.py:199-216): out_i = RQS^-1(in_i; theta_i(out_<i)) -- the spline parameters depend on
    // the flow's own OUTPUTS.  Rebuild the outputs, then walk i downwards: the gradient reaching out_i is complete
    // once every i' > i has pushed its conditioner gradient back (same scheme as the sequential MADE direction).
    const int nb = 3 * op.K - 1;
    int sizes[MNF_MAX_LIN + 1];
    for (int l = 0; l <= op.n_lin; ++l) sizes[l] = op.sizes[l];
    int woff[MNF_MAX_DIM];
    float outv[MNF_MAX_DIM], gin[MNF_MAX_DIM];
    {
        int w = op.net_off[0];
        for (int i = 0; i < D; ++i) {
            woff[i] = w;
            float ld_dummy = 0.f, x = in[i];
            if (i == 0) {
                rq_spline<0, false>(P + op.aux_off, op.K, op.bound, op.edge_deriv, true, x, ld_dummy);
            } else {
                sizes[0] = i;
                NetRef net{P + w, G + w, op.n_lin, sizes, false};
                mlp_forward_tape(net, outv, sc.acts, sc.out);
                rq_spline<0, false>(sc.out, op.K, op.bound, op.edge_deriv, true, x, ld_dummy);
                for (int l = 0; l < op.n_lin; ++l) w += sizes[l + 1] * sizes[l] + sizes[l + 1];
            }
            outv[i] = x;
        }
    }
    for (int i = D - 1; i >= 0; --i) {
        for (int o = 0; o < nb; ++o) sc.gout[o] = 0.f;
        if (i == 0) {
            gin[0] = rq_spline_backward(P + op.aux_off, op.K, op.bound, op.edge_deriv, true, in[0], g[0], gl, sc.gout);
            for (int o = 0; o < nb; ++o) red_add(&G[op.aux_off + o], sc.gout[o]);
        } else {
            sizes[0] = i;
            NetRef net{P + woff[i], G + woff[i], op.n_lin, sizes, false};
            mlp_forward_tape(net, outv, sc.acts, sc.out);
            gin[i] = rq_spline_backward(sc.out, op.K, op.bound, op.edge_deriv, true, in[i], g[i], gl, sc.gout);
            mlp_backward(net, outv, sc.acts, sc.gout, g);  // adds to d loss / d out_<i
        }
    }
    for (int i = 0; i < D; ++i) g[i] = gin[i];
}

__device__ void bw_made(const mnf_flow_op &op, const float *P, float *G, const float *in, float *g, float gl, int D,
                        bool inverse, Scratch &sc) {
    const bool parity = op.flags & MNF_FLAG_PARITY;
    const bool sequential = (op.flags & MNF_FLAG_MADE_SEQ) ? !inverse : inverse;
    NetRef net{P + op.net_off[0], G + op.net_off[0], op.n_lin, op.sizes, true};
    float gin[MNF_MAX_DIM];
    if (!sequential) {
        // out[flip(i)] = in_i e^{s_i} + t_i, ld += sum s   (maf.py:53-62)
        mlp_forward_tape(net, in, sc.acts, sc.out);
        for (int i = 0; i < D; ++i) {
            const float go = g[parity ? D - 1 - i : i], e = expf(sc.out[i]);
            sc.gout[i] = go * in[i] * e + gl;  // d/ds_i
            sc.gout[D + i] = go;               // d/dt_i
            gin[i] = go * e;
        }
        mlp_backward(net, in, sc.acts, sc.gout, gin);
        for (int i = 0; i < D; ++i) g[i] = gin[i];
    } else {
        // x_i = (z_{f(i)} - t_i(x_<i)) e^{-s_i(x_<i)}, ld -= s_i   (maf.py:39-51).  Rebuild x, then walk i downwards:
        // the gradient reaching x_i is complete once all i' > i have been processed (autoregressive masks).
        float xs[MNF_MAX_DIM];
        for (int i = 0; i < D; ++i) xs[i] = 0.f;
        for (int i = 0; i < D; ++i) {
            mlp_forward_tape(net, xs, sc.acts, sc.out);
            xs[i] = (in[parity ? D - 1 - i : i] - sc.out[D + i]) * expf(-sc.out[i]);
        }
        float gx[MNF_MAX_DIM];
        for (int i = 0; i < D; ++i) {
            gx[i] = g[i];
            gin[i] = 0.f;
        }
        mlp_forward_tape(net, xs, sc.acts, sc.out);  // s_i, t_i of every i from the completed x
        for (int i = D - 1; i >= 0; --i) {
            const float e = expf(-sc.out[i]);
            for (int o = 0; o < 2 * D; ++o) sc.gout[o] = 0.f;
            sc.gout[i] = -gx[i] * xs[i] - gl;
            sc.gout[D + i] = -gx[i] * e;
            gin[parity ? D - 1 - i : i] = gx[i] * e;
            mlp_backward(net, xs, sc.acts, sc.gout, gx);  // reaches only x_<i
        }
        for (int i = 0; i < D; ++i) g[i] = gin[i];
    }
}

__global__ void __launch_bounds__(128)
flow_backward_kernel(const __grid_constant__ FlowProgram prog, const float *__restrict__ params,
                     float *gparams, const float *__restrict__ x, const float *__restrict__ inter,
                     const float *__restrict__ gy, const float *__restrict__ gld, const float *__restrict__ ginter,
                     float *__restrict__ gx, long long n_rows, int D, int inverse, int n_params, int grads_in_smem) {
    // parameter gradients are summed per block in shared memory when the blob fits (one flush of global atomics
    // per block instead of one global atomic per point and weight)
    extern __shared__ float smem_grad[];
    float *const gout = gparams;
    if (grads_in_smem) {
        for (int i = threadIdx.x; i < n_params; i += blockDim.x) smem_grad[i] = 0.f;
        __syncthreads();
        gparams = smem_grad;
    }
    Scratch sc;
    float g[MNF_MAX_DIM], in[MNF_MAX_DIM];
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n_rows; base += (long long)gridDim.x * blockDim.x) {
        const long long row = base + threadIdx.x;
        const bool valid = row < n_rows;  // lanes past the end run on zeros so that the warp stays converged
        for (int d = 0; d < D; ++d) g[d] = (valid && gy) ? gy[row * D + d] : 0.f;
        const float gl = (valid && gld) ? gld[row] : 0.f;
        for (int kk = prog.n_ops - 1; kk >= 0; --kk) {
            const mnf_flow_op &op = prog.ops[inverse ? prog.n_ops - 1 - kk : kk];
            if (ginter && valid)  // the caller also used this flow's output directly (an element of the returned list)
                for (int d = 0; d < D; ++d) g[d] += ginter[((size_t)kk * n_rows + row) * D + d];
            const float *src = kk == 0 ? x + row * D : inter + ((size_t)(kk - 1) * n_rows + row) * D;
            for (int d = 0; d < D; ++d) in[d] = valid ? src[d] : 0.f;
            switch (op.type) {
                case MNF_OP_AFFINE_CONST: bw_affine_const(op, params, gparams, in, g, gl, D, inverse); break;
                case MNF_OP_GLOW: bw_glow(op, params, gparams, in, g, gl, D, inverse); break;
                case MNF_OP_AFFINE_HALF: bw_affine_half(op, params, gparams, in, g, gl, D, inverse, sc); break;
                case MNF_OP_NSF_CL: bw_nsf_cl(op, params, gparams, in, g, gl, D, inverse, sc); break;
                case MNF_OP_NSF_AR:
                    if (inverse) bw_nsf_ar_inverse(op, params, gparams, in, g, gl, D, sc);
                    else bw_nsf_ar_forward(op, params, gparams, in, g, gl, D, sc);
                    break;
                case MNF_OP_MADE: bw_made(op, params, gparams, in, g, gl, D, inverse, sc); break;
                default: break;
            }
        }
        if (gx && valid)
            for (int d = 0; d < D; ++d) gx[row * D + d] = g[d];
    }
    if (grads_in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_params; i += blockDim.x)
            if (smem_grad[i] != 0.f) atomicAdd(&gout[i], smem_grad[i]);
    }
}

int validate_program(const mnf_flow_op *ops, int n_ops, int dim, int64_t n_params);

}  // namespace mnf

using namespace mnf;

extern "C" {

int mnf_flow_stack_backward(const mnf_flow_op *ops_host, int n_ops, const float *params, int64_t n_params,
                            float *grad_params, const float *x, const float *intermediates, const float *grad_y,
                            const float *grad_log_det, const float *grad_intermediates, float *grad_x,
                            int64_t n_rows, int dim, int flags, void *stream) {
    int rc = validate_program(ops_host, n_ops, dim, n_params);
    if (rc) return rc;
    MNF_REQUIRE(params && grad_params && x && (intermediates || n_ops <= 1), MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE((flags & ~MNF_RUN_INVERSE) == 0, MNF_E_ARG, "only MNF_RUN_INVERSE is meaningful here");
    MNF_REQUIRE(n_rows >= 0, MNF_E_ARG, "negative n_rows");
    const int inverse = flags & MNF_RUN_INVERSE;
    for (int k = 0; k < n_ops; ++k) {
        const mnf_flow_op &op = ops_host[k];
        if (op.type >= MNF_OP_AFFINE_HALF)
            MNF_REQUIRE(op.sizes[op.n_lin] <= kMaxNetOut || op.type == MNF_OP_NSF_AR, MNF_E_SHAPE,
                        "conditioner output %d wider than %d", op.sizes[op.n_lin], kMaxNetOut);
    }
    if (n_rows == 0 || n_ops == 0) return 0;
    FlowProgram prog;
    prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) prog.ops[k] = ops_host[k];
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr, MNF_E_DEVICE, "no CUDA device");
    const size_t smem = sizeof(float) * (size_t)n_params;
    const int in_smem = smem <= 96 * 1024;
    if (in_smem && smem > 48 * 1024)
        MNF_CUDA(cudaFuncSetAttribute(flow_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the kernel is latency-bound (local-memory scratch, shuffles): 8 blocks x 4 warps per SM where the per-block gradient
    // copy allows it (ncu at 2 warps x 4 blocks: issue slots 15 % busy, profiles/r01_flow_backward_ncu_full.md)
    long long blocks = (n_rows + 127) / 128;
    long long per_sm = in_smem ? (long long)(200 * 1024 / (smem > 0 ? smem : 1)) : 8;
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    const long long cap = (long long)dp->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    flow_backward_kernel<<<(unsigned)blocks, 128, in_smem ? smem : 0, (cudaStream_t)stream>>>(
        prog, params, grad_params, x, intermediates, grad_y, grad_log_det, grad_intermediates, grad_x, n_rows, dim, inverse,
        (int)n_params, in_smem);
    return launch_status("flow_backward_kernel");
}

}  // extern "C"
