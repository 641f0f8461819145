// mnf_kl.cu -- the weight-space part of MNFLinear.kl_div / MNFConv2d.kl_div
// (mnf_linear.py:66-90, mnf_conv.py:90-133).  The two RNVP stacks (flow_q before, flow_r after)
// run through mnf_rnvp_forward with one row; this file does the HBM-bound pass over the weights
// and the final scalar:
//   pass 1 (one CTA per matrix row, float4-coalesced, warp-shuffle + smem reduction):
//       kl_row  = sum(-log_var + exp(log_var) + (z * W_mean)^2 - 1)
//       act_row = tanh( r0_c . (z*W_mean + sqrt(exp(log_var)) * eps_w) )            (linear)
//       act_row =       r0_c . (z*W_mean) + (r0_c . sqrt(exp(log_var))) * eps_w[row] (conv, on the
//                 reference's view(-1, n_out) reinterpretation of the [n_out,n_in,k,k] memory)
//   pass 2 (one CTA): kl_W, kl_b, log_q, mean(act), log_r and the final sum.
#include <cooperative_groups.h>

#include "mnf_common.cuh"

namespace mnf {

__device__ __forceinline__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (w == 0) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    return red[0];
}

// three block sums with one shared-memory exchange (the weight pass ends every row with them)
__device__ __forceinline__ void block_sum3(float &a, float &b, float &c, float *red) {
    a = warp_sum(a), b = warp_sum(b), c = warp_sum(c);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) red[w] = a, red[32 + w] = b, red[64 + w] = c;
    __syncthreads();
    if (w == 0) {
        float x = lane < nw ? red[lane] : 0.f, y = lane < nw ? red[32 + lane] : 0.f, z = lane < nw ? red[64 + lane] : 0.f;
        x = warp_sum(x), y = warp_sum(y), z = warp_sum(z);
        if (lane == 0) red[0] = x, red[32] = y, red[64] = z;
    }
    __syncthreads();
    a = red[0], b = red[32], c = red[64];
}

struct KlRowArgs {
    const float *W_mean, *W_log_var;  // flat [rows * cols]
    const float *z;                   // multiplicative noise after flow_q
    const float *r0_c;                // [cols]
    const float *eps;                 // linear: [rows*cols]; conv: [rows]; nullptr -> Philox
    uint64_t seed;
    uint32_t noise_stream;
    int rows, cols;
    int conv;        // 0: z indexed by column; 1: z indexed by flat / z_block (output channel of the element)
    int z_block;     // conv: n_in * k * k
    float *kl_rows;  // [rows]
    float *act;      // [rows]
};

__device__ __forceinline__ void kl_rows_body(const KlRowArgs &a, int row, float *red) {
    const size_t base = (size_t)row * a.cols;
    const Philox rng(a.seed);
    float kl = 0.f, dot_mean = 0.f, dot_noise = 0.f;
    auto element = [&](size_t f, int c, float wmean, float lv, float e) {
        // sd = exp(lv / 2) with one EX2, var = sd^2, and -log(var) = -lv: the pass is bound by this arithmetic and the
        // Philox draws, not by HBM (2^-22 relative error per term against the 2e-5 tolerance of the sum)
        float sd;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(0.7213475204444817f * lv));
        const float var = sd * sd;
        const float zz = a.conv ? a.z[f / a.z_block] : a.z[c];
        const float wm = wmean * zz;
        kl += (var - lv) + fmaf(wm, wm, -1.f);  // mnf_linear.py:73 / mnf_conv.py:99 (-W_var.log())
        const float rc = a.r0_c[c];
        dot_mean = fmaf(rc, wm, dot_mean);
        dot_noise = fmaf(rc, a.conv ? sd : sd * e, dot_noise);  // conv: W_std row . r0_c, scaled by eps_w[row] below
    };
    if ((a.cols & 3) == 0) {
        // float4 path: one pass over W_mean / W_log_var (and eps_w) at full sector width; one Philox block gives
        // the four normals of four consecutive elements (same mapping as philox_normal)
        for (int c = 4 * threadIdx.x; c < a.cols; c += 4 * blockDim.x) {
            const size_t f = base + c;
            const float4 wm4 = *reinterpret_cast<const float4 *>(a.W_mean + f);
            const float4 lv4 = *reinterpret_cast<const float4 *>(a.W_log_var + f);
            float e4[4] = {0.f, 0.f, 0.f, 0.f};
            if (!a.conv) {
                if (a.eps) {
                    const float4 q = *reinterpret_cast<const float4 *>(a.eps + f);
                    e4[0] = q.x, e4[1] = q.y, e4[2] = q.z, e4[3] = q.w;
                } else {
                    const uint4 q = rng((uint64_t)f >> 2, a.noise_stream);
                    const float2 n01 = box_muller(q.x, q.y), n23 = box_muller(q.z, q.w);
                    e4[0] = n01.x, e4[1] = n01.y, e4[2] = n23.x, e4[3] = n23.y;
                }
            }
            element(f, c, wm4.x, lv4.x, e4[0]);
            element(f + 1, c + 1, wm4.y, lv4.y, e4[1]);
            element(f + 2, c + 2, wm4.z, lv4.z, e4[2]);
            element(f + 3, c + 3, wm4.w, lv4.w, e4[3]);
        }
    } else {
        for (int c = threadIdx.x; c < a.cols; c += blockDim.x) {
            const size_t f = base + c;
            const float e = a.conv ? 0.f : (a.eps ? a.eps[f] : philox_normal(rng, f, a.noise_stream));
            element(f, c, a.W_mean[f], a.W_log_var[f], e);
        }
    }
    block_sum3(kl, dot_mean, dot_noise, red);
    if (threadIdx.x == 0) {
        a.kl_rows[row] = kl;
        if (a.conv) {
            const float e = a.eps ? a.eps[row] : philox_normal(rng, (uint64_t)row, a.noise_stream);
            a.act[row] = dot_mean + dot_noise * e;  // linear activation for conv layers, mnf_conv.py:111-115
        } else {
            a.act[row] = tanhf(dot_mean + dot_noise);  // mnf_linear.py:81
        }
    }
}

__global__ void __launch_bounds__(256) kl_rows_kernel(const KlRowArgs a) {
    __shared__ float red[96];
    kl_rows_body(a, blockIdx.x, red);
}

struct KlFinalArgs {
    const float *kl_rows, *act;
    int rows;
    const float *b_mean;  // nullptr for conv (b_mean is the zero tensor of mnf_conv.py:45)
    const float *b_log_var;
    int n_b;
    const float *q0_log_var, *r0_b1, *r0_b2, *zT;
    int n_z;
    const float *ld_q, *ld_r;  // device scalars from the two RNVP stacks
    int conv;
    const float *r0_c;   // conv: bias term of the auxiliary activation
    const float *eps_b;  // conv: scalar noise (device) or nullptr -> Philox
    uint64_t seed;
    uint32_t noise_stream;
    float *out;  // [1] result; out[1..4] = kl_W, kl_b, log_q, log_r for inspection
};

__device__ __forceinline__ void kl_final_body(const KlFinalArgs &a, float *red) {
    float s = 0.f, sa = 0.f;
    for (int r = threadIdx.x; r < a.rows; r += blockDim.x) {
        s += a.kl_rows[r];
        sa += a.act[r];
    }
    const float kl_W = 0.5f * block_sum(s, red);
    float mean_act = block_sum(sa, red) / (float)a.rows;

    float sb = 0.f, sbv = 0.f;
    for (int i = threadIdx.x; i < a.n_b; i += blockDim.x) {
        const float lv = a.b_log_var[i], bm = a.b_mean ? a.b_mean[i] : 0.f;
        const float var = expf(lv);
        sb += (a.conv ? -logf(var) : -lv) + var + bm * bm - 1.f;  // mnf_linear.py:74-76 / mnf_conv.py:100
        if (a.conv) sbv = fmaf(var, a.r0_c[i] * a.r0_c[i], sbv);  // mnf_conv.py:117
    }
    const float kl_b = 0.5f * block_sum(sb, red);
    if (a.conv) {
        const float bv = block_sum(sbv, red);
        const float e = a.eps_b ? a.eps_b[0] : philox_normal(Philox(a.seed), 0, a.noise_stream);
        mean_act += sqrtf(bv) * e;  // act += b_mean(=0) + sqrt(b_var) * eps_b, same shift for every row
    }
    float sq = 0.f;
    for (int i = threadIdx.x; i < a.n_z; i += blockDim.x) sq += a.q0_log_var[i];
    const float log_q = -a.ld_q[0] - 0.5f * block_sum(sq, red);

    float sr = 0.f;
    for (int i = threadIdx.x; i < a.n_z; i += blockDim.x) {
        const float mean_r = a.r0_b1[i] * mean_act;     // ger(b1, act).mean(1) == b1 * mean(act)
        const float log_var_r = a.r0_b2[i] * mean_act;  // eq. (10)
        const float d = a.zT[i] - mean_r;
        sr += -expf(log_var_r) * d * d + log_var_r;
    }
    const float log_r = a.ld_r[0] + 0.5f * block_sum(sr, red);
    if (threadIdx.x == 0) {
        a.out[0] = kl_W + kl_b + log_q - log_r;
        a.out[1] = kl_W;
        a.out[2] = kl_b;
        a.out[3] = log_q;
        a.out[4] = log_r;
    }
}


__global__ void __launch_bounds__(1024) kl_final_kernel(const KlFinalArgs a) {
    __shared__ float red[32];
    kl_final_body(a, red);
}

// ---------------------------------------------------------------------------------------------------------
// z0, flow_q and flow_r of one kl_div() call on their single row, in ONE launch (they were 1 + n_q + n_r launches
// plus a copy): a thread-block cluster of 8 CTAs; per flow the CTAs split the conditioner outputs, meet at a cluster
// barrier, split the dims for the gate, and meet again before the next flow reads z.
// ---------------------------------------------------------------------------------------------------------
struct KlFlowDev {
    const float *W0, *b0, *Wt, *bt, *Ws, *bs, *mask;
    uint32_t mask_stream;
    int Hn;
};
struct KlFlowsArgs {
    int nq, nr, dim;
    uint32_t z_stream;
    uint64_t seed;
    const float *q0_mean, *q0_log_var, *eps_z;
    float *z, *zT, *ld_q, *ld_r, *ybuf;
    KlFlowDev f[2 * MNF_KL_MAX_FLOWS];
};
constexpr int kKlCluster = 8;

__device__ __forceinline__ void kl_flows_body(const KlFlowsArgs &a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, dim = a.dim;
    extern __shared__ __align__(16) float sm_kl[];
    float *mz = sm_kl, *ys = sm_kl + dim, *red = ys + 64;
    const int per = (dim + kKlCluster - 1) / kKlCluster, d0 = c * per, d1 = min(dim, d0 + per);
    const Philox rng(a.seed);
    // The flows' weights (2.4 MB per flow at dim 4096) are read in a chain of dependent phases by 8 SMs: cold, every
    // phase pays DRAM latency with little parallelism (r02 ncu: 120 us, issue 7 %, long-scoreboard stalls).  Ask for all
    // of them now -- one L2 prefetch per 128-byte line, spread over the cluster's 2048 threads -- so that the phases
    // below find them in L2.
    {
        const int gtid = c * 256 + tid, gthreads = kKlCluster * 256;
        for (int f = 0; f < a.nq + a.nr; ++f) {
            const KlFlowDev &fl = a.f[f];
            const size_t lines = ((size_t)fl.Hn * dim * sizeof(float) + 127) / 128;
            for (size_t l = gtid; l < lines; l += gthreads) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(fl.W0) + 128 * l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(fl.Wt) + 128 * l));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(fl.Ws) + 128 * l));
            }
        }
    }
    for (int d = d0 + tid; d < d1; d += 256) {  // z0 = q0_mean + q0_std * eps (mnf_linear.py:58-62), this CTA's dims
        const float nz = a.eps_z ? a.eps_z[d] : philox_normal(rng, (uint64_t)d, a.z_stream);
        a.z[d] = a.q0_mean[d] + sqrtf(expf(a.q0_log_var[d])) * nz;
    }
    if (c == 0 && tid == 0) *a.ld_q = 0.f, *a.ld_r = 0.f;
    __threadfence();
    cluster.sync();
    float *cur = a.z;
    for (int f = 0; f < a.nq + a.nr; ++f) {
        const KlFlowDev &fl = a.f[f];
        if (f == a.nq) {  // flow_r runs on a copy: z itself feeds the weight pass (mnf_linear.py:83)
            for (int d = d0 + tid; d < d1; d += 256) a.zT[d] = __ldcg(a.z + d);
            cur = a.zT;
            __threadfence();
            cluster.sync();
        }
        const NoiseSrc mask{fl.mask, a.seed, fl.mask_stream, 0};
        // (every loop of this kernel is latency-bound on a few warps: unrolled so that several loads are in flight)
#pragma unroll 8
        for (int d = tid; d < dim; d += 256) mz[d] = noise_bernoulli(mask, d, (long long)d) * __ldcg(cur + d);
        __syncthreads();
        // phase 1: conditioner outputs y = W0 (mask z) + b0, one WARP per output (j = c, c + 8, ... over the cluster's
        // 64 warps), 16-byte loads, no block-level synchronisation inside the phase
        for (int j = c + kKlCluster * warp; j < fl.Hn; j += kKlCluster * 8) {
            const float *w0 = fl.W0 + (size_t)j * dim;
            float acc = 0.f;
            if ((dim & 3) == 0 && (reinterpret_cast<uintptr_t>(fl.W0) & 15) == 0) {
#pragma unroll 8
                for (int d = 4 * lane; d < dim; d += 128) {
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(w0 + d)), m = *reinterpret_cast<const float4 *>(mz + d);
                    acc = fmaf(w.x, m.x, fmaf(w.y, m.y, fmaf(w.z, m.z, fmaf(w.w, m.w, acc))));
                }
            } else {
                for (int d = lane; d < dim; d += 32) acc = fmaf(w0[d], mz[d], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) a.ybuf[j] = acc + fl.b0[j];
        }
        __threadfence();
        cluster.sync();  // every y[j] is written, every CTA has finished reading z
        for (int j = tid; j < fl.Hn; j += 256) ys[j] = __ldcg(a.ybuf + j);
        __syncthreads();
        float ldsum = 0.f;
        for (int d = d0 + tid; d < d1; d += 256) {  // phase 2: this CTA's dims, one thread per dim
            float shift = fl.bt[d], scale = fl.bs[d];
            const float *wt = fl.Wt + (size_t)d * fl.Hn, *ws = fl.Ws + (size_t)d * fl.Hn;
            if ((fl.Hn & 1) == 0 && ((reinterpret_cast<uintptr_t>(fl.Wt) | reinterpret_cast<uintptr_t>(fl.Ws)) & 7) == 0) {  // rows of Wt / Ws are 8-byte aligned: paired loads, independent of each other
                const float2 *wt2 = reinterpret_cast<const float2 *>(wt), *ws2 = reinterpret_cast<const float2 *>(ws);
                float sh1 = 0.f, sc1 = 0.f;
#pragma unroll 5
                for (int j = 0; j < fl.Hn / 2; ++j) {
                    const float2 t2 = __ldg(wt2 + j), s2 = __ldg(ws2 + j);
                    shift = fmaf(t2.x, ys[2 * j], shift), sh1 = fmaf(t2.y, ys[2 * j + 1], sh1);
                    scale = fmaf(s2.x, ys[2 * j], scale), sc1 = fmaf(s2.y, ys[2 * j + 1], sc1);
                }
                shift += sh1, scale += sc1;
            } else {
                for (int j = 0; j < fl.Hn; ++j) {
                    shift = fmaf(wt[j], ys[j], shift);
                    scale = fmaf(ws[j], ys[j], scale);
                }
            }
            const float mk = noise_bernoulli(mask, d, (long long)d);
            const float gate = 1.f / (1.f + expf(-scale)), zz = __ldcg(cur + d);
            cur[d] = ((1.f - mk) * zz * gate + (1.f - gate) * shift) + mk * zz;  // rnvp.py:37
            ldsum += (1.f - mk) * logf(gate);                                     // rnvp.py:36
        }
        ldsum = warp_sum(ldsum);
        if (lane == 0) red[warp] = ldsum;
        __syncthreads();
        if (tid == 0) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += red[w];
            atomicAdd(f < a.nq ? a.ld_q : a.ld_r, v);
        }
        __threadfence();
        cluster.sync();  // z of this flow is complete (and y, ys are free) before the next flow starts
    }
}

__global__ void __cluster_dims__(kKlCluster, 1, 1) __launch_bounds__(256) kl_flows_kernel(const __grid_constant__ KlFlowsArgs a) {
    kl_flows_body(a);
}

// ---------------------------------------------------------------------------------------------------------
// The kl_div() of SEVERAL layers in the same three launches (MNFLeNet.kl_div, mnf_lenet.py:28-32, sums four): the layers
// are independent and each one's flow kernel is latency-bound on its 8 SMs, so cluster l of the grid runs layer l's
// flows, grid row l of the weight pass layer l's rows, CTA l of the final kernel layer l's reduction.
// ---------------------------------------------------------------------------------------------------------
constexpr int kKlMultiMax = 4;
struct KlFlowsMulti { KlFlowsArgs a[kKlMultiMax]; };
struct KlRowsMulti { KlRowArgs a[kKlMultiMax]; };
struct KlFinalMulti { KlFinalArgs a[kKlMultiMax]; };

__global__ void __cluster_dims__(kKlCluster, 1, 1) __launch_bounds__(256) kl_flows_multi_kernel(const __grid_constant__ KlFlowsMulti m) {
    kl_flows_body(m.a[blockIdx.x / kKlCluster]);
}
__global__ void __launch_bounds__(256) kl_rows_multi_kernel(const __grid_constant__ KlRowsMulti m) {
    __shared__ float red[96];
    const KlRowArgs &a = m.a[blockIdx.y];
    if ((int)blockIdx.x >= a.rows) return;
    kl_rows_body(a, blockIdx.x, red);
}
__global__ void __launch_bounds__(1024) kl_final_multi_kernel(const __grid_constant__ KlFinalMulti m) {
    __shared__ float red[32];
    kl_final_body(m.a[blockIdx.x], red);
}

// argument blocks of the three kernels from the ABI structs (shared by the one-layer and the multi-layer entry points)
static int fill_rows_final(const mnf_kl_args *a, KlRowArgs &ra, KlFinalArgs &fa) {
    MNF_REQUIRE(a != nullptr, MNF_E_ARG, "args is NULL");
    MNF_REQUIRE(a->W_mean && a->W_log_var && a->z && a->zT && a->r0_c && a->r0_b1 && a->r0_b2 && a->b_log_var &&
                    a->q0_log_var && a->ld_q && a->ld_r && a->workspace && a->out,
                MNF_E_ARG, "NULL pointer in mnf_kl_args");
    MNF_REQUIRE(a->n_out >= 1 && a->n_in >= 1 && a->ksize >= 1, MNF_E_ARG, "bad shape");
    MNF_REQUIRE(a->conv == 0 || a->conv == 1, MNF_E_ARG, "conv must be 0 or 1");
    const int fan = a->n_in * a->ksize * a->ksize;
    const int rows = a->conv ? fan : a->n_out;   // mnf_conv.py:107: view(-1, n_out)
    const int cols = a->conv ? a->n_out : fan;
    float *kl_rows = a->workspace, *act = a->workspace + rows;
    ra = KlRowArgs{a->W_mean, a->W_log_var, a->z, a->r0_c, a->eps_w, a->seed, a->noise_stream, rows, cols,
                   a->conv, fan, kl_rows, act};
    fa = KlFinalArgs{kl_rows, act, rows, a->conv ? nullptr : a->b_mean, a->b_log_var, a->n_out, a->q0_log_var,
                     a->r0_b1, a->r0_b2, a->zT, a->conv ? a->n_out : a->n_in, a->ld_q, a->ld_r, a->conv, a->r0_c,
                     a->eps_b, a->seed, a->noise_stream + 1, a->out};
    return 0;
}

static int fill_flows(const mnf_kl_fused_args *a, KlFlowsArgs &fa) {
    MNF_REQUIRE(a != nullptr, MNF_E_ARG, "args is NULL");
    const mnf_kl_args &k = a->kl;
    MNF_REQUIRE(k.z && k.zT && k.ld_q && k.ld_r && k.workspace && k.q0_log_var && a->q0_mean, MNF_E_ARG, "NULL pointer in mnf_kl_fused_args");
    MNF_REQUIRE(a->n_flows_q >= 0 && a->n_flows_q <= MNF_KL_MAX_FLOWS && a->n_flows_r >= 0 && a->n_flows_r <= MNF_KL_MAX_FLOWS,
                MNF_E_SHAPE, "at most %d flows per stack", MNF_KL_MAX_FLOWS);
    MNF_REQUIRE(k.n_out >= 1 && k.n_in >= 1 && k.ksize >= 1, MNF_E_ARG, "bad shape");
    const int dim = k.conv ? k.n_out : k.n_in;
    MNF_REQUIRE(dim <= 11000, MNF_E_SHAPE, "dim=%d does not fit the one-row cluster kernel", dim);
    const int fan = k.n_in * k.ksize * k.ksize, rows = k.conv ? fan : k.n_out;
    fa = KlFlowsArgs{};
    fa.nq = a->n_flows_q, fa.nr = a->n_flows_r, fa.dim = dim, fa.z_stream = a->z_stream, fa.seed = k.seed;
    fa.q0_mean = a->q0_mean, fa.q0_log_var = k.q0_log_var, fa.eps_z = a->eps_z;
    fa.z = const_cast<float *>(k.z), fa.zT = const_cast<float *>(k.zT);
    fa.ld_q = const_cast<float *>(k.ld_q), fa.ld_r = const_cast<float *>(k.ld_r);
    fa.ybuf = k.workspace + 2 * (size_t)rows;
    for (int f = 0; f < fa.nq + fa.nr; ++f) {
        const int src = f < fa.nq ? f : MNF_KL_MAX_FLOWS + (f - fa.nq);
        const mnf_rnvp_flow &fl = a->flows[src];
        MNF_REQUIRE(fl.n_net == 1 && fl.net_sizes[0] >= 1 && fl.net_sizes[0] <= 64, MNF_E_SHAPE,
                    "fused kl_div needs single-Linear RNVP conditioners of width <= 64");
        MNF_REQUIRE(fl.net_w[0] && fl.net_b[0] && fl.t_w && fl.t_b && fl.s_w && fl.s_b, MNF_E_ARG, "NULL pointer in flow %d", f);
        fa.f[f] = KlFlowDev{fl.net_w[0], fl.net_b[0], fl.t_w, fl.t_b, fl.s_w, fl.s_b, a->masks[src], a->mask_streams[src], fl.net_sizes[0]};
    }
    return 0;
}


}  // namespace mnf

using namespace mnf;

extern "C" {

int mnf_kl_div(const mnf_kl_args *a, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    KlRowArgs ra;
    KlFinalArgs fa;
    int rc = fill_rows_final(a, ra, fa);
    if (rc) return rc;
    kl_rows_kernel<<<ra.rows, 256, 0, st>>>(ra);
    rc = launch_status("kl_rows_kernel");
    if (rc) return rc;
    kl_final_kernel<<<1, 1024, 0, st>>>(fa);
    return launch_status("kl_final_kernel");
}

int mnf_kl_div_fused(const mnf_kl_fused_args *a, void *stream) {
    KlFlowsArgs fa;
    int rc = fill_flows(a, fa);
    if (rc) return rc;
    // (a one-CTA, all-shared-memory form of this kernel was measured too: 462 us at dim 4096, 46-89 us on MNF-LeNet's
    // layers against 120 / 41-57 us for the cluster -- the time is cold-miss latency of the dependent weight reads)
    kl_flows_kernel<<<kKlCluster, 256, sizeof(float) * (fa.dim + 64 + 8), (cudaStream_t)stream>>>(fa);
    rc = launch_status("kl_flows_kernel");
    if (rc) return rc;
    return mnf_kl_div(&a->kl, stream);
}

int mnf_kl_div_fused_multi(const mnf_kl_fused_args *const *layers, int n_layers, void *stream) {
    MNF_REQUIRE(layers != nullptr && n_layers >= 1, MNF_E_ARG, "no layers");
    cudaStream_t st = (cudaStream_t)stream;
    for (int first = 0; first < n_layers; first += kKlMultiMax) {  // groups of up to kKlMultiMax layers per launch triple
        const int n = n_layers - first < kKlMultiMax ? n_layers - first : kKlMultiMax;
        KlFlowsMulti fm{};
        KlRowsMulti rm{};
        KlFinalMulti lm{};
        int max_dim = 0, max_rows = 0;
        for (int l = 0; l < n; ++l) {
            int rc = fill_flows(layers[first + l], fm.a[l]);
            if (rc) return rc;
            rc = fill_rows_final(&layers[first + l]->kl, rm.a[l], lm.a[l]);
            if (rc) return rc;
            max_dim = fm.a[l].dim > max_dim ? fm.a[l].dim : max_dim;
            max_rows = rm.a[l].rows > max_rows ? rm.a[l].rows : max_rows;
        }
        kl_flows_multi_kernel<<<kKlCluster * n, 256, sizeof(float) * (max_dim + 64 + 8), st>>>(fm);
        int rc = launch_status("kl_flows_multi_kernel");
        if (rc) return rc;
        kl_rows_multi_kernel<<<dim3((unsigned)max_rows, (unsigned)n), 256, 0, st>>>(rm);
        rc = launch_status("kl_rows_multi_kernel");
        if (rc) return rc;
        kl_final_multi_kernel<<<n, 1024, 0, st>>>(lm);
        rc = launch_status("kl_final_multi_kernel");
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
