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

__global__ void __launch_bounds__(256) kl_rows_kernel(const KlRowArgs a) {
    __shared__ float red[32];
    const int row = blockIdx.x;
    const size_t base = (size_t)row * a.cols;
    const Philox rng(a.seed);
    float kl = 0.f, dot_mean = 0.f, dot_noise = 0.f;
    auto element = [&](size_t f, int c, float wmean, float lv, float e) {
        const float var = expf(lv);
        const float zz = a.conv ? a.z[f / a.z_block] : a.z[c];
        const float wm = wmean * zz;
        kl += -logf(var) + var + wm * wm - 1.f;  // mnf_linear.py:73 / mnf_conv.py:99 (-W_var.log())
        const float rc = a.r0_c[c];
        dot_mean = fmaf(rc, wm, dot_mean);
        const float sd = sqrtf(var);
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
    kl = block_sum(kl, red);
    dot_mean = block_sum(dot_mean, red);
    dot_noise = block_sum(dot_noise, red);
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

__global__ void __launch_bounds__(1024) kl_final_kernel(const KlFinalArgs a) {
    __shared__ float red[32];
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

}  // namespace mnf

using namespace mnf;

extern "C" {

int mnf_kl_div(const mnf_kl_args *a, void *stream) {
    MNF_REQUIRE(a != nullptr, MNF_E_ARG, "args is NULL");
    MNF_REQUIRE(a->W_mean && a->W_log_var && a->z && a->zT && a->r0_c && a->r0_b1 && a->r0_b2 && a->b_log_var &&
                    a->q0_log_var && a->ld_q && a->ld_r && a->workspace && a->out,
                MNF_E_ARG, "NULL pointer in mnf_kl_args");
    MNF_REQUIRE(a->n_out >= 1 && a->n_in >= 1 && a->ksize >= 1, MNF_E_ARG, "bad shape");
    MNF_REQUIRE(a->conv == 0 || a->conv == 1, MNF_E_ARG, "conv must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int fan = a->n_in * a->ksize * a->ksize;
    const int rows = a->conv ? fan : a->n_out;   // mnf_conv.py:107: view(-1, n_out)
    const int cols = a->conv ? a->n_out : fan;
    float *kl_rows = a->workspace, *act = a->workspace + rows;
    KlRowArgs ra{a->W_mean, a->W_log_var, a->z, a->r0_c, a->eps_w, a->seed, a->noise_stream, rows, cols,
                 a->conv, fan, kl_rows, act};
    kl_rows_kernel<<<rows, 256, 0, st>>>(ra);
    int rc = launch_status("kl_rows_kernel");
    if (rc) return rc;
    KlFinalArgs fa{kl_rows, act, rows, a->conv ? nullptr : a->b_mean, a->b_log_var, a->n_out, a->q0_log_var,
                   a->r0_b1, a->r0_b2, a->zT, a->conv ? a->n_out : a->n_in, a->ld_q, a->ld_r, a->conv, a->r0_c,
                   a->eps_b, a->seed, a->noise_stream + 1, a->out};
    kl_final_kernel<<<1, 1024, 0, st>>>(fa);
    return launch_status("kl_final_kernel");
}

}  // extern "C"
