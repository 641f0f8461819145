// mnf_layers.cu -- MNF layer kernels, exact-fp32 path (arbitrary shapes):
//   * z0 = q0_mean + sqrt(exp(q0_log_var)) * eps                      (mnf_linear.py:59-62)
//   * RNVP forward with injected or Philox masks                       (rnvp.py:25-39)
//   * MNFLinear.forward: dual GEMM sharing the x tile + noise epilogue (mnf_linear.py:46-56)
//   * MNFConv2d.forward as an implicit GEMM, optional fused ReLU + 2x2 max-pool epilogue
//                                                                      (mnf_conv.py:67-78)
// All GEMM-shaped work goes through the functor skeleton in mnf_common.cuh.
#include "mnf_common.cuh"

namespace mnf {

struct NoiseSrc {
    const float *ptr;  // injected tensor or nullptr -> Philox(seed, stream) indexed by global element
    uint64_t seed;
    uint32_t stream;
    uint64_t row_offset;  // global index of local row 0 (sharded runs draw the same numbers)
};

__device__ __forceinline__ float noise_normal(const NoiseSrc &s, long long local_idx, long long global_idx) {
    return s.ptr ? s.ptr[local_idx] : philox_normal(Philox(s.seed), (uint64_t)global_idx, s.stream);
}
__device__ __forceinline__ float noise_bernoulli(const NoiseSrc &s, long long local_idx, long long global_idx) {
    return s.ptr ? s.ptr[local_idx] : philox_bernoulli(Philox(s.seed), (uint64_t)global_idx, s.stream);
}

// ---------------------------------------------------------------------------------------
__global__ void sample_z0_kernel(const float *__restrict__ q0_mean, const float *__restrict__ q0_log_var,
                                 NoiseSrc eps, float *__restrict__ z, long long n_rows, int dim) {
    const long long total = n_rows * dim;
    const Philox rng(eps.seed);
    // four consecutive elements per step: one Philox block yields exactly their four normals
    // (same element -> number mapping as philox_normal)
    for (long long e0 = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); e0 < total;
         e0 += 4LL * gridDim.x * blockDim.x) {
        float n4[4];
        if (eps.ptr == nullptr) {
            const uint4 q = rng((uint64_t)((long long)eps.row_offset * dim + e0) >> 2, eps.stream);
            const float2 a = box_muller(q.x, q.y), b = box_muller(q.z, q.w);
            n4[0] = a.x, n4[1] = a.y, n4[2] = b.x, n4[3] = b.y;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long e = e0 + u;
            if (e >= total) break;
            const int d = (int)(e % dim);
            const float std = sqrtf(expf(q0_log_var[d]));  // .exp().sqrt(), mnf_linear.py:59
            // (row_offset * dim) % 4 == 0 is required for the block alignment; otherwise fall back per element
            const float nz = eps.ptr ? eps.ptr[e]
                                     : ((((long long)eps.row_offset * dim) & 3) == 0
                                            ? n4[u]
                                            : philox_normal(rng, (uint64_t)((long long)eps.row_offset * dim + e), eps.stream));
            z[e] = q0_mean[d] + std * nz;
        }
    }
}

// ---------------------------------------------------------------------------------------
// plain Linear (+ optional LeakyReLU 0.2): extra conditioner layers when h_sizes has > 1 entry
struct LinearProb {
    int M, N, K;
    const float *A, *W, *b;
    float *C;
    int leaky;
    static constexpr bool kRowReduce = false;
    using RowCtx = size_t;
    __device__ RowCtx row_ctx(int m) const { return (size_t)m * K; }
    __device__ void load_a(const RowCtx &r, int, int k, float (&a)[1]) const { a[0] = A[r + k]; }
    __device__ void load_b(int n, int k, float (&v)[1]) const { v[0] = W[(size_t)n * K + k]; }
    __device__ void epilogue4(int m0, int n, const float (&acc)[1][4], float (&)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (m0 + i >= M) continue;
            float v = acc[0][i] + b[n];
            if (leaky) v = fmaxf(v, 0.2f * v);
            C[(size_t)(m0 + i) * N + n] = v;
        }
    }
    __device__ void row_out(int, float) const {}
};

// RNVP stage 1: y = (mask * z) @ Wn^T + bn        (rnvp.py:30-31)
struct RnvpHiddenProb {
    int M, N, K;  // rows, hidden, dim
    const float *z, *Wn, *bn;
    float *y;
    NoiseSrc mask;
    int leaky;
    static constexpr bool kRowReduce = false;
    using RowCtx = long long;
    __device__ RowCtx row_ctx(int m) const { return (long long)m * K; }
    __device__ void load_a(const RowCtx &r, int, int k, float (&a)[1]) const {
        const long long e = r + k;
        a[0] = noise_bernoulli(mask, e, (long long)mask.row_offset * K + e) * z[e];
    }
    __device__ void load_b(int n, int k, float (&v)[1]) const { v[0] = Wn[(size_t)n * K + k]; }
    __device__ void epilogue4(int m0, int n, const float (&acc)[1][4], float (&)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (m0 + i >= M) continue;
            float v = acc[0][i] + bn[n];
            if (leaky) v = fmaxf(v, 0.2f * v);
            y[(size_t)(m0 + i) * N + n] = v;
        }
    }
    __device__ void row_out(int, float) const {}
};

// RNVP stage 2: shift = y Wt^T + bt, scale = y Ws^T + bs, gated update + log-det (rnvp.py:32-39)
struct RnvpOutProb {
    int M, N, K;  // rows, dim, hidden
    const float *y, *Wt, *bt, *Ws, *bs;
    float *z;       // in/out [M, N]
    float *logdet;  // [M], accumulated with atomics over column tiles (zeroed by the caller)
    NoiseSrc mask;
    static constexpr bool kRowReduce = true;
    using RowCtx = size_t;
    __device__ RowCtx row_ctx(int m) const { return (size_t)m * K; }
    __device__ void load_a(const RowCtx &r, int, int k, float (&a)[2]) const { a[0] = a[1] = y[r + k]; }
    __device__ void load_b(int n, int k, float (&v)[2]) const {
        v[0] = Wt[(size_t)n * K + k];
        v[1] = Ws[(size_t)n * K + k];
    }
    __device__ void epilogue4(int m0, int n, const float (&acc)[2][4], float (&rowsum)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m >= M) continue;
            const long long e = (long long)m * N + n;
            const float mk = noise_bernoulli(mask, e, (long long)mask.row_offset * N + e);
            const float shift = acc[0][i] + bt[n], scale = acc[1][i] + bs[n];
            const float gate = 1.f / (1.f + expf(-scale));  // torch.sigmoid
            const float zz = z[e];
            const float z1 = (1.f - mk) * zz, z2 = mk * zz;
            z[e] = (z1 * gate + (1.f - gate) * shift) + z2;  // rnvp.py:37 (shift reaches kept dims too)
            rowsum[i] += (1.f - mk) * logf(gate);           // rnvp.py:36
        }
    }
    __device__ void row_out(int m, float sum) const { atomicAdd(&logdet[m], sum); }
};

// MNFLinear.forward: mean = (x*z) Wm^T + bm ; var = x^2 exp(Wlv)^T + exp(blv) ; out = mean + sqrt(var) eps
struct MnfLinearProb {
    int M, N, K;  // rows, n_out, n_in
    const float *x;
    int x_rows;      // row m reads x[m % x_rows] (MC replication without materialising x.repeat)
    const float *z;  // [M, K]
    const float *Wm, *Wlv, *bm, *blv;
    float *out;
    NoiseSrc eps;
    int relu;
    static constexpr bool kRowReduce = false;
    struct RowCtx {
        size_t xo, zo;
    };
    __device__ RowCtx row_ctx(int m) const { return RowCtx{(size_t)(m % x_rows) * K, (size_t)m * K}; }
    __device__ void load_a(const RowCtx &r, int, int k, float (&a)[2]) const {
        const float xv = x[r.xo + k];
        a[0] = xv * z[r.zo + k];
        a[1] = xv * xv;
    }
    __device__ void load_b(int n, int k, float (&v)[2]) const {
        v[0] = Wm[(size_t)n * K + k];
        v[1] = expf(Wlv[(size_t)n * K + k]);
    }
    __device__ void epilogue4(int m0, int n, const float (&acc)[2][4], float (&)[4]) const {
        const float b0 = bm[n], b1 = expf(blv[n]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m >= M) continue;
            const long long e = (long long)m * N + n;
            const float mean = acc[0][i] + b0, var = acc[1][i] + b1;
            const float v = mean + sqrtf(var) * noise_normal(eps, e, (long long)eps.row_offset * N + e);
            out[e] = relu ? fmaxf(v, 0.f) : v;
        }
    }
    __device__ void row_out(int, float) const {}
};

// MNFConv2d.forward as an implicit GEMM: rows = output pixels, cols = output channels,
// K = n_in*k*k.  With `pool` the row index is pool-major (4 consecutive rows = one 2x2 window)
// and the epilogue applies ReLU + max over the window (the nn.ReLU / nn.MaxPool2d(2) that
// follow every MNFConv2d in MNFLeNet, mnf_lenet.py:16-21).
struct MnfConvProb {
    int M, N, K;
    const float *x;  // [R, C, H, W]
    int x_imgs;      // image r reads x[r % x_imgs]
    int C, H, W, ks, OH, OW;
    const float *Wm, *Wlv, *blv;  // [N, C, ks, ks], [N]
    const float *z;               // [N] multiplicative noise, shared by the batch (mnf_conv.py:72)
    float *out;                   // [R, N, OH, OW] or pooled [R, N, OH/2, OW/2]
    NoiseSrc eps;                 // indexed like the un-pooled output [R, N, OH, OW]
    int pool;
    static constexpr bool kRowReduce = false;
    __device__ __forceinline__ void decode(int m, int &r, int &oy, int &ox) const {
        if (pool) {
            const int q = m & 3, w = m >> 2;
            const int PW = OW >> 1, PH = OH >> 1;
            const int px = w % PW, py = (w / PW) % PH;
            r = w / (PW * PH);
            oy = 2 * py + (q >> 1);
            ox = 2 * px + (q & 1);
        } else {
            ox = m % OW;
            oy = (m / OW) % OH;
            r = m / (OW * OH);
        }
    }
    // offset of filter tap k inside an image, (ci*H + ky)*W + kx, tabulated on the host (K <= kMaxTaps)
    static constexpr int kMaxTaps = 640;
    int koff[kMaxTaps];
    using RowCtx = size_t;  // offset of the output pixel's receptive-field origin in x
    __device__ RowCtx row_ctx(int m) const {
        int r, oy, ox;
        decode(m, r, oy, ox);
        return ((size_t)(r % x_imgs) * C * H + oy) * W + ox;
    }
    __device__ void load_a(const RowCtx &r, int, int k, float (&a)[2]) const {
        const float xv = x[r + koff[k]];
        a[0] = xv;
        a[1] = xv * xv;
    }
    __device__ void load_b(int n, int k, float (&v)[2]) const {
        v[0] = Wm[(size_t)n * K + k] * z[n];
        v[1] = expf(Wlv[(size_t)n * K + k]);
    }
    __device__ void epilogue4(int m0, int n, const float (&acc)[2][4], float (&)[4]) const {
        const float bvar = expf(blv[n]);
        float best = 0.f;  // ReLU floor
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m >= M) continue;
            int r, oy, ox;
            decode(m, r, oy, ox);
            const long long e = (((long long)r * N + n) * OH + oy) * OW + ox;
            const long long ge = ((((long long)r + (long long)eps.row_offset) * N + n) * OH + oy) * OW + ox;
            const float v = acc[0][i] + sqrtf(acc[1][i] + bvar) * noise_normal(eps, e, ge);  // b_mean == 0
            if (pool)
                best = fmaxf(best, v);
            else
                out[e] = v;
        }
        if (pool && m0 < M) {
            int r, oy, ox;
            decode(m0, r, oy, ox);
            out[(((size_t)r * N + n) * (OH >> 1) + (oy >> 1)) * (OW >> 1) + (ox >> 1)] = best;
        }
    }
    __device__ void row_out(int, float) const {}
};

}  // namespace mnf

using namespace mnf;

extern "C" {

int mnf_sample_z0(const float *q0_mean, const float *q0_log_var, const float *eps, uint64_t seed,
                  uint32_t noise_stream, uint64_t row_offset, float *z, int64_t n_rows, int dim, void *stream) {
    MNF_REQUIRE(q0_mean && q0_log_var && z, MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE(n_rows >= 0 && dim >= 1, MNF_E_ARG, "bad shape");
    if (n_rows == 0) return 0;
    const long long total = (long long)n_rows * dim;
    long long blocks = (total / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    sample_z0_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        q0_mean, q0_log_var, NoiseSrc{eps, seed, noise_stream, row_offset}, z, n_rows, dim);
    return launch_status("sample_z0_kernel");
}

int mnf_rnvp_forward(const mnf_rnvp_flow *flows_host, int n_flows, float *z, float *log_det,
                     const float *const *masks_host, uint64_t seed, uint32_t first_noise_stream,
                     uint64_t row_offset, int64_t n_rows, int dim, float *workspace, float *intermediates,
                     void *stream) {
    MNF_REQUIRE(flows_host && z && log_det && workspace, MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE(n_flows >= 0 && n_rows >= 0 && dim >= 1, MNF_E_ARG, "bad shape");
    MNF_REQUIRE(n_rows <= 0x7fffffff / 64, MNF_E_SHAPE, "too many rows for one call (%lld)", (long long)n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) return 0;
    MNF_CUDA(cudaMemsetAsync(log_det, 0, sizeof(float) * n_rows, st));
    for (int f = 0; f < n_flows; ++f) {
        const mnf_rnvp_flow &fl = flows_host[f];
        MNF_REQUIRE(fl.n_net >= 1 && fl.n_net <= MNF_RNVP_MAX_NET, MNF_E_SHAPE, "flow %d: n_net=%d", f, fl.n_net);
        const NoiseSrc mask{masks_host ? masks_host[f] : nullptr, seed, first_noise_stream + (uint32_t)f, row_offset};
        // conditioner net (MLP(dim, *h_sizes), last activation dropped -- mlp.py:12)
        int maxh = 1;
        for (int g = 0; g < n_flows; ++g)
            for (int l = 0; l < flows_host[g].n_net && l < MNF_RNVP_MAX_NET; ++l)
                maxh = flows_host[g].net_sizes[l] > maxh ? flows_host[g].net_sizes[l] : maxh;
        float *ya = workspace, *yb = workspace + (size_t)n_rows * maxh;
        RnvpHiddenProb hp{(int)n_rows, fl.net_sizes[0], dim, z, fl.net_w[0], fl.net_b[0], ya, mask, fl.n_net > 1};
        int rc = launch_simt_gemm<RnvpHiddenProb, 1>(hp, st, "rnvp_hidden");
        if (rc) return rc;
        for (int l = 1; l < fl.n_net; ++l) {
            LinearProb lp{(int)n_rows, fl.net_sizes[l], fl.net_sizes[l - 1], ya, fl.net_w[l], fl.net_b[l], yb,
                          l + 1 < fl.n_net};
            rc = launch_simt_gemm<LinearProb, 1>(lp, st, "rnvp_net_layer");
            if (rc) return rc;
            float *t = ya;
            ya = yb;
            yb = t;
        }
        RnvpOutProb op{(int)n_rows, dim, fl.net_sizes[fl.n_net - 1], ya, fl.t_w, fl.t_b, fl.s_w, fl.s_b, z, log_det, mask};
        rc = launch_simt_gemm<RnvpOutProb, 2>(op, st, "rnvp_out");
        if (rc) return rc;
        if (intermediates)
            MNF_CUDA(cudaMemcpyAsync(intermediates + (size_t)f * n_rows * dim, z, sizeof(float) * n_rows * dim,
                                     cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int mnf_linear_forward(const float *x, int64_t x_rows, const float *z, const float *W_mean, const float *W_log_var,
                       const float *b_mean, const float *b_log_var, const float *eps, uint64_t seed,
                       uint32_t noise_stream, uint64_t row_offset, float *out, int64_t n_rows, int n_in, int n_out,
                       int relu, void *stream) {
    MNF_REQUIRE(x && z && W_mean && W_log_var && b_mean && b_log_var && out, MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE(n_rows >= 0 && n_in >= 1 && n_out >= 1 && x_rows >= 1, MNF_E_ARG, "bad shape");
    MNF_REQUIRE(n_rows <= 0x7fffffff / 64, MNF_E_SHAPE, "too many rows for one call (%lld)", (long long)n_rows);
    MnfLinearProb p{(int)n_rows, n_out, n_in, x, (int)x_rows, z, W_mean, W_log_var, b_mean, b_log_var, out,
                    NoiseSrc{eps, seed, noise_stream, row_offset}, relu};
    return launch_simt_gemm<MnfLinearProb, 2>(p, (cudaStream_t)stream, "mnf_linear_forward");
}

int mnf_conv2d_forward(const float *x, int64_t x_imgs, const float *z, const float *W_mean, const float *W_log_var,
                       const float *b_log_var, const float *eps, uint64_t seed, uint32_t noise_stream,
                       uint64_t row_offset, float *out, int64_t n_imgs, int c_in, int height, int width, int c_out,
                       int ksize, int relu_pool, void *stream) {
    MNF_REQUIRE(x && z && W_mean && W_log_var && b_log_var && out, MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE(n_imgs >= 0 && c_in >= 1 && c_out >= 1 && ksize >= 1 && x_imgs >= 1, MNF_E_ARG, "bad shape");
    const int OH = height - ksize + 1, OW = width - ksize + 1;
    MNF_REQUIRE(OH >= 1 && OW >= 1, MNF_E_SHAPE, "kernel %d larger than input %dx%d", ksize, height, width);
    MNF_REQUIRE(!relu_pool || (OH % 2 == 0 && OW % 2 == 0), MNF_E_SHAPE,
                "fused 2x2 max-pool needs even output size, got %dx%d", OH, OW);
    const long long M = (long long)n_imgs * OH * OW;
    MNF_REQUIRE(M <= 0x7fffffff - 64, MNF_E_SHAPE, "too many output pixels for one call (%lld): chunk the batch", M);
    MNF_REQUIRE(c_in * ksize * ksize <= MnfConvProb::kMaxTaps, MNF_E_SHAPE, "c_in*k*k = %d exceeds %d filter taps",
                c_in * ksize * ksize, MnfConvProb::kMaxTaps);
    MnfConvProb p{(int)M, c_out, c_in * ksize * ksize, x, (int)x_imgs, c_in, height, width, ksize, OH, OW,
                  W_mean, W_log_var, b_log_var, z, out, NoiseSrc{eps, seed, noise_stream, row_offset}, relu_pool, {}};
    for (int k = 0; k < p.K; ++k) {
        const int kx = k % ksize, ky = (k / ksize) % ksize, ci = k / (ksize * ksize);
        p.koff[k] = (ci * height + ky) * width + kx;
    }
    return launch_simt_gemm<MnfConvProb, 2>(p, (cudaStream_t)stream, "mnf_conv2d_forward");
}

}  // extern "C"
