// made_fast.cu -- exact-fp32 density pass of a stack of MAF / IAF flows (maf.py:53-62) for the BASELINE config-3
// shape (dim 64, MADE 64-24-24-24-128): one launch per flow, one persistent 12-warp block per SM, two rows per
// lane, every warp owning a private 64-row tile (no block barriers in the steady state), packed
// FFMA2 with the masked weights broadcast from shared memory (LDS.128: four weights feed four FFMA2).  A first
// version kept the weights in the constant bank as uniform operands like the dim-2 flow kernel; its 115 KB of
// straight-line code ran at 2.8 stalled cycles per issue waiting for instructions (profiles/r01_made_fast_*).  The K=64 / N=24 contractions are far too small for a tensor-core tile to pay:
// at 5760 FMA per row and flow the fp32 pipe finishes 2^20 rows x 9 flows in ~2 ms, the TF32 GEMM chain
// (36 launches over HBM-resident activations) needs ~7.5 ms and is only TF32-accurate.
//
// Data movement per launch: a block stages 256 rows x 64 floats in shared memory with coalesced 16-byte
// streaming loads (row stride 65 floats: conflict-free per-thread row walks), every thread keeps only the 24
// hidden activations of its two rows in registers, writes z_i over x_i in place and the block stores the tile
// back coalesced (flipped when the flow has parity).  Flows after the first run in place on the output buffer.
#include "flow_math.cuh"

namespace mnf {

template <int D, int H>
struct MadeLayout {  // float offsets inside the constant bank; weights transposed to [in][out]
    static constexpr int W1 = 0;            // [D][H]
    static constexpr int B1 = W1 + D * H;   // [H]
    static constexpr int W2 = B1 + H;       // [H][H]
    static constexpr int B2 = W2 + H * H;
    static constexpr int W3 = B2 + H;
    static constexpr int B3 = W3 + H * H;
    static constexpr int W4 = B3 + H;          // [H][D][2]: (s_i, t_i) pairs
    static constexpr int B4 = W4 + H * 2 * D;  // [D][2]
    static constexpr int kFloats = B4 + 2 * D;
    static_assert(H % 4 == 0 && D % 8 == 0, "pairs of pairs");
};

__device__ __forceinline__ float4 lds4(const float *W, int i) { return reinterpret_cast<const float4 *>(W)[i >> 2]; }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// blob layout (torch): W1[H][D] b1 W2[H][H] b2 W3[H][H] b3 W4[2D][H] b4[2D]  ->  MadeLayout in shared memory
template <int D, int H>
__device__ __forceinline__ void made_stage_net(const float *__restrict__ src, float *dst) {
    using L = MadeLayout<D, H>;
    const float *w1 = src, *b1 = w1 + H * D, *w2 = b1 + H, *b2 = w2 + H * H, *w3 = b2 + H, *b3 = w3 + H * H,
                *w4 = b3 + H, *b4 = w4 + 2 * D * H;
#pragma unroll 4
    for (int e = threadIdx.x; e < D * H; e += blockDim.x) dst[L::W1 + (e % D) * H + e / D] = w1[e];
#pragma unroll 4
    for (int e = threadIdx.x; e < H * H; e += blockDim.x) {
        dst[L::W2 + (e % H) * H + e / H] = w2[e];
        dst[L::W3 + (e % H) * H + e / H] = w3[e];
    }
#pragma unroll 4
    for (int e = threadIdx.x; e < H; e += blockDim.x) {
        dst[L::B1 + e] = b1[e];
        dst[L::B2 + e] = b2[e];
        dst[L::B3 + e] = b3[e];
    }
#pragma unroll 4
    for (int e = threadIdx.x; e < 2 * D * H; e += blockDim.x) {
        const int row = e / H, k = e % H, c = row / D, i = row % D;  // c = 0: s rows, 1: t rows
        dst[L::W4 + k * 2 * D + 2 * i + c] = w4[e];
    }
#pragma unroll 4
    for (int e = threadIdx.x; e < 2 * D; e += blockDim.x) dst[L::B4 + 2 * (e % D) + e / D] = b4[e];
}

// out[j] = relu(bias[j] + sum_i Wt[i][j] in[i]) for two rows; weights broadcast from shared memory
template <int H>
__device__ __forceinline__ void dense_relu(const float *W, int wt_off, int bias_off, const float2 (&inA)[H / 2],
                                           const float2 (&inB)[H / 2], float2 (&outA)[H / 2], float2 (&outB)[H / 2]) {
#pragma unroll
    for (int j = 0; j < H / 2; j += 2) {
        const float4 b = lds4(W, bias_off + 2 * j);
        outA[j] = outB[j] = make_float2(b.x, b.y);
        outA[j + 1] = outB[j + 1] = make_float2(b.z, b.w);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float a = (i & 1) ? inA[i >> 1].y : inA[i >> 1].x;
        const float b = (i & 1) ? inB[i >> 1].y : inB[i >> 1].x;
        const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
        for (int j = 0; j < H / 2; j += 2) {
            const float4 w = lds4(W, wt_off + i * H + 2 * j);
            outA[j] = ffma2(make_float2(w.x, w.y), aa, outA[j]);
            outB[j] = ffma2(make_float2(w.x, w.y), bb, outB[j]);
            outA[j + 1] = ffma2(make_float2(w.z, w.w), aa, outA[j + 1]);
            outB[j + 1] = ffma2(make_float2(w.z, w.w), bb, outB[j + 1]);
        }
    }
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        outA[j] = make_float2(fmaxf(outA[j].x, 0.f), fmaxf(outA[j].y, 0.f));
        outB[j] = make_float2(fmaxf(outB[j].x, 0.f), fmaxf(outB[j].y, 0.f));
    }
}

constexpr int kMadeWarps = 12;      // warps per block; one block per SM (shared memory: net + 12 private tiles)
constexpr int kMadeWarpRows = 64;   // rows per warp tile: 2 per lane

// Persistent block, one per SM: the net is staged once, then every WARP walks its own tiles of 64 rows with no
// block-level synchronisation, so the load / store phases of one warp overlap the arithmetic of the others.
// lp_mode: 0 none, 1 base log-density of z, 2 log-density + log-det
template <int D, int H>
__global__ void __launch_bounds__(32 * kMadeWarps, 1)
made_fast_kernel(const float *__restrict__ net, const float *__restrict__ x, float *__restrict__ z,
                 const float *__restrict__ ld_in, float *__restrict__ ld_out, float *__restrict__ lp_out,
                 long long n_rows, int parity, int lp_mode) {
    using L = MadeLayout<D, H>;
    constexpr int S = D + 1;  // padded row stride
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *W = smem, *tile = smem + L::kFloats + warp * (kMadeWarpRows * S);
    made_stage_net<D, H>(net, W);
    __syncthreads();
    const long long n_tiles = (n_rows + kMadeWarpRows - 1) / kMadeWarpRows;
    float *ra = tile + lane * S, *rb = ra + 32 * S;

#pragma unroll 1
    for (long long t = (long long)blockIdx.x * kMadeWarps + warp; t < n_tiles; t += (long long)gridDim.x * kMadeWarps) {
        const long long row0 = t * kMadeWarpRows;
        __syncwarp();  // previous tile fully stored
#pragma unroll 8
        for (int e = lane; e < kMadeWarpRows * (D / 4); e += 32) {
            const int r = e / (D / 4), c = (e % (D / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n_rows) v = ld_stream4(reinterpret_cast<const float4 *>(x + (row0 + r) * D + c));
            float *d = tile + r * S + c;
            d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
        }
        __syncwarp();

        float2 hA[H / 2], hB[H / 2], gA[H / 2], gB[H / 2];
#pragma unroll
        for (int j = 0; j < H / 2; j += 2) {
            const float4 b = lds4(W, L::B1 + 2 * j);
            hA[j] = hB[j] = make_float2(b.x, b.y);
            hA[j + 1] = hB[j + 1] = make_float2(b.z, b.w);
        }
#pragma unroll 4
        for (int i = 0; i < D; ++i) {
            const float xa = ra[i], xb = rb[i];
            const float2 aa = make_float2(xa, xa), bb = make_float2(xb, xb);
#pragma unroll
            for (int j = 0; j < H / 2; j += 2) {
                const float4 w = lds4(W, L::W1 + i * H + 2 * j);
                hA[j] = ffma2(make_float2(w.x, w.y), aa, hA[j]);
                hB[j] = ffma2(make_float2(w.x, w.y), bb, hB[j]);
                hA[j + 1] = ffma2(make_float2(w.z, w.w), aa, hA[j + 1]);
                hB[j + 1] = ffma2(make_float2(w.z, w.w), bb, hB[j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < H / 2; ++j) {
            hA[j] = make_float2(fmaxf(hA[j].x, 0.f), fmaxf(hA[j].y, 0.f));
            hB[j] = make_float2(fmaxf(hB[j].x, 0.f), fmaxf(hB[j].y, 0.f));
        }
        dense_relu<H>(W, L::W2, L::B2, hA, hB, gA, gB);
        dense_relu<H>(W, L::W3, L::B3, gA, gB, hA, hB);

        // output layer in chunks of 8 dims: acc = (s_i, t_i); z_i = x_i exp(s_i) + t_i replaces x_i in the tile
        float ldA = 0.f, ldB = 0.f, ssA = 0.f, ssB = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < D; c0 += 8) {
            float2 oA[8], oB[8];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const float4 b = lds4(W, L::B4 + 2 * (c0 + u));
                oA[u] = oB[u] = make_float2(b.x, b.y);
                oA[u + 1] = oB[u + 1] = make_float2(b.z, b.w);
            }
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float a = (k & 1) ? hA[k >> 1].y : hA[k >> 1].x;
                const float b = (k & 1) ? hB[k >> 1].y : hB[k >> 1].x;
                const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const float4 w = lds4(W, L::W4 + k * 2 * D + 2 * (c0 + u));
                    oA[u] = ffma2(make_float2(w.x, w.y), aa, oA[u]);
                    oB[u] = ffma2(make_float2(w.x, w.y), bb, oB[u]);
                    oA[u + 1] = ffma2(make_float2(w.z, w.w), aa, oA[u + 1]);
                    oB[u + 1] = ffma2(make_float2(w.z, w.w), bb, oB[u + 1]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float za = fmaf(ra[c0 + u], sm_exp<true>(oA[u].x), oA[u].y);
                const float zb = fmaf(rb[c0 + u], sm_exp<true>(oB[u].x), oB[u].y);
                ra[c0 + u] = za, rb[c0 + u] = zb;
                ldA += oA[u].x, ldB += oB[u].x;
                ssA = fmaf(za, za, ssA), ssB = fmaf(zb, zb, ssB);
            }
        }
        const long long rowA = row0 + lane, rowB = rowA + 32;
        if (rowA < n_rows) {
            const float l = (ld_in ? ld_in[rowA] : 0.f) + ldA;
            if (ld_out) ld_out[rowA] = l;
            const float lp = -0.5f * ssA - 0.5f * (float)D * 1.8378770664093453f;
            if (lp_mode) lp_out[rowA] = lp_mode == 2 ? lp + l : lp;
        }
        if (rowB < n_rows) {
            const float l = (ld_in ? ld_in[rowB] : 0.f) + ldB;
            if (ld_out) ld_out[rowB] = l;
            const float lp = -0.5f * ssB - 0.5f * (float)D * 1.8378770664093453f;
            if (lp_mode) lp_out[rowB] = lp_mode == 2 ? lp + l : lp;
        }
        if (z == nullptr) continue;
        __syncwarp();
#pragma unroll 4
        for (int e = lane; e < kMadeWarpRows * (D / 4); e += 32) {
            const int r = e / (D / 4), c = (e % (D / 4)) * 4;
            if (row0 + r >= n_rows) continue;
            const float *d = tile + r * S;
            const float4 v = parity ? make_float4(d[D - 1 - c], d[D - 2 - c], d[D - 3 - c], d[D - 4 - c])
                                    : make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]);
            st_stream4(reinterpret_cast<float4 *>(z + (row0 + r) * D + c), v);
        }
    }
}

__device__ __forceinline__ float2 lds2(const float *W, int i) { return reinterpret_cast<const float2 *>(W)[i >> 1]; }

// Sequential direction (MAF.forward / IAF.inverse, maf.py:39-51): x_i = (z_f(i) - t_i(x_<i)) exp(-s_i(x_<i)), D passes.
// The reference re-evaluates the whole MADE on the partially filled x in every pass (D x the density cost); here a
// pass costs 1/4.7 of that: the first layer's pre-activations are kept in registers and updated with the rank-1
// term W1[:, i] x_i as soon as x_i is known (the not-yet-filled inputs are zero, so this is exactly the reference's
// arithmetic), the two hidden layers are recomputed, and only the two outputs (s_i, t_i) of the last layer are formed.
constexpr int kMadeSeqWarps = 8;  // the sequential kernel keeps three activation sets live: fewer warps, 255 registers

template <int D, int H>
__global__ void __launch_bounds__(32 * kMadeSeqWarps, 1)
made_seq_kernel(const float *__restrict__ net, const float *__restrict__ z, float *__restrict__ x,
                const float *__restrict__ ld_in, float *__restrict__ ld_out, float *__restrict__ lp_out,
                long long n_rows, int parity, int lp_mode) {
    using L = MadeLayout<D, H>;
    constexpr int S = D + 1;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *W = smem, *tile = smem + L::kFloats + warp * (kMadeWarpRows * S);
    made_stage_net<D, H>(net, W);
    __syncthreads();
    const long long n_tiles = (n_rows + kMadeWarpRows - 1) / kMadeWarpRows;
    float *ra = tile + lane * S, *rb = ra + 32 * S;

#pragma unroll 1
    for (long long t = (long long)blockIdx.x * kMadeSeqWarps + warp; t < n_tiles; t += (long long)gridDim.x * kMadeSeqWarps) {
        const long long row0 = t * kMadeWarpRows;
        __syncwarp();
#pragma unroll 8
        for (int e = lane; e < kMadeWarpRows * (D / 4); e += 32) {  // the flip (maf.py:41) happens while staging
            const int r = e / (D / 4), c = (e % (D / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n_rows) v = ld_stream4(reinterpret_cast<const float4 *>(z + (row0 + r) * D + c));
            float *d = tile + r * S;
            if (parity) d[D - 1 - c] = v.x, d[D - 2 - c] = v.y, d[D - 3 - c] = v.z, d[D - 4 - c] = v.w;
            else d[c] = v.x, d[c + 1] = v.y, d[c + 2] = v.z, d[c + 3] = v.w;
        }
        __syncwarp();

        float2 pA[H / 2], pB[H / 2];  // first-layer pre-activations of the partially filled x
#pragma unroll
        for (int j = 0; j < H / 2; j += 2) {
            const float4 b = lds4(W, L::B1 + 2 * j);
            pA[j] = pB[j] = make_float2(b.x, b.y);
            pA[j + 1] = pB[j + 1] = make_float2(b.z, b.w);
        }
        float ldA = 0.f, ldB = 0.f, ssA = 0.f, ssB = 0.f;
#pragma unroll 1
        for (int i = 0; i < D; ++i) {
            float2 hA[H / 2], hB[H / 2], gA[H / 2], gB[H / 2];
#pragma unroll
            for (int j = 0; j < H / 2; ++j) {
                gA[j] = make_float2(fmaxf(pA[j].x, 0.f), fmaxf(pA[j].y, 0.f));
                gB[j] = make_float2(fmaxf(pB[j].x, 0.f), fmaxf(pB[j].y, 0.f));
            }
            dense_relu<H>(W, L::W2, L::B2, gA, gB, hA, hB);
            dense_relu<H>(W, L::W3, L::B3, hA, hB, gA, gB);
            // (s_i, t_i): four partial sums per row keep the dependent FFMA2 chains short
            float2 oA[4], oB[4];
            oA[0] = oB[0] = lds2(W, L::B4 + 2 * i);
#pragma unroll
            for (int q = 1; q < 4; ++q) oA[q] = oB[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float a = (k & 1) ? gA[k >> 1].y : gA[k >> 1].x;
                const float b = (k & 1) ? gB[k >> 1].y : gB[k >> 1].x;
                const float2 w = lds2(W, L::W4 + k * 2 * D + 2 * i);
                oA[k & 3] = ffma2(w, make_float2(a, a), oA[k & 3]);
                oB[k & 3] = ffma2(w, make_float2(b, b), oB[k & 3]);
            }
            const float sA = (oA[0].x + oA[1].x) + (oA[2].x + oA[3].x), tA = (oA[0].y + oA[1].y) + (oA[2].y + oA[3].y);
            const float sB = (oB[0].x + oB[1].x) + (oB[2].x + oB[3].x), tB = (oB[0].y + oB[1].y) + (oB[2].y + oB[3].y);
            const float xa = (ra[i] - tA) * sm_exp<true>(-sA), xb = (rb[i] - tB) * sm_exp<true>(-sB);
            ra[i] = xa, rb[i] = xb;
            ldA -= sA, ldB -= sB;
            ssA = fmaf(xa, xa, ssA), ssB = fmaf(xb, xb, ssB);
            const float2 aa = make_float2(xa, xa), bb = make_float2(xb, xb);
#pragma unroll
            for (int j = 0; j < H / 2; j += 2) {  // rank-1 update: x_i enters the first layer
                const float4 w = lds4(W, L::W1 + i * H + 2 * j);
                pA[j] = ffma2(make_float2(w.x, w.y), aa, pA[j]);
                pB[j] = ffma2(make_float2(w.x, w.y), bb, pB[j]);
                pA[j + 1] = ffma2(make_float2(w.z, w.w), aa, pA[j + 1]);
                pB[j + 1] = ffma2(make_float2(w.z, w.w), bb, pB[j + 1]);
            }
        }
        const long long rowA = row0 + lane, rowB = rowA + 32;
        if (rowA < n_rows) {
            const float l = (ld_in ? ld_in[rowA] : 0.f) + ldA;
            if (ld_out) ld_out[rowA] = l;
            const float lp = -0.5f * ssA - 0.5f * (float)D * 1.8378770664093453f;
            if (lp_mode) lp_out[rowA] = lp_mode == 2 ? lp + l : lp;
        }
        if (rowB < n_rows) {
            const float l = (ld_in ? ld_in[rowB] : 0.f) + ldB;
            if (ld_out) ld_out[rowB] = l;
            const float lp = -0.5f * ssB - 0.5f * (float)D * 1.8378770664093453f;
            if (lp_mode) lp_out[rowB] = lp_mode == 2 ? lp + l : lp;
        }
        if (x == nullptr) continue;
        __syncwarp();
#pragma unroll 4
        for (int e = lane; e < kMadeWarpRows * (D / 4); e += 32) {
            const int r = e / (D / 4), c = (e % (D / 4)) * 4;
            if (row0 + r >= n_rows) continue;
            const float *d = tile + r * S;
            st_stream4(reinterpret_cast<float4 *>(x + (row0 + r) * D + c), make_float4(d[c], d[c + 1], d[c + 2], d[c + 3]));
        }
    }
}

template <int D, int H>
static bool made_shape_is(const mnf_flow_op &op) {
    return op.n_lin == 4 && op.sizes[0] == D && op.sizes[1] == H && op.sizes[2] == H && op.sizes[3] == H &&
           op.sizes[4] == 2 * D;
}

// One (dim, hidden) instantiation: returns 0 when launched, 1 when the program is not an all-MADE stack of THIS shape,
// other values on error.
template <int D, int H>
static int launch_made_fast_t(const mnf_flow_op *ops, int n_ops, const float *params, const float *x, float *y, float *log_det,
                              float *base_lp, float *inter, int64_t n_rows, int dim, int dir_flags, float *scratch,
                              cudaStream_t stream, bool plan_only) {
    using L = MadeLayout<D, H>;
    if (dim != D || n_ops < 1) return 1;
    const int inverse = dir_flags & 1;
    for (int k = 0; k < n_ops; ++k) {
        const mnf_flow_op &op = ops[k];
        if (op.type != MNF_OP_MADE || !made_shape_is<D, H>(op)) return 1;
    }
    if (!y && !inter && n_ops > 1 && !scratch) return 1;
    if (plan_only) return 0;
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr, MNF_E_DEVICE, "no CUDA device");

    const size_t smem = sizeof(float) * (L::kFloats + kMadeWarps * kMadeWarpRows * (D + 1));
    int dev = 0;
    MNF_CUDA(cudaGetDevice(&dev));
    static bool attr_set[64] = {};  // per instantiation
    if (!attr_set[dev & 63]) {
        MNF_CUDA(cudaFuncSetAttribute(made_fast_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MNF_CUDA(cudaFuncSetAttribute(made_seq_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 63] = true;
    }
    const long long n_tiles = (n_rows + kMadeWarpRows - 1) / kMadeWarpRows;
    const long long want = (n_tiles + kMadeWarps - 1) / kMadeWarps;
    const unsigned blocks = (unsigned)(want < dp->sm_count ? want : dp->sm_count);  // one persistent block per SM
    const long long seq_want = (n_tiles + kMadeSeqWarps - 1) / kMadeSeqWarps;
    const unsigned seq_blocks = (unsigned)(seq_want < dp->sm_count ? seq_want : dp->sm_count);
    const float *src = x;
    int rc = 0;
    for (int k = 0; k < n_ops; ++k) {
        const mnf_flow_op &op = ops[inverse ? n_ops - 1 - k : k];
        const bool last = k == n_ops - 1;
        float *dst = inter ? inter + (size_t)k * n_rows * D : (y ? y : (last ? nullptr : scratch));
        // the running log-det lives in log_det when requested, else (log-prob mode) in base_lp until the last flow
        float *ld_dst = log_det ? log_det : (last ? nullptr : (base_lp && (dir_flags & 2) ? base_lp : nullptr));
        const float *ld_src = k == 0 ? nullptr : (log_det ? log_det : ((dir_flags & 2) ? base_lp : nullptr));
        const int lp_mode = (last && base_lp) ? ((dir_flags & 2) ? 2 : 1) : 0;
        const bool sequential = (op.flags & MNF_FLAG_MADE_SEQ) ? !inverse : inverse;
        const int parity = (op.flags & MNF_FLAG_PARITY) ? 1 : 0;
        if (sequential)
            made_seq_kernel<D, H><<<seq_blocks, 32 * kMadeSeqWarps, smem, stream>>>(params + op.net_off[0], src, dst, ld_src,
                                                                                ld_dst, base_lp, n_rows, parity, lp_mode);
        else
            made_fast_kernel<D, H><<<blocks, 32 * kMadeWarps, smem, stream>>>(params + op.net_off[0], src, dst, ld_src, ld_dst,
                                                                          base_lp, n_rows, parity, lp_mode);
        rc = launch_status(sequential ? "made_seq_kernel" : "made_fast_kernel");
        if (rc) break;
        if (last && inter && y)
            MNF_CUDA(cudaMemcpyAsync(y, dst, sizeof(float) * n_rows * D, cudaMemcpyDeviceToDevice, stream));
        src = dst;
    }
    return rc;
}

// Returns 0 when launched, 1 when the program is not an all-MADE stack of one of the instantiated shapes (three hidden
// layers of equal width; the caller falls through to the interpreter), other values on error.  dir_flags: bit0 inverse,
// bit1 sum log-det into base_lp.  scratch: [n_rows * dim] floats, needed only when y == NULL and n_ops > 1.
// Shapes: BASELINE config 3 (64, 24) plus a grid of common ones -- (dim, hidden) in {8, 16, 32, 64} x {16, 32} and (64, 24);
// anything else (other depths, hidden > 32, which would not leave two rows' activations in registers) is interpreted.
int launch_made_fast(const mnf_flow_op *ops, int n_ops, const float *params, const float *x, float *y, float *log_det,
                     float *base_lp, float *inter, int64_t n_rows, int dim, int dir_flags, float *scratch,
                     cudaStream_t stream, bool plan_only) {
    if (n_ops < 1 || ops[0].type != MNF_OP_MADE || ops[0].n_lin != 4) return 1;
    const int h = ops[0].sizes[1];
#define MNF_MADE_TRY(DD, HH)                                                                                              \
    if (dim == DD && h == HH)                                                                                             \
        return launch_made_fast_t<DD, HH>(ops, n_ops, params, x, y, log_det, base_lp, inter, n_rows, dim, dir_flags, scratch, \
                                          stream, plan_only);
    MNF_MADE_TRY(64, 24)
    MNF_MADE_TRY(64, 32)
    MNF_MADE_TRY(64, 16)
    MNF_MADE_TRY(32, 32)
    MNF_MADE_TRY(32, 16)
    MNF_MADE_TRY(16, 32)
    MNF_MADE_TRY(16, 16)
    MNF_MADE_TRY(8, 32)
    MNF_MADE_TRY(8, 16)
#undef MNF_MADE_TRY
    return 1;
}

}  // namespace mnf
