// flow_tc.cu -- dim-2 spline stacks (BASELINE config 2: [ActNormFlow, Glow, NSF_CL(K=8, n_h=16)] x n) with the
// conditioner MLPs (spline_flow.py:252-257 over mlp.py:4-12) on the tcgen05 tensor cores at fp32-class accuracy.
//
// Whole stack in ONE launch, nothing but the algorithmic bytes touches HBM (8 B in, 4..20 B out per point), and the
// library keeps no device state: the conditioner weights travel as a per-call image in the caller's workspace.
//
// The three 16-wide layers of a conditioner (16x16, 16x16, 16x23) are 3xTF32 products: every operand is split into a
// TF32 "hi" part and the exact fp32 remainder "lo", and hi*hi + lo*hi + hi*lo is accumulated in fp32 by the tensor core
// (the dropped lo*lo term is 2^-22 of the product).  One 128-point tile = the 128 TMEM lanes:
//   A operand  [128 x 32] = (hi[16] | lo[16]) of the layer input, written by the epilogue threads straight into TMEM
//              with tcgen05.st (TS-mode MMA: the operand never touches shared memory);
//   B operands pre-swizzled K-major weight tiles, resident in shared memory for the whole kernel:
//              B1 = (W_hi | W_hi) [N x 32], B2 = (W_lo | b_hi, b_lo, 0...) [N x 32]; the bias rides on a constant
//              A fragment of ones in TMEM, so the epilogue never adds it;
//   D          fp32 accumulator in TMEM, read back with tcgen05.ld by the thread that owns the point.
// Per layer: 7 MMAs of M128 x N16 (N32 for the output layer) x K8.
//
// One CTA per SM: 4 control warps (lane 0 of warp w issues the MMAs of tile slots w and w + 4) and T epilogue groups of
// 128 threads.  A thread owns one point for the whole stack: ActNorm / Glow in registers, layer 0 of each conditioner
// (1 -> 16) on the FMA pipe, LeakyReLU + hi/lo split between the layers, then the rational-quadratic spline
// (flow_math.cuh) on the 23 raw outputs.  T tiles in flight per SM hide the MMA round trips.
#include <stdlib.h>

#include "flow_math.cuh"
#include "tc_common.cuh"

namespace mnf {
namespace ftc {
using namespace tc;

constexpr int H = 16, KBINS = 8, NB = 3 * KBINS - 1, NOUT = 32;
constexpr uint32_t WL_BYTES = 2 * H * 128;               // hidden layer: B1 | B2, 16 rows x 128 B each
constexpr uint32_t WO_BYTES = 2 * NOUT * 128;            // output layer: B1 | B2, 32 rows x 128 B each
constexpr uint32_t NET_BYTES = 2 * WL_BYTES + WO_BYTES;  // 16 KB per conditioner
constexpr uint32_t L0_BYTES = 2 * H * 4;                 // layer 0: w0[16], b0[16]
constexpr int MAX_NETS = 12, MAX_T = 7;
constexpr uint32_t SLOT_COLS = 64, ONES_COL = 448;  // TMEM: slot s owns columns [64 s, +32) = A, [64 s + 32, +32) = D
constexpr uint32_t IDESC_H = tf32_instr_desc(BM, H), IDESC_O = tf32_instr_desc(BM, NOUT);

struct Params {
    FlowProgram prog;  // module order
    const float *params, *x;
    const uint8_t *image;  // [n_nets][NET_BYTES] weight tiles, then [n_nets][L0_BYTES]
    float *y, *log_det, *base_lp, *inter;
    long long n_rows;
    int dir_flags, n_nets, debug;
    mnf_gather_out gather;
};

__host__ __device__ constexpr uint32_t image_bytes(int n_nets) { return (uint32_t)n_nets * (NET_BYTES + L0_BYTES); }

// ---------------------------------------------------------------------------------------------------------
// image builder: one CTA per conditioner (execution order); writes the swizzled B tiles and the layer-0 block
// ---------------------------------------------------------------------------------------------------------
struct NetList {
    int n;
    int off[MAX_NETS];  // float offset of each conditioner in the parameter blob
};

__device__ __forceinline__ float tf32_hi(float v) { return rn_tf32(v); }

__global__ void flow_tc_image_kernel(const __grid_constant__ NetList nets, const float *__restrict__ params,
                                     uint8_t *__restrict__ image) {
    const int net = blockIdx.x;
    const float *src = params + nets.off[net];
    float *tiles = reinterpret_cast<float *>(image + (size_t)net * NET_BYTES);
    // blob layout of the MLP 1 -> 16 -> 16 -> 16 -> 23: per Linear weight[out][in], bias[out]
    constexpr int OFF1 = 2 * H, OFF2 = OFF1 + H * H + H, OFF3 = OFF2 + H * H + H;
    for (int e = threadIdx.x; e < (int)(NET_BYTES / 4); e += blockDim.x) {
        int layer, r = e;
        if (r < (int)(WL_BYTES / 4)) layer = 0;
        else if (r < (int)(2 * WL_BYTES / 4)) layer = 1, r -= WL_BYTES / 4;
        else layer = 2, r -= 2 * WL_BYTES / 4;
        const int n_rows_tile = layer == 2 ? NOUT : H, n_out = layer == 2 ? NB : H;
        const int which = r / (n_rows_tile * 32);  // 0 = B1, 1 = B2
        r -= which * n_rows_tile * 32;
        const int n = r >> 5, pos = (r >> 2) & 7, j = r & 3;
        const int k = ((pos ^ (n & 7)) << 2) + j;  // logical K index of this float (128-byte swizzle: chunk ^= row & 7)
        const float *W = src + (layer == 0 ? OFF1 : layer == 1 ? OFF2 : OFF3);
        const float *bias = W + n_out * H;
        float v = 0.f;
        if (n < n_out) {
            if (which == 0) {
                v = tf32_hi(W[n * H + (k & 15)]);
            } else if (k < 16) {
                const float w = W[n * H + k];
                v = tf32_hi(w - tf32_hi(w));
            } else if (k == 16) {
                v = tf32_hi(bias[n]);
            } else if (k == 17) {
                v = tf32_hi(bias[n] - tf32_hi(bias[n]));
            }
        }
        tiles[e] = v;
    }
    float *l0 = reinterpret_cast<float *>(image + (size_t)nets.n * NET_BYTES + (size_t)net * L0_BYTES);
    for (int e = threadIdx.x; e < 2 * H; e += blockDim.x) l0[e] = src[e];
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// D[tmem] (+)= A[tmem] * B[smem]  (TS mode: A is a [128 lanes x 8 columns] TF32 fragment)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// LeakyReLU(0.2) (mlp.py:9) of 16 pre-activations, then the 3xTF32 split: hi = value rounded to TF32 (nearest, ties
// away: integer add of half a TF32 ulp, then mask), lo = exact fp32 remainder (the tensor core reads its top 11 bits)
template <bool TRUNC>
__device__ __forceinline__ void act_split16(const float (&pre)[16], uint32_t (&r)[32]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float a = fmaxf(pre[j], 0.2f * pre[j]);
        const uint32_t b = __float_as_uint(a);
        const uint32_t hi = TRUNC ? (b & 0xffffe000u) : ((b + 0x1000u) & 0xffffe000u);
        r[j] = hi;
        r[16 + j] = __float_as_uint(a - __uint_as_float(hi));
    }
}

template <int T>
__global__ void __launch_bounds__(128 + 128 * T, 1) flow_tc_kernel(const __grid_constant__ Params p) {
    constexpr int THREADS = 128 + 128 * T;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int n_nets = p.n_nets;
    const uint32_t off_l0 = (uint32_t)n_nets * NET_BYTES, off_bar = off_l0 + (uint32_t)n_nets * L0_BYTES;
    const uint32_t bars = base + off_bar;
    const uint32_t w_bar = bars;
    auto a_ready = [&](int s) { return bars + 8u * (1 + s); };
    auto acc_ready = [&](int s) { return bars + 8u * (1 + T + s); };
    const uint32_t tmem_slot = bars + 8u * (1 + 2 * T);
    const uint32_t *tmem_slot_ptr = reinterpret_cast<const uint32_t *>(base_ptr + off_bar + 8u * (1 + 2 * T));
    const float *sl0 = reinterpret_cast<const float *>(base_ptr + off_l0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inverse = p.dir_flags & 1;
    const bool sum_lp = p.dir_flags & 2;
    const long long n_tiles = (p.n_rows + BM - 1) / BM;

    if (warp == 1 && lane == 0) {
        mbar_init(w_bar, 1);
        for (int s = 0; s < T; ++s) {
            mbar_init(a_ready(s), 4);  // one arrival per epilogue warp of the group
            mbar_init(acc_ready(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0 && lane == 0) {  // the weight image: one bulk copy per conditioner + the layer-0 block
        mbar_expect_tx(w_bar, image_bytes(n_nets));
        for (int n = 0; n < n_nets; ++n) bulk_load(base + (uint32_t)n * NET_BYTES, p.image + (size_t)n * NET_BYTES, NET_BYTES, w_bar);
        bulk_load(base + off_l0, p.image + (size_t)n_nets * NET_BYTES, (uint32_t)n_nets * L0_BYTES, w_bar);
    }

    if (warp < 4) {
        if (lane == 0) {
            // ---------------- MMA issuer of tile slots `warp` and `warp + 4` ----------------
            // slot s walks tiles blockIdx.x + (s + j T) gridDim.x; every tile takes 3 MMA steps per conditioner
            auto steps_of = [&](int s) -> uint32_t {
                if (s >= T) return 0u;
                const long long first = (long long)blockIdx.x + (long long)s * gridDim.x;
                if (first >= n_tiles) return 0u;
                const long long stride = (long long)T * gridDim.x;
                return (uint32_t)((n_tiles - first + stride - 1) / stride) * (uint32_t)(3 * n_nets);
            };
            const int s0 = warp, s1 = warp + 4;
            const uint32_t n0 = steps_of(s0), n1 = steps_of(s1);
            uint32_t i0 = 0, i1 = 0;
            auto issue = [&](int s, uint32_t i) {
                const uint32_t net = (i / 3u) % (uint32_t)n_nets, layer = i % 3u;
                const uint32_t wb = base + net * NET_BYTES + layer * WL_BYTES;
                const uint32_t n_tile_rows = layer == 2 ? NOUT : H;
                const uint64_t b1 = make_smem_desc(wb), b2 = make_smem_desc(wb + n_tile_rows * 128u);
                const uint32_t idesc = layer == 2 ? IDESC_O : IDESC_H;
                const uint32_t a = tmem_base + SLOT_COLS * (uint32_t)s, d = a + 32u, ones = tmem_base + ONES_COL;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (!(p.debug & 2)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, a + 8u * k, b1 + (uint64_t)(2 * k), k != 0, idesc);  // (hi | lo) x (W_hi | W_hi)
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_tf32_ts(d, a + 8u * k, b2 + (uint64_t)(2 * k), 1u, idesc);  // hi x W_lo
                    umma_tf32_ts(d, ones, b2 + 4u, 1u, idesc);  // (1, 1, 0...) x (b_hi, b_lo, 0...)
                }
                umma_commit(acc_ready(s));
            };
            mbar_wait(w_bar, 0);
            if (n1 == 0) {
                for (; i0 < n0; ++i0) {
                    mbar_wait_parked(a_ready(s0), i0 & 1u);
                    issue(s0, i0);
                }
            } else {
                while (i0 < n0 || i1 < n1) {
                    if (i0 < n0 && mbar_test(a_ready(s0), i0 & 1u)) issue(s0, i0), ++i0;
                    if (i1 < n1 && mbar_test(a_ready(s1), i1 & 1u)) issue(s1, i1), ++i1;
                }
            }
        }
    } else {
        // ---------------- epilogue groups: thread = one point of the slot's current tile ----------------
        const int slot = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a_addr = lane_base + SLOT_COLS * (uint32_t)slot, d_addr = a_addr + 32u;
        {  // the constant A fragment of the bias MMA: identical values from every group (the first MMA of a slot follows
           // that slot's own stores)
            uint32_t one[8] = {0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(lane_base + ONES_COL, one);
        }
        mbar_wait_parked(w_bar, 0);
        uint32_t acc_phase = 0;
        auto publish = [&](const uint32_t(&r)[32]) {
            tmem_st32(a_addr, r);
            tmem_st_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(slot));
        };
        auto acquire = [&]() {
            mbar_wait_parked(acc_ready(slot), acc_phase);
            acc_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        };
        // conditioner `net` on the scalar c -> raw[23] (spline_flow.py:252-253 / :258-259)
        auto conditioner = [&](int net, float c, float(&raw)[NB]) {
            uint32_t r[32];
            {
                const float4 *wb = reinterpret_cast<const float4 *>(sl0 + net * 2 * H);
                float pre[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = wb[j], b = wb[4 + j];
                    pre[4 * j] = fmaf(w.x, c, b.x), pre[4 * j + 1] = fmaf(w.y, c, b.y);
                    pre[4 * j + 2] = fmaf(w.z, c, b.z), pre[4 * j + 3] = fmaf(w.w, c, b.w);
                }
                if (p.debug & 4) act_split16<true>(pre, r);
                else act_split16<false>(pre, r);
            }
            publish(r);
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
                acquire();
                uint32_t t[16];
                tmem_ld16(d_addr, t);
                tmem_ld_wait();
                float pre[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pre[j] = __uint_as_float(t[j]);
                if (p.debug & 4) act_split16<true>(pre, r);
                else act_split16<false>(pre, r);
                publish(r);
            }
            acquire();
            uint32_t t0[16], t1[8];
            tmem_ld16(d_addr, t0);
            tmem_ld8(d_addr + 16u, t1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) raw[j] = __uint_as_float(t0[j]);
#pragma unroll
            for (int j = 0; j < NB - 16; ++j) raw[16 + j] = __uint_as_float(t1[j]);
        };

#pragma unroll 1
        for (long long tile = (long long)blockIdx.x + (long long)slot * gridDim.x; tile < n_tiles; tile += (long long)T * gridDim.x) {
            const long long r_glob = tile * BM + row;
            const bool live = r_glob < p.n_rows;
            const long long r_ld = live ? r_glob : p.n_rows - 1;
            float v0, v1, ld = 0.f;
            {
                const float2 xin = ld_stream2(reinterpret_cast<const float2 *>(p.x) + r_ld);
                v0 = xin.x, v1 = xin.y;
            }
            int net = 0;
#pragma unroll 1
            for (int kk = 0; kk < p.prog.n_ops; ++kk) {
                const mnf_flow_op &op = p.prog.ops[inverse ? p.prog.n_ops - 1 - kk : kk];
                if (op.type == MNF_OP_AFFINE_CONST) {
                    const float4 st = *reinterpret_cast<const float4 *>(p.params + op.aux_off);  // s0 s1 t0 t1
                    if (inverse) {  // affine_constant_flow.py:24
                        v0 = (v0 - st.z) * expf(-st.x);
                        v1 = (v1 - st.w) * expf(-st.y);
                        ld -= st.x + st.y;
                    } else {  // affine_constant_flow.py:19
                        v0 = v0 * expf(st.x) + st.z;
                        v1 = v1 * expf(st.y) + st.w;
                        ld += st.x + st.y;
                    }
                } else if (op.type == MNF_OP_GLOW) {
                    const float4 W = *reinterpret_cast<const float4 *>(p.params + op.aux_off + (inverse ? 4 : 0));  // glow.py:28,36: v @ W
                    const float lg = p.params[op.aux_off + 8];
                    const float n0 = fmaf(v1, W.z, v0 * W.x), n1 = fmaf(v1, W.w, v0 * W.y);
                    v0 = n0, v1 = n1;
                    ld += inverse ? -lg : lg;
                } else {  // NSF_CL: f1 on (lower -> upper) then f2 on (upper -> lower) going forward (spline_flow.py:249-266);
                          // f2 first, then f1, both with the spline inverse, going backward (:268-285)
#pragma unroll 1
                    for (int step = 0; step < 2; ++step) {
                        const bool use_f1 = (step == 0) != (inverse != 0);
                        float raw[NB];
                        conditioner(net++, use_f1 ? v0 : v1, raw);
                        float tr = use_f1 ? v1 : v0, l = 0.f;
                        if (!(p.debug & 1)) rq_spline<KBINS, true>(raw, KBINS, op.bound, op.edge_deriv, inverse != 0, tr, l);
                        else tr += raw[0] + raw[22], l = raw[11];
                        ld += l;
                        if (use_f1) v1 = tr; else v0 = tr;
                    }
                }
                if (p.inter && live)
                    st_stream2(reinterpret_cast<float2 *>(p.inter + ((size_t)kk * p.n_rows + r_glob) * 2), make_float2(v0, v1));
            }
            if (!live) continue;
            float lp = fmaf(-0.5f, fmaf(v0, v0, v1 * v1), -1.8378770664093453f);  // -(D/2) log(2 pi), D = 2
            if (sum_lp) lp += ld;
            if (p.y) st_stream2(reinterpret_cast<float2 *>(p.y) + r_glob, make_float2(v0, v1));
            if (p.log_det) p.log_det[r_glob] = ld;
            if (p.base_lp) p.base_lp[r_glob] = lp;
            // fused gather: the result also goes straight to the other ranks over NVLink
            if (p.gather.multicast_ptr) {
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.gather.multicast_ptr + p.gather.row_offset + r_glob), "f"(lp)
                             : "memory");
            } else {
                for (int g = 0; g < p.gather.n_peers; ++g) p.gather.peer_ptrs[g][p.gather.row_offset + r_glob] = lp;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int T>
static int launch_t(const Params &p, unsigned grid, size_t smem, cudaStream_t st) {
    MNF_CUDA(cudaFuncSetAttribute(flow_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flow_tc_kernel<T><<<grid, 128 + 128 * T, smem, st>>>(p);
    return launch_status("flow_tc_kernel");
}

}  // namespace ftc

// floats of caller workspace the tensor-core kernel needs for its weight image (any eligible program)
int64_t flow_tc_workspace_floats() { return (int64_t)ftc::image_bytes(ftc::MAX_NETS) / 4; }

// returns 1 if the program is not eligible (caller falls back to the other dim-2 kernels)
int launch_flow_tc(const mnf_flow_op *ops, int n_ops, const float *params, const float *x, float *y, float *log_det,
                   float *base_lp, float *inter, int64_t n_rows, int dim, int dir_flags, float *workspace,
                   const mnf_gather_out *gather, cudaStream_t stream, bool plan_only) {
    using namespace ftc;
    if (dim != 2 || n_ops < 1 || n_ops > MNF_MAX_OPS) return 1;
    const int inverse = dir_flags & 1;
    NetList nets;
    nets.n = 0;
    for (int kk = 0; kk < n_ops; ++kk) {
        const mnf_flow_op &op = ops[inverse ? n_ops - 1 - kk : kk];
        if (op.type == MNF_OP_AFFINE_CONST || op.type == MNF_OP_GLOW) continue;
        if (op.type != MNF_OP_NSF_CL || op.K != KBINS || op.n_lin != 4 || op.sizes[0] != 1 || op.sizes[1] != H ||
            op.sizes[2] != H || op.sizes[3] != H || op.sizes[4] != NB)
            return 1;
        if (nets.n + 2 > MAX_NETS) return 1;
        // forward: f1 then f2 (spline_flow.py:249-266); inverse: f2 then f1 (:268-285)
        nets.off[nets.n++] = op.net_off[inverse ? 1 : 0];
        nets.off[nets.n++] = op.net_off[inverse ? 0 : 1];
    }
    if (nets.n == 0) return 1;
    if (plan_only) return 0;
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr && dp->cc_major == 10, MNF_E_DEVICE, "tcgen05 path needs an sm_100 device");
    MNF_REQUIRE(workspace != nullptr && ((uintptr_t)workspace % 16) == 0, MNF_E_ARG,
                "the tensor-core flow kernel needs the 16-byte aligned workspace of mnf_flow_stack_workspace()");
    MNF_REQUIRE(((uintptr_t)x % 8) == 0 && (!y || ((uintptr_t)y % 8) == 0) && (!inter || ((uintptr_t)inter % 8) == 0), MNF_E_ALIGN,
                "x, y and intermediates must be 8-byte aligned");
    MNF_REQUIRE(n_rows > 0 && n_rows <= (int64_t)0x7fffffff * 64, MNF_E_ARG, "bad row count");
    uint8_t *image = reinterpret_cast<uint8_t *>(workspace);
    flow_tc_image_kernel<<<nets.n, 256, 0, stream>>>(nets, params, image);
    int rc = launch_status("flow_tc_image_kernel");
    if (rc) return rc;
    Params p{};
    p.prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) p.prog.ops[k] = ops[k];
    p.params = params, p.x = x, p.image = image, p.y = y, p.log_det = log_det, p.base_lp = base_lp, p.inter = inter;
    p.n_rows = n_rows, p.dir_flags = dir_flags & 3, p.n_nets = nets.n;
    const char *dbg = getenv("MNF_FTC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    if (gather) {
        MNF_REQUIRE(gather->n_peers >= 0 && gather->n_peers <= MNF_MAX_PEERS, MNF_E_ARG, "bad n_peers");
        p.gather = *gather;
    }
    const size_t smem = image_bytes(nets.n) + 8 * (2 + 2 * MAX_T) + 1024;
    MNF_REQUIRE(smem <= (size_t)dp->smem_optin, MNF_E_SHAPE, "weight image does not fit in shared memory");
    const long long n_tiles = (n_rows + tc::BM - 1) / tc::BM;
    const char *tenv = getenv("MNF_FTC_T");
    const int T = tenv ? atoi(tenv) : 6;
    const long long per_cta = (n_tiles + T - 1) / T;
    const unsigned grid = (unsigned)(per_cta < dp->sm_count ? (per_cta < 1 ? 1 : per_cta) : dp->sm_count);
    switch (T) {
        case 4: return launch_t<4>(p, grid, smem, stream);
        case 5: return launch_t<5>(p, grid, smem, stream);
        case 6: return launch_t<6>(p, grid, smem, stream);
        case 7: return launch_t<7>(p, grid, smem, stream);
        default: return fail(MNF_E_ARG, "MNF_FTC_T must be 4..7");
    }
}

}  // namespace mnf
