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
// One CTA per SM of T groups of 128 threads, one 128-point tile in flight per group.  A thread owns one point for the whole
// stack: ActNorm / Glow in registers, layer 0 of each conditioner (1 -> 16) on the FMA pipe, LeakyReLU + hi/lo split
// between the layers, then the rational-quadratic spline (flow_math.cuh) on the 23 raw outputs.  After a group has
// stored its A operand (named barrier over its 4 warps) its first thread issues the layer's MMAs itself and commits them
// to the group's mbarrier: no separate issuer warp sits on the critical path.  T tiles in flight hide the round trips.
#include <cuda_fp16.h>
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
constexpr int VEC_FLOATS = 4 * H + NOUT;                 // per conditioner: w0[16] b0[16] | b1[16] | b2[16] | b3[23 (+9 pad)]
constexpr uint32_t VEC_BYTES = VEC_FLOATS * 4;
constexpr int MAX_NETS = 12, MAX_T = 7;
constexpr uint32_t SLOT_COLS = 64, ONES_COL = 448;  // TMEM: slot s owns columns [64 s, +32) = A, [64 s + 32, +32) = D
constexpr uint32_t IDESC_H = tf32_instr_desc(BM, H), IDESC_O = tf32_instr_desc(BM, NOUT);

struct Params {
    FlowProgram prog;  // module order
    const float *params, *x;
    const uint8_t *image;  // [n_nets][NET_BYTES] weight tiles, then [n_nets][VEC_BYTES]
    float *y, *log_det, *base_lp, *inter;
    long long n_rows;
    int dir_flags, n_nets, debug, sleep_ns;
    mnf_gather_out gather;
};

__host__ __device__ constexpr uint32_t image_bytes(int n_nets) { return (uint32_t)n_nets * (NET_BYTES + VEC_BYTES); }

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
    float *vec = reinterpret_cast<float *>(image + (size_t)nets.n * NET_BYTES + (size_t)net * VEC_BYTES);
    for (int e = threadIdx.x; e < VEC_FLOATS; e += blockDim.x) {
        float v = 0.f;
        if (e < 2 * H) v = src[e];                                     // layer 0: weight[16][1], bias[16]
        else if (e < 3 * H) v = src[OFF1 + H * H + (e - 2 * H)];       // b1
        else if (e < 4 * H) v = src[OFF2 + H * H + (e - 3 * H)];       // b2
        else if (e < 4 * H + NB) v = src[OFF3 + NB * H + (e - 4 * H)];  // b3
        vec[e] = v;
    }
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
// wait with an explicit back-off: a warp that spins on try_wait competes for issue slots with the warps doing the work
// (ncu: 40% of the issued instructions of the first version of this kernel were such spins)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns = 40) {
    while (!mbar_test(bar, parity)) __nanosleep(ns);
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

// G groups of 128 threads; a thread carries P points (one per tile slot g * P + i) through the stack in lock step, so the
// MMA round trip of one slot is covered by the epilogue work of the group's other slot(s) as well as by the other groups.
// BIAS_MMA: biases ride on the ones fragment (needs 8 spare TMEM columns: G * P <= 7); otherwise the epilogue adds them.
template <int G, int P, bool BIAS_MMA>
__global__ void __launch_bounds__(128 * G, 1) flow_tc_kernel(const __grid_constant__ Params p) {
    constexpr int S = G * P;
    static_assert(S * (int)SLOT_COLS + (BIAS_MMA ? 8 : 0) <= 512 && S <= 15, "TMEM columns / named barriers");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int n_nets = p.n_nets;
    const uint32_t off_vec = (uint32_t)n_nets * NET_BYTES, off_desc = off_vec + (uint32_t)n_nets * VEC_BYTES,
                   off_bar = off_desc + (uint32_t)n_nets * 3u * 16u;
    const uint32_t bars = base + off_bar;
    const uint32_t w_bar = bars;
    auto acc_ready = [&](int s) { return bars + 8u * (1 + s); };
    const uint32_t tmem_slot = bars + 8u * (1 + S);
    const uint32_t *tmem_slot_ptr = reinterpret_cast<const uint32_t *>(base_ptr + off_bar + 8u * (1 + S));
    const float *svec = reinterpret_cast<const float *>(base_ptr + off_vec);
    // per (conditioner, layer): the two B-tile descriptors (B1, B2) of its MMAs
    ulonglong2 *sdesc = reinterpret_cast<ulonglong2 *>(base_ptr + off_desc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inverse = p.dir_flags & 1;
    const bool sum_lp = p.dir_flags & 2;
    const long long n_tiles = (p.n_rows + BM - 1) / BM;

    if (warp == 1 && lane == 0) {
        mbar_init(w_bar, 1);
        for (int s = 0; s < S; ++s) mbar_init(acc_ready(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 3 * n_nets; i += 128 * G) {
        const uint32_t net = (uint32_t)i / 3u, layer = (uint32_t)i % 3u;
        const uint32_t wb = base + net * NET_BYTES + layer * WL_BYTES;
        sdesc[i] = make_ulonglong2(make_smem_desc(wb), make_smem_desc(wb + (layer == 2 ? NOUT : H) * 128u));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0 && lane == 0) {  // the weight image: one bulk copy per conditioner + the vector block
        mbar_expect_tx(w_bar, image_bytes(n_nets));
        for (int n = 0; n < n_nets; ++n) bulk_load(base + (uint32_t)n * NET_BYTES, p.image + (size_t)n * NET_BYTES, NET_BYTES, w_bar);
        bulk_load(base + off_vec, p.image + (size_t)n_nets * NET_BYTES, (uint32_t)n_nets * VEC_BYTES, w_bar);
    }

    const int grp = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const bool issuer = (threadIdx.x & 127) == 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (BIAS_MMA) {  // the constant A fragment of the bias MMA: identical values from every group (a group's
                               // first MMA follows its own stores)
        uint32_t one[8] = {0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(lane_base + ONES_COL, one);
    }
    mbar_wait_backoff(w_bar, 0);
    uint32_t acc_phase = 0;  // bit i: phase of slot grp * P + i

    // A operand of the next layer of slot i -> TMEM; once the whole tile is there the group's first thread issues the layer
    auto publish = [&](int i, const uint32_t(&r)[32], int step) {
        const int slot = grp * P + i;
        tmem_st32(lane_base + SLOT_COLS * (uint32_t)slot, r);
        tmem_st_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");
        if (issuer) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const ulonglong2 bd = sdesc[step];
            const uint32_t idesc = (step % 3 == 2) ? IDESC_O : IDESC_H;
            const uint32_t mma_a = tmem_base + SLOT_COLS * (uint32_t)slot, mma_d = mma_a + 32u;
            if (p.debug & 8) {  // timing experiment: 4 of the 7 MMAs
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_ts(mma_d, mma_a + 8u * k, bd.x + (uint64_t)(2 * k), k != 0, idesc);
            } else if ((p.debug & 16) && step % 3 != 2) {  // timing experiment: two independent accumulation chains (D, D + 16)
                umma_tf32_ts(mma_d, mma_a, bd.x, 0u, IDESC_H);
                umma_tf32_ts(mma_d + 16u, mma_a, bd.y, 0u, IDESC_H);
                umma_tf32_ts(mma_d, mma_a + 8u, bd.x + 2u, 1u, IDESC_H);
                umma_tf32_ts(mma_d + 16u, mma_a + 8u, bd.y + 2u, 1u, IDESC_H);
                umma_tf32_ts(mma_d, mma_a + 16u, bd.x + 4u, 1u, IDESC_H);
                umma_tf32_ts(mma_d + 16u, tmem_base + ONES_COL, bd.y + 4u, 1u, IDESC_H);
                umma_tf32_ts(mma_d, mma_a + 24u, bd.x + 6u, 1u, IDESC_H);
            } else if (p.debug & 64) {  // timing experiment: one MMA
                umma_tf32_ts(mma_d, mma_a, bd.x, 0u, idesc);
            } else if (!(p.debug & 2)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_ts(mma_d, mma_a + 8u * k, bd.x + (uint64_t)(2 * k), k != 0, idesc);  // (hi | lo) x (W_hi | W_hi)
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_tf32_ts(mma_d, mma_a + 8u * k, bd.y + (uint64_t)(2 * k), 1u, idesc);  // hi x W_lo
                if constexpr (BIAS_MMA) umma_tf32_ts(mma_d, tmem_base + ONES_COL, bd.y + 4u, 1u, idesc);  // (1, 1, 0...) x (b_hi, b_lo, 0...)
            }
            umma_commit(acc_ready(slot));
        }
        __syncwarp();
    };
    auto acquire = [&](int i) {
        mbar_wait_backoff(acc_ready(grp * P + i), (acc_phase >> i) & 1u);
        acc_phase ^= 1u << i;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    auto split = [&](const float(&pre)[16], uint32_t(&r)[32]) {
        if (p.debug & 4) act_split16<true>(pre, r);
        else act_split16<false>(pre, r);
    };

#pragma unroll 1
    for (long long tile0 = (long long)blockIdx.x + (long long)(grp * P) * gridDim.x; tile0 < n_tiles; tile0 += (long long)S * gridDim.x) {
        // slot grp * P + i works on tile0 + i * gridDim.x; the trailing slots of the last round may be idle (group-uniform)
        float v0[P], v1[P], ld[P];
        bool active[P], live[P];
        long long r_glob[P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const long long tile = tile0 + (long long)i * gridDim.x;
            active[i] = tile < n_tiles;
            r_glob[i] = tile * BM + row;
            live[i] = active[i] && r_glob[i] < p.n_rows;
            const float2 xin = ld_stream2(reinterpret_cast<const float2 *>(p.x) + (live[i] ? r_glob[i] : p.n_rows - 1));
            v0[i] = xin.x, v1[i] = xin.y, ld[i] = 0.f;
        }
        int net = 0;
#pragma unroll 1
        for (int kk = 0; kk < p.prog.n_ops; ++kk) {
            const mnf_flow_op &op = p.prog.ops[inverse ? p.prog.n_ops - 1 - kk : kk];
            if (op.type == MNF_OP_AFFINE_CONST) {
                const float4 st = *reinterpret_cast<const float4 *>(p.params + op.aux_off);  // s0 s1 t0 t1
                if (inverse) {  // affine_constant_flow.py:24
                    const float e0 = expf(-st.x), e1 = expf(-st.y);
#pragma unroll
                    for (int i = 0; i < P; ++i) v0[i] = (v0[i] - st.z) * e0, v1[i] = (v1[i] - st.w) * e1, ld[i] -= st.x + st.y;
                } else {  // affine_constant_flow.py:19
                    const float e0 = expf(st.x), e1 = expf(st.y);
#pragma unroll
                    for (int i = 0; i < P; ++i) v0[i] = v0[i] * e0 + st.z, v1[i] = v1[i] * e1 + st.w, ld[i] += st.x + st.y;
                }
            } else if (op.type == MNF_OP_GLOW) {
                const float4 W = *reinterpret_cast<const float4 *>(p.params + op.aux_off + (inverse ? 4 : 0));  // glow.py:28,36: v @ W
                const float lg = p.params[op.aux_off + 8];
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const float n0 = fmaf(v1[i], W.z, v0[i] * W.x), n1 = fmaf(v1[i], W.w, v0[i] * W.y);
                    v0[i] = n0, v1[i] = n1;
                    ld[i] += inverse ? -lg : lg;
                }
            } else {  // NSF_CL: f1 on (lower -> upper) then f2 on (upper -> lower) going forward (spline_flow.py:249-266);
                      // f2 first, then f1, both with the spline inverse, going backward (:268-285)
#pragma unroll 1
                for (int step = 0; step < 2; ++step, ++net) {
                    const bool use_f1 = (step == 0) != (inverse != 0);
                    const float *vec = svec + net * VEC_FLOATS;
                    // layer 0 (1 -> 16) on the FMA pipe, conditioner of spline_flow.py:252-253 / :258-259
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        if (!active[i]) continue;
                        const float c = use_f1 ? v0[i] : v1[i];
                        const float4 *wb = reinterpret_cast<const float4 *>(vec);
                        float pre[16];
                        uint32_t r[32];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w = wb[j], b = wb[4 + j];
                            pre[4 * j] = fmaf(w.x, c, b.x), pre[4 * j + 1] = fmaf(w.y, c, b.y);
                            pre[4 * j + 2] = fmaf(w.z, c, b.z), pre[4 * j + 3] = fmaf(w.w, c, b.w);
                        }
                        split(pre, r);
                        publish(i, r, 3 * net);
                    }
#pragma unroll 1
                    for (int l = 0; l < 2; ++l) {
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            if (!active[i]) continue;
                            acquire(i);
                            uint32_t t[16], r[32];
                            tmem_ld16(lane_base + SLOT_COLS * (uint32_t)(grp * P + i) + 32u, t);
                            tmem_ld_wait();
                            float pre[16];
                            if constexpr (BIAS_MMA) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) pre[j] = __uint_as_float(t[j]);
                            } else {
                                const float4 *bb = reinterpret_cast<const float4 *>(vec + 2 * H + l * H);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 b = bb[j];
                                    pre[4 * j] = __uint_as_float(t[4 * j]) + b.x, pre[4 * j + 1] = __uint_as_float(t[4 * j + 1]) + b.y;
                                    pre[4 * j + 2] = __uint_as_float(t[4 * j + 2]) + b.z, pre[4 * j + 3] = __uint_as_float(t[4 * j + 3]) + b.w;
                                }
                            }
                            split(pre, r);
                            publish(i, r, 3 * net + 1 + l);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        if (!active[i]) continue;
                        acquire(i);
                        uint32_t t0[16], t1[8];
                        const uint32_t d_addr = lane_base + SLOT_COLS * (uint32_t)(grp * P + i) + 32u;
                        tmem_ld16(d_addr, t0);
                        tmem_ld8(d_addr + 16u, t1);
                        tmem_ld_wait();
                        float raw[NB];
#pragma unroll
                        for (int j = 0; j < 16; ++j) raw[j] = __uint_as_float(t0[j]);
#pragma unroll
                        for (int j = 0; j < NB - 16; ++j) raw[16 + j] = __uint_as_float(t1[j]);
                        if constexpr (!BIAS_MMA) {
#pragma unroll
                            for (int j = 0; j < NB; ++j) raw[j] += vec[4 * H + j];
                        }
                        float tr = use_f1 ? v1[i] : v0[i], l = 0.f;
                        if (!(p.debug & 1)) rq_spline<KBINS, true>(raw, KBINS, op.bound, op.edge_deriv, inverse != 0, tr, l);
                        else tr += raw[0] + raw[22], l = raw[11];
                        ld[i] += l;
                        if (use_f1) v1[i] = tr; else v0[i] = tr;
                    }
                }
            }
            if (p.inter) {
#pragma unroll
                for (int i = 0; i < P; ++i)
                    if (live[i]) st_stream2(reinterpret_cast<float2 *>(p.inter + ((size_t)kk * p.n_rows + r_glob[i]) * 2), make_float2(v0[i], v1[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
            if (!live[i]) continue;
            float lp = fmaf(-0.5f, fmaf(v0[i], v0[i], v1[i] * v1[i]), -1.8378770664093453f);  // -(D/2) log(2 pi), D = 2
            if (sum_lp) lp += ld[i];
            if (p.y) st_stream2(reinterpret_cast<float2 *>(p.y) + r_glob[i], make_float2(v0[i], v1[i]));
            if (p.log_det) p.log_det[r_glob[i]] = ld[i];
            if (p.base_lp) p.base_lp[r_glob[i]] = lp;
            // fused gather: the result also goes straight to the other ranks over NVLink
            if (p.gather.multicast_ptr) {
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.gather.multicast_ptr + p.gather.row_offset + r_glob[i]), "f"(lp)
                             : "memory");
            } else {
                for (int g = 0; g < p.gather.n_peers; ++g) p.gather.peer_ptrs[g][p.gather.row_offset + r_glob[i]] = lp;
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

template <int G, int P, bool BIAS_MMA>
static int launch_v(const Params &p, long long n_tiles, int sm_count, size_t smem, cudaStream_t st) {
    const long long per_cta = (n_tiles + G * P - 1) / (G * P);
    const unsigned grid = (unsigned)(per_cta < sm_count ? (per_cta < 1 ? 1 : per_cta) : sm_count);
    MNF_CUDA(cudaFuncSetAttribute(flow_tc_kernel<G, P, BIAS_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flow_tc_kernel<G, P, BIAS_MMA><<<grid, 128 * G, smem, st>>>(p);
    return launch_status("flow_tc_kernel");
}

// =========================================================================================================
// fp16-split variant: the same 3-term product with fp16 operands (kind::f16, 11-bit significands like TF32, but 16 values
// per MMA instead of 8): 3 MMAs per layer instead of 7, an A operand of 16 TMEM columns instead of 32, 8 tiles in flight.
// Measured on the TF32 version: a 128 x 16 x 8 MMA costs ~20 cycles of the (serial) tensor queue whatever its size, so
// the instruction count is what matters.
//   hi = fp16(a), lo = fp16(a - hi): a = hi + lo to 2^-24 relative, or 3e-8 absolute where lo is subnormal (|a| < 0.25);
//   an activation beyond the fp16 range becomes inf, which reaches every output of the conditioner as inf / NaN: the
//   owning thread then re-evaluates ITS point's conditioner in plain fp32 from the parameter blob (rare, divergent).
// Weight tile row n (K-major, 128-byte swizzle): (W_hi[16] | W_hi[16] | W_lo[16] | 0[16]) fp16; biases are added by the
// epilogue from shared memory.  Only the second hidden layer and the output layer are GEMMs (see the table below).
// =========================================================================================================
constexpr uint32_t W16L_BYTES = H * 128, W16O_BYTES = NOUT * 128, NET16_BYTES = W16L_BYTES + W16O_BYTES;  // 6 KB: hidden layer 2, output layer
// Layers 0 and 1 of a conditioner need no GEMM at all: its input is ONE scalar c, so W1 leaky(w0 c + b0) + b1 is a
// piecewise-linear function of c with the H breakpoints -b0_j / w0_j -- on each of the H + 1 intervals it is A_i c + B_i.
// The image holds the sorted breakpoints and the (A_i | B_i) rows (built in fp64 per parameter version); a thread finds its
// interval with H compares and forms the layer-1 pre-activations with H FMAs: 16 instead of 16 + 256 multiply-adds per
// point, and one MMA round trip per conditioner less.
constexpr int TBL_STRIDE = 36;                                    // floats per (A_i | B_i) row, padded off the bank period
constexpr int VEC16_B2 = 0, VEC16_B3 = H, VEC16_TB = H + NOUT, VEC16_TBL = VEC16_TB + H;
constexpr int VEC16_FLOATS = VEC16_TBL + (H + 1) * TBL_STRIDE;   // b2[16] b3[32] breakpoints[16] table[17][36]
constexpr uint32_t VEC16_BYTES = VEC16_FLOATS * 4;
static_assert(VEC16_BYTES % 16 == 0 && (VEC16_TBL * 4) % 16 == 0 && (TBL_STRIDE * 4) % 16 == 0, "16-byte loads / bulk copies");
constexpr uint32_t SLOT16_COLS = 48;  // A: 8 columns hi, 8 columns lo; D: 32 columns at + 16
__host__ __device__ constexpr uint32_t image16_bytes(int n_nets) { return (uint32_t)n_nets * (NET16_BYTES + VEC16_BYTES); }
__host__ __device__ constexpr uint32_t f16_instr_desc(int m, int n) {  // D = F32, A = B = F16, both K-major
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t IDESC16_H = f16_instr_desc(BM, H), IDESC16_O = f16_instr_desc(BM, NOUT);

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {  // `lo` lands in the low half (K index 2j)
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t h) {
    float2 r;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(h));
    return r;
}

__global__ void flow_tc16_image_kernel(const __grid_constant__ NetList nets, const float *__restrict__ params,
                                       uint8_t *__restrict__ image) {
    const int net = blockIdx.x;
    const float *src = params + nets.off[net];
    uint16_t *tiles = reinterpret_cast<uint16_t *>(image + (size_t)net * NET16_BYTES);
    constexpr int OFF1 = 2 * H, OFF2 = OFF1 + H * H + H, OFF3 = OFF2 + H * H + H;
    for (int e = threadIdx.x; e < (int)(NET16_BYTES / 2); e += blockDim.x) {
        int layer = 1, r = e;  // MLP layer index: 1 = second hidden layer (blob offset OFF2), 2 = output layer (OFF3)
        if (r >= (int)(W16L_BYTES / 2)) layer = 2, r -= W16L_BYTES / 2;
        const int n_out = layer == 2 ? NB : H;
        const int n = r >> 6, pos = (r >> 3) & 7, j = r & 7;
        const int k = ((pos ^ (n & 7)) << 3) + j;  // logical K index of this half (128-byte swizzle: 16-byte chunk ^= row & 7)
        const float *W = src + (layer == 1 ? OFF2 : OFF3);
        __half v = __float2half_rn(0.f);
        if (n < n_out && k < 48) {
            const float w = W[n * H + (k & 15)];
            const __half hi = __float2half_rn(w);
            v = k < 32 ? hi : __float2half_rn(w - __half2float(hi));
        }
        tiles[e] = __half_as_ushort(v);
    }
    float *vec = reinterpret_cast<float *>(image + (size_t)nets.n * NET16_BYTES + (size_t)net * VEC16_BYTES);
    for (int e = threadIdx.x; e < VEC16_TB; e += blockDim.x) {
        float v = 0.f;
        if (e < H) v = src[OFF2 + H * H + e];                   // b2
        else if (e < H + NB) v = src[OFF3 + NB * H + (e - H)];   // b3
        vec[e] = v;
    }
    // layers 0 + 1 as a piecewise-linear table: thread i builds interval i (between the sorted breakpoints i - 1 and i)
    if (threadIdx.x <= H) {
        const int i = threadIdx.x;
        const float *w0 = src, *b0 = src + H, *W1 = src + OFF1, *b1 = src + OFF1 + H * H;
        double ts[H];
        for (int j = 0; j < H; ++j) ts[j] = w0[j] != 0.f ? -(double)b0[j] / (double)w0[j] : INFINITY;
        for (int a = 1; a < H; ++a) {  // insertion sort, ascending (+inf = "no breakpoint" last)
            const double v = ts[a];
            int b = a - 1;
            while (b >= 0 && ts[b] > v) ts[b + 1] = ts[b], --b;
            ts[b + 1] = v;
        }
        if (i < H) vec[VEC16_TB + i] = (float)ts[i];
        const double lo = i > 0 ? ts[i - 1] : -INFINITY, hi = i < H ? ts[i] : INFINITY;
        double c = 0.0;  // a point inside the interval: fixes the sign of every first-layer pre-activation on it
        if (isfinite(lo) && isfinite(hi)) c = 0.5 * (lo + hi);
        else if (isfinite(hi)) c = hi - 1.0 - fabs(hi);
        else if (isfinite(lo)) c = lo + 1.0 + fabs(lo);
        float *row = vec + VEC16_TBL + i * TBL_STRIDE;
        for (int k = 0; k < H; ++k) {
            double A = 0.0, B = (double)b1[k];
            for (int j = 0; j < H; ++j) {
                const double pre = (double)w0[j] * c + (double)b0[j];
                const double slope = pre > 0.0 ? 1.0 : 0.2;  // LeakyReLU(0.2), mlp.py:9
                A += (double)W1[k * H + j] * slope * (double)w0[j];
                B += (double)W1[k * H + j] * slope * (double)b0[j];
            }
            row[k] = (float)A, row[H + k] = (float)B;
        }
        for (int k = 2 * H; k < TBL_STRIDE; ++k) row[k] = 0.f;
    }
}

// the three MMAs of one layer and their commit, one instruction stream (issued by one thread):
//   D  = hi x W_hi ; D += lo x W_hi ; D += hi x W_lo
__device__ __forceinline__ void issue_layer16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred pt, pf;\n\t.reg .b64 b1, b2;\n\t.reg .b32 a1;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\tsetp.eq.u32 pt, 0, 0;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u32 a1, %1, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, pf;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, pt;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b2, %3, pt;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t}" ::"r"(d),
        "r"(a), "l"(b), "r"(idesc), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): the kernel is bound by instruction issue, one instruction per pair
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
// LeakyReLU(0.2) (mlp.py:9) of 16 pre-activations (8 pairs) -> 8 columns of packed fp16 hi, 8 columns of packed fp16 lo
__device__ __forceinline__ void act_split16_h(const float2 (&pre)[8], uint32_t (&r)[16]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 s = f2_mul(pre[j], make_float2(0.2f, 0.2f));
        const float2 a = make_float2(fmaxf(pre[j].x, s.x), fmaxf(pre[j].y, s.y));
        const uint32_t h = pack_h2(a.x, a.y);
        const float2 lo = f2_fma(unpack_h2(h), make_float2(-1.f, -1.f), a);  // a - hi, exact
        r[j] = h;
        r[8 + j] = pack_h2(lo.x, lo.y);
    }
}
// exact-fp32 re-evaluation of one point's conditioner (activations outside the fp16 range): plain loops over the
// parameter blob, per Linear weight[out][in], bias[out]
__device__ __noinline__ void conditioner_fp32(const float *__restrict__ net, float c, float *raw) {
    float h[H], g[H];
    for (int j = 0; j < H; ++j) h[j] = leaky02(fmaf(net[j], c, net[H + j]));
    const float *W = net + 2 * H;
    for (int l = 0; l < 2; ++l, W += H * H + H) {
        for (int j = 0; j < H; ++j) {
            float acc = W[H * H + j];
            for (int i = 0; i < H; ++i) acc = fmaf(W[j * H + i], h[i], acc);
            g[j] = leaky02(acc);
        }
        for (int j = 0; j < H; ++j) h[j] = g[j];
    }
    for (int j = 0; j < NB; ++j) {
        float acc = W[NB * H + j];
        for (int i = 0; i < H; ++i) acc = fmaf(W[j * H + i], h[i], acc);
        raw[j] = acc;
    }
}

struct NetOffsets {
    int off[MAX_NETS];  // execution order
};

template <int G>
__global__ void __launch_bounds__(128 * G, 1) flow_tc16_kernel(const __grid_constant__ Params p, const __grid_constant__ NetOffsets noff) {
    static_assert(G * (int)SLOT16_COLS <= 512 && G <= 15, "TMEM columns / named barriers");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int n_nets = p.n_nets;
    const uint32_t off_vec = (uint32_t)n_nets * NET16_BYTES, off_bar = off_vec + (uint32_t)n_nets * VEC16_BYTES;
    const uint32_t bars = base + off_bar;
    const uint32_t w_bar = bars;
    auto acc_ready = [&](int s) { return bars + 8u * (1 + s); };
    const uint32_t tmem_slot = bars + 8u * (1 + G);
    const uint32_t *tmem_slot_ptr = reinterpret_cast<const uint32_t *>(base_ptr + off_bar + 8u * (1 + G));
    const float *svec = reinterpret_cast<const float *>(base_ptr + off_vec);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inverse = p.dir_flags & 1;
    const bool sum_lp = p.dir_flags & 2;
    const long long n_tiles = (p.n_rows + BM - 1) / BM;

    if (warp == 1 && lane == 0) {
        mbar_init(w_bar, 1);
        for (int s = 0; s < G; ++s) mbar_init(acc_ready(s), 1);
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

    if (warp == 0 && lane == 0) {
        mbar_expect_tx(w_bar, image16_bytes(n_nets));
        for (int n = 0; n < n_nets; ++n) bulk_load(base + (uint32_t)n * NET16_BYTES, p.image + (size_t)n * NET16_BYTES, NET16_BYTES, w_bar);
        bulk_load(base + off_vec, p.image + (size_t)n_nets * NET16_BYTES, (uint32_t)n_nets * VEC16_BYTES, w_bar);
    }

    const int slot = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const bool issuer = (threadIdx.x & 127) == 0;
    const uint32_t a_addr = tmem_base + ((uint32_t)(q * 32) << 16) + SLOT16_COLS * (uint32_t)slot, d_addr = a_addr + 16u;
    const uint32_t mma_a = tmem_base + SLOT16_COLS * (uint32_t)slot, mma_d = mma_a + 16u;
    const uint32_t my_bar = acc_ready(slot);
    const uint64_t desc0 = make_smem_desc(base);
    mbar_wait_backoff(w_bar, 0);
    uint32_t acc_phase = 0;

    // A operand of the next layer -> TMEM; once the whole tile is there the group's first thread issues the layer's MMAs
    auto publish = [&](const uint32_t(&r)[16], uint32_t tile_off, uint32_t idesc) {
        if (!(p.debug & 256)) {  // (timing experiment: no TMEM stores)
            tmem_st16(a_addr, r);
            tmem_st_wait();
        }
        if (p.debug & 1024) return;  // (timing experiment: no synchronisation at all)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");
        if (issuer) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!(p.debug & 2)) issue_layer16(mma_d, mma_a, desc0 + (uint64_t)(tile_off >> 4), idesc, my_bar);
            else if (p.debug & 512) mbar_arrive(my_bar);  // (timing experiment: plain arrive instead of tcgen05.commit)
            else umma_commit(my_bar);
        }
        __syncwarp();
    };
    auto acquire = [&]() {
        if (p.debug & 1024) return;
        mbar_wait_backoff(my_bar, acc_phase, (uint32_t)p.sleep_ns);
        acc_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

#pragma unroll 1
    for (long long tile = (long long)blockIdx.x + (long long)slot * gridDim.x; tile < n_tiles; tile += (long long)G * gridDim.x) {
        const long long r_glob = tile * BM + row;
        const bool live = r_glob < p.n_rows;
        float v0, v1, ld = 0.f;
        {
            const float2 xin = ld_stream2(reinterpret_cast<const float2 *>(p.x) + (live ? r_glob : p.n_rows - 1));
            v0 = xin.x, v1 = xin.y;
        }
        int net = 0;
#pragma unroll 1
        for (int kk = 0; kk < p.prog.n_ops; ++kk) {
            const mnf_flow_op &op = p.prog.ops[inverse ? p.prog.n_ops - 1 - kk : kk];
            if (op.type == MNF_OP_AFFINE_CONST) {
                const float4 st = *reinterpret_cast<const float4 *>(p.params + op.aux_off);  // s0 s1 t0 t1
                if (inverse) {  // affine_constant_flow.py:24
                    v0 = (v0 - st.z) * expf(-st.x), v1 = (v1 - st.w) * expf(-st.y), ld -= st.x + st.y;
                } else {  // affine_constant_flow.py:19
                    v0 = v0 * expf(st.x) + st.z, v1 = v1 * expf(st.y) + st.w, ld += st.x + st.y;
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
                for (int step = 0; step < 2; ++step, ++net) {
                    const bool use_f1 = (step == 0) != (inverse != 0);
                    const float c = use_f1 ? v0 : v1;
                    const float *vec = svec + net * VEC16_FLOATS;
                    const uint32_t tiles = (uint32_t)net * NET16_BYTES;
                    uint32_t r[16];
                    {  // layers 0 and 1: interval of c among the sorted breakpoints, then pre1 = A_i c + B_i
                        const float4 *tb = reinterpret_cast<const float4 *>(vec + VEC16_TB);
                        int iv = 0;
#pragma unroll
                        for (int j = 0; j < H / 4; ++j) {
                            const float4 t4 = tb[j];
                            iv += (c > t4.x) + (c > t4.y) + (c > t4.z) + (c > t4.w);
                        }
                        const float4 *row = reinterpret_cast<const float4 *>(vec + VEC16_TBL + iv * TBL_STRIDE);
                        float2 pre[8];
                        const float2 cc = make_float2(c, c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 a4 = row[j], b4 = row[4 + j];
                            pre[2 * j] = f2_fma(make_float2(a4.x, a4.y), cc, make_float2(b4.x, b4.y));
                            pre[2 * j + 1] = f2_fma(make_float2(a4.z, a4.w), cc, make_float2(b4.z, b4.w));
                        }
                        act_split16_h(pre, r);
                    }
                    publish(r, tiles, IDESC16_H);  // second hidden layer
                    {
                        acquire();
                        uint32_t t[16];
                        if (!(p.debug & 128)) {  // (timing experiment: no TMEM loads)
                            tmem_ld16(d_addr, t);
                            tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) t[j] = r[j];
                        }
                        const float4 *bb = reinterpret_cast<const float4 *>(vec + VEC16_B2);
                        float2 pre[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = bb[j];
                            pre[2 * j] = f2_add(make_float2(__uint_as_float(t[4 * j]), __uint_as_float(t[4 * j + 1])), make_float2(b.x, b.y));
                            pre[2 * j + 1] = f2_add(make_float2(__uint_as_float(t[4 * j + 2]), __uint_as_float(t[4 * j + 3])), make_float2(b.z, b.w));
                        }
                        act_split16_h(pre, r);
                        publish(r, tiles + W16L_BYTES, IDESC16_O);  // output layer
                    }
                    acquire();
                    float raw[NB];
                    {
                        uint32_t t0[16], t1[8];
                        if (!(p.debug & 128)) {
                            tmem_ld16(d_addr, t0);
                            tmem_ld8(d_addr + 16u, t1);
                            tmem_ld_wait();
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) t0[j] = r[j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) t1[j] = r[j + 8];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) raw[j] = __uint_as_float(t0[j]) + vec[VEC16_B3 + j];
#pragma unroll
                        for (int j = 0; j < NB - 16; ++j) raw[16 + j] = __uint_as_float(t1[j]) + vec[VEC16_B3 + 16 + j];
                    }
                    if (!(fabsf(raw[0] + raw[KBINS] + raw[2 * KBINS]) < 3.0e38f)) {  // an activation left the fp16 range
                        float exact[NB];
                        conditioner_fp32(p.params + noff.off[net], c, exact);
#pragma unroll
                        for (int j = 0; j < NB; ++j) raw[j] = exact[j];
                    }
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int G>
static int launch_16(const Params &p, const NetOffsets &noff, long long n_tiles, int sm_count, cudaStream_t st) {
    const size_t smem = image16_bytes(p.n_nets) + 8 * (2 + G) + 1024;
    const long long per_cta = (n_tiles + G - 1) / G;
    const unsigned grid = (unsigned)(per_cta < sm_count ? (per_cta < 1 ? 1 : per_cta) : sm_count);
    MNF_CUDA(cudaFuncSetAttribute(flow_tc16_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flow_tc16_kernel<G><<<grid, 128 * G, smem, st>>>(p, noff);
    return launch_status("flow_tc16_kernel");
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
    const char *venv = getenv("MNF_FTC_V");  // tuning: 200 + groups = fp16 split; else TF32 split: groups * 10 + tiles per
                                              // thread (+ 100: biases added by the epilogue)
    const int v = venv ? atoi(venv) : 207;  // measured at config 2: 207 3.85 ms, 208 3.90, 206 3.93; TF32 split 71 4.17
    int rc;
    if (v >= 200) {
        flow_tc16_image_kernel<<<nets.n, 256, 0, stream>>>(nets, params, image);
        rc = launch_status("flow_tc16_image_kernel");
    } else {
        flow_tc_image_kernel<<<nets.n, 256, 0, stream>>>(nets, params, image);
        rc = launch_status("flow_tc_image_kernel");
    }
    if (rc) return rc;
    Params p{};
    p.prog.n_ops = n_ops;
    for (int k = 0; k < n_ops; ++k) p.prog.ops[k] = ops[k];
    p.params = params, p.x = x, p.image = image, p.y = y, p.log_det = log_det, p.base_lp = base_lp, p.inter = inter;
    p.n_rows = n_rows, p.dir_flags = dir_flags & 3, p.n_nets = nets.n;
    const char *dbg = getenv("MNF_FTC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    const char *slp = getenv("MNF_FTC_SLEEP");
    p.sleep_ns = slp ? atoi(slp) : 40;
    if (gather) {
        MNF_REQUIRE(gather->n_peers >= 0 && gather->n_peers <= MNF_MAX_PEERS, MNF_E_ARG, "bad n_peers");
        p.gather = *gather;
    }
    const size_t smem = image_bytes(nets.n) + (size_t)nets.n * 3 * 16 + 8 * (2 + 8) + 1024;
    MNF_REQUIRE(smem <= (size_t)dp->smem_optin, MNF_E_SHAPE, "weight image does not fit in shared memory");
    const long long n_tiles = (n_rows + tc::BM - 1) / tc::BM;
    if (v >= 200) {
        NetOffsets noff{};
        for (int i = 0; i < nets.n; ++i) noff.off[i] = nets.off[i];
        switch (v) {
            case 206: return launch_16<6>(p, noff, n_tiles, dp->sm_count, stream);
            case 207: return launch_16<7>(p, noff, n_tiles, dp->sm_count, stream);
            case 208: return launch_16<8>(p, noff, n_tiles, dp->sm_count, stream);
            default: return fail(MNF_E_ARG, "MNF_FTC_V: unknown variant %d", v);
        }
    }
    switch (v) {
        case 41: return launch_v<4, 1, true>(p, n_tiles, dp->sm_count, smem, stream);
        case 61: return launch_v<6, 1, true>(p, n_tiles, dp->sm_count, smem, stream);
        case 71: return launch_v<7, 1, true>(p, n_tiles, dp->sm_count, smem, stream);
        case 32: return launch_v<3, 2, true>(p, n_tiles, dp->sm_count, smem, stream);
        case 23: return launch_v<2, 3, true>(p, n_tiles, dp->sm_count, smem, stream);
        case 142: return launch_v<4, 2, false>(p, n_tiles, dp->sm_count, smem, stream);
        case 181: return launch_v<8, 1, false>(p, n_tiles, dp->sm_count, smem, stream);
        case 171: return launch_v<7, 1, false>(p, n_tiles, dp->sm_count, smem, stream);
        default: return fail(MNF_E_ARG, "MNF_FTC_V: unknown variant %d", v);
    }
}

}  // namespace mnf
