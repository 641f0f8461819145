// made_fused.cu -- MAF.inverse (density direction, maf.py:53-62 over made.py:22-23) for a whole STACK of MAF flows as ONE
// persistent tcgen05 kernel: the activations of the masked MLPs never touch HBM.
//
// Shape class: dim 64, hidden widths <= 31 (padded to 32; column 31 is a constant one that carries the biases), 1..4
// hidden layers, up to 16 flows -- BASELINE config 3 is 9 x MADE(64-24-24-24-128).
//
// Operand precision: fp16 (tcgen05.mma.kind::f16, fp32 accumulation in TMEM).  fp16 has the 11-bit significand of TF32,
// i.e. the same 2e-3 tolerance class BASELINE.json states for the MADE GEMMs, at half the bytes: the kernel is bound by
// shared-memory bandwidth (operand tiles written by the epilogue threads and read by the tensor core), and the TF32
// version of this kernel measured 1.8x slower (profiles/r02_made_fused_history.md).  Operands are converted with
// round-to-nearest and saturate at +-65504; the running point z itself stays in exact fp32 registers.
//
// One CTA per SM, T tiles of 128 rows in flight, S threads per row:
//   warps 0..T-1   MMA issuers, one per tile: a single thread issues the tile's MMAs (M128; N32 for the hidden layers,
//                  N128 for the (s, t) output layer) and commits the accumulator to an mbarrier.  Separate issuers keep
//                  the tiles' layer chains independent, so one tile's MMAs overlap the other tiles' epilogues.  Issuer 0
//                  also streams the weights: one cp.async.bulk per flow brings the flow's pre-swizzled fp16 matrices
//                  (packed once per parameter version by the host) into a 2-slot shared-memory ring.
//   warp 2         also the TMEM allocator (512 columns; a tile's layers alias the same 128 columns)
//   warps 4..      T epilogue groups of S x 128 threads.  A thread owns (1/S of) one row: the exact fp32 coordinates stay
//                  in REGISTERS across all flows.  Per layer it reads its accumulator columns with tcgen05.ld, applies
//                  bias / ReLU, converts to fp16 and writes them into the 128-byte-swizzled K-major operand tile of the
//                  next MMA (fence.proxy.async + mbarrier); after the output layer it applies z_i = x_i exp(s_i) + t_i,
//                  adds s_i to its log-det and restages fp16(z) for the next flow.
// A tile arrives by TMA (two 128 x 32 fp32 boxes) and leaves by TMA store from the same 32 KB region, which holds the
// operand tiles in between; HBM traffic is the algorithmic 516 B/row.  The parity flips (maf.py:60) are folded into the
// packed weights (a reversed flow has its input columns and output pairs permuted), so nothing is permuted on chip; only
// the final store may reverse the row.
#include <stdlib.h>

#include "tc_common.cuh"

namespace mnf {
namespace madef {
using namespace tc;

constexpr int D = 64, HP = 32, MAX_FLOWS = 16, MAX_HIDDEN = 4;
constexpr uint32_t ZBLK = BM * 32 * 4;      // one fp32 K-block of the staged tile: 128 rows x 128 B
constexpr uint32_t TILE_BYTES = 2 * ZBLK;   // per tile slot: fp32 staging [128 x 64]; later fp16 z [128 x 64] | fp16 h [128 x 32 (+32 pad)]
constexpr uint32_t OPER_H = ZBLK;           // offset of the hidden-activation operand inside the slot
constexpr uint32_t W1_BYTES = HP * 128;     // [32 rows x 64 fp16]
constexpr uint32_t WH_BYTES = HP * 128;     // [32 rows x 32 fp16, rows padded to 128 B]
constexpr uint32_t WO_BYTES = 2 * D * 128;  // [128 rows x 32 fp16, padded], rows interleaved (s_0, t_0, s_1, t_1, ...)
constexpr uint32_t WSLOT = W1_BYTES + (MAX_HIDDEN - 1) * WH_BYTES + WO_BYTES;
// T = tiles in flight per CTA (one epilogue group each), S = threads per row (the S warps with the same warp % 4 split a
// row's columns).  Register split between the 4 control warps and the epilogue warps (setmaxnreg; 0 = leave as compiled):
//   T S  threads  launch regs  control / epilogue registers
//   2 1    384       168            56 / 224
//   3 1    512       128            56 / 152
//   4 1    640        96            32 / 112
//   3 2    896        72            -- / 72 (uniform)
template <int T, int S>
struct Layout {
    static constexpr int THREADS = 128 + T * S * 128;
    static constexpr int REG_CTRL = (S == 2) ? 0 : (T == 4 ? 32 : 56);
    static constexpr int REG_EPI = (S == 2) ? 0 : (T == 2 ? 224 : (T == 3 ? 152 : 112));
    // registers the CTA holds at launch (per-thread count rounded down to a multiple of 8) must cover the split
    static constexpr int REG_LAUNCH = (65536 / THREADS) / 8 * 8;
    static_assert(REG_EPI == 0 || 128 * REG_CTRL + (THREADS - 128) * REG_EPI <= THREADS * REG_LAUNCH, "setmaxnreg split exceeds the CTA's registers");
    static_assert(T <= 4, "one issuer warp per tile slot");
    static constexpr uint32_t OFF_W = T * TILE_BYTES, OFF_B1 = OFF_W + 2 * WSLOT, OFF_RED = OFF_B1 + MAX_FLOWS * HP * 4,
                              OFF_BAR = OFF_RED + T * 256 * 4, SMEM_BYTES = OFF_BAR + 256 + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};
// tcgen05 instruction descriptor, kind::f16: D = F32 (1 << 4), A = B = F16 (format 0), both K-major
__host__ __device__ constexpr uint32_t f16_instr_desc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t IDESC_H = f16_instr_desc(BM, HP), IDESC_O = f16_instr_desc(BM, 2 * D);
constexpr int UK = 16;  // K per kind::f16 MMA (32 bytes)
constexpr float LN2 = 0.6931471805599453f, HALF_LOG_2PI = 0.9189385332046727f;

struct Params {
    const void *wimg;   // [n_flows][flow_bytes] pre-swizzled fp16 weight images (device)
    const float *b1;    // [n_flows][32] first-layer biases (device)
    float *log_det;     // [n_rows] or NULL
    float *log_prob;    // [n_rows] or NULL: log_det + standard-normal log-density of the result
    long long n_rows;
    int n_flows, n_hidden, final_reversed, store_z;
    int debug;  // timing experiments only (MNF_MADE_DEBUG): 1 = no proxy fence, 4 = no MMAs, 8 = no epilogue math
};

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// two fp32 -> packed fp16 (round to nearest, saturating at +-65504); `lo` lands at the lower address
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of re-issuing try_wait in a tight loop:
// the epilogue warps of a CTA spend most of their time here and their spinning competes for issue slots
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(0x989680u)
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
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int T, int S>
__global__ void __launch_bounds__(Layout<T, S>::THREADS, 1)
made_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z, const Params p) {
    using L = Layout<T, S>;
    constexpr int THREADS = L::THREADS;
    constexpr uint32_t OFF_W = L::OFF_W, OFF_B1 = L::OFF_B1, OFF_BAR = L::OFF_BAR, OFF_RED = L::OFF_RED;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + OFF_BAR;
    auto w_full = [&](int s) { return bars + 8u * s; };
    auto w_empty = [&](int s) { return bars + 8u * (2 + s); };
    auto in_full = [&](int t) { return bars + 8u * (4 + t); };
    auto a_ready = [&](int t) { return bars + 8u * (4 + T + t); };
    auto acc_ready = [&](int t) { return bars + 8u * (4 + 2 * T + t); };
    const uint32_t tmem_slot = bars + 8u * (4 + 3 * T);
    float *sb1 = reinterpret_cast<float *>(smem_raw + (base + OFF_B1 - smem_u32(smem_raw)));
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = p.n_flows, NH = p.n_hidden;
    const uint32_t flow_bytes = W1_BYTES + (uint32_t)(NH - 1) * WH_BYTES + WO_BYTES;
    const int n_tiles = (int)((p.n_rows + BM - 1) / BM);
    // this CTA owns tiles blockIdx.x + j * gridDim.x, j < my_tiles; tile slot t takes j = t, t + T, ...
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int iters = (my_tiles + T - 1) / T;

    for (int i = threadIdx.x; i < F * HP; i += THREADS) sb1[i] = p.b1[i];
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        if (p.store_z) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), T);  // one arrival per issuer
        }
        for (int t = 0; t < T; ++t) {
            mbar_init(in_full(t), 1);
            mbar_init(a_ready(t), 4 * S);  // one arrival per epilogue warp of the group
            mbar_init(acc_ready(t), 1);
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

    if (warp < 4) {
        // control warps hand most of their registers to the epilogue groups (a thread there keeps its row resident)
        if constexpr (L::REG_CTRL != 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(L::REG_CTRL));
        if (warp < T && lane == 0) {
            // ---------------- MMA issuer of tile slot `t` (issuer 0 also streams the weights) ----------------
            const int t = warp;
            const uint32_t d = tmem_base + (uint32_t)(t * 128);
            const uint64_t az = make_smem_desc(base + t * TILE_BYTES), ah = make_smem_desc(base + t * TILE_BYTES + OPER_H);
            const uint32_t total = (uint32_t)(iters * F);  // every issuer walks the same sequence of (iteration, flow) steps
            uint32_t loaded = 0, a_ph = 0;                 // weight images requested so far (issuer 0)
            auto load_weights = [&](uint32_t n) {          // image of step n into ring slot n & 1
                const int ws = n & 1;
                mbar_expect_tx(w_full(ws), flow_bytes);
                bulk_load(base + OFF_W + ws * WSLOT, (const uint8_t *)p.wimg + (size_t)(n % F) * flow_bytes, flow_bytes, w_full(ws));
            };
#pragma unroll 1
            for (uint32_t n = 0; n < total; ++n) {
                const int ws = n & 1;
                const bool active = (int)(n / F) * T + t < my_tiles;  // idle issuers only keep the weight-ring counts right
                if (t == 0) {
                    // step n's image must be on its way before this thread blocks on it; step n + 1 is prefetched as soon as
                    // its slot has been released by every issuer (they finished step n - 1)
                    while (loaded <= n) {
                        mbar_wait(w_empty(loaded & 1), ((loaded >> 1) & 1u) ^ 1u);
                        load_weights(loaded++);
                    }
                    if (loaded == n + 1 && loaded < total && mbar_test(w_empty(loaded & 1), ((loaded >> 1) & 1u) ^ 1u))
                        load_weights(loaded++);
                }
                mbar_wait(w_full(ws), (n >> 1) & 1u);
                if (!active) {
                    mbar_arrive(w_empty(ws));
                    continue;
                }
                const uint64_t b1d = make_smem_desc(base + OFF_W + ws * WSLOT);
#pragma unroll 1
                for (int l = 0; l <= NH; ++l) {
                    mbar_wait_parked(a_ready(t), a_ph);
                    a_ph ^= 1u;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (p.debug & 4) {
                    } else if (l == 0) {
#pragma unroll
                        for (int k = 0; k < D / UK; ++k)  // 32 bytes per k-step inside the 128-byte swizzle atom
                            umma_f16(d, az + (uint64_t)(2 * k), b1d + (uint64_t)(2 * k), k != 0, IDESC_H);
                    } else {
                        const uint64_t b0 = b1d + (uint64_t)((W1_BYTES + (uint32_t)(l - 1) * WH_BYTES) >> 4);
                        const uint32_t idesc = l == NH ? IDESC_O : IDESC_H;
#pragma unroll
                        for (int k = 0; k < HP / UK; ++k) umma_f16(d, ah + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), k != 0, idesc);
                    }
                    umma_commit(acc_ready(t));
                    if (t == 0 && loaded == n + 1 && loaded < total && mbar_test(w_empty(loaded & 1), ((loaded >> 1) & 1u) ^ 1u))
                        load_weights(loaded++);  // another chance to prefetch the next flow's weights
                }
                umma_commit(w_empty(ws));  // this issuer's reads of the weight slot have completed when this fires
            }
        }
    } else {
        if constexpr (L::REG_EPI != 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(L::REG_EPI));
        // ---------------- epilogue groups: one tile in flight per group, S threads per row ----------------
        // a warp may only touch TMEM lanes [32 * (warp % 4), +32): the S warps with the same warp % 4 share those rows and
        // split their columns -- thread `part` owns dims [part * D / S, (part + 1) * D / S) of row q * 32 + lane
        constexpr int WPS = 4 * S;  // warps per group
        constexpr int ZC = 16 / S;  // 16-byte fp32 chunks of the staged row per thread
        constexpr int DPT = D / S;  // dims per thread
        const int w = (warp - 4) % WPS, slot = (warp - 4) / WPS, q = w & 3, part = w >> 2, row = q * 32 + lane;
        const bool leader = (w == 0 && lane == 0);
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t rowoff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t zrow = base + slot * TILE_BYTES + rowoff;  // fp32 staging row (K-block 0; K-block 1 at + ZBLK) = fp16 z row
        const uint32_t hrow = zrow + OPER_H;                      // fp16 hidden row (first 64 of its 128 bytes)
        const uint32_t trow = tmem_base + (uint32_t)(slot * 128) + ((uint32_t)(q * 32) << 16);
        float *red = reinterpret_cast<float *>(smem_raw + (base + OFF_RED - smem_u32(smem_raw))) + slot * 256;
        // fp32 staging: address of 16-byte chunk cc of the row (8 chunks per 16 KB K-block)
        auto saddr = [&](int cc) { return zrow + (uint32_t)(cc >> 3) * ZBLK + (((uint32_t)(cc & 7) ^ swz) << 4); };
        auto publish = [&]() {  // operand rows written: make them visible to the tensor core, one arrival per warp
            if (!(p.debug & 1)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(slot));
        };
        // fp16 z operand: 8 dims per 16-byte chunk; the thread's dims [DPT * part, +DPT) are chunks (DPT / 8) * part + i
        auto stage_z = [&](const float(&z)[DPT]) {
#pragma unroll
            for (int i = 0; i < DPT / 8; ++i)
                sts128u(zrow + (((uint32_t)((DPT / 8) * part + i) ^ swz) << 4), pack_h2(z[8 * i], z[8 * i + 1]),
                        pack_h2(z[8 * i + 2], z[8 * i + 3]), pack_h2(z[8 * i + 4], z[8 * i + 5]), pack_h2(z[8 * i + 6], z[8 * i + 7]));
        };
        uint32_t in_phase = 0, acc_phase = 0;
        int tile = (int)blockIdx.x + slot * (int)gridDim.x;
        if (leader && tile < n_tiles) {
            mbar_expect_tx(in_full(slot), TILE_BYTES);
            tma_load_2d(base + slot * TILE_BYTES, &map_x, in_full(slot), 0, tile * BM);
            tma_load_2d(base + slot * TILE_BYTES + ZBLK, &map_x, in_full(slot), 32, tile * BM);
        }
#pragma unroll 1
        for (; tile < n_tiles; tile += T * (int)gridDim.x) {
            float z[DPT];
            mbar_wait_parked(in_full(slot), in_phase);
            in_phase ^= 1u;
#pragma unroll
            for (int i = 0; i < ZC; ++i) {
                const float4 v = lds128(saddr(ZC * part + i));
                z[4 * i] = v.x, z[4 * i + 1] = v.y, z[4 * i + 2] = v.z, z[4 * i + 3] = v.w;
            }
            // the fp16 operand rows overwrite the staged fp32 rows: with two threads per row the partner must have read first
            if constexpr (S == 2) asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(WPS * 32) : "memory");
            stage_z(z);
            publish();
            float ld4[4] = {0.f, 0.f, 0.f, 0.f};  // four partial sums: the additions must not form one dependent chain
#pragma unroll 1
            for (int f = 0; f < F; ++f) {
#pragma unroll 1
                for (int l = 0; l < NH; ++l) {
                    mbar_wait_parked(acc_ready(slot), acc_phase);
                    acc_phase ^= 1u;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int g = 0; g < 2 / S; ++g) {  // 16 hidden columns per load: bias (first layer), ReLU, fp16, operand store
                        const int col0 = 16 * (part * (2 / S) + g);
                        uint32_t r[16];
                        tmem_ld16(trow + (uint32_t)col0, r);
                        if (p.debug & 8) continue;
                        float h[16];
                        if (l == 0) {
                            const float4 *bb = reinterpret_cast<const float4 *>(sb1 + f * HP + col0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4 b = bb[c];
                                h[4 * c] = fmaxf(__uint_as_float(r[4 * c]) + b.x, 0.f), h[4 * c + 1] = fmaxf(__uint_as_float(r[4 * c + 1]) + b.y, 0.f);
                                h[4 * c + 2] = fmaxf(__uint_as_float(r[4 * c + 2]) + b.z, 0.f), h[4 * c + 3] = fmaxf(__uint_as_float(r[4 * c + 3]) + b.w, 0.f);
                            }
                            if (col0 + 15 == HP - 1) h[15] = 1.f;  // the constant-one column that carries the later layers' biases
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) h[j] = fmaxf(__uint_as_float(r[j]), 0.f);
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c)  // 8 fp16 per 16-byte chunk
                            sts128u(hrow + (((uint32_t)(col0 / 8 + c) ^ swz) << 4), pack_h2(h[8 * c], h[8 * c + 1]), pack_h2(h[8 * c + 2], h[8 * c + 3]),
                                    pack_h2(h[8 * c + 4], h[8 * c + 5]), pack_h2(h[8 * c + 6], h[8 * c + 7]));
                    }
                    publish();
                }
                // output layer: 128 accumulator columns = 64 (s, t) pairs; z_i = x_i exp(s_i) + t_i (maf.py:58), log_det += sum s (maf.py:61)
                mbar_wait_parked(acc_ready(slot), acc_phase);
                acc_phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool last = f == F - 1;
#pragma unroll
                for (int g = 0; g < 8 / S; ++g) {  // 16 accumulator columns = 8 (s, t) pairs = 8 dims per load
                    uint32_t r[16];
                    tmem_ld16(trow + (uint32_t)(part * (128 / S) + g * 16), r);
                    if (p.debug & 8) continue;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float sv = __uint_as_float(r[2 * u]), tv = __uint_as_float(r[2 * u + 1]);
                        z[8 * g + u] = fmaf(z[8 * g + u], ex2(sv), tv);  // the packer scaled the s rows by log2(e)
                        ld4[u & 3] += sv;
                    }
                }
                if (!last) {
                    stage_z(z);
                    publish();
                } else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
            }
            // ---- results ----
            float ld = ((ld4[0] + ld4[1]) + (ld4[2] + ld4[3])) * LN2, ss = 0.f;  // the accumulators hold s * log2(e)
            if (p.log_prob) {
#pragma unroll
                for (int i = 0; i < DPT; ++i) ss = fmaf(z[i], z[i], ss);
            }
            if constexpr (S == 2) {  // the two threads of a row add up their halves through shared memory
                if (part == 1) red[row] = ld, red[128 + row] = ss;
                asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(WPS * 32) : "memory");
                if (part == 0) ld += red[row], ss += red[128 + row];
            }
            const long long grow = (long long)tile * BM + row;
            if (part == 0 && grow < p.n_rows) {
                if (p.log_det) p.log_det[grow] = ld;
                if (p.log_prob) p.log_prob[grow] = ld - 0.5f * ss - (float)D * HALF_LOG_2PI;
            }
            if (p.store_z) {
                // exact fp32 row back into the staging layout (the operand tiles are dead: the last MMA has completed)
                if (p.final_reversed) {  // physical position j holds logical dim 63 - j: chunk cc goes to chunk 15 - cc, reversed
#pragma unroll
                    for (int i = 0; i < ZC; ++i) sts128(saddr(15 - (ZC * part + i)), z[4 * i + 3], z[4 * i + 2], z[4 * i + 1], z[4 * i]);
                } else {
#pragma unroll
                    for (int i = 0; i < ZC; ++i) sts128(saddr(ZC * part + i), z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(WPS * 32) : "memory");
                if (leader) {
                    tma_store_2d(&map_z, base + slot * TILE_BYTES, 0, tile * BM);
                    tma_store_2d(&map_z, base + slot * TILE_BYTES + ZBLK, 32, tile * BM);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile buffer may be overwritten
                }
            } else if constexpr (S == 2) {
                // the partner's last operand stores to this slot are ordered before the next TMA load by this barrier
                asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(WPS * 32) : "memory");
            }
            const int next = tile + T * (int)gridDim.x;
            if (leader && next < n_tiles) {
                mbar_expect_tx(in_full(slot), TILE_BYTES);
                tma_load_2d(base + slot * TILE_BYTES, &map_x, in_full(slot), 0, next * BM);
                tma_load_2d(base + slot * TILE_BYTES + ZBLK, &map_x, in_full(slot), 32, next * BM);
            }
        }
        if (leader && p.store_z) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace madef
}  // namespace mnf

using namespace mnf;

template <int T, int S>
static int launch_fused(const CUtensorMap &mx, const CUtensorMap &mz, const madef::Params &p, unsigned grid, cudaStream_t st) {
    using L = madef::Layout<T, S>;
    MNF_CUDA(cudaFuncSetAttribute(madef::made_fused_kernel<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM_BYTES));
    madef::made_fused_kernel<T, S><<<grid, L::THREADS, L::SMEM_BYTES, st>>>(mx, mz, p);
    return launch_status("made_fused_kernel");
}

extern "C" {

int64_t mnf_made_fused_image_bytes(int n_hidden) {
    if (n_hidden < 1 || n_hidden > madef::MAX_HIDDEN) return 0;
    return (int64_t)(madef::W1_BYTES + (uint32_t)(n_hidden - 1) * madef::WH_BYTES + madef::WO_BYTES);
}

int mnf_made_density_fused(const void *wimg, const float *b1, int n_flows, int n_hidden, int final_reversed,
                           const float *x, float *z, float *log_det, float *log_prob, int64_t n_rows, int dim,
                           int variant, void *stream) {
    MNF_REQUIRE(wimg && b1 && x && (z || log_det || log_prob), MNF_E_ARG, "NULL pointer");
    MNF_REQUIRE(dim == madef::D, MNF_E_SHAPE, "fused MADE kernel is built for dim %d (got %d)", madef::D, dim);
    MNF_REQUIRE(n_flows >= 1 && n_flows <= madef::MAX_FLOWS && n_hidden >= 1 && n_hidden <= madef::MAX_HIDDEN, MNF_E_SHAPE,
                "fused MADE kernel: 1..%d flows, 1..%d hidden layers (got %d, %d)", madef::MAX_FLOWS, madef::MAX_HIDDEN, n_flows,
                n_hidden);
    MNF_REQUIRE(n_rows >= 0 && n_rows <= 0x7fffffff - 256, MNF_E_ARG, "bad row count");
    MNF_REQUIRE(((uintptr_t)wimg % 16) == 0 && ((uintptr_t)x % 16) == 0 && (!z || ((uintptr_t)z % 16) == 0), MNF_E_ALIGN,
                "pointers must be 16-byte aligned");
    if (n_rows == 0) return 0;
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr && dp->cc_major == 10, MNF_E_DEVICE, "tcgen05 path needs an sm_100 device");
    CUtensorMap mx, mz;
    int rc = tc::make_map(&mx, x, (int)n_rows, dim, tc::BM);
    if (rc) return rc;
    rc = tc::make_map(&mz, z ? z : x, (int)n_rows, dim, tc::BM);
    if (rc) return rc;
    const char *dbg = getenv("MNF_MADE_DEBUG");
    madef::Params p{wimg, b1, log_det, log_prob, (long long)n_rows, n_flows, n_hidden, final_reversed, z ? 1 : 0,
                    dbg ? atoi(dbg) : 0};
    const long long n_tiles = (n_rows + tc::BM - 1) / tc::BM;
    const unsigned grid = (unsigned)(n_tiles < dp->sm_count ? n_tiles : dp->sm_count);
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {  // 10 * tiles in flight + threads per row
        case 21: return launch_fused<2, 1>(mx, mz, p, grid, st);
        case 31: return launch_fused<3, 1>(mx, mz, p, grid, st);
        case 0:  // measured at config 3 (2^20 rows x 9 flows): 41 0.59 ms, 31 0.64, 32 0.70, 21 0.83
        case 41: return launch_fused<4, 1>(mx, mz, p, grid, st);
        case 32: return launch_fused<3, 2>(mx, mz, p, grid, st);
        default: return fail(MNF_E_ARG, "variant must be 0 (default), 21, 31, 41 or 32 (10 * tiles in flight + threads per row)");
    }
}

}  // extern "C"
