// made_fused.cu -- MAF.inverse (density direction, maf.py:53-62 over made.py:22-23) for a whole STACK of MAF flows as ONE
// persistent tcgen05 kernel: the activations of the masked MLPs never touch HBM.
//
// Shape class: dim 64, hidden widths <= 31 (padded to 32; column 31 is a constant one that carries the biases), 1..4
// hidden layers, up to 16 flows -- BASELINE config 3 is 9 x MADE(64-24-24-24-128).
//
// One CTA per SM, 512 threads:
//   warp 0      weight producer: one cp.async.bulk per flow brings the flow's four pre-swizzled weight matrices (32 KB,
//               packed once per parameter version by the host) into a 2-slot shared-memory ring
//   warps 1-3   MMA issuers, one per tile in flight: one thread each issues its tile's tcgen05.mma.kind::tf32 (M128; N32 for
//               the hidden layers, N128 for the (s, t) output layer); tcgen05.commit publishes the accumulator
//   warp 2      also the TMEM allocator (512 columns; a tile's four layers alias the same 128 columns)
//   warps 4-15  three epilogue groups of 128 threads, one 128-row tile in flight each.  A thread owns one row: its 64
//               exact fp32 coordinates stay in REGISTERS across all flows.  Per layer it reads its accumulator row
//               with tcgen05.ld, applies bias / ReLU, rounds to TF32 and writes the row into the 128-byte-swizzled
//               K-major operand tile of the next MMA (fence.proxy.async + mbarrier); after the output layer it applies
//               z_i = x_i exp(s_i) + t_i, adds s_i to its log-det and restages tf32(z) for the next flow.
// The tile arrives by TMA (two 128 x 32 boxes) and the result leaves by TMA store; HBM traffic is the algorithmic
// 516 B/row.  The parity flips (maf.py:60) are folded into the packed weights (a reversed flow has its input columns and
// output pairs permuted), so nothing is permuted on chip; only the final store may reverse the row.
#include <stdlib.h>

#include "tc_common.cuh"

namespace mnf {
namespace madef {
using namespace tc;

constexpr int D = 64, HP = 32, MAX_FLOWS = 16, MAX_HIDDEN = 4;
constexpr uint32_t ZBLK = BM * 32 * 4;          // one K-block of the point tile: 128 rows x 128 B
constexpr uint32_t Z_BYTES = 2 * ZBLK;          // 128 x 64 fp32
constexpr uint32_t H_BYTES = BM * HP * 4;       // 128 x 32 hidden activations
constexpr uint32_t W1_BYTES = HP * D * 4;       // [32, 64] as two K-blocks of [32 x 32]
constexpr uint32_t WH_BYTES = HP * HP * 4;      // [32, 32]
constexpr uint32_t WO_BYTES = 2 * D * HP * 4;   // [128, 32], rows interleaved (s_0, t_0, s_1, t_1, ...)
constexpr uint32_t WSLOT = W1_BYTES + (MAX_HIDDEN - 1) * WH_BYTES + WO_BYTES;
// T = tiles in flight per CTA (one epilogue group each), S = threads per row (the S warps with the same warp % 4 split a
// row's columns).  Register split between the 4 control warps and the epilogue warps (setmaxnreg; 0 = leave as compiled):
//   T S  threads  control / epilogue registers
//   2 1    384       56 / 224
//   3 1    512       56 / 152
//   2 2    640       40 / 104   (the CTA is launched with 640 x 96 registers: 128 x 56 freed = 512 x 14 -> +8)
//   3 2    896       -- / 72 (uniform)
template <int T, int S>
struct Layout {
    static constexpr int THREADS = 128 + T * S * 128;
    static constexpr int REG_CTRL = (S == 1) ? 56 : (T == 2 ? 40 : 0);
    static constexpr int REG_EPI = (S == 1) ? (T == 2 ? 224 : 152) : (T == 2 ? 104 : 0);
    // registers the CTA holds at launch (per-thread count rounded down to a multiple of 8) must cover the split
    static constexpr int REG_LAUNCH = (65536 / THREADS) / 8 * 8;
    static_assert(REG_EPI == 0 || 128 * REG_CTRL + (THREADS - 128) * REG_EPI <= THREADS * REG_LAUNCH, "setmaxnreg split exceeds the CTA's registers");
    static constexpr uint32_t OFF_H = T * Z_BYTES, OFF_W = OFF_H + T * H_BYTES, OFF_B1 = OFF_W + 2 * WSLOT,
                              OFF_RED = OFF_B1 + MAX_FLOWS * HP * 4, OFF_BAR = OFF_RED + T * 256 * 4,
                              SMEM_BYTES = OFF_BAR + 256 + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};
constexpr uint32_t IDESC_H = tf32_instr_desc(BM, HP), IDESC_O = tf32_instr_desc(BM, 2 * D);
constexpr float LN2 = 0.6931471805599453f, HALF_LOG_2PI = 0.9189385332046727f;

struct Params {
    const float *wimg;  // [n_flows][flow_floats] pre-swizzled weight images (device)
    const float *b1;    // [n_flows][32] first-layer biases (device)
    float *log_det;     // [n_rows] or NULL
    float *log_prob;    // [n_rows] or NULL: log_det + standard-normal log-density of the result
    long long n_rows;
    int n_flows, n_hidden, final_reversed, store_z;
    int debug;  // timing experiments only (MNF_MADE_DEBUG): 1 = no proxy fence, 2 = no operand stores, 4 = no MMAs, 8 = no epilogue math
};

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// round-to-nearest TF32 for an operand the tensor core will truncate: adding half a TF32 ulp to the bit pattern and
// letting the MMA drop the low 13 bits is cvt.rna (ties away from zero) in ONE integer add; the cvt instruction itself
// expands to four (it also handles inf / nan, which cannot survive a flow anyway)
__device__ __forceinline__ float rnt(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }
// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of re-issuing try_wait in a tight loop:
// the 24 epilogue warps of a CTA spend most of their time here and their spinning competes for issue slots
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
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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
    constexpr uint32_t OFF_H = L::OFF_H, OFF_W = L::OFF_W, OFF_B1 = L::OFF_B1, OFF_BAR = L::OFF_BAR, OFF_RED = L::OFF_RED;
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
    // this CTA owns tiles blockIdx.x + j * gridDim.x, j < my_tiles; epilogue group t takes j = t, t + T, ...
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

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
        // control warps hand most of their registers to the epilogue groups (a thread there keeps a 64-wide row resident)
        if constexpr (L::REG_CTRL != 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(L::REG_CTRL));
        if (warp == 0 && lane == 0) {
            // ---------------- weight producer ----------------
            uint32_t wcount = 0;
            for (int j0 = 0; j0 < my_tiles; j0 += T) {
                for (int f = 0; f < F; ++f, ++wcount) {
                    const int ws = wcount & 1;
                    mbar_wait(w_empty(ws), ((wcount >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(w_full(ws), flow_bytes);
                    bulk_load(base + OFF_W + ws * WSLOT, p.wimg + (size_t)f * (flow_bytes / 4), flow_bytes, w_full(ws));
                }
            }
        } else if (warp >= 1 && warp <= T && lane == 0) {
            // ---------------- MMA issuers: one per tile slot, so the tiles' layer chains run independently ----------------
            // (a single issuer walking the tiles round-robin kept them in lockstep: its per-tile wake-up / descriptor /
            // commit latency was paid T times per layer and the MMA time never overlapped the epilogues)
            const int t = warp - 1;
            const uint32_t d = tmem_base + (uint32_t)(t * 128);
            const uint64_t az = make_smem_desc(base + t * Z_BYTES), ah = make_smem_desc(base + OFF_H + t * H_BYTES);
            uint32_t wcount = 0, a_ph = 0;
#pragma unroll 1
            const int iters = (my_tiles + T - 1) / T;
            for (int i = 0; i < iters; ++i) {              // every issuer walks the same (iteration, flow) sequence ...
                const bool active = i * T + t < my_tiles;  // ... idle ones only keep the weight-ring counts right
#pragma unroll 1
                for (int f = 0; f < F; ++f, ++wcount) {
                    const int ws = wcount & 1;
                    if (!active) {
                        mbar_wait(w_full(ws), (wcount >> 1) & 1u);
                        mbar_arrive(w_empty(ws));
                        continue;
                    }
                    mbar_wait(w_full(ws), (wcount >> 1) & 1u);
                    const uint32_t wb = base + OFF_W + ws * WSLOT;
                    const uint64_t b1d = make_smem_desc(wb);
#pragma unroll 1
                    for (int l = 0; l <= NH; ++l) {
                        mbar_wait_parked(a_ready(t), a_ph);
                        a_ph ^= 1u;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (p.debug & 4) {
                        } else if (l == 0) {
#pragma unroll
                            for (int k = 0; k < D / UMMA_K; ++k) {
                                // k-block kb = k / 4 (16 KB apart in A, 4 KB in B), 32 B per k-step inside the swizzle atom
                                const uint64_t ao = (uint64_t)((k >> 2) * (ZBLK >> 4) + 2 * (k & 3));
                                const uint64_t bo = (uint64_t)((k >> 2) * ((HP * 128) >> 4) + 2 * (k & 3));
                                umma_tf32(d, az + ao, b1d + bo, k != 0, IDESC_H);
                            }
                        } else {
                            const uint64_t b0 = b1d + (uint64_t)((W1_BYTES + (uint32_t)(l - 1) * WH_BYTES) >> 4);
                            const uint32_t idesc = l == NH ? IDESC_O : IDESC_H;
#pragma unroll
                            for (int k = 0; k < HP / UMMA_K; ++k) umma_tf32(d, ah + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), k != 0, idesc);
                        }
                        umma_commit(acc_ready(t));
                    }
                    umma_commit(w_empty(ws));  // this issuer's reads of the weight slot have completed when this fires
                }
            }
        }
    } else {
        if constexpr (Layout<T, S>::REG_EPI != 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Layout<T, S>::REG_EPI));
        // ---------------- epilogue groups: one tile in flight per group, S threads per row ----------------
        // a warp may only touch TMEM lanes [32 * (warp % 4), +32): the S warps with the same warp % 4 share those rows and
        // split their columns -- thread `part` owns dims [part * D / S, (part + 1) * D / S) of row q * 32 + lane
        constexpr int WPS = 4 * S;            // warps per group
        constexpr int ZC = 16 / S;            // 16-byte chunks of the point row per thread
        constexpr int DPT = D / S;            // dims per thread
        const int w = (warp - 4) % WPS, slot = (warp - 4) / WPS, q = w & 3, part = w >> 2, row = q * 32 + lane;
        const bool leader = (w == 0 && lane == 0);
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t zrow = base + slot * Z_BYTES + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t hrow = base + OFF_H + slot * H_BYTES + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t trow = tmem_base + (uint32_t)(slot * 128) + ((uint32_t)(q * 32) << 16);
        float *red = reinterpret_cast<float *>(smem_raw + (base + OFF_RED - smem_u32(smem_raw))) + slot * 256;
        // address of the thread's i-th chunk of the point row (chunks ZC * part + i; 8 chunks per 16 KB K-block)
        auto zaddr = [&](int i) {
            const int cc = ZC * part + i;
            return zrow + (uint32_t)(cc >> 3) * ZBLK + (((uint32_t)(cc & 7) ^ swz) << 4);
        };
        auto publish = [&]() {  // operand rows written: make them visible to the tensor core, one arrival per warp
            if (!(p.debug & 1)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(slot));
        };
        uint32_t in_phase = 0, acc_phase = 0;
        int tile = (int)blockIdx.x + slot * (int)gridDim.x;
        if (leader && tile < n_tiles) {
            mbar_expect_tx(in_full(slot), Z_BYTES);
            tma_load_2d(base + slot * Z_BYTES, &map_x, in_full(slot), 0, tile * BM);
            tma_load_2d(base + slot * Z_BYTES + ZBLK, &map_x, in_full(slot), 32, tile * BM);
        }
#pragma unroll 1
        for (; tile < n_tiles; tile += T * (int)gridDim.x) {
            float z[DPT];
            mbar_wait_parked(in_full(slot), in_phase);
            in_phase ^= 1u;
#pragma unroll
            for (int i = 0; i < ZC; ++i) {
                const uint32_t a = zaddr(i);
                const float4 v = lds128(a);
                z[4 * i] = v.x, z[4 * i + 1] = v.y, z[4 * i + 2] = v.z, z[4 * i + 3] = v.w;
                sts128(a, rnt(v.x), rnt(v.y), rnt(v.z), rnt(v.w));
            }
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
                    for (int g = 0; g < 2 / S; ++g) {  // 16 hidden columns per load: bias (first layer), ReLU, TF32 rounding, operand store
                        const int col0 = 16 * (part * (2 / S) + g);
                        uint32_t r[16];
                        tmem_ld16(trow + (uint32_t)col0, r);
                        if (p.debug & 8) {
                        } else if (l == 0) {
                            const float4 *bb = reinterpret_cast<const float4 *>(sb1 + f * HP + col0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4 b = bb[c];
                                const float h0 = rnt(fmaxf(__uint_as_float(r[4 * c]) + b.x, 0.f));
                                const float h1 = rnt(fmaxf(__uint_as_float(r[4 * c + 1]) + b.y, 0.f));
                                const float h2 = rnt(fmaxf(__uint_as_float(r[4 * c + 2]) + b.z, 0.f));
                                float h3 = rnt(fmaxf(__uint_as_float(r[4 * c + 3]) + b.w, 0.f));
                                if (col0 + 4 * c + 3 == HP - 1) h3 = 1.f;  // the constant-one column that carries the later biases
                                sts128(hrow + (((uint32_t)(col0 / 4 + c) ^ swz) << 4), h0, h1, h2, h3);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                sts128(hrow + (((uint32_t)(col0 / 4 + c) ^ swz) << 4), rnt(fmaxf(__uint_as_float(r[4 * c]), 0.f)),
                                       rnt(fmaxf(__uint_as_float(r[4 * c + 1]), 0.f)), rnt(fmaxf(__uint_as_float(r[4 * c + 2]), 0.f)),
                                       rnt(fmaxf(__uint_as_float(r[4 * c + 3]), 0.f)));
                        }
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
                    if (!last) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int i = 2 * g + c;
                            sts128(zaddr(i), rnt(z[4 * i]), rnt(z[4 * i + 1]), rnt(z[4 * i + 2]), rnt(z[4 * i + 3]));
                        }
                    }
                }
                if (!last) publish();
                else asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
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
                if (p.final_reversed) {  // physical position j holds logical dim 63 - j: chunk cc goes to chunk 15 - cc, reversed
#pragma unroll
                    for (int i = 0; i < ZC; ++i) {
                        const int cc = 15 - (ZC * part + i);
                        sts128(zrow + (uint32_t)(cc >> 3) * ZBLK + (((uint32_t)(cc & 7) ^ swz) << 4), z[4 * i + 3], z[4 * i + 2],
                               z[4 * i + 1], z[4 * i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < ZC; ++i) sts128(zaddr(i), z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(WPS * 32) : "memory");
                if (leader) {
                    tma_store_2d(&map_z, base + slot * Z_BYTES, 0, tile * BM);
                    tma_store_2d(&map_z, base + slot * Z_BYTES + ZBLK, 32, tile * BM);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile buffer may be overwritten
                }
            } else if constexpr (S == 2) {
                // no store: the partner may still be reading `red` / nothing else is shared; the next tile's load must only
                // wait for this group's last operand reads, which completed before the output accumulator was published
            }
            const int next = tile + T * (int)gridDim.x;
            if (leader && next < n_tiles) {
                mbar_expect_tx(in_full(slot), Z_BYTES);
                tma_load_2d(base + slot * Z_BYTES, &map_x, in_full(slot), 0, next * BM);
                tma_load_2d(base + slot * Z_BYTES + ZBLK, &map_x, in_full(slot), 32, next * BM);
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

int64_t mnf_made_fused_image_floats(int n_hidden) {
    if (n_hidden < 1 || n_hidden > madef::MAX_HIDDEN) return 0;
    return (int64_t)(madef::W1_BYTES + (uint32_t)(n_hidden - 1) * madef::WH_BYTES + madef::WO_BYTES) / 4;
}

int mnf_made_density_fused(const float *wimg, const float *b1, int n_flows, int n_hidden, int final_reversed,
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
        case 22: return launch_fused<2, 2>(mx, mz, p, grid, st);
        case 0:
        case 32: return launch_fused<3, 2>(mx, mz, p, grid, st);
        default: return fail(MNF_E_ARG, "variant must be 0 (default), 21, 31, 22 or 32 (10 * tiles in flight + threads per row)");
    }
}

}  // extern "C"
