// rnvp_fused.cu -- second half of an RNVP flow (rnvp.py:32-39) for wide dims, fused on tcgen05:
//     [shift | scale] = y [Wt; Ws]^T + b      (y = net(mask * z) [R, 64], the conditioner output of the first GEMM)
//     gate = sigmoid(scale);  z <- (1 - mask) z gate + (1 - gate) shift + mask z;  log_det += sum (1 - mask) log gate
// as ONE persistent kernel, so that the [R, 2 * dim] shift/scale matrix (8.6 GB per 131 072 rows at dim 4096, written
// and re-read by the separate GEMM + gate kernels this replaces) never exists.  HBM traffic per flow: read z, write z,
// write the next GEMM operand (tf32(mask' z') for the next flow or tf32(x z') for MNFLinear's mean GEMM).
//
// A CTA walks row tiles of 128 rows; for each it keeps the y tile (A operand, 32 KB, one TMA load) and sweeps the dim/32
// column chunks.  Per chunk: the 64 interleaved (shift_n, scale_n) weight rows arrive by TMA (L2-resident), one thread
// issues 8 tcgen05.mma.kind::tf32 (M128 x N64 x K8) into one of four TMEM accumulators, and the z tile [128 x 32] arrives
// by TMA into a 4-stage ring -- every byte of z, z' and the next operand moves through TMA boxes, because a thread owns a
// ROW of the accumulator (TMEM lane = row) and row-per-lane global accesses are 32 separate cache lines per instruction
// (the first version of this kernel did that and ran at 2.1 TB/s, L1TEX-wavefront bound).  Eight epilogue warps (two per
// TMEM lane quarter, 16 dims each) read their z segment from the swizzled tile, apply the gate, write z' back in place
// and the next operand into a staging tile; one thread issues the TMA stores.  The biases ride in column 63 of y (set to
// one by the first GEMM's packed bias).  Log-det partials stay in registers across the chunks of a row tile.
#include "rnvp_fused.cuh"
#include "tc_common.cuh"

namespace mnf {
namespace rnvpf {
using namespace tc;

constexpr int HP = 64;            // conditioner width padded (column 63 = constant one)
constexpr int CH = 32;            // dims per chunk = 64 accumulator columns = one 128-byte row of the z tile
constexpr int B_STAGES = 2, Z_STAGES = 4, O_STAGES = 2, ACC = 4;
constexpr int EPI_WARPS = 8, THREADS = 128 + EPI_WARPS * 32;
constexpr int X_MAX_ROWS = 64;    // TMA path for the x multiplier needs x_rows <= 64 and 128 % x_rows == 0
constexpr uint32_t A_BYTES = BM * HP * 4;      // y tile: two K-blocks of [128 x 32]
constexpr uint32_t KBLK = BM * 32 * 4;         // [128 rows x 128 B]
constexpr uint32_t B_BYTES = 2 * CH * HP * 4;  // weight chunk [64 rows x 64]: two K-blocks of [64 x 32]
constexpr uint32_t BBLK = 2 * CH * 32 * 4;     // 8 KB
constexpr uint32_t Z_BYTES = BM * CH * 4;      // 16 KB
constexpr uint32_t X_BYTES = X_MAX_ROWS * CH * 4;
constexpr uint32_t OFF_B = A_BYTES, OFF_Z = OFF_B + B_STAGES * B_BYTES, OFF_O = OFF_Z + Z_STAGES * Z_BYTES,
                   OFF_X = OFF_O + O_STAGES * Z_BYTES, OFF_RED = OFF_X + Z_STAGES * X_BYTES, OFF_BAR = OFF_RED + 2 * 128 * 4,
                   SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t IDESC = tf32_instr_desc(BM, 2 * CH);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// round-to-nearest TF32 for an operand the tensor core will truncate: half a TF32 ulp added to the bit pattern
__device__ __forceinline__ float rnt(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(THREADS, 1)
rnvp_out_gate_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_o,
                     const __grid_constant__ CUtensorMap map_x, const Params p, const int x_tma) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + OFF_BAR;
    auto a_full = [&]() { return bars; };
    auto a_empty = [&]() { return bars + 8u; };
    auto b_full = [&](int s) { return bars + 8u * (2 + s); };
    auto b_empty = [&](int s) { return bars + 8u * (2 + B_STAGES + s); };
    auto z_full = [&](int s) { return bars + 8u * (2 + 2 * B_STAGES + s); };
    auto z_empty = [&](int s) { return bars + 8u * (2 + 2 * B_STAGES + Z_STAGES + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 + 2 * B_STAGES + 2 * Z_STAGES + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 + 2 * B_STAGES + 2 * Z_STAGES + ACC + a); };
    const uint32_t tmem_slot = bars + 8u * (2 + 2 * B_STAGES + 2 * Z_STAGES + 2 * ACC);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float *red = reinterpret_cast<float *>(smem_raw + (base + OFF_RED - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full(), 1), mbar_init(a_empty(), 1);
        for (int s = 0; s < B_STAGES; ++s) mbar_init(b_full(s), 1), mbar_init(b_empty(s), 1);
        for (int s = 0; s < Z_STAGES; ++s) mbar_init(z_full(s), 1), mbar_init(z_empty(s), 1);
        for (int a = 0; a < ACC; ++a) mbar_init(acc_full(a), 1), mbar_init(acc_empty(a), EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int n_tiles = (int)((p.n_rows + BM - 1) / BM), n_chunks = p.dim / CH;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer: y tile per row tile; weight chunk + z tile (+ x tile) per column chunk ----------------
        int bs = 0, zs = 0, it = 0;
        uint32_t bph = 0, zph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            mbar_wait(a_empty(), (it & 1u) ^ 1u);
            mbar_expect_tx(a_full(), A_BYTES);
            tma_load_2d(base, &map_y, a_full(), 0, tile * BM);
            tma_load_2d(base + KBLK, &map_y, a_full(), 32, tile * BM);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(b_empty(bs), bph ^ 1u);
                mbar_expect_tx(b_full(bs), B_BYTES);
                const uint32_t sb = base + OFF_B + bs * B_BYTES;
                tma_load_2d(sb, &map_w, b_full(bs), 0, c * 2 * CH);
                tma_load_2d(sb + BBLK, &map_w, b_full(bs), 32, c * 2 * CH);
                if (++bs == B_STAGES) bs = 0, bph ^= 1u;
                mbar_wait(z_empty(zs), zph ^ 1u);
                mbar_expect_tx(z_full(zs), Z_BYTES + (x_tma ? (uint32_t)p.xmul_rows * CH * 4 : 0u));
                tma_load_2d(base + OFF_Z + zs * Z_BYTES, &map_z, z_full(zs), c * CH, tile * BM);
                if (x_tma) tma_load_2d(base + OFF_X + zs * X_BYTES, &map_x, z_full(zs), c * CH, 0);
                if (++zs == Z_STAGES) zs = 0, zph ^= 1u;
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer ----------------
        int bs = 0, acc = 0, it = 0;
        uint32_t bph = 0, aph = 0;
        const uint64_t adesc = make_smem_desc(base);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            mbar_wait(a_full(), it & 1u);
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(b_full(bs), bph);
                mbar_wait(acc_empty(acc), aph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bdesc = make_smem_desc(base + OFF_B + bs * B_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(acc * 2 * CH);
#pragma unroll
                for (int k = 0; k < HP / UMMA_K; ++k)
                    umma_tf32(d, adesc + (uint64_t)((k >> 2) * (KBLK >> 4) + 2 * (k & 3)),
                              bdesc + (uint64_t)((k >> 2) * (BBLK >> 4) + 2 * (k & 3)), k != 0, IDESC);
                umma_commit(b_empty(bs));
                umma_commit(acc_full(acc));
                if (++bs == B_STAGES) bs = 0, bph ^= 1u;
                if (++acc == ACC) acc = 0, aph ^= 1u;
            }
            umma_commit(a_empty());  // every MMA that read this y tile has completed when this fires
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: 8 warps, thread = (row, 16 dims of the chunk) ----------------
        const int q = warp & 3, sub = (warp - 4) >> 2, row = q * 32 + lane;
        const bool leader = warp == 4 && lane == 0;
        const bool have_o = p.mz_next != nullptr || p.xz_out != nullptr;
        const uint32_t swz = (uint32_t)(row & 7), rowoff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const int xr = p.xmul_rows > 0 ? row % p.xmul_rows : 0;
        const uint32_t xoff = (uint32_t)((xr >> 3) * 1024 + (xr & 7) * 128), xswz = (uint32_t)(xr & 7);
        const Philox rng(p.seed);
        int acc = 0, zs = 0;
        uint32_t aph = 0, zph = 0, g = 0;  // g: chunks processed by this CTA so far (TMA store groups issued)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long m = (long long)tile * BM + row;
            const bool live = m < p.n_rows;
            const long long mm = live ? m : p.n_rows - 1;  // rows past the end compute on a valid row; TMA clips their stores
            float ld = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++g) {
                const int d0 = c * CH + sub * 16;
                const size_t e0 = (size_t)mm * p.dim + d0;
                const uint64_t g0 = (uint64_t)p.row_offset * p.dim + e0;  // global element index (a multiple of 16)
                uint32_t mbits = 0, nbits = 0;  // Bernoulli masks of this / the next flow for the 16 dims, one bit each
                if (p.mask) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = *reinterpret_cast<const float4 *>(p.mask + e0 + 4 * j);
                        mbits |= (v.x != 0.f ? 1u : 0u) << (4 * j) | (v.y != 0.f ? 1u : 0u) << (4 * j + 1) |
                                 (v.z != 0.f ? 1u : 0u) << (4 * j + 2) | (v.w != 0.f ? 1u : 0u) << (4 * j + 3);
                    }
                } else {
                    mbits = philox_bits16(rng, g0, p.stream);
                }
                if (p.mz_next) {
                    if (p.mask_next) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 v = *reinterpret_cast<const float4 *>(p.mask_next + e0 + 4 * j);
                            nbits |= (v.x != 0.f ? 1u : 0u) << (4 * j) | (v.y != 0.f ? 1u : 0u) << (4 * j + 1) |
                                     (v.z != 0.f ? 1u : 0u) << (4 * j + 2) | (v.w != 0.f ? 1u : 0u) << (4 * j + 3);
                        }
                    } else {
                        nbits = philox_bits16(rng, g0, p.next_stream);
                    }
                }
                // this thread's 16 dims of z (and of x) from the TMA-staged tiles: chunks 4 * sub .. 4 * sub + 3 of its row
                mbar_wait(z_full(zs), zph);
                const uint32_t zt = base + OFF_Z + zs * Z_BYTES + rowoff;
                float zi[16], xv[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = lds128(zt + (((uint32_t)(4 * sub + j) ^ swz) << 4));
                    zi[4 * j] = v.x, zi[4 * j + 1] = v.y, zi[4 * j + 2] = v.z, zi[4 * j + 3] = v.w;
                }
                if (p.xz_out) {
                    if (x_tma) {
                        const uint32_t xt = base + OFF_X + zs * X_BYTES + xoff;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 v = lds128(xt + (((uint32_t)(4 * sub + j) ^ xswz) << 4));
                            xv[4 * j] = v.x, xv[4 * j + 1] = v.y, xv[4 * j + 2] = v.z, xv[4 * j + 3] = v.w;
                        }
                    } else {
                        const float *xrow = p.xmul + (size_t)(mm % p.xmul_rows) * p.dim + d0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 v = *reinterpret_cast<const float4 *>(xrow + 4 * j);
                            xv[4 * j] = v.x, xv[4 * j + 1] = v.y, xv[4 * j + 2] = v.z, xv[4 * j + 3] = v.w;
                        }
                    }
                }
                mbar_wait(acc_full(acc), aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tcol = tmem_base + (uint32_t)(acc * 2 * CH + sub * 32) + ((uint32_t)(q * 32) << 16);
                float zn[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // 16 accumulator columns = 8 (shift, scale) pairs per load
                    uint32_t r[16];
                    tmem_ld16(tcol + (uint32_t)(h * 16), r);
                    if (h == 1) {  // the accumulator is in registers: release it before the arithmetic
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(acc));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = 8 * h + u;
                        const float sh = __uint_as_float(r[2 * u]), sc = __uint_as_float(r[2 * u + 1]);
                        const float mk = (float)((mbits >> j) & 1u);
                        const float den = 1.f + ex2f(-1.4426950408889634f * sc);  // 1 + exp(-scale)
                        const float gate = rcpf(den);                              // torch.sigmoid, rnvp.py:35
                        zn[j] = ((1.f - mk) * zi[j] * gate + (1.f - gate) * sh) + mk * zi[j];  // rnvp.py:37
                        ld -= (1.f - mk) * lg2f(den);                              // log gate = -log(den), rnvp.py:36 (x ln 2 at the end)
                    }
                }
                if (++acc == ACC) acc = 0, aph ^= 1u;
                // ---- outputs through TMA: z' in place in the z tile, the next operand in a staging tile ----
                if (leader) {
                    // all store groups but the newest have finished reading shared memory: chunk g - 2's z stage and this
                    // chunk's staging buffer are free
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    if (g >= 2) mbar_arrive(z_empty((int)((g - 2) % Z_STAGES)));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                const uint32_t ot = base + OFF_O + (g & 1u) * Z_BYTES + rowoff;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t cs = (((uint32_t)(4 * sub + j) ^ swz) << 4);
                    const float a0 = zn[4 * j], a1 = zn[4 * j + 1], a2 = zn[4 * j + 2], a3 = zn[4 * j + 3];
                    if (p.write_z) sts128(zt + cs, a0, a1, a2, a3);
                    if (p.mz_next)
                        sts128(ot + cs, rnt((float)((nbits >> (4 * j)) & 1u) * a0), rnt((float)((nbits >> (4 * j + 1)) & 1u) * a1),
                               rnt((float)((nbits >> (4 * j + 2)) & 1u) * a2), rnt((float)((nbits >> (4 * j + 3)) & 1u) * a3));
                    else if (p.xz_out)
                        sts128(ot + cs, rnt(xv[4 * j] * a0), rnt(xv[4 * j + 1] * a1), rnt(xv[4 * j + 2] * a2), rnt(xv[4 * j + 3] * a3));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (leader) {
                    if (p.write_z) tma_store_2d(&map_z, base + OFF_Z + zs * Z_BYTES, c * CH, tile * BM);
                    if (have_o) tma_store_2d(&map_o, base + OFF_O + (g & 1u) * Z_BYTES, c * CH, tile * BM);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (++zs == Z_STAGES) zs = 0, zph ^= 1u;
            }
            // the two warps that share a row add up their log-det partials (one row tile = one CTA: no atomics)
            red[sub * 128 + row] = ld;
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            if (sub == 0 && live) {
                const float t = (red[row] + red[128 + row]) * 0.6931471805599453f;
                p.log_det[m] = p.accumulate_ld ? p.log_det[m] + t : t;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
        // every store must have left shared memory AND reached global memory before the CTA exits
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

bool eligible(int dim, int h) { return dim % CH == 0 && dim >= 64 && h >= 1 && h <= HP - 1; }

int launch(const float *y, const float *Wts, const Params &p, cudaStream_t stream) {
    const DeviceProps *dp = device_props();
    MNF_REQUIRE(dp != nullptr && dp->cc_major == 10, MNF_E_DEVICE, "tcgen05 path needs an sm_100 device");
    MNF_REQUIRE(p.dim % CH == 0 && p.n_rows >= 1 && p.n_rows <= 0x7fffffff - 256, MNF_E_SHAPE, "bad shape for the fused RNVP gate");
    MNF_REQUIRE(((uintptr_t)p.z % 16) == 0 && (!p.mz_next || ((uintptr_t)p.mz_next % 16) == 0) &&
                    (!p.xz_out || ((uintptr_t)p.xz_out % 16) == 0) && (!p.xmul || ((uintptr_t)p.xmul % 16) == 0),
                MNF_E_ALIGN, "pointers must be 16-byte aligned");
    MNF_REQUIRE(!(p.mz_next && p.xz_out), MNF_E_ARG, "one next-operand output at a time");
    CUtensorMap my, mw, mz, mo, mx;
    int rc = make_map(&my, y, (int)p.n_rows, HP, BM);
    if (rc) return rc;
    rc = make_map(&mw, Wts, 2 * p.dim, HP, 2 * CH);
    if (rc) return rc;
    rc = make_map(&mz, p.z, (int)p.n_rows, p.dim, BM);
    if (rc) return rc;
    const float *optr = p.mz_next ? p.mz_next : (p.xz_out ? p.xz_out : p.z);
    rc = make_map(&mo, optr, (int)p.n_rows, p.dim, BM);
    if (rc) return rc;
    const int x_tma = (p.xz_out && p.xmul && p.xmul_rows >= 1 && p.xmul_rows <= X_MAX_ROWS && BM % p.xmul_rows == 0) ? 1 : 0;
    rc = x_tma ? make_map(&mx, p.xmul, p.xmul_rows, p.dim, p.xmul_rows) : make_map(&mx, p.z, (int)p.n_rows, p.dim, BM);
    if (rc) return rc;
    MNF_CUDA(cudaFuncSetAttribute(rnvp_out_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const int n_tiles = (int)((p.n_rows + BM - 1) / BM);
    const unsigned grid = (unsigned)(n_tiles < dp->sm_count ? n_tiles : dp->sm_count);
    rnvp_out_gate_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(my, mw, mz, mo, mx, p, x_tma);
    return launch_status("rnvp_out_gate_kernel");
}

}  // namespace rnvpf
}  // namespace mnf
