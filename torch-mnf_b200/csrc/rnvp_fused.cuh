// rnvp_fused.cuh -- host interface of the fused RNVP output-GEMM + gate kernel (rnvp_fused.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mnf {
namespace rnvpf {

struct Params {
    float *z;               // [R, dim] in / out
    float *log_det;         // [R]
    const float *mask;      // [R, dim] injected Bernoulli mask of THIS flow, or NULL (Philox, stream `stream`)
    const float *mask_next; // injected mask of the next flow or NULL (Philox, stream `next_stream`)
    float *mz_next;         // [R, dim] tf32(mask_next * z') or NULL
    const float *xmul;      // [xmul_rows, dim] or NULL
    float *xz_out;          // [R, dim] tf32(x[m % xmul_rows] * z') or NULL
    long long n_rows;
    int dim, xmul_rows, write_z, accumulate_ld;
    uint64_t seed, row_offset;
    uint32_t stream, next_stream;
};

// dim must be a multiple of 64 and the conditioner width at most 63 (column 63 of the padded conditioner output is the
// constant one that carries the shift / scale biases)
bool eligible(int dim, int h);
// y: [R, 64] TF32-rounded conditioner output whose column 63 is one; Wts: [2 * dim, 64] interleaved (shift_n, scale_n) rows,
// column 63 = bias.  Enqueues one launch on `stream`.
int launch(const float *y, const float *Wts, const Params &p, cudaStream_t stream);

}  // namespace rnvpf
}  // namespace mnf
