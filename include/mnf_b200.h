/*
 * mnf_b200.h -- C ABI of libmnf_b200.so: the sm_100a (B200) implementation of the
 * torch-mnf hot path (flow forward/inverse + log-det, MNF layer forward, kl_div).
 *
 * The reference (janosh/torch-mnf) is pure Python on ATen and has no FFI layer; the
 * boundary a maintainer binds is therefore the set of module methods listed beside
 * each entry point (file:line relative to the reference checkout).  The ctypes stub
 * that binds these symbols is shown in INTEGRATION.md and lives in
 * torch-mnf_b200/torch_mnf/_lib.py.
 *
 * Conventions (all entry points):
 *   - plain C: pointers, sizes, POD structs.  No torch / C++ types cross the boundary.
 *   - every data pointer is a DEVICE pointer to contiguous fp32 unless its name ends
 *     in `_host`.  The caller owns all memory (inputs, outputs, packed parameters,
 *     workspaces); the library allocates nothing persistent and frees nothing.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no default-stream use.
 *   - return value: 0 = OK, negative = argument/shape error (MNF_E_*), positive =
 *     cudaError_t.  mnf_last_error() returns a thread-local message.  Nothing throws
 *     or exits across the boundary.
 */
#ifndef MNF_B200_H
#define MNF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNF_ABI_VERSION 1

#define MNF_E_ARG (-1)     /* null pointer / bad enum / negative size        */
#define MNF_E_SHAPE (-2)   /* shape outside what the kernels support          */
#define MNF_E_ALIGN (-3)   /* pointer not aligned as documented               */
#define MNF_E_DEVICE (-4)  /* no sm_100 device / wrong architecture           */

int mnf_abi_version(void);
const char *mnf_last_error(void);
/* Fills SM count, max opt-in shared memory per block, compute capability major/minor of
 * the current device.  Any pointer may be NULL. */
int mnf_device_info(int *sm_count, int *smem_optin, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------
 * Flow stacks: NormalizingFlow.forward / .inverse (flows/core.py:17-35) over
 * AffineConstantFlow/ActNormFlow (affine_constant_flow.py:18-26), AffineHalfFlow
 * (affine_half_flow.py:44-66), Glow (glow.py:26-37), MAF/IAF (maf.py:39-62),
 * NSF_CL / NSF_AR (spline_flow.py:199-285).  One launch runs the whole stack: each
 * point stays in registers across all flows and its log-det is accumulated on chip.
 * ---------------------------------------------------------------------------------- */
enum mnf_flow_type {
    MNF_OP_AFFINE_CONST = 1, /* aux: s[D], t[D]                                         */
    MNF_OP_GLOW = 2,         /* aux: W[D*D], Winv[D*D], logdet[1] (mnf_glow_assemble)    */
    MNF_OP_AFFINE_HALF = 3,  /* net[0] = s_net, net[1] = t_net (MLP, LeakyReLU 0.2)      */
    MNF_OP_NSF_CL = 4,       /* net[0] = f1, net[1] = f2                                 */
    MNF_OP_NSF_AR = 5,       /* aux: init_param[3K-1]; net[0] = layers.0, rest follow    */
    MNF_OP_MADE = 6,         /* net[0] = MADE with masks folded in (ReLU); see flags     */
};

#define MNF_FLAG_PARITY 1u    /* AffineHalfFlow / MAF parity                              */
#define MNF_FLAG_SCALE 2u     /* AffineHalfFlow has s_net                                 */
#define MNF_FLAG_SHIFT 4u     /* AffineHalfFlow has t_net                                 */
#define MNF_FLAG_MADE_SEQ 8u  /* MADE op runs the D-pass sequential direction when the    */
                              /* stack runs forward (MAF) -- if clear, when it runs inverse (IAF) */

#define MNF_MAX_OPS 32
#define MNF_MAX_LIN 6
#define MNF_MAX_DIM 64
#define MNF_MAX_HIDDEN 128
#define MNF_MAX_BINS 32

/* One flow of the stack.  Offsets are in floats into the packed parameter blob.  A net is
 * stored as, per Linear layer l: weight[sizes[l+1]][sizes[l]] (torch's [out][in] order)
 * followed by bias[sizes[l+1]]. */
typedef struct mnf_flow_op {
    int32_t type;                   /* enum mnf_flow_type                               */
    uint32_t flags;                 /* MNF_FLAG_*                                       */
    int32_t K;                      /* spline bins (NSF_*)                              */
    float bound;                    /* spline tail bound B (NSF_*)                      */
    int32_t n_lin;                  /* Linear layers per conditioner net                */
    int32_t sizes[MNF_MAX_LIN + 1]; /* layer widths, sizes[0] = net input width         */
    int32_t net_off[2];             /* offsets of the conditioner nets                  */
    int32_t aux_off;                /* offset of the op's auxiliary parameters          */
    float edge_deriv;               /* NSF: min_deriv + softplus(log(exp(1-min_deriv)-1)) */
} mnf_flow_op;

/* Runs `n_ops` flows over `n_rows` points of dimension `dim`.
 *   ops_host     : host array of descriptors, in module order (flows[0] first).
 *   params       : device fp32 blob the offsets refer to, `n_params` floats.
 *   x            : input  [n_rows, dim];   y: output [n_rows, dim] (may alias x; NULL = do not
 *                  store the transformed points, e.g. when only log-probabilities are wanted).
 *   log_det      : output [n_rows] -- sum of the flows' log|det J| (core.py:19,23).
 *   base_log_prob: optional output [n_rows]: standard-normal log-density of y
 *                  (NormalizingFlowModel.base_log_prob, core.py:46-49, for a N(0,I) base).
 *   intermediates: optional output [n_ops, n_rows, dim]: the output of every flow in
 *                  execution order (core.py:20-25 returns them as a list).
 *   flags        : MNF_RUN_INVERSE = inverse direction (flows[n-1] first), else forward;
 *                  MNF_RUN_GENERIC forces the generic interpreter; MNF_RUN_VARIANT(v) picks a
 *                  code variant of the dim-2 kernel (tests / tuning; 0 = library default);
 *                  MNF_RUN_LOGPROB makes base_log_prob receive log_det + base log-density, i.e.
 *                  log p(x) of tests/test_flows.py:22-24 in one pass.
 * Replaces NormalizingFlow.forward/inverse (flows/core.py:17-35). */
int mnf_flow_stack_run(const mnf_flow_op *ops_host, int n_ops, const float *params,
                       int64_t n_params, const float *x, float *y, float *log_det,
                       float *base_log_prob, float *intermediates, int64_t n_rows, int dim,
                       int flags, void *stream);

#define MNF_RUN_INVERSE 1
#define MNF_RUN_GENERIC 2
#define MNF_RUN_LOGPROB 4
#define MNF_RUN_VARIANT_MASK 0x30
#define MNF_RUN_VARIANT(v) ((((v) + 1) << 4) & MNF_RUN_VARIANT_MASK) /* v in {0,1,2} */

/* Which kernel mnf_flow_stack_run would pick: 0 = generic interpreter, 1 = specialised
 * D=2 register-resident kernel.  Host-only, no launch. */
int mnf_flow_stack_plan(const mnf_flow_op *ops_host, int n_ops, int dim, int64_t n_params);

/* Glow._assemble_W + torch.inverse (glow.py:20-24, 34-35):  W = P (tril(L,-1)+I)(triu(U,1)+diag S),
 * W^-1 by triangular solves, logdet = sum log|S|.  out = [W (D*D) | Winv (D*D) | logdet (1)].
 * All pointers device; D <= MNF_MAX_DIM. */
int mnf_glow_assemble(const float *P, const float *L, const float *U, const float *S, float *out,
                      int dim, void *stream);

/* ActNormFlow data-dependent init (affine_constant_flow.py:44-49):
 * s = log(std(x, dim 0, unbiased)), t = mean(x * exp(s), dim 0).  `do_s` / `do_t` select which
 * are (re)computed (the reference skips an all-zero parameter).  workspace: 4*dim doubles,
 * zero-initialised by the call. */
int mnf_actnorm_init(const float *x, int64_t n_rows, int dim, float *s, float *t, int do_s,
                     int do_t, double *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MNF_B200_H */
