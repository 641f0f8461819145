"""Masked autoregressive MLP (reference: layers/made.py:11-94; Germain et al. 2015)."""

from __future__ import annotations

import numpy as np
import torch
from torch import nn


class MaskedLinear(nn.Linear):
    """Dense layer whose weight is multiplied by a fixed 0/1 connectivity mask ``[n_in, n_out]``.
    Inside MAF/IAF the mask is folded into the weight when the flow program is packed."""

    def __init__(self, n_in, n_out, bias=True):
        super().__init__(n_in, n_out, bias)
        self.register_buffer("mask", torch.ones(n_in, n_out, dtype=torch.bool))

    def set_mask(self, mask: np.ndarray) -> None:
        # in place: bumps the buffer's version so cached flow programs re-fold the mask
        self.mask.copy_(torch.from_numpy(np.ascontiguousarray(mask)).to(self.mask.device, torch.bool))

    def forward(self, x):
        """x @ (W.T * mask) + b (made.py:22-23) on the exact-fp32 GEMM kernel (mnf_gemm_f32); differentiable through
        the hand-written dgrad / wgrad GEMMs when grad is enabled.  Inside MAF / IAF the whole MADE runs in the fused
        flow kernels instead (mask folded into the packed weights)."""
        from .. import _lib
        from . import _train

        lead = x.shape[:-1]
        x2 = _lib.require_cuda_f32(x, "input").reshape(-1, self.in_features)
        w = self.weight * self.mask.to(self.weight.dtype).T  # [out, in]; folding the mask is O(params), not O(batch)
        out = _train.LinearFn.apply(x2, w, self.bias)
        return out.reshape(*lead, self.out_features)


class MADE(nn.Sequential):
    """MaskedLinear / ReLU chain with autoregressive masks.

    Degrees m(k) follow the reference: inputs keep their natural order (or a permutation),
    hidden unit degrees are drawn with numpy's legacy ``RandomState(seed)`` from
    [min degree of previous layer, n_in - 1), hidden masks use ``<=``, the output mask ``<``
    and is tiled when ``n_out`` is a multiple of ``n_in`` (made.py:59-94)."""

    def __init__(self, n_in, hidden_sizes, n_out, num_masks=1, natural_ordering=False):
        if n_out % n_in:
            raise AssertionError("n_out must be integer multiple of n_in")
        widths = [n_in, *hidden_sizes, n_out]
        mods = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [MaskedLinear(a, b), nn.ReLU()]
        super().__init__(*mods[:-1])
        self.n_in, self.n_out, self.hidden_sizes = n_in, n_out, list(hidden_sizes)
        self.natural_ordering, self.num_masks = natural_ordering, num_masks
        self.seed = 0
        self.m = {}
        self.update_masks()

    def forward(self, x):
        """MaskedLinear / ReLU chain (made.py:43-57 builds it as an nn.Sequential); the ReLUs run in the library's
        elementwise kernel rather than ATen so that a standalone MADE(x) stays on the CUDA path end to end."""
        from . import _train

        h = x
        for m in self:
            h = m(h) if isinstance(m, MaskedLinear) else _train.ReluFn.apply(h)
        return h

    def update_masks(self):
        if self.m and self.num_masks == 1:
            return
        rng = np.random.RandomState(self.seed)
        self.seed = (self.seed + 1) % self.num_masks
        degrees = {-1: np.arange(self.n_in) if self.natural_ordering else rng.permutation(self.n_in)}
        for i, width in enumerate(self.hidden_sizes):
            degrees[i] = rng.randint(degrees[i - 1].min(), self.n_in - 1, size=width)
        self.m = degrees
        n_hidden = len(self.hidden_sizes)
        masks = [degrees[i - 1][:, None] <= degrees[i][None, :] for i in range(n_hidden)]
        out_mask = degrees[n_hidden - 1][:, None] < degrees[-1][None, :]
        masks.append(np.tile(out_mask, (1, self.n_out // self.n_in)))
        for layer, mask in zip((m for m in self if isinstance(m, MaskedLinear)), masks):
            layer.set_mask(mask)


# ---------------------------------------------------------------------------------------
# tensor-core density path for stacks of MAF flows (mnf_made_density_tc)
# ---------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402

from .. import _lib  # noqa: E402

MADE_MAX_HIDDEN = 4
TC_MIN_ELEMS = 1 << 18  # rows * dim below which the exact-fp32 interpreter is used under precision="auto"


class MadeLayer(C.Structure):
    """struct mnf_made_layer."""

    _fields_ = [
        ("n_hidden", C.c_int32), ("hidden", C.c_int32 * MADE_MAX_HIDDEN), ("parity", C.c_int32),
        ("w", C.c_void_p * MADE_MAX_HIDDEN), ("b", C.c_void_p * MADE_MAX_HIDDEN),
        ("w_out", C.c_void_p), ("b_out", C.c_void_p),
    ]


_lib.register({
    "mnf_made_workspace": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "mnf_made_density_tc": (C.c_int, [C.POINTER(MadeLayer), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "mnf_made_fused_image_bytes": (C.c_int64, [C.c_int]),
    "mnf_made_density_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
})


def _round_tf32(t):
    """Round to nearest TF32 (10 mantissa bits), ties away from zero like cvt.rna.tf32.f32."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def made_tc_eligible(flow, n_rows=None) -> bool:
    net = flow.net
    if not isinstance(net, MADE):
        return False
    d = flow.dim
    lin = [m for m in net if isinstance(m, MaskedLinear)]
    ok = d % 4 == 0 and 4 <= d <= 128 and 2 <= len(lin) <= MADE_MAX_HIDDEN + 1 and net.n_out == 2 * d
    return ok and all(m.out_features <= 256 for m in lin[:-1])


class MadeStackPlan:
    """Packed weights of a stack of MAF flows: mask folded, hidden widths zero-padded to multiples of 32
    (TMA rows of 128 B), output rows interleaved (s_0, t_0, s_1, t_1, ...), everything TF32-rounded."""

    def __init__(self, flows):
        self.flows = list(flows)
        self.key = None

    def _tensors(self):
        out = []
        for f in self.flows:
            out += list(f.net.parameters()) + list(f.net.buffers())
        return out

    def build(self, device):
        ts = self._tensors()
        from .._program import param_epoch

        # versions catch in-place updates, the data pointers of EVERY tensor catch re-assigned storage, the epoch
        # catches CUDA-graph replays of an optimizer step (torch_mnf.graphs); the list is re-read from the modules
        key = (str(device), tuple(t._version for t in ts), tuple(t.data_ptr() for t in ts), param_epoch())
        if key == self.key:
            return
        structs, keep, self.max_hidden = [], [], 32
        with torch.no_grad():
            for f in self.flows:
                lin = [m for m in f.net if isinstance(m, MaskedLinear)]
                d = f.dim
                st = MadeLayer()
                st.n_hidden, st.parity = len(lin) - 1, int(bool(f.parity))
                k_in = d
                for l, m in enumerate(lin[:-1]):
                    h = (m.out_features + 31) // 32 * 32
                    w = torch.zeros(h, k_in, device=device)
                    w[: m.out_features, : m.in_features] = (m.weight * m.mask.to(m.weight.dtype).T).to(device)
                    b = torch.zeros(h, device=device)
                    b[: m.out_features] = m.bias.to(device)
                    w = _round_tf32(w)
                    keep += [w, b]
                    st.hidden[l], st.w[l], st.b[l] = h, w.data_ptr(), b.data_ptr()
                    self.max_hidden = max(self.max_hidden, h)
                    k_in = h
                m = lin[-1]
                wo = torch.zeros(2 * d, k_in, device=device)
                full = (m.weight * m.mask.to(m.weight.dtype).T).to(device)  # rows: s_0..s_{d-1}, t_0..t_{d-1} (maf.py:57)
                wo[0::2, : m.in_features] = full[:d]
                wo[1::2, : m.in_features] = full[d:]
                bo = torch.empty(2 * d, device=device)
                bo[0::2], bo[1::2] = m.bias[:d].to(device), m.bias[d:].to(device)
                wo = _round_tf32(wo)
                keep += [wo, bo]
                st.w_out, st.b_out = wo.data_ptr(), bo.data_ptr()
                structs.append(st)
        self.arr = (MadeLayer * len(structs))(*structs)
        self.keep, self.key = keep, key


@torch.no_grad()
def made_density(plan: MadeStackPlan, x, want_inter=False):
    """Density direction of a MAF stack on the tensor cores -> (z, log_det, intermediates or None)."""
    x = _lib.require_cuda_f32(x, "input")
    dev = x.device
    B, D = x.shape
    plan.build(dev)
    lib = _lib.lib()
    n = len(plan.flows)
    z = torch.empty_like(x)
    ld = torch.empty(B, device=dev, dtype=torch.float32)
    inter = torch.empty((n, B, D), device=dev, dtype=torch.float32) if want_inter else None
    ws = torch.empty(lib.mnf_made_workspace(B, D, plan.max_hidden), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = lib.mnf_made_density_tc(plan.arr, n, x.data_ptr(), z.data_ptr(), ld.data_ptr(), _lib.ptr(inter), B, D,
                                     ws.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "mnf_made_density_tc")
    return z, ld, inter


# ---------------------------------------------------------------------------------------
# fused tcgen05 kernel for stacks of dim-64 MAF flows (mnf_made_density_fused, csrc/made_fused.cu)
# ---------------------------------------------------------------------------------------
FUSED_DIM, FUSED_HP, FUSED_MAX_FLOWS = 64, 32, 16
LOG2E = 1.4426950408889634
VARIANT = 0  # 0 = library default; 10 * tiles in flight + threads per row (21, 31, 41, 32) for tuning / tests


def made_fused_eligible(flow) -> bool:
    net = flow.net
    if not isinstance(net, MADE) or flow.dim != FUSED_DIM or net.n_out != 2 * FUSED_DIM:
        return False
    hs = list(net.hidden_sizes)
    return 1 <= len(hs) <= MADE_MAX_HIDDEN and all(1 <= h <= FUSED_HP - 1 for h in hs)


def _swizzled_image(W):
    """[N, K] (K <= 64) -> flat fp16 image: rows of 128 bytes (K padded to 64), K-major, 128-byte swizzle
    (see include/mnf_b200.h).  fp16 conversion rounds to nearest and saturates at +-65504."""
    N, K = W.shape
    n = torch.arange(N, device=W.device)[:, None]
    k = torch.arange(K, device=W.device)[None, :]
    off = (n // 8) * 512 + (n % 8) * 64 + (((k // 8) ^ (n % 8)) * 8) + k % 8
    img = torch.zeros(N * 64, device=W.device, dtype=torch.float16)
    img[off.reshape(-1)] = W.clamp(-65504.0, 65504.0).to(torch.float16).reshape(-1)
    return img


class FusedMadePlan:
    """Packed weight images of a stack of dim-64 MAF flows for mnf_made_density_fused, in execution order.
    chained=True: one launch runs the whole stack, the parity flips folded into the following flows' weights;
    chained=False: one launch per flow (every flow starts on an un-reversed row) -- used when the per-flow outputs
    are wanted (core.py:20-25 returns every intermediate)."""

    def __init__(self, flows, chained=True):
        self.flows, self.chained, self.key = list(flows), chained, None

    def _tensors(self):
        return [t for f in self.flows for t in (*f.net.parameters(), *f.net.buffers())]

    def build(self, device):
        from .._program import param_epoch

        ts = self._tensors()
        key = (str(device), tuple(t._version for t in ts), tuple(t.data_ptr() for t in ts), param_epoch())
        if key == self.key:
            return
        D, HP = FUSED_DIM, FUSED_HP
        nh = {len(f.net.hidden_sizes) for f in self.flows}
        if len(nh) != 1:
            raise ValueError("fused MADE kernel needs the same number of hidden layers in every flow")
        self.n_hidden = nh.pop()
        imgs, b1s, self.reversed_after = [], [], []
        rev = False
        flip = torch.arange(D - 1, -1, -1, device=device)
        with torch.no_grad():
            for f in self.flows:
                if not self.chained:
                    rev = False
                lin = [m for m in f.net if isinstance(m, MaskedLinear)]
                ws = [(m.weight * m.mask.to(m.weight.dtype).T).to(device, torch.float32) for m in lin]
                bs = [m.bias.to(device, torch.float32) for m in lin]
                h = [w.size(0) for w in ws[:-1]]
                w1 = torch.zeros(HP, D, device=device)
                w1[: h[0]] = ws[0][:, flip] if rev else ws[0]
                b1 = torch.zeros(HP, device=device)
                b1[: h[0]] = bs[0]
                parts = [_swizzled_image(w1)]
                for l in range(1, len(h)):
                    wp = torch.zeros(HP, HP, device=device)
                    wp[: h[l], : h[l - 1]] = ws[l]
                    wp[: h[l], HP - 1] = bs[l]
                    wp[HP - 1, HP - 1] = 1.0  # keeps the constant-one column alive
                    parts.append(_swizzled_image(wp))
                wo, bo = ws[-1], bs[-1]  # rows: s_0..s_{D-1}, t_0..t_{D-1} (maf.py:57)
                s_rows, t_rows, s_b, t_b = wo[:D], wo[D:], bo[:D], bo[D:]
                if rev:
                    s_rows, t_rows, s_b, t_b = s_rows[flip], t_rows[flip], s_b[flip], t_b[flip]
                wop = torch.zeros(2 * D, HP, device=device)
                # the s rows carry log2(e): the kernel evaluates exp(s) as ex2(s') and rescales the log-det sum by ln 2
                wop[0::2, : h[-1]], wop[1::2, : h[-1]] = s_rows * LOG2E, t_rows
                wop[0::2, HP - 1], wop[1::2, HP - 1] = s_b * LOG2E, t_b
                parts.append(_swizzled_image(wop))
                imgs.append(torch.cat(parts))
                b1s.append(b1)
                rev ^= bool(f.parity)  # maf.py:60: z.flip(dims=[1]) after the transform
                self.reversed_after.append(rev)
        self.images = torch.stack(imgs).contiguous()
        self.b1 = torch.stack(b1s).contiguous()
        expect = _lib.lib().mnf_made_fused_image_bytes(self.n_hidden)
        if self.images.size(1) * 2 != expect:
            raise RuntimeError(f"weight image has {self.images.size(1) * 2} bytes, the library expects {expect}")
        self.key = key


@torch.no_grad()
def made_density_fused(plans, x, want_inter=False, want_z=True, want_log_prob=False, log_prob_out=None):
    """Density direction of a dim-64 MAF stack through the fused tcgen05 kernel.
    plans = (chained plan, per-flow plan).  -> (z or None, log_det, intermediates or None, log_prob or None)."""
    x = _lib.require_cuda_f32(x, "input")
    dev = x.device
    B, D = x.shape
    lib = _lib.lib()
    stream = _lib.stream_ptr(dev)
    ld = torch.empty(B, device=dev, dtype=torch.float32)

    def launch(plan, lo, hi, src, dst, ld_out, lp_out, rev_before):
        # flows [lo, hi) of the plan; rev_before: orientation the packer assumed at flow lo (chained plans only)
        rc = lib.mnf_made_density_fused(plan.images[lo:hi].data_ptr(), plan.b1[lo:hi].data_ptr(), hi - lo, plan.n_hidden,
                                        int(plan.reversed_after[hi - 1]), src.data_ptr(), _lib.ptr(dst), _lib.ptr(ld_out),
                                        _lib.ptr(lp_out), B, D, VARIANT, stream)
        _lib.check(rc, "mnf_made_density_fused")
        _lib.launch_count += 1

    with _lib.on_device(dev):
        if want_inter:
            plan = plans[1]
            plan.build(dev)
            n = len(plan.flows)
            inter = torch.empty((n, B, D), device=dev, dtype=torch.float32)
            tmp = torch.empty(B, device=dev, dtype=torch.float32)
            src = x
            for i in range(n):
                launch(plan, i, i + 1, src, inter[i], ld if i == 0 else tmp, None, False)
                if i:
                    ld += tmp
                src = inter[i]
            lp = None
            if want_log_prob:
                zf = inter[-1]
                lp = ld - 0.5 * zf.square().sum(1) - 0.5 * D * 1.8378770664093453
            return inter[-1], ld, inter, lp
        plan = plans[0]
        plan.build(dev)
        n = len(plan.flows)
        if n <= FUSED_MAX_FLOWS:
            z = torch.empty_like(x) if want_z else None
            lp = None
            if want_log_prob:
                lp = log_prob_out if log_prob_out is not None else torch.empty(B, device=dev, dtype=torch.float32)
            launch(plan, 0, n, x, z, ld, lp, False)
            return z, ld, None, lp
    # longer stacks: chunks of <= 16 flows, each chunk packed as its own chained plan (orientation restarts per chunk)
    raise NotImplementedError(f"fused MADE kernel: stacks of more than {FUSED_MAX_FLOWS} flows are not packed yet")
