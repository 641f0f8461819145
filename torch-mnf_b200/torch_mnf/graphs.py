"""CUDA-graph capture of a whole training step (additive helper, not part of the reference API).

The differentiable paths of the drop-in modules are chains of small launches -- an MNF-LeNet step with the reference's
loss (nll + 1e-3 * kl_div) is ~600 kernels of a few microseconds each, so at the reference's batch size of 32 the
step is bound by launch overhead on the host, not by the GPU.  Every launch goes to torch's current stream, allocates
through torch's caching allocator and draws noise from torch's graph-safe generator, so forward + backward +
optimizer step can be captured once and replayed."""

from __future__ import annotations

import torch

from ._program import bump_param_epoch


def graphed_training_step(model, loss_fn, optimizer, example_inputs, warmup: int = 3):
    """Returns ``step(*inputs) -> loss`` that replays one captured ``loss_fn(model, *inputs).backward();
    optimizer.step()``.  ``example_inputs``: CUDA tensors with the shapes / dtypes every later call will use (they are
    trained on during warm-up).  The optimizer must be capturable (e.g. ``torch.optim.Adam(..., capturable=True)``).
    Data-dependent initialisation (ActNormFlow) has to be done before calling this."""
    static_inputs = tuple(t.clone() for t in example_inputs)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):  # lazy initialisation (kernel attributes, optimizer state) happens outside the capture
            optimizer.zero_grad(set_to_none=True)
            loss_fn(model, *static_inputs).backward()
            optimizer.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    optimizer.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph):
        static_loss = loss_fn(model, *static_inputs)
        static_loss.backward()
        optimizer.step()

    def step(*inputs):
        for dst, src in zip(static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        bump_param_epoch()  # the replay updated parameters without touching their version counters: drop packed blobs
        return static_loss.detach()

    step.graph = graph
    return step


def graphed_inference(fn, example_inputs, warmup: int = 3):
    """Returns ``call(*inputs) -> outputs`` replaying one captured ``fn(*inputs)`` under ``torch.no_grad()``.
    Useful where a call is a chain of small launches (e.g. MNF layers at small batch sizes); a single fused flow
    kernel gains nothing (BASELINE config 1 measures 80 us per call either way: the time is inside the kernel).
    The outputs are static buffers that the next call overwrites.  Stochastic modules (MNF layers, RNVP) draw their
    noise from torch's graph-safe CUDA generator while a capture is in progress (layers/_mnf_ops.py::Noise), so every
    replay sees fresh z, masks and eps; the in-kernel Philox path takes its seed by value and is not used under capture."""
    static_inputs = tuple(t.clone() for t in example_inputs)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.no_grad(), torch.cuda.stream(side):
        for _ in range(warmup):
            fn(*static_inputs)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(graph):
        static_out = fn(*static_inputs)

    def call(*inputs):
        for dst, src in zip(static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return static_out

    call.graph = graph
    return call
