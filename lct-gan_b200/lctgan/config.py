"""Run-time switches of the lctgan kernels."""

#: Tensor-core mode (BASELINE.json configs[2]: "LCT-GAN training bf16"): the dense 1024->1024 convolution (MSD
#: convs.5) runs on tcgen05 with bf16 operands and the grouped discriminator convolutions on TF32 mma.sync, all
#: with fp32 accumulation.  False = the fp32 SIMT kernels everywhere (used by the tight fp32 parity tests).
dense_tensor_cores = True


#: Tensor-core mode only: the grouped / first-layer discriminator convolutions on tcgen05 (conv_tc.cu: UMMA kind::tf32,
#: operands read from shared memory through descriptors, accumulators in TMEM).  False = the round-1 TF32 mma.sync
#: kernels (conv_mma.cu), kept for A/B measurements.
grouped_conv_tcgen05 = True
#: ... including the data gradient of the period discriminators' grouped layers (P > 1: staged epilogue)
tc_dgrad_periods = False


def set_precision(mode: str) -> None:
    """"bf16": tensor-core paths (default): tcgen05 bf16 dense contraction, TF32 grouped discriminator convolutions,
    3xTF32 (error-compensated, ~fp32 accuracy) generator GEMMs / convolutions; "fp32": fp32 SIMT kernels only."""
    global dense_tensor_cores, gconv_tensor_cores
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"unknown precision mode {mode!r}")
    dense_tensor_cores = (mode == "bf16")
    gconv_tensor_cores = dense_tensor_cores
    from ._lib import call_ret
    call_ret("lct_set_rowgemm", int(dense_tensor_cores))


#: Generator encoder / decoder convolutions as implicit row GEMMs on the tensor cores (3xTF32 error-compensated, fp32-level
#: accuracy, rowgemm.cu).  False = the SIMT kernels of gen_conv.cu.
gconv_tensor_cores = True

#: Run the independent sub-discriminators (5 periods, 3 scales) on parallel CUDA streams (forked from and joined
#: to the caller's stream; autograd replays the same streams in backward).  Each sub-discriminator is a chain of
#: kernels that fill at most one wave of the 148 SMs at batch 8, so the chains are overlapped; inside a captured
#: CUDA graph they become parallel branches.
concurrent_discriminators = True

#: CUDA stream priorities (lower = more urgent; kernel nodes of a captured graph inherit them).  The serial chains that
#: the step waits for - the generator (priority_generator) and the sub-discriminators' forward / data-gradient chains
#: (chain_priority) - outrank the helper streams that carry weight-gradient kernels (priority 0), so a wide
#: weight-gradient grid cannot park its CTAs in front of a critical-path kernel.
chain_priority = -1
priority_generator = -2

_STREAMS = {}
_GEN_STREAM = {}


def generator_stream(device):
    """The high-priority stream the generator's forward / backward kernels run on."""
    import torch
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _GEN_STREAM.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=priority_generator)
        _GEN_STREAM[key] = st
    return st


_GEN_SIDE = {}


def generator_side_stream(device):
    """The stream of the generator's parameter-gradient kernels (weight-gradient GEMMs, bias column sums): below the
    generator's data-gradient chain, ABOVE the helper streams (priority 0) that carry the discriminators' weight gradients -
    at equal priority the G step's dead discriminator gradients were served first and the generator's own parameter
    gradients finished 0.25 ms after its backward chain, alone on the GPU (CUPTI timeline)."""
    import torch
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _GEN_SIDE.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=chain_priority)
        _GEN_SIDE[key] = st
    return st


def side_streams(n: int, device):
    """n persistent side streams for `device` (created once)."""
    import torch
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device, priority=chain_priority))
    return pool[:n]


_AUX = {}


def aux_stream_for(stream):
    """A persistent helper stream paired with `stream` (used to run weight-gradient kernels beside the data-gradient
    chain of the same sub-discriminator)."""
    import torch
    key = (stream.device.index, stream.cuda_stream)
    aux = _AUX.get(key)
    if aux is None:
        aux = torch.cuda.Stream(device=stream.device, priority=0)
        _AUX[key] = aux
    return aux


#: G step only (a sub-discriminator whose INPUT requires grad): do not run the discriminator parameter gradients - weight
#: gradient kernels, weight-norm backward and the accumulation into ``.grad`` - inside the sub-discriminator's backward.
#: train.py computes these gradients and never reads them (the next ``d_opt.zero_grad`` discards them), so nothing on the
#: critical path waits for them.  The backward hands them over as one closure per sub-discriminator (``defer``); they are
#: launched on low-priority helper streams when the generator's backward begins (``launch_deferred_param_grads``, called by
#: lctgan.gen_impl) - its chain of small kernels leaves most SMs idle - and joined at the end of the G phase
#: (``join_deferred_param_grads``, called by lctgan.training).  Launched earlier, as soon as their inputs exist, the wide
#: weight-gradient grids sat in front of the loss-gradient / iSTFT-backward kernels the generator's backward waits for
#: (CUPTI timeline, profiles/timeline_r2_*.txt: 0.7 ms).  Values of every ``.grad`` after the join are unchanged.  Off by
#: default because the gradients are only valid after the join; lctgan.training.StepArgs.defer_dead_d_grads turns it on.
defer_dead_param_grads = False
_LATE = []        # closures not launched yet
_PENDING = []     # (helper stream, closure) launched, not joined yet
_LATE_STREAMS = {}
late_param_grad_streams = 3
late_param_grad_ctas = 1


def defer(fn) -> None:
    """`fn()` enqueues, on the current stream, everything one sub-discriminator owes its parameters."""
    _LATE.append(fn)


def launch_deferred_param_grads() -> None:
    """Start the deferred parameter-gradient work behind the current point of the current stream."""
    import torch
    if not _LATE:
        return
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    pool = _LATE_STREAMS.setdefault(key, [])
    while len(pool) < late_param_grad_streams:
        pool.append(torch.cuda.Stream(device=cur.device, priority=0))
    for st in pool:
        st.wait_stream(cur)
    from ._lib import call_ret
    call_ret("lct_set_wgrad_ctas", late_param_grad_ctas)     # narrow grids: the generator's kernels need the other half of an SM
    try:
        for i, fn in enumerate(_LATE):
            st = pool[i % len(pool)]
            with torch.cuda.stream(st):
                fn()
            _PENDING.append((st, fn))
    finally:
        call_ret("lct_set_wgrad_ctas", 0)
        _LATE.clear()


#: Data parallel (lctgan.parallel): called with the flat buffer that holds ALL parameter gradients of a sub-discriminator
#: the moment its backward has produced them (D step only; lctgan.training sets and clears it around d_loss.backward()).
stack_grad_hook = None


def join_deferred_param_grads() -> None:
    """Order the current stream after every deferred parameter-gradient chain and release the tensors kept for them."""
    import torch
    launch_deferred_param_grads()
    if not _PENDING:
        return
    cur = torch.cuda.current_stream()
    for st in {id(st): st for st, _ in _PENDING}.values():
        cur.wait_stream(st)
    _PENDING.clear()
