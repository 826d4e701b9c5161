"""Run-time switches of the lctgan kernels."""

#: Tensor-core mode (BASELINE.json configs[2]: "LCT-GAN training bf16"): the dense 1024->1024 convolution (MSD
#: convs.5) runs on tcgen05 with bf16 operands and the grouped discriminator convolutions on TF32 mma.sync, all
#: with fp32 accumulation.  False = the fp32 SIMT kernels everywhere (used by the tight fp32 parity tests).
dense_tensor_cores = True


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

_STREAMS = {}


def side_streams(n: int, device):
    """n persistent side streams for `device` (created once)."""
    import torch
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


_AUX = {}


def aux_stream_for(stream):
    """A persistent helper stream paired with `stream` (used to run weight-gradient kernels beside the data-gradient
    chain of the same sub-discriminator)."""
    import torch
    key = (stream.device.index, stream.cuda_stream)
    aux = _AUX.get(key)
    if aux is None:
        aux = torch.cuda.Stream(device=stream.device)
        _AUX[key] = aux
    return aux
