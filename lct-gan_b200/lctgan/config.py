"""Run-time switches of the lctgan kernels."""

#: Run the dense 1024->1024 convolution (MSD convs.5) on the tcgen05 tensor cores with bf16 operands and fp32
#: accumulation (BASELINE.json configs[2]: "LCT-GAN training bf16").  False = the fp32 SIMT kernels everywhere
#: (used by the tight fp32 parity tests).
dense_tensor_cores = True


def set_precision(mode: str) -> None:
    """"bf16": tensor-core path for the dense contraction (default); "fp32": fp32 SIMT kernels only."""
    global dense_tensor_cores
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"unknown precision mode {mode!r}")
    dense_tensor_cores = (mode == "bf16")
