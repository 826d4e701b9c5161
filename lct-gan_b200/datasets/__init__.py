"""Drop-in `datasets` package for the LCT-GAN hot path (B200 build).

Only the signal front end lives here (`datasets.stft`, `datasets.tf_features`).  The reference's
file-I/O dataset (`datasets/datasets.py`: LCTScpDataset, collate_fn - torchaudio loading) is out of
scope of this build; when the reference checkout is reachable (``$LCT_REF`` or /root/reference) the
two names are forwarded to it so that the reference's train.py / infer.py import unchanged.
"""
import importlib.util
import os
import sys

_FORWARDED = ("LCTScpDataset", "collate_fn")


def _load_reference_datasets():
    for root in (os.environ.get("LCT_REF"), "/root/reference"):
        if not root:
            continue
        path = os.path.join(root, "datasets", "datasets.py")
        if os.path.exists(path):
            spec = importlib.util.spec_from_file_location("datasets.datasets", path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules["datasets.datasets"] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("datasets.LCTScpDataset / collate_fn are file-I/O components of the reference and are not part "
                      "of the B200 hot-path build; set LCT_REF to a reference checkout to forward them")


def __getattr__(name):
    if name in _FORWARDED:
        return getattr(_load_reference_datasets(), name)
    raise AttributeError(name)
