"""Small host helpers mirroring the hot-path parts of the reference's utils/functions.py."""
import torch

_COMP = str.maketrans("ACGTRYKMBVDHacgtrykmbvdh", "TGCAYRMKVBHDtgcayrmkvbhd")


def precision2dtype(precision_str: str) -> torch.dtype:
    """Lightning precision string -> dtype (utils/functions.py:12-32)."""
    p = precision_str.lower().strip()
    if "bf16" in p:
        return torch.bfloat16
    if "16" in p:
        return torch.float16
    if "32" in p:
        return torch.float32
    raise ValueError(f"Unknown precision string: {precision_str}")


def reverse_complement(sequence: str) -> str:
    """Host-side convenience (utils/functions.py:129-172); the batched path does this inside vf_encode_windows."""
    return sequence[::-1].translate(_COMP)
