"""ALiBi slopes (seq2gene/modules/layers.py:15-37, seq2reg/modules.py:13-33; same formula flash_attn's MHA uses)."""
import math

import torch


def alibi_slopes(n: int) -> torch.Tensor:
    def pow2(m):
        start = 2 ** (-(2 ** -(math.log2(m) - 3)))
        return [start * start ** i for i in range(m)]

    if math.log2(n).is_integer():
        return torch.tensor(pow2(n), dtype=torch.float32)
    c = 2 ** math.floor(math.log2(n))
    return torch.tensor(pow2(c) + alibi_slopes(2 * c)[0::2][: n - c].tolist(), dtype=torch.float32)
