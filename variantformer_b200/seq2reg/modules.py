"""seq2reg encoder layer (parameter layout of the reference's seq2reg/modules.py:129-147)."""
import torch.nn as nn

from .._params import Affine, MHAParams

FFN_HIDDEN = 2048


class FlashTransformerLayer(nn.Module):
    """LN1 -> self-MHA -> +src -> LN2 -> GeGLU FFN -> +src(layer input); executed by Engine.seq2reg."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, use_alibi=False, mlp_dout=0.1):
        super().__init__()
        self.MHA = MHAParams(d_model)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.nhead, self.use_alibi = nhead, use_alibi
