"""seq2reg encoder layer (parameter layout of the reference's seq2reg/modules.py:129-147; forward signature :149)."""
import torch
import torch.nn as nn

from .._params import Affine, MHAParams
from ..utils.alibi import alibi_slopes

FFN_HIDDEN = 2048


class FlashTransformerLayer(nn.Module):
    """LN1 -> self-MHA -> +src -> LN2 -> GeGLU FFN -> +src(layer input).  The batched path runs it inside
    Engine.seq2reg; `forward` serves callers of the reference's layer-level API through the same kernels."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, use_alibi=False, mlp_dout=0.1):
        super().__init__()
        self.MHA = MHAParams(d_model)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.nhead, self.use_alibi = nhead, use_alibi
        self._folded = None

    @torch.no_grad()
    def forward(self, src, src_key_padding_mask=None, precision=torch.float32):
        """src [batch, seqlen, d_model]; src_key_padding_mask bool [batch, seqlen], True = padding (seq2reg/modules.py:149).
        `precision` is accepted for signature parity: the kernels compute in bf16 with fp32 accumulation."""
        from .. import layer_ops as LO
        if self._folded is None:
            self._folded = (LO._Cache(), LO.Workspace(src.device))
        cache, ws = self._folded
        L = cache.get(self, lambda sd, dev: LO.seq2reg_layer_weights(sd, "", dev))
        slopes = alibi_slopes(self.nhead).to(src.device) if self.use_alibi else None
        return LO.seq2reg_layer_forward(L, ws, self.nhead, slopes, src, src_key_padding_mask)
