"""seq2reg encoder layer (parameter layout of the reference's seq2reg/modules.py:129-147; forward signature :149)."""
import torch
import torch.nn as nn

from .._params import Affine, AttnBlock, MHAParams
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


class ContextFlashAttentionEncoderLayer(nn.Module):
    """seq2reg's context layer (seq2reg/modules.py:42-126), used when Seq2RegPredictor(use_context=True): self-MHA,
    cross-MHA against the window's label context (same padding mask), GeGLU FFN; residual = layer input."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, batch_first=True, use_alibi=False,
                 make_data_kv=False, mlp_dout=0.0):
        super().__init__()
        if make_data_kv:
            raise NotImplementedError("make_data_kv=True is not implemented")
        self.mixer = AttnBlock(d_model)
        self.crossMHA = AttnBlock(d_model, cross=True)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model); self.norm3 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.use_alibi, self.num_heads = use_alibi, nhead
        if use_alibi:
            self.register_buffer("m", alibi_slopes(nhead))
        self._folded = None

    @torch.no_grad()
    def forward(self, src, context, key_padding_mask=None, precision=torch.float32):
        """seq2reg/modules.py:76: src, context [batch, seqlen, d_model] of the same shape; key_padding_mask True = pad."""
        from .. import layer_ops as LO
        assert src.shape == context.shape, "src and context must have the same shape"
        if self._folded is None:
            self._folded = (LO._Cache(), LO.Workspace(src.device))
        cache, ws = self._folded
        L = cache.get(self, lambda sd, dev: LO.context_layer_weights(sd, "", dev))
        slopes = alibi_slopes(self.num_heads).to(src.device) if self.use_alibi else None
        return LO.context_layer_forward(L, ws, src.shape[-1], self.num_heads, slopes, src, context, key_padding_mask,
                                        key_padding_mask)
