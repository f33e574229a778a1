"""Seq2RegPredictor — CRE / gene-window encoder behind the reference's signature
(seq2reg/model.py:40-64 ctor, :193-201 forward).  Token embedding (+ sinusoidal PE or ALiBi), N transformer
layers, masked mean-pool; computed by the sm_100a kernels through Engine.seq2reg."""
import numpy as np
import torch
import torch.nn as nn

from .._params import Affine, Table
from .modules import ContextFlashAttentionEncoderLayer, FlashTransformerLayer


class _HParams(dict):
    __getattr__ = dict.get


class Seq2RegPredictor(nn.Module):
    def __init__(self, vocab_size: int, embedding_dim: int, num_heads: int, num_layers: int, num_tissues: int,
                 num_classes: int, learning_rate: float = 1e-4, loss_fn=("cross_entropy", 0), seq_pool: str = "mean",
                 cre_type: str = "multi", token_length: int = None, use_context: bool = False,
                 positional_encoding: str = "sinusoidal", use_flash: bool = False, majority_weight: float = None,
                 weight_decay: float = 0.0, lr_scale: float = 1.0, strand_agg: str = "mean",
                 expand_context: bool = False, mlp_dout: float = 0.1, tissues: list = None, **kwargs):
        super().__init__()
        assert use_flash, "Only Flash is supported"                       # seq2reg/model.py:72
        assert positional_encoding in ("sinusoidal", "alibi"), "Position encoding must be either 'sinusoidal' or 'alibi'"
        if seq_pool != "mean":
            raise NotImplementedError("only seq_pool='mean' is implemented")
        hp = dict(vocab_size=vocab_size, embedding_dim=embedding_dim, num_heads=num_heads, num_layers=num_layers,
                  num_tissues=num_tissues, num_classes=num_classes, learning_rate=learning_rate, loss_fn=loss_fn,
                  seq_pool=seq_pool, cre_type=cre_type, token_length=token_length, use_context=use_context,
                  positional_encoding=positional_encoding, use_flash=use_flash, strand_agg=strand_agg,
                  mlp_dout=mlp_dout, tissues=tissues, **kwargs)
        self.hparams = _HParams(hp)
        self.token_embedding = Table(vocab_size, embedding_dim)
        self.pos_encoding_type = positional_encoding
        use_alibi = positional_encoding == "alibi"
        if use_context:
            # seq2reg/model.py:93-121: label-context variant — a learned embedding per reference cCRE class, optionally
            # expanded along the token axis by Linear(1, token_length); layers with a cross-attention to that context.
            # Not the configured tokenizers (vf_model.yaml): runs layer by layer through the layers' own forwards.
            from ..utils.constants import REF_CREs
            self.context_embedding = Table(len(REF_CREs), embedding_dim)
            self.expand_context_type = expand_context
            if expand_context:
                self.expand_context = Affine(token_length, 1)
            self.transformer_encoder = nn.ModuleList(
                [ContextFlashAttentionEncoderLayer(d_model=embedding_dim, nhead=num_heads, batch_first=True,
                                                   use_alibi=use_alibi, mlp_dout=mlp_dout) for _ in range(num_layers)])
        else:
            self.transformer_encoder = nn.ModuleList(
                [FlashTransformerLayer(embedding_dim, num_heads, use_alibi=use_alibi) for _ in range(num_layers)])
        out_dim = embedding_dim * 2 if strand_agg == "concat" else embedding_dim
        self.tissue_classifiers = nn.ModuleDict({str(t): Affine(num_classes, out_dim) for t in range(num_tissues)})
        self.use_context = use_context
        self._engine = None

    # -- engine plumbing ---------------------------------------------------------------------------
    def _weights(self):
        from ..engine import Seq2RegWeights, Workspace
        dev = self.token_embedding.weight.device
        if self._engine is None or self._engine[0] != dev:
            if dev.type != "cuda":
                raise RuntimeError("Seq2RegPredictor runs on a B200 only: move the module to CUDA first (no CPU path)")
            sd = {"t." + k: v for k, v in self.state_dict().items()}
            self._engine = (dev, Seq2RegWeights(sd, "t.", dict(self.hparams), dev), Workspace(dev))
        return self._engine[1], self._engine[2]

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x, padding_mask, tissue_vector, context=None, only_embed=False, precision=torch.float32):
        """x int64 [batch, strands, L]; padding_mask bool (True = pad) -> [batch, strands, embedding_dim].
        Only `only_embed=True` (the inference hot path, model_combined_modulator.py:776-783) is implemented."""
        if not only_embed:
            raise NotImplementedError("tissue-classifier logits are a training-time output, outside the hot path")
        if self.use_context:
            return self._forward_with_context(x, padding_mask, context)
        from .. import ops
        from ..engine import AttnPlan, Engine
        W, ws = self._weights()
        b, s, L = x.shape
        dev = self.token_embedding.weight.device
        tok = x.reshape(b * s, L).to(device=dev, dtype=torch.int32).contiguous()
        msk = padding_mask.reshape(b * s, L).to(device=dev, dtype=torch.uint8).contiguous()
        lens = ops.window_lengths(msk).cpu().numpy().astype(np.int64)
        eng = Engine.__new__(Engine); eng.device = dev; eng.ws = ws
        pooled = eng.seq2reg(W, tok, msk, lens, ops.cu_seqlens(lens, dev), AttnPlan(lens, dev, W.hd))
        return pooled.float().view(b, s, -1)

    @torch.no_grad()
    def _forward_with_context(self, x, padding_mask, context):
        """use_context=True (seq2reg/model.py:222-250): context int [batch] reference-cCRE class of every window."""
        from .. import ops
        from ..engine import sinusoidal_pe
        assert context is not None, "context (reference cCRE class per window) is required when use_context is True"
        b, s, L = x.shape
        dev = self.token_embedding.weight.device
        if dev.type != "cuda":
            raise RuntimeError("Seq2RegPredictor runs on a B200 only: move the module to CUDA first (no CPU path)")
        tok = x.reshape(b * s, L).to(dev).long()
        mask = padding_mask.reshape(b * s, L).to(dev).bool()
        h = self.token_embedding.weight[tok]
        if self.pos_encoding_type == "sinusoidal":
            h = h + sinusoidal_pe(h.shape[-1], L).to(dev)
        ctx = self.context_embedding.weight[context.to(dev).long()]                    # [b, d]
        ctx = ctx[:, None, :].expand(b, s, -1).reshape(b * s, 1, -1)
        if self.expand_context_type:                                                     # Linear(1, token_length) per element
            ctx = ctx * self.expand_context.weight.reshape(1, L, 1) + self.expand_context.bias.reshape(1, L, 1)
        else:
            ctx = ctx.expand(b * s, L, -1)
        ctx = ctx.contiguous()
        for layer in self.transformer_encoder:
            h = layer(h, ctx, key_padding_mask=mask)
        lens = (~mask).sum(1)
        cu = ops.cu_seqlens(lens.cpu().numpy(), dev)
        pooled = ops.masked_meanpool(h[~mask].float().contiguous(), cu, b * s, want_f32=True)[1]
        return pooled.view(b, s, -1)
