"""seq2gene layers (parameter layout of the reference's seq2gene/modules/layers.py:47-86, :502-524, :1012-1111; forward
signatures :88-98, :1113).  The batched path executes them inside Engine._layer / Engine.run; the `forward` methods
serve callers of the reference's layer-level API through the same kernels (variantformer_b200/layer_ops.py)."""
import torch
import torch.nn as nn

from ..._params import Affine, AttnBlock, Table
from ...utils.alibi import alibi_slopes

FFN_HIDDEN = 2048


class ContextFlashAttentionEncoderLayer(nn.Module):
    """LN1 -> self-MHA(+ALiBi) -> +src -> LN2 -> cross-MHA(context) -> + -> LN3 -> GeGLU FFN -> +src(layer input)."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, batch_first=True, use_alibi=False,
                 make_data_kv=False, mlp_dout=0.0, cross_alibi=False, flash_attn_3=False):
        super().__init__()
        if make_data_kv or flash_attn_3:
            raise NotImplementedError("make_data_kv / flash_attn_3 variants are not implemented")
        self.cross_alibi = cross_alibi
        self.mixer = AttnBlock(d_model)
        self.crossMHA = AttnBlock(d_model, cross=True)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model); self.norm3 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.use_alibi, self.num_heads = use_alibi, nhead
        if use_alibi:
            self.register_buffer("m", alibi_slopes(nhead))          # persistent, like the reference (unused in forward)
        self._folded = None

    @torch.no_grad()
    def forward(self, src, context, src_key_padding_mask=None, context_padding_mask=None, precision=torch.float32,
                unpad_info=None, context_unpad_info=None, gene_unpad_info=None):
        """layers.py:88-98.  Padded mode: src [B, S, D], context [B, Sc, D], masks True = padding.  Unpadded mode:
        src [rows, D] / context [rows_c, D] with `unpad_info` (or `gene_unpad_info`) / `context_unpad_info` dicts holding
        `cu_seqlens`.  `precision` is accepted for signature parity (bf16 operands, fp32 accumulation)."""
        from ... import layer_ops as LO
        if self._folded is None:
            self._folded = (LO._Cache(), LO.Workspace(src.device))
        cache, ws = self._folded
        L = cache.get(self, lambda sd, dev: LO.context_layer_weights(sd, "", dev))
        D = self.norm1.weight.shape[0]
        slopes = alibi_slopes(self.num_heads).to(src.device) if self.use_alibi else None
        return LO.context_layer_forward(L, ws, D, self.num_heads, slopes, src, context, src_key_padding_mask,
                                        context_padding_mask, unpad_info, context_unpad_info, gene_unpad_info,
                                        cross_slopes=alibi_slopes(self.num_heads).to(src.device) if self.cross_alibi else None)


class FlashAttentionEncoderLayer(nn.Module):
    """LN1 -> self-MHA(+ALiBi) -> +src -> LN2 -> GeGLU FFN -> +src(layer input): the CRE layer of a model built with
    use_context=False (layers.py:166-228; norm3 exists in the reference's state_dict but its forward never reads it)."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, batch_first=True, use_alibi=False,
                 make_data_kv=False, mlp_dout=0.0):
        super().__init__()
        self.mixer = AttnBlock(d_model)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model); self.norm3 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.use_alibi, self.num_heads = use_alibi, nhead
        if use_alibi:
            self.register_buffer("m", alibi_slopes(nhead))
        self._folded = None

    @torch.no_grad()
    def forward(self, src, src_key_padding_mask=None, precision=torch.float32, unpad_info=None):
        """layers.py:197-228: padded [B, S, D] + mask (True = padding) or unpadded [rows, D] + unpad_info."""
        from ... import layer_ops as LO
        if self._folded is None:
            self._folded = (LO._Cache(), LO.Workspace(src.device))
        cache, ws = self._folded
        L = cache.get(self, lambda sd, dev: LO.seq2reg_layer_weights(sd, "", dev, mha="mixer.MHA."))
        slopes = alibi_slopes(self.num_heads).to(src.device) if self.use_alibi else None
        return LO.seq2reg_layer_forward(L, ws, self.num_heads, slopes, src, src_key_padding_mask, unpad_info)


class ContextFlashCrossAttentionEncoderLayer(nn.Module):
    """LN1 -> cross-MHA(context) -> +src -> LN2 -> GeGLU FFN -> +src(layer input): the gene layer of a model built with
    only_cross_attention=True (layers.py:231-325)."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, batch_first=True, use_alibi=False,
                 make_data_kv=False, mlp_dout=0.0, cross_alibi=False, flash_attn_3=False):
        super().__init__()
        if make_data_kv or flash_attn_3:
            raise NotImplementedError("make_data_kv / flash_attn_3 variants are not implemented")
        self.crossMHA = AttnBlock(d_model, cross=True)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.use_alibi, self.cross_alibi, self.num_heads = use_alibi, cross_alibi, nhead
        if use_alibi:
            self.register_buffer("m", alibi_slopes(nhead))
        self._folded = None

    @torch.no_grad()
    def forward(self, src, context, context_padding_mask=None, src_key_padding_mask=None, precision=torch.float32,
                gene_unpad_info=None, context_unpad_info=None):
        """layers.py:268-277 (note the reference's argument order: context_padding_mask before src_key_padding_mask)."""
        from ... import layer_ops as LO
        if self._folded is None:
            self._folded = (LO._Cache(), LO.Workspace(src.device))
        cache, ws = self._folded
        L = cache.get(self, lambda sd, dev: LO.cross_layer_weights(sd, "", dev))
        D = self.norm1.weight.shape[0]
        slopes = alibi_slopes(self.num_heads).to(src.device) if self.cross_alibi else None
        return LO.cross_layer_forward(L, ws, D, self.num_heads, slopes, src, context, src_key_padding_mask,
                                      context_padding_mask, gene_unpad_info, context_unpad_info)


class StartToken(nn.Module):
    """One learned start token prepended to every sequence (layers.py:491-499)."""

    def __init__(self, emb_dim):
        super().__init__()
        self.start_token = nn.Parameter(torch.empty(1, 1, emb_dim), requires_grad=False)

    def forward(self, x):
        return self.start_token.expand(x.size(0), 1, x.size(2)).clone()


class AddContext(nn.Module):
    """Tissue embedding ADDED to every position (layers.py:558-576)."""

    def __init__(self, num_tissues, emb_dim):
        super().__init__()
        self.num_registry_tokens = num_tissues
        self.registry_tokens = Table(num_tissues, emb_dim)

    def forward(self, x, tissue_vector):
        t = torch.as_tensor(tissue_vector, device=x.device).reshape(x.size(0), -1)[:, 0].long()
        return x + self.registry_tokens.weight[t][:, None, :]

    def get_registry_tokens(self):
        return self.registry_tokens.weight


class ConcatTissueContext(nn.Module):
    """Tissue embedding prepended as a token, padding mask extended by one valid position (layers.py:527-555)."""

    def __init__(self, num_tissues, emb_dim):
        super().__init__()
        self.num_registry_tokens = num_tissues
        self.registry_tokens = Table(num_tissues, emb_dim)

    def forward(self, x, tissue_vector, padding_mask):
        t = torch.as_tensor(tissue_vector, device=x.device).reshape(x.size(0), -1)[:, 0].long()
        combined = torch.cat((self.registry_tokens.weight[t][:, None, :], x), dim=1)
        start = torch.zeros((padding_mask.size(0), 1), dtype=padding_mask.dtype, device=padding_mask.device)
        return combined, torch.cat((start, padding_mask), dim=1)

    def get_registry_tokens(self):
        return self.registry_tokens.weight


class MultiRegistry(nn.Module):
    """One learned registry token per tissue, prepended to the gene stream (layers.py:502-524)."""

    def __init__(self, num_tissues, emb_dim):
        super().__init__()
        self.num_registry_tokens = num_tissues
        self.registry_tokens = Table(num_tissues, emb_dim)

    def forward(self, x, tissue_vector):
        """layers.py:508-521: the tissue's registry token prepended to x [batch, seq, emb] -> (combined, residual)."""
        t = torch.as_tensor(tissue_vector, device=x.device).reshape(x.size(0), -1)[:, 0].long()
        combined = torch.cat((self.registry_tokens.weight[t][:, None, :], x), dim=1)
        return combined, combined.clone()

    def get_registry_tokens(self):
        return self.registry_tokens.weight


class TissueExpressionHeads(nn.Module):
    """Shared 'bigger' MLP head + Softplus: Sequential indices 0 Linear, 1 LayerNorm, 4 Linear, 6 Linear(emb,1)."""

    def __init__(self, emb_dim, num_tissues, use_bigger_head=False, multi_head=True, mlp_dout=0.1, loss_fn="poisson",
                 head_type="mlp"):
        super().__init__()
        if not (use_bigger_head and not multi_head and head_type == "mlp" and loss_fn == "poisson"):
            raise NotImplementedError("only the vf_model.yaml head (use_bigger_head, shared, mlp, poisson) is implemented")
        self.multi_head = multi_head
        self.tissue_expressions = nn.ModuleDict({"0": Affine(emb_dim, emb_dim), "1": Affine(emb_dim),
                                                 "4": Affine(emb_dim, emb_dim), "6": Affine(1, emb_dim)})
        self._folded = None

    @torch.no_grad()
    def forward(self, g_exp, tissue_vector):
        """layers.py:1113-1144: g_exp [batch, emb_dim], tissue_vector [batch, 1] -> predictions [batch, 1].  With the
        shared head every row runs the same MLP; tissue_vector is checked like upstream (one unique id per row)."""
        from ... import layer_ops as LO
        tv = torch.as_tensor(tissue_vector).reshape(g_exp.shape[0], -1)
        assert bool((tv == tv[:, :1]).all()), "Tissue vector not unique"
        if self._folded is None:
            self._folded = LO._Cache()
        W = self._folded.get(self, lambda sd, dev: LO.head_weights(sd, dev))
        return LO.head_forward(W, g_exp)
