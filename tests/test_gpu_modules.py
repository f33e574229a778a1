"""Layer-level API of the mirror classes (the module forward signatures) on the GPU box: each class's
`forward`, called the way the reference calls it, against the fp32 oracle's restatement of the same layer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model_fp32 as O  # noqa: E402  (checker only)
from tests.common import parity_report  # noqa: E402
from variantformer_b200.utils import random_init  # noqa: E402

CFG = dict(random_init.V4_PCG_MODEL, emb_dim=384, gene_emb_dim=256, num_heads=8, num_layers=3, token_dim=256)
HP = dict(random_init.SEQ2REG_HP, embedding_dim=256, num_heads=4, num_layers=2)
DEV = "cuda"


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _ok(got, want, what):
    rep = parity_report(got.float().cpu().numpy(), want.numpy())
    assert rep["ok"], (what, rep)


def test_flash_transformer_layer_forward():
    from variantformer_b200.seq2reg.modules import FlashTransformerLayer
    sd = random_init.make_state_dict(CFG, HP, seed=11)
    p = "cre_tokenizer.transformer_encoder.1."
    layer = FlashTransformerLayer(HP["embedding_dim"], HP["num_heads"]).to(DEV)
    layer.load_state_dict(_sub(sd, p))
    g = torch.Generator().manual_seed(0)
    B, S, d = 5, 200, HP["embedding_dim"]
    src = torch.randn(B, S, d, generator=g)
    lens = [200, 97, 1, 130, 64]
    mask = torch.arange(S)[None, :] >= torch.tensor(lens)[:, None]               # True = padding
    out = layer(src.to(DEV), src_key_padding_mask=mask.to(DEV))
    assert out.shape == (B, S, d)
    num = O._Num()
    x = src[~mask]
    cu = [0] + np.cumsum(lens).tolist()
    a = O._mha_self(num, sd, p + "MHA.", O._ln(sd, p + "norm1.", x), cu, HP["num_heads"], None)
    x1 = a + x
    want = O._geglu_ffn(num, sd, p, O._ln(sd, p + "norm2.", x1)) + x
    _ok(out[~mask.to(DEV)], want, "FlashTransformerLayer")
    assert bool((out[mask.to(DEV)] == 0).all())
    # no mask: every position valid
    out2 = layer(src[:2].to(DEV))
    a = O._mha_self(num, sd, p + "MHA.", O._ln(sd, p + "norm1.", src[:2].reshape(-1, d)), [0, S, 2 * S], HP["num_heads"], None)
    x1 = a + src[:2].reshape(-1, d)
    want2 = O._geglu_ffn(num, sd, p, O._ln(sd, p + "norm2.", x1)) + src[:2].reshape(-1, d)
    _ok(out2.reshape(-1, d), want2, "FlashTransformerLayer (no mask)")


def test_context_encoder_layer_forward_padded_and_unpadded():
    from variantformer_b200.seq2gene.modules.layers import ContextFlashAttentionEncoderLayer
    sd = random_init.make_state_dict(CFG, HP, seed=12)
    p = "combined_modulator.gene_layers.1."
    D, H = CFG["emb_dim"], CFG["num_heads"]
    layer = ContextFlashAttentionEncoderLayer(D, H, use_alibi=True).to(DEV)
    layer.load_state_dict(_sub(sd, p))
    g = torch.Generator().manual_seed(1)
    q_lens, k_lens = [201, 40, 130], [300, 64, 9]
    B, S, Sc = 3, 201, 300
    src = torch.randn(B, S, D, generator=g); ctx = torch.randn(B, Sc, D, generator=g)
    qmask = torch.arange(S)[None, :] >= torch.tensor(q_lens)[:, None]
    kmask = torch.arange(Sc)[None, :] >= torch.tensor(k_lens)[:, None]
    cu_q = [0] + np.cumsum(q_lens).tolist(); cu_k = [0] + np.cumsum(k_lens).tolist()
    want = O._context_layer(O._Num(), sd, p, src[~qmask], cu_q, ctx[~kmask], cu_k, H, O.alibi_slopes(H))
    out = layer(src.to(DEV), ctx.to(DEV), src_key_padding_mask=qmask.to(DEV), context_padding_mask=kmask.to(DEV))
    assert out.shape == (B, S, D)
    _ok(out[~qmask.to(DEV)], want, "ContextFlashAttentionEncoderLayer (padded)")
    # unpadded mode, the way CombinedModulator.forward calls its layers (gene_unpad_info / context_unpad_info)
    info = lambda cu, m: {"cu_seqlens": torch.tensor(cu, dtype=torch.int32, device=DEV), "max_seqlen": m}
    out_u = layer(src[~qmask].to(DEV), ctx[~kmask].to(DEV), gene_unpad_info=info(cu_q, S), context_unpad_info=info(cu_k, Sc))
    assert out_u.shape == (sum(q_lens), D)
    assert torch.equal(out_u, out[~qmask.to(DEV)])                                # same kernels, same rows: bit-identical


def test_tissue_expression_heads_forward():
    from variantformer_b200.seq2gene.modules.layers import TissueExpressionHeads
    sd = random_init.make_state_dict(CFG, HP, seed=13)
    head = TissueExpressionHeads(CFG["emb_dim"], 63, use_bigger_head=True, multi_head=False).to(DEV)
    head.load_state_dict(_sub(sd, "tissue_heads."))
    g = torch.Generator().manual_seed(2)
    e = torch.randn(37, CFG["emb_dim"], generator=g)
    tv = torch.randint(0, 63, (37, 1), generator=g)
    out = head(e.to(DEV), tv.to(DEV))
    want = O.head(O._Num(), sd, e)
    assert out.shape == (37, 1)
    assert torch.allclose(out.cpu(), want, atol=2e-2, rtol=2e-2)
    with pytest.raises(AssertionError, match="not unique"):
        head(e[:2].to(DEV), torch.tensor([[1, 2], [3, 3]]))


def test_combined_modulator_forward():
    """CombinedModulator.forward with the reference's signature on a padded batch (one row per (gene, tissue) copy, the
    CRE stream repeated per copy as prepare_input does) against the oracle's layer loop."""
    from variantformer_b200.seq2gene.model_combined_modulator import CombinedModulator
    cfg = dict(CFG, num_layers=2)
    sd = random_init.make_state_dict(cfg, HP, seed=14)
    D, H = cfg["emb_dim"], cfg["num_heads"]
    mod = CombinedModulator(D, H, 2, True, 0.0, True, num_ref_cres=9, only_cross_attention=False).to(DEV)
    mod.load_state_dict(_sub(sd, "combined_modulator."))
    g = torch.Generator().manual_seed(3)
    B, Sc, Sg = 2, 150, 41
    c_lens, g_lens = [150, 33], [41, 12]
    cre = torch.randn(B, Sc, D, generator=g); gene = torch.randn(B, Sg, D, generator=g)
    labels = torch.randint(0, 9, (B, Sc), generator=g)
    cmask = torch.arange(Sc)[None, :] >= torch.tensor(c_lens)[:, None]
    gmask = torch.arange(Sg)[None, :] >= torch.tensor(g_lens)[:, None]
    gpos = torch.tensor([[5], [0]]); cpos = torch.tensor([[149], [7]])
    out, gtok, ctok = mod(cre.to(DEV), gene.to(DEV), context=labels.to(DEV), cre_padding_mask=cmask.to(DEV),
                          gene_padding_mask=gmask.to(DEV), context_padding_mask=cmask.to(DEV),
                          cre_token_position=cpos.to(DEV), gene_token_position=gpos.to(DEV))
    num, slopes = O._Num(), O.alibi_slopes(H)
    cu_c = [0] + np.cumsum(c_lens).tolist(); cu_g = [0] + np.cumsum(g_lens).tolist()
    cx, gx = cre[~cmask], gene[~gmask]
    lab = sd["combined_modulator.second_level_context_embedding.weight"][labels[~cmask]]
    gx = O._context_layer(num, sd, "combined_modulator.gene_layers.0.", gx, cu_g, cx, cu_c, H, slopes)
    cx = O._context_layer(num, sd, "combined_modulator.cre_layers.0.", cx, cu_c, lab, cu_c, H, slopes)
    gx = O._context_layer(num, sd, "combined_modulator.gene_layers.1.", gx, cu_g, cx, cu_c, H, slopes)
    _ok(out[~gmask.to(DEV)], gx, "CombinedModulator gene output")
    _ok(gtok, torch.stack([gx[cu_g[0] + 5], gx[cu_g[1] + 0]]), "gene token embedding")
    _ok(ctok, torch.stack([cx[cu_c[0] + 149], cx[cu_c[1] + 7]]), "cre token embedding")


# ---- layer variants the reference can build from other config flags (a16, a25) --------------------------
def _init(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("norm1.weight") or n.endswith("norm2.weight") or n.endswith("norm3.weight"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) / (p.shape[-1] ** 0.5))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return {k: v.detach().clone().float().cpu() for k, v in module.state_dict().items()}


def _ffn_tail(sd, p, x1, src, norm):
    return O._geglu_ffn(O._Num(), sd, p, O._ln(sd, p + norm + ".", x1)) + src


def test_self_only_and_cross_only_layer_variants():
    from variantformer_b200.seq2gene.modules.layers import (ContextFlashCrossAttentionEncoderLayer,
                                                            FlashAttentionEncoderLayer)
    D, H = 384, 8
    g = torch.Generator().manual_seed(5)
    q_lens, k_lens = [130, 201, 7], [64, 300, 9]
    B, S, Sc = 3, 201, 300
    src = torch.randn(B, S, D, generator=g); ctx = torch.randn(B, Sc, D, generator=g)
    qmask = torch.arange(S)[None, :] >= torch.tensor(q_lens)[:, None]
    kmask = torch.arange(Sc)[None, :] >= torch.tensor(k_lens)[:, None]
    cu_q = [0] + np.cumsum(q_lens).tolist(); cu_k = [0] + np.cumsum(k_lens).tolist()
    num, slopes = O._Num(), O.alibi_slopes(H)
    x = src[~qmask]
    # FlashAttentionEncoderLayer (use_context=False CRE layers): self-attention + FFN(norm2); norm3 is never read
    layer = FlashAttentionEncoderLayer(D, H, use_alibi=True)
    sd = _init(layer, 1); layer = layer.to(DEV)
    a = O._mha_self(num, sd, "mixer.MHA.", O._ln(sd, "norm1.", x), cu_q, H, slopes)
    want = _ffn_tail(sd, "", a + x, x, "norm2")
    out = layer(src.to(DEV), src_key_padding_mask=qmask.to(DEV))
    _ok(out[~qmask.to(DEV)], want, "FlashAttentionEncoderLayer")
    # ContextFlashCrossAttentionEncoderLayer (only_cross_attention gene layers), without and with cross ALiBi
    for cross_alibi in (False, True):
        layer = ContextFlashCrossAttentionEncoderLayer(D, H, use_alibi=True, cross_alibi=cross_alibi)
        sd = _init(layer, 2 + cross_alibi); layer = layer.to(DEV)
        xn = O._ln(sd, "norm1.", x)
        n = xn.shape[0]
        qq = F_linear(xn, sd, "crossMHA.MHA.Wq").view(n, H, D // H)
        kv = F_linear(ctx[~kmask], sd, "crossMHA.MHA.Wkv").view(-1, 2, H, D // H)
        o = O._attention(num, qq, kv[:, 0], kv[:, 1], cu_q, cu_k, slopes if cross_alibi else None).reshape(n, D)
        c = F_linear(o, sd, "crossMHA.MHA.out_proj")
        want = _ffn_tail(sd, "", c + x, x, "norm2")
        out = layer(src.to(DEV), ctx.to(DEV), context_padding_mask=kmask.to(DEV), src_key_padding_mask=qmask.to(DEV))
        _ok(out[~qmask.to(DEV)], want, f"ContextFlashCrossAttentionEncoderLayer cross_alibi={cross_alibi}")


def F_linear(x, sd, name):
    return torch.nn.functional.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def test_combined_modulator_other_flag_combinations():
    """use_context=False (self-only CRE layers), only_cross_attention=True gene layers, use_res=True."""
    from variantformer_b200.seq2gene.model_combined_modulator import CombinedModulator
    D, H = 384, 8
    mod = CombinedModulator(D, H, 2, True, 0.0, False, only_cross_attention=True, use_res=True)
    sd = _init(mod, 7); mod = mod.to(DEV)
    g = torch.Generator().manual_seed(6)
    B, Sc, Sg = 2, 140, 33
    c_lens, g_lens = [140, 20], [33, 5]
    cre = torch.randn(B, Sc, D, generator=g); gene = torch.randn(B, Sg, D, generator=g)
    cmask = torch.arange(Sc)[None, :] >= torch.tensor(c_lens)[:, None]
    gmask = torch.arange(Sg)[None, :] >= torch.tensor(g_lens)[:, None]
    out, _, _ = mod(cre.to(DEV), gene.to(DEV), cre_padding_mask=cmask.to(DEV), gene_padding_mask=gmask.to(DEV))
    num, slopes = O._Num(), O.alibi_slopes(H)
    cu_c = [0] + np.cumsum(c_lens).tolist(); cu_g = [0] + np.cumsum(g_lens).tolist()
    cx, gx = cre[~cmask], gene[~gmask]
    gres = gx.clone()

    def gene_layer(i, gx, cx):
        p = f"gene_layers.{i}."
        xn = O._ln(sd, p + "norm1.", gx); n = xn.shape[0]
        qq = F_linear(xn, sd, p + "crossMHA.MHA.Wq").view(n, H, D // H)
        kv = F_linear(cx, sd, p + "crossMHA.MHA.Wkv").view(-1, 2, H, D // H)
        o = O._attention(num, qq, kv[:, 0], kv[:, 1], cu_g, cu_c, None).reshape(n, D)
        x1 = F_linear(o, sd, p + "crossMHA.MHA.out_proj") + gx
        return _ffn_tail(sd, p, x1, gx, "norm2") + gres
    gx = gene_layer(0, gx, cx)
    a = O._mha_self(num, sd, "cre_layers.0.mixer.MHA.", O._ln(sd, "cre_layers.0.norm1.", cx), cu_c, H, slopes)
    cx = _ffn_tail(sd, "cre_layers.0.", a + cx, cx, "norm2")
    gx = gene_layer(1, gx, cx)
    _ok(out[~gmask.to(DEV)], gx, "CombinedModulator(use_context=False, only_cross_attention, use_res)")


@pytest.mark.parametrize("expand", [False, True])
def test_seq2reg_with_label_context(expand):
    """Seq2RegPredictor(use_context=True) (seq2reg/model.py:222-250): label embedding as cross-attention context."""
    from variantformer_b200.seq2reg.model import Seq2RegPredictor
    hp = dict(HP, use_context=True, expand_context=expand, token_length=200, num_layers=2)
    model = Seq2RegPredictor(**hp)
    sd = _init(model, 11 + expand); model = model.to(DEV)
    g = torch.Generator().manual_seed(8)
    b, L, d, H = 6, 200, hp["embedding_dim"], hp["num_heads"]
    lens = [200, 97, 3, 150, 64, 1]
    tok = torch.randint(4, 500, (b, 1, L), generator=g)
    mask = (torch.arange(L)[None, :] >= torch.tensor(lens)[:, None])[:, None, :]
    labels = torch.randint(0, 9, (b,), generator=g)
    out = model(tok.to(DEV), mask.to(DEV), None, context=labels.to(DEV), only_embed=True)
    assert out.shape == (b, 1, d)
    num = O._Num()
    keep = ~mask[:, 0]
    cu = [0] + np.cumsum(lens).tolist()
    x = sd["token_embedding.weight"][tok[:, 0]] + O.sinusoidal_pe(d, L)
    ctx = sd["context_embedding.weight"][labels][:, None, :]
    ctx = ctx * sd["expand_context.weight"].reshape(1, L, 1) + sd["expand_context.bias"].reshape(1, L, 1) if expand \
        else ctx.expand(b, L, d)
    x, ctx = x[keep], ctx[keep]
    for l in range(2):
        x = O._context_layer(num, sd, f"transformer_encoder.{l}.", x, cu, ctx, cu, H, None)
    want = torch.stack([x[cu[i]:cu[i + 1]].mean(0) for i in range(b)])
    _ok(out[:, 0], want, f"Seq2RegPredictor use_context expand={expand}")


def test_generic_top_level_forward_agrees_with_the_engine_and_serves_other_variants():
    """Seq2GenePredictorCombinedModulator._forward_generic restates the reference's forward (:540-720) module by module.
    On the vf_model.yaml architecture it must agree with the batched engine (same kernels, different schedule: no tissue
    de-duplication, padded batches); other gene_pooling / context flags must run through it."""
    from tests.common import synth_batch
    from variantformer_b200.seq2gene.model_combined_modulator import Seq2GenePredictorCombinedModulator, attach_trainer
    from variantformer_b200.seq2reg.model import Seq2RegPredictor
    sd = random_init.make_state_dict(CFG, HP, seed=21)
    model = Seq2GenePredictorCombinedModulator(cre_tokenizer=Seq2RegPredictor(**HP), gene_tokenizer=Seq2RegPredictor(**HP), **CFG)
    model.load_state_dict(sd); model.eval().to(DEV); attach_trainer(model)
    batch = synth_batch(5, 2, [40, 150], [3, 5], [[62, 3], [7]])
    args = (batch["cre_sequences"], batch["cre_attention_masks"], batch["tissue_context"], batch["ref_cre_labels"],
            batch["strand_val"], batch["gene_embeddings"], batch["gene_attention_masks"])
    assert model._engine_variant()
    pred_e, _, emb_e, _, _ = model(*args, return_embedding=True)
    pred_g, _, emb_g, _, _ = model._forward_generic(args[0], args[1], args[2], args[3], args[5], args[6], True)
    rep = parity_report(emb_g.cpu().numpy(), emb_e.cpu().numpy())
    assert rep["ok"], rep
    assert torch.allclose(pred_g.cpu(), pred_e.cpu(), atol=2e-2, rtol=2e-2)
    # another architecture: start token pooling, cross-attention-only gene layers, use_res, tissue context added to the CREs
    cfg = dict(CFG, gene_pooling="start_token", only_cross_attention=True, use_res=True, add_context_to_cres=True, num_layers=2)
    other = Seq2GenePredictorCombinedModulator(cre_tokenizer=Seq2RegPredictor(**HP), gene_tokenizer=Seq2RegPredictor(**HP), **cfg)
    _init(other, 22); other.eval().to(DEV); attach_trainer(other)
    assert not other._engine_variant()
    pred, donors, emb, _, _ = other(*args, return_embedding=True)
    assert pred.shape == (3, 1) and emb.shape == (3, CFG["emb_dim"]) and donors == [0, 1]
    assert bool(torch.isfinite(pred).all()) and bool(torch.isfinite(emb).all())
    assert (emb[0] - emb[1]).abs().max() > 0                    # the two tissues of gene 0 differ (AddContext)
    mx = Seq2GenePredictorCombinedModulator(cre_tokenizer=Seq2RegPredictor(**HP), gene_tokenizer=Seq2RegPredictor(**HP),
                                            **dict(CFG, gene_pooling="max", num_layers=2))
    _init(mx, 23); mx.eval().to(DEV); attach_trainer(mx)
    pred, _ = mx(*args)
    assert pred.shape == (3, 1) and bool(torch.isfinite(pred).all())
