"""Deterministic random-init `state_dict` with the reference's key names and shapes.

There are no checkpoints offline, so benchmarks, tests and golden fixtures use
random weights (BASELINE.json: "random-init vf_model.yaml weights").  The key
layout is the drop-in contract of SURVEY.md §8(b) — it is asserted equal to the
reference classes' own `state_dict()` by tests/golden/make_model_golden.py.
Distributions follow PyTorch's defaults (Linear: U(±1/sqrt(fan_in)), Embedding:
N(0,1); LayerNorm: 1+0.1·N / 0.1·N) so activations stay in a benign range.
"""
import math

import torch

from .alibi import alibi_slopes

V4_PCG_MODEL = dict(  # configs/vf_model.yaml v4_pcg.model (shape-relevant keys)
    num_tissues=63, emb_dim=1536, gene_emb_dim=512, num_heads=32, num_layers=25, use_alibi=True,
    mlp_dout=0.1, use_context=True, use_batching=True, use_res=False, gene_pooling="multi_registry",
    token_dim=512, use_bigger_head=True, multi_head=False, only_cross_attention=False,
    cross_alibi=False, add_context_to_cres=False, train_gene_tokenizer=True, precision="bf16-mixed",
)
SEQ2REG_HP = dict(  # not in the repo (lives in the checkpoint); SURVEY.md §8(d) documented assumption
    vocab_size=500, embedding_dim=512, num_heads=8, num_layers=6, num_tissues=1, num_classes=11,
    learning_rate=1e-4, loss_fn=["cross_entropy", 0], token_length=200, use_context=False,
    positional_encoding="sinusoidal", use_flash=True,
)
FFN_HIDDEN = 2048  # default hidden_dim of every encoder layer, independent of d_model
NUM_REF_CRES = 9   # len(utils.constants.REF_CREs)


def _linear(sd, g, name, out_f, in_f, dtype):
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = ((torch.rand(out_f, in_f, generator=g, device=g.device) * 2 - 1) * bound).to(dtype)
    sd[name + ".bias"] = ((torch.rand(out_f, generator=g, device=g.device) * 2 - 1) * bound).to(dtype)


def _norm(sd, g, name, d, dtype):
    # perturbed around (1, 0) so that gamma/beta handling is actually exercised by parity tests
    sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(d, generator=g, device=g.device)).to(dtype)
    sd[name + ".bias"] = (0.1 * torch.randn(d, generator=g, device=g.device)).to(dtype)


def seq2reg_state_dict(hp, prefix, g, dtype=torch.float32):
    sd = {}
    d = hp["embedding_dim"]
    sd[prefix + "token_embedding.weight"] = torch.randn(hp["vocab_size"], d, generator=g, device=g.device).to(dtype)
    for l in range(hp["num_layers"]):
        p = f"{prefix}transformer_encoder.{l}."
        _linear(sd, g, p + "MHA.Wqkv", 3 * d, d, dtype)
        _linear(sd, g, p + "MHA.out_proj", d, d, dtype)
        _norm(sd, g, p + "norm1", d, dtype); _norm(sd, g, p + "norm2", d, dtype)
        _linear(sd, g, p + "linear_geglu_1", FFN_HIDDEN, d, dtype)
        _linear(sd, g, p + "linear_geglu_2", d, FFN_HIDDEN // 2, dtype)
    for t in range(hp.get("num_tissues", 1)):
        _linear(sd, g, f"{prefix}tissue_classifiers.{t}", hp.get("num_classes", 11), d, dtype)
    return sd


def make_state_dict(cfg=None, seq2reg_hp=None, seed=0, dtype=torch.float32, device="cpu"):
    """device='cpu' is the bit-reproducible stream the golden fixtures use; device='cuda' draws from the CUDA
    generator instead (different values, same distributions) for fast full-size benchmark set-up."""
    cfg = dict(V4_PCG_MODEL if cfg is None else cfg)
    hp = dict(SEQ2REG_HP if seq2reg_hp is None else seq2reg_hp)
    g = torch.Generator(device=device).manual_seed(seed)
    D, H = cfg["emb_dim"], cfg["num_heads"]
    sd = {}
    sd["start_tkn.registry_tokens.weight"] = torch.randn(cfg["num_tissues"], D, generator=g, device=g.device).to(dtype)
    sd.update(seq2reg_state_dict(hp, "cre_tokenizer.", g, dtype))
    sd.update(seq2reg_state_dict(hp, "gene_tokenizer.", g, dtype))
    _linear(sd, g, "gene_map", D, cfg["gene_emb_dim"], dtype)
    if cfg["token_dim"] != D:
        _linear(sd, g, "cre_map", D, cfg["token_dim"], dtype)
    sd["combined_modulator.second_level_context_embedding.weight"] = torch.randn(NUM_REF_CRES, D, generator=g, device=g.device).to(dtype)
    slopes = alibi_slopes(H)
    for stream, n in (("cre_layers", cfg["num_layers"] - 1), ("gene_layers", cfg["num_layers"])):
        for l in range(n):
            p = f"combined_modulator.{stream}.{l}."
            if cfg.get("use_alibi", True):
                sd[p + "m"] = slopes.clone().to(g.device)
            _linear(sd, g, p + "mixer.MHA.Wqkv", 3 * D, D, dtype)
            _linear(sd, g, p + "mixer.MHA.out_proj", D, D, dtype)
            _linear(sd, g, p + "crossMHA.MHA.Wq", D, D, dtype)
            _linear(sd, g, p + "crossMHA.MHA.Wkv", 2 * D, D, dtype)
            _linear(sd, g, p + "crossMHA.MHA.out_proj", D, D, dtype)
            for k in ("norm1", "norm2", "norm3"):
                _norm(sd, g, p + k, D, dtype)
            _linear(sd, g, p + "linear_geglu_1", FFN_HIDDEN, D, dtype)
            _linear(sd, g, p + "linear_geglu_2", D, FFN_HIDDEN // 2, dtype)
    p = "tissue_heads.tissue_expressions."
    _linear(sd, g, p + "0", D, D, dtype); _norm(sd, g, p + "1", D, dtype)
    _linear(sd, g, p + "4", D, D, dtype); _linear(sd, g, p + "6", 1, D, dtype)
    return sd
