"""The reference's OWN schedule on the same B200 with LIBRARY kernels — the end-to-end "kernel to beat" of BASELINE.md
section 4.5: bf16 autocast nn.Linear (cuBLASLt), flash_attn 2.8.x varlen attention, fp32 LayerNorm, and the reference's
data flow: the CRE stream and the gene stream are both repeated once per tissue (model_combined_modulator.py:622-649),
the CRE x label cross-attention runs over all C keys, every LayerNorm is a separate pass.

A self-contained restatement for measurement only (the reference checkout is not on the GPU box and its Lightning /
omegaconf dependencies are not installable): random-init weights of the full vf_model.yaml architecture, token-level
synthetic input of the benchmark's shape.  Not product code; imports nothing from oracle/ and no kernel of this repo.

    python tools/ref_schedule_gpu.py [genes_per_step=8] [C=1024] [G=200] [T=63] [steps=3]
"""
import json
import math
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from variantformer_b200.utils import random_init, synth  # noqa: E402  (weights + synthetic tokens only)

from flash_attn import flash_attn_varlen_func  # noqa: E402

DEV = "cuda"


def alibi_slopes(n):
    return torch.tensor([2 ** (-8 * (h + 1) / n) for h in range(n)], device=DEV, dtype=torch.float32)


def lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])          # under autocast: bf16 cuBLASLt GEMM


def ln(sd, name, x):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def mha_self(sd, p, x, cu, maxlen, H, slopes):
    n, d = x.shape
    qkv = lin(sd, p + "Wqkv", x).view(n, 3, H, d // H)
    o = flash_attn_varlen_func(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, maxlen, maxlen, causal=False, alibi_slopes=slopes)
    return lin(sd, p + "out_proj", o.reshape(n, d))


def mha_cross(sd, p, x, ctx, cu_q, cu_k, mq, mk, H):
    n, d = x.shape
    q = lin(sd, p + "Wq", x).view(n, H, d // H)
    kv = lin(sd, p + "Wkv", ctx).view(ctx.shape[0], 2, H, d // H)
    o = flash_attn_varlen_func(q, kv[:, 0], kv[:, 1], cu_q, cu_k, mq, mk, causal=False)
    return lin(sd, p + "out_proj", o.reshape(n, d))


def ffn(sd, p, x):
    u, gate = lin(sd, p + "linear_geglu_1", x).chunk(2, dim=-1)
    return lin(sd, p + "linear_geglu_2", u * F.gelu(gate))


def context_layer(sd, p, src, cu, maxlen, ctx, cu_k, mk, H, slopes):
    x = mha_self(sd, p + "mixer.MHA.", ln(sd, p + "norm1", src), cu, maxlen, H, slopes) + src
    x = mha_cross(sd, p + "crossMHA.MHA.", ln(sd, p + "norm2", x), ctx, cu, cu_k, maxlen, mk, H) + x
    return ffn(sd, p, ln(sd, p + "norm3", x)) + src


def seq2reg(sd, pre, hp, tok, mask):
    keep = ~mask
    lens = keep.sum(1)
    cu = F.pad(lens.cumsum(0), (1, 0)).int()
    n, L = tok.shape
    pos = torch.arange(L, device=DEV).expand(n, L)[keep]
    d = hp["embedding_dim"]
    pe = torch.zeros(L, d, device=DEV)
    div = torch.exp(torch.arange(0, d, 2, device=DEV).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(torch.arange(L, device=DEV)[:, None] * div); pe[:, 1::2] = torch.cos(torch.arange(L, device=DEV)[:, None] * div)
    x = sd[pre + "token_embedding.weight"][tok[keep]] + pe[pos]
    for l in range(hp["num_layers"]):
        p = f"{pre}transformer_encoder.{l}."
        x1 = mha_self(sd, p + "MHA.", ln(sd, p + "norm1", x), cu, int(lens.max()), hp["num_heads"], None) + x
        x = ffn(sd, p, ln(sd, p + "norm2", x1)) + x
    seg = torch.repeat_interleave(torch.arange(n, device=DEV), lens)
    return torch.zeros(n, d, device=DEV, dtype=x.dtype).index_add_(0, seg, x) / lens[:, None]


@torch.no_grad()
def forward(sd, cfg, hp, batch):
    """One batch of genes, reference schedule: every (gene, tissue) copy carries its own CRE stream."""
    H, NL, D = cfg["num_heads"], cfg["num_layers"], cfg["emb_dim"]
    slopes = alibi_slopes(H)
    cres, genes, labels, tissues = [], [], [], []
    for g in range(len(batch["cre_sequences"])):
        ct = batch["cre_sequences"][g][:, 0].to(DEV); cm = batch["cre_attention_masks"][g][:, 0].to(DEV)
        gt = batch["gene_embeddings"][g][:, 0].to(DEV); gm = batch["gene_attention_masks"][g][:, 0].to(DEV)
        cres.append(lin(sd, "cre_map", seq2reg(sd, "cre_tokenizer.", hp, ct, cm)))
        genes.append(lin(sd, "gene_map", seq2reg(sd, "gene_tokenizer.", hp, gt, gm)))
        labels.append(batch["ref_cre_labels"][g].to(DEV)); tissues.append(batch["tissue_context"][g].to(DEV))
    cx, cctx, gx, cu_c, cu_g = [], [], [], [0], [0]
    for c, gseq, lab, tis in zip(cres, genes, labels, tissues):
        T = len(tis)
        cx.append(c.repeat(T, 1)); cctx.append(sd["combined_modulator.second_level_context_embedding.weight"][lab].repeat(T, 1))
        reg = sd["start_tkn.registry_tokens.weight"][tis]
        gx.append(torch.cat([reg[:, None, :], gseq[None].expand(T, -1, -1)], 1).reshape(-1, D))
        for _ in range(T):
            cu_c.append(cu_c[-1] + c.shape[0]); cu_g.append(cu_g[-1] + gseq.shape[0] + 1)
    cx, cctx, gx = torch.cat(cx), torch.cat(cctx), torch.cat(gx)
    cu_c = torch.tensor(cu_c, device=DEV, dtype=torch.int32); cu_g = torch.tensor(cu_g, device=DEV, dtype=torch.int32)
    mc = int((cu_c[1:] - cu_c[:-1]).max()); mg = int((cu_g[1:] - cu_g[:-1]).max())
    gx = context_layer(sd, "combined_modulator.gene_layers.0.", gx, cu_g, mg, cx, cu_c, mc, H, slopes)
    for i in range(NL - 1):
        cx = context_layer(sd, f"combined_modulator.cre_layers.{i}.", cx, cu_c, mc, cctx, cu_c, mc, H, slopes)
        gx = context_layer(sd, f"combined_modulator.gene_layers.{i + 1}.", gx, cu_g, mg, cx, cu_c, mc, H, slopes)
    emb = gx[cu_g[:-1].long()]
    p = "tissue_heads.tissue_expressions."
    h = F.gelu(ln(sd, p + "1", lin(sd, p + "0", emb)))
    h = F.gelu(lin(sd, p + "4", h))
    return F.softplus(lin(sd, p + "6", h).float()), emb.float()


def main():
    a = [int(x) for x in sys.argv[1:]]
    B, C, G, T, steps = (a + [8, 1024, 200, 63, 3][len(a):])[:5]
    cfg, hp = dict(random_init.V4_PCG_MODEL), dict(random_init.SEQ2REG_HP)
    sd = random_init.make_state_dict(cfg, hp, seed=0, device=DEV)
    batches = [synth.token_batch(s, B, C, G, T) for s in range(2)]
    # the reference batches T tissue copies of every gene: memory grows with B x T x C; fall back to fewer genes per call
    per_call = B
    with torch.autocast("cuda", dtype=torch.bfloat16):
        while True:
            try:
                sub = {k: (v[:per_call] if isinstance(v, list) else v) for k, v in batches[0].items()}
                forward(sd, cfg, hp, sub); torch.cuda.synchronize()
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache(); per_call = max(1, per_call // 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            b = batches[s % 2]
            for g0 in range(0, B, per_call):
                pred, emb = forward(sd, cfg, hp, {k: (v[g0:g0 + per_call] if isinstance(v, list) else v) for k, v in b.items()})
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"what": "reference schedule with library kernels (cuBLASLt bf16 autocast + flash_attn varlen), same B200",
           "genes_per_step": B, "C": C, "G": G, "T": T, "genes_per_call": per_call, "ms_per_step": ms,
           "predictions_per_s": B * T / (ms / 1e3), "finite": bool(torch.isfinite(pred).all())}
    print(json.dumps(out))
    json.dump(out, open("gpurun_out/ref_schedule_gpu.json", "w"), indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
