"""Coarse entry points of the C ABI (vf_seq2reg_forward, vf_seq2gene_forward, vf_attention_build_slots): the layer loop
inside the library must give the bits of the per-kernel path (same kernels, same order) — GPU box."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.common import GOLD_CFG, GOLD_HP, GOLD_SEED, load_model_golden, parity_report, synth_batch  # noqa: E402
from variantformer_b200 import _lib, engine as engine_mod, ops  # noqa: E402
from variantformer_b200.engine import Engine  # noqa: E402
from variantformer_b200.utils import random_init  # noqa: E402


def _run(engine, batch, **kw):
    sq = lambda xs: [x[:, 0, :] for x in xs]
    out = engine.forward_tokens(sq(batch["cre_sequences"]), sq(batch["cre_attention_masks"]), sq(batch["gene_embeddings"]),
                                sq(batch["gene_attention_masks"]), batch["tissue_context"], batch["ref_cre_labels"], **kw)
    torch.cuda.synchronize()
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}


@pytest.mark.parametrize("cre_stream", [True, False])
def test_coarse_forward_is_bit_identical_to_the_per_kernel_path(monkeypatch, cre_stream):
    cfg = dict(random_init.V4_PCG_MODEL, emb_dim=384, gene_emb_dim=256, num_heads=8, num_layers=4, token_dim=256)
    hp = dict(random_init.SEQ2REG_HP, embedding_dim=256, num_heads=4, num_layers=2)
    sd = random_init.make_state_dict(cfg, hp, seed=3)
    eng = Engine(sd, cfg, hp)
    batch = synth_batch(21, 3, [150, 40, 300], [5, 2, 9], [[62, 0, 14, 7], [3], [5, 9]])
    kw = dict(cre_token_position=[7, 0, 299], gene_token_position=[4, 1, 0])
    monkeypatch.setattr(engine_mod, "CRE_STREAM", cre_stream)
    monkeypatch.setattr(engine_mod, "COARSE", True)
    n0 = ops.launch_count()
    coarse = _run(eng, batch, **kw)
    n_coarse = ops.launch_count() - n0
    monkeypatch.setattr(engine_mod, "COARSE", False)
    n0 = ops.launch_count()
    fine = _run(eng, batch, **kw)
    # the same launches, counted inside the library — except that each of the two seq2reg passes (CRE windows, gene
    # chunks) embeds and centres its tokens in ONE kernel on the coarse path (embed_center_kernel; same arithmetic)
    assert ops.launch_count() - n0 == n_coarse + 2 and n_coarse > 100
    for k in ("pred", "emb", "gene_token_embedding", "cre_token_embedding"):
        assert torch.equal(coarse[k], fine[k]), f"{k}: max diff {(coarse[k] - fine[k]).abs().max().item()}"
    assert coarse["T"] == fine["T"]


def test_coarse_forward_matches_the_reference_golden():
    batch, want, _ = load_model_golden("model_golden_large.npz")
    eng = Engine(random_init.make_state_dict(GOLD_CFG, GOLD_HP, seed=GOLD_SEED), GOLD_CFG, GOLD_HP)
    assert engine_mod.COARSE
    out = _run(eng, batch)
    rep = parity_report(out["emb"].cpu().numpy(), np.concatenate(want["embeddings"]))
    assert rep["ok"], rep
