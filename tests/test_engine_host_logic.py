"""CPU check of the engine's host-side bookkeeping (unpadding, tile maps, tissue stacking, gather indices, GeGLU
weight interleave, label counts) with tests/fake_ops.py standing in for the C-ABI kernels.  The numerical parity
of the kernels themselves is the job of the `-m gpu` tests."""
import numpy as np
import pytest
import torch

from oracle import model_fp32
from tests import fake_ops
from tests.common import GOLD_CFG, GOLD_HP, GOLD_SEED, load_model_golden, pearson, rel_err
from variantformer_b200 import engine as engine_mod
from variantformer_b200.utils import random_init


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(engine_mod, "ops", fake_ops)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)

    class _NullCtx:
        def __init__(self, *a): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    monkeypatch.setattr(torch.cuda, "device", _NullCtx)
    sd = random_init.make_state_dict(GOLD_CFG, GOLD_HP, seed=GOLD_SEED)
    return engine_mod.Engine(sd, GOLD_CFG, GOLD_HP, device="cpu")


def test_engine_schedule_matches_golden(cpu_engine):
    batch, want, _ = load_model_golden()
    sq = lambda xs: [x[:, 0, :] for x in xs]
    out = cpu_engine.forward_tokens(sq(batch["cre_sequences"]), sq(batch["cre_attention_masks"]),
                                    sq(batch["gene_embeddings"]), sq(batch["gene_attention_masks"]),
                                    batch["tissue_context"], batch["ref_cre_labels"],
                                    cre_token_position=[2, 0], gene_token_position=[1, 0])
    emb = out["emb"].numpy(); w_emb = np.concatenate(want["embeddings"])
    assert rel_err(emb, w_emb) <= 1e-2 and pearson(emb, w_emb) >= 0.9999
    pred = out["pred"].numpy(); w_pred = np.concatenate(want["pred_gene_exp"]).ravel()
    assert rel_err(pred, w_pred) <= 1e-2
    assert out["T"] == [3, 1]
    assert out["gene_token_embedding"].shape == (4, GOLD_CFG["emb_dim"])
    assert out["cre_token_embedding"].shape == (4, GOLD_CFG["emb_dim"])


def test_interleave_geglu_layout():
    w = torch.arange(2048 * 3, dtype=torch.float32).view(2048, 3)
    p = engine_mod.interleave_geglu(w)
    assert torch.equal(p[:128], w[:128]) and torch.equal(p[128:256], w[1024:1152])
    assert torch.equal(p[256:384], w[128:256]) and torch.equal(p[1920:2048], w[1920:2048])
    b = engine_mod.interleave_geglu(torch.arange(2048, dtype=torch.float32))
    assert b[128].item() == 1024 and b[256].item() == 128
