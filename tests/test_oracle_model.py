"""Pins oracle/model_fp32.py against the fixture produced by the REFERENCE's own
classes (tests/golden/make_model_golden.py), and checks the schedule identities
the CUDA path relies on."""
import numpy as np
import torch

from oracle import model_fp32
from tests.common import GOLD_CFG, GOLD_HP, GOLD_SEED, load_model_golden, pearson, rel_err
from variantformer_b200.utils import random_init


def _sd():
    return random_init.make_state_dict(GOLD_CFG, GOLD_HP, seed=GOLD_SEED)


def test_oracle_matches_reference_golden():
    batch, want, _ = load_model_golden()
    got = model_fp32.predict_step(_sd(), GOLD_CFG, GOLD_HP, batch, schedule="reference")
    for g in range(len(want["pred_gene_exp"])):
        # fp32 vs fp32: only summation-order noise is allowed
        np.testing.assert_allclose(got["pred_gene_exp"][g], want["pred_gene_exp"][g], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(got["embeddings"][g], want["embeddings"][g], rtol=2e-4, atol=2e-5)


def test_oracle_matches_large_reference_golden():
    # C = 300 / 150 CREs, G = 130 / 40 chunks, T = 5 / 63, strand flags 0 / 1 — from the reference's own predict_step
    batch, want, _ = load_model_golden("model_golden_large.npz")
    got = model_fp32.predict_step(_sd(), GOLD_CFG, GOLD_HP, batch, schedule="dedup")
    assert [e.shape[0] for e in got["embeddings"]] == [5, 63]
    for g in range(2):
        np.testing.assert_allclose(got["pred_gene_exp"][g], want["pred_gene_exp"][g], rtol=5e-5, atol=5e-6)
        np.testing.assert_allclose(got["embeddings"][g], want["embeddings"][g], rtol=5e-4, atol=5e-5)


def test_dedup_schedule_is_exact():
    # the CRE stream is tissue independent
    batch, _, _ = load_model_golden()
    a = model_fp32.predict_step(_sd(), GOLD_CFG, GOLD_HP, batch, schedule="reference")
    b = model_fp32.predict_step(_sd(), GOLD_CFG, GOLD_HP, batch, schedule="dedup")
    for g in range(len(a["embeddings"])):
        np.testing.assert_allclose(a["embeddings"][g], b["embeddings"][g], rtol=1e-5, atol=1e-6)


def test_bf16_emulation_within_stated_tolerance():
    # numerical model of the CUDA path (bf16 operands, fp32 accumulate/stream): the tolerance the GPU tests use
    batch, want, _ = load_model_golden()
    got = model_fp32.predict_step(_sd(), GOLD_CFG, GOLD_HP, batch, schedule="dedup", emulate_bf16=True)
    e = np.concatenate([x.ravel() for x in got["embeddings"]]); w = np.concatenate([x.ravel() for x in want["embeddings"]])
    assert rel_err(e, w) <= 1e-2 and pearson(e, w) >= 0.9999


def test_full_size_key_contract():
    _, _, z = load_model_golden()
    sd = random_init.make_state_dict(dict(random_init.V4_PCG_MODEL, num_layers=2), dict(random_init.SEQ2REG_HP, num_layers=1))
    mine = sorted(f"{k}:{tuple(v.shape)}" for k, v in sd.items())
    assert mine == list(z["full_keys"])


def test_alibi_slopes():
    s = model_fp32.alibi_slopes(32)
    assert torch.allclose(s, torch.tensor([2 ** (-0.25 * (h + 1)) for h in range(32)]))
    s = model_fp32.alibi_slopes(8)
    assert torch.allclose(s, torch.tensor([2.0 ** -(h + 1) for h in range(8)]))
