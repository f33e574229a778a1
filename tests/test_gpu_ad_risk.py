"""AD-risk head: vf_forest_predict vs the GBDT semantics it replaces.  The reference evaluates
treelite models (S3 artifacts, treelite not installable here): parity is pinned against scikit-learn's
GradientBoostingClassifier.predict_proba — the model family those treelite checkpoints are exported from — GPU box."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fit(seed, n_trees, depth, d=64, n=400):
    from sklearn.ensemble import GradientBoostingClassifier
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d)).astype(np.float32)
    y = (x[:, :5].sum(1) + 0.5 * rng.normal(size=n) > 0).astype(int)
    return GradientBoostingClassifier(n_estimators=n_trees, max_depth=depth, learning_rate=0.1, random_state=seed).fit(x, y), rng


def test_forest_bank_matches_sklearn_predict_proba(tmp_path):
    from variantformer_b200.processors.ad_risk import ADrisk, ForestBank, LocalPredictorManifest, forest_from_sklearn, save_forest
    models = [_fit(1, 50, 3), _fit(2, 120, 4), _fit(3, 7, 6)]
    forests = [forest_from_sklearn(m) for m, _ in models]
    bank = ForestBank(forests)
    rng = np.random.default_rng(9)
    x = rng.normal(size=(1000, 64)).astype(np.float32)
    ids = rng.integers(0, 3, 1000)
    got = bank.predict_proba(x, ids)
    want = np.empty(1000)
    for k, (m, _) in enumerate(models):
        want[ids == k] = m.predict_proba(x[ids == k])[:, 1]
    assert np.allclose(got, want, atol=2e-6, rtol=1e-5), np.abs(got - want).max()
    # the single-predictor surface of the reference (ADrisk(gene, tissue)(embeddings))
    save_forest(str(tmp_path / "ENSG000001.1_62.npz"), forests[1])
    risk = ADrisk("ENSG000001.1", 62, manifest=LocalPredictorManifest(str(tmp_path)))
    assert np.allclose(risk(x[:50]), models[1][0].predict_proba(x[:50])[:, 1], atol=2e-6)
    with pytest.raises(FileNotFoundError):
        ADrisk("ENSG000001.1", 3, manifest=LocalPredictorManifest(str(tmp_path)))


def test_treelite_json_import_and_missing_values():
    from variantformer_b200.processors.ad_risk import ForestBank, forest_from_treelite_json
    model = {"trees": [
        {"root_id": 0, "nodes": [
            {"node_id": 0, "split_feature_id": 2, "threshold": 0.5, "comparison_op": "<", "default_left": True,
             "left_child": 1, "right_child": 2},
            {"node_id": 1, "leaf_value": 0.25}, {"node_id": 2, "leaf_value": -0.75}]},
        {"root_id": 0, "nodes": [{"node_id": 0, "leaf_value": [1.0]}]}],
        "base_scores": [0.125]}
    bank = ForestBank([forest_from_treelite_json(json.dumps(model))])
    x = np.zeros((4, 8), np.float32)
    x[0, 2] = 0.4; x[1, 2] = 0.5; x[2, 2] = 0.6; x[3, 2] = np.nan
    raw = np.array([0.125 + 0.25 + 1.0, 0.125 - 0.75 + 1.0, 0.125 - 0.75 + 1.0, 0.125 + 0.25 + 1.0])   # "<": 0.5 goes right
    assert np.allclose(bank.predict_proba(x, np.zeros(4, np.int32)), 1 / (1 + np.exp(-raw)), atol=1e-6)
