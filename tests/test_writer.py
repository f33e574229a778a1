"""Result sink (variantformer_b200/writer.py): atomic per-slab Parquet files, resume bookkeeping, schema — CPU only."""
import os

import numpy as np
import pytest

from variantformer_b200.writer import ResultWriter


def test_writer_round_trip_and_resume(tmp_path):
    D = 8
    w = ResultWriter(str(tmp_path / "out"), D, meta={"model": "x"})
    rng = np.random.default_rng(0)
    released = []
    for slab in (0, 2):
        rows = [(i // 3, "S1", f"gene{i // 3}", i % 3) for i in range(6)]
        pred = rng.random(6).astype(np.float32); emb = rng.random((6, D)).astype(np.float32)
        w.submit(slab, rows, pred, emb, release=lambda s=slab: released.append(s))
        if slab == 0:
            first = (pred.copy(), emb.copy())
    w.close()
    assert sorted(released) == [0, 2] and w.done() == {0, 2}
    assert not [f for f in os.listdir(tmp_path / "out") if f.endswith(".tmp")]
    df = w.read_all()
    assert list(df.columns) == ["slab", "item", "sample", "gene_id", "tissue", "predicted_expression", "embeddings"]
    assert len(df) == 12 and df["slab"].tolist() == [0] * 6 + [2] * 6
    assert np.allclose(df["predicted_expression"].to_numpy()[:6], first[0])
    assert np.allclose(np.stack(df["embeddings"].to_numpy()[:6]), first[1])
    # a second writer on the same directory sees the finished slabs; a different run is refused
    w2 = ResultWriter(str(tmp_path / "out"), D, meta={"model": "x"})
    assert w2.done() == {0, 2}
    w2.close()
    with pytest.raises(ValueError, match="different run"):
        ResultWriter(str(tmp_path / "out"), D + 1, meta={"model": "x"})
    # expression-only sink (cohort runs)
    w3 = ResultWriter(str(tmp_path / "expr"), D, embeddings=False)
    w3.submit(5, [(0, "S", "g", 1)], np.ones(1, np.float32), None)
    w3.close()
    assert list(w3.read_all().columns) == ["slab", "item", "sample", "gene_id", "tissue", "predicted_expression"]
