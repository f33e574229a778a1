"""CPU checks of the BPE merge schedule the stage-1 kernels rely on (variantformer_b200/stage1.py: merge_batches): the
batches are symbol-disjoint, interacting ranks keep their order, and applying the batches — each one as ONE
simultaneous, order-free pass — gives the tokens of the rank-by-rank sweeps (the reference tokenizer's semantics:
utils/data_process.py BPE encoding through the `tokenizers` merge table)."""
import numpy as np

from variantformer_b200.stage1 import load_merge_table, merge_batches


def _apply_rank(sym, a, b, c):
    """One rank, left to right, non-overlapping (the classic BPE sweep)."""
    out, i = [], 0
    while i < len(sym):
        if i + 1 < len(sym) and sym[i] == a and sym[i + 1] == b:
            out.append(c); i += 2
        else:
            out.append(sym[i]); i += 1
    return out


def _apply_batch_simultaneously(sym, ranks, a, b, c, rng):
    """All ranks of a symbol-disjoint batch at once: every adjacent pair that matches some rank of the batch is merged;
    the pairs are visited in RANDOM order (the kernel's threads race), which must not matter."""
    table = {(int(a[r]), int(b[r])): int(c[r]) for r in ranks}
    hits = [i for i in range(len(sym) - 1) if (sym[i], sym[i + 1]) in table]
    sym = list(sym); dead = [False] * len(sym)
    for i in rng.permutation(hits):
        assert not dead[i] and not dead[i + 1]             # disjoint symbol sets: matches never share a slot
        sym[i] = table[(sym[i], sym[i + 1])]; dead[i + 1] = True
    return [s for s, d in zip(sym, dead) if not d]


def test_batches_are_symbol_disjoint_and_keep_interacting_ranks_in_order():
    a, b, c, _ = load_merge_table()
    order, bid = merge_batches(a, b, c)
    assert sorted(order.tolist()) == list(range(len(a)))
    assert (np.diff(bid.astype(int)) >= 0).all() and np.bincount(bid).max() <= 16
    pa, pb, pc = a[order], b[order], c[order]
    for k in np.unique(bid):
        rs = np.nonzero(bid == k)[0]
        syms = [s for r in rs for s in (int(pa[r]), int(pb[r]), int(pc[r]))]
        if len(rs) > 1:
            assert len(set(syms)) == len(syms), f"batch {k} is not symbol-disjoint"
            assert all(pa[r] != pb[r] for r in rs), "a self pair must be a batch of its own"
    pos = np.empty(len(a), int); pos[order] = np.arange(len(a))
    for r in range(len(a)):
        sr = {int(a[r]), int(b[r]), int(c[r])}
        for r2 in range(r + 1, len(a)):
            if sr & {int(a[r2]), int(b[r2]), int(c[r2])}:
                assert bid[pos[r]] < bid[pos[r2]], (r, r2)


def test_batched_schedule_gives_the_tokens_of_the_rank_by_rank_sweeps():
    a, b, c, vocab = load_merge_table()
    order, bid = merge_batches(a, b, c)
    rng = np.random.default_rng(5)
    base = [vocab[ch] for ch in "ACGT"]
    words = [rng.choice(base, n, p=p).tolist() for n, p in ((300, None), (700, [0.4, 0.1, 0.1, 0.4]), (1200, [0.7, 0.1, 0.1, 0.1]))]
    words.append([base[0]] * 257 + [base[0], base[1], base[2], base[3]] * 50 + [base[3]] * 130)
    for w in words:
        want = list(w)
        for r in range(len(a)):
            want = _apply_rank(want, int(a[r]), int(b[r]), int(c[r]))
        got = list(w)
        for k in np.unique(bid):
            ranks = order[bid == k]
            if len(ranks) == 1:                              # (self pairs: the sequential sweep is the definition)
                r = int(ranks[0]); got = _apply_rank(got, int(a[r]), int(b[r]), int(c[r]))
            else:
                got = _apply_batch_simultaneously(got, ranks, a, b, c, rng)
        assert got == want
