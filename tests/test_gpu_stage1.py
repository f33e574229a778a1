"""Stage-1 kernels vs the CPU oracle (bit-exact) through the C ABI — GPU box."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import stage1 as O  # noqa: E402  (checker only)
from variantformer_b200 import ops  # noqa: E402
from variantformer_b200.stage1 import Genome, SampleVariants, WindowTokenizer, cre_window, gene_window  # noqa: E402
from variantformer_b200.utils import synth  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage1_golden.npz"), allow_pickle=True)
DEV = "cuda"


def _pack(seqs):
    lens = np.array([len(s) for s in seqs], np.int32)
    pitch = max(16, ((int(lens.max()) + 15) // 16) * 16)
    buf = np.zeros((len(seqs), pitch), np.uint8)
    for i, s in enumerate(seqs):
        buf[i, :len(s)] = np.frombuffer(s.encode(), np.uint8)
    return torch.from_numpy(buf).to(DEV), torch.from_numpy(lens).to(DEV), int(lens.max())


def test_bpe_golden_vectors_bit_exact():
    tk = WindowTokenizer(DEV)
    seqs = [str(s) for s in G["seqs"]]
    flat, off = G["tok_flat"].astype(np.int64), G["tok_off"]
    short = [i for i, s in enumerate(seqs) if len(s) <= 8192]
    long_ = [i for i, s in enumerate(seqs) if len(s) > 8192]
    for group in (short, long_):
        buf, lens, mx = _pack([seqs[i] for i in group])
        cap = max(mx, 1)
        tok, cnt, starts = ops.bpe_tokenize(buf, lens, max(mx, 1), tk.merges, cap, cap, want_starts=True)
        tok, cnt, starts = tok.cpu().numpy(), cnt.cpu().numpy(), starts.cpu().numpy()
        bpe = O.OracleBPE()
        for k, i in enumerate(group):
            want = flat[off[i]:off[i + 1]]
            assert cnt[k] == len(want), f"seq {i}: count {cnt[k]} != {len(want)}"
            assert (tok[k, :len(want)] == want).all(), f"seq {i} (len {len(seqs[i])}) tokens differ"
            assert (tok[k, len(want):] == 0).all()
            _, st = bpe.encode_with_starts(seqs[i])
            assert (starts[k, :len(want)] == st).all(), f"seq {i} token starts differ"


def test_bpe_rank_batches_do_not_change_a_single_token():
    """Applying the ranks of a batch (pairwise disjoint symbol sets, no self pair) in one sweep must give exactly the
    tokens of the rank-by-rank sweeps: golden sequences, homopolymer / repeat-rich words, and 301 kb windows through
    the cluster kernel, batched vs unbatched (merge_batch = NULL) vs the oracle."""
    from variantformer_b200.stage1 import load_merge_table, merge_batches
    tk = WindowTokenizer(DEV)
    a, b, c, _ = load_merge_table()
    order, bid = merge_batches(a, b, c)
    assert sorted(order.tolist()) == list(range(len(a))) and len(np.unique(bid)) < 130
    assert (np.diff(bid.astype(int)) >= 0).all() and np.bincount(bid).max() <= 16
    pa, pb, pc = a[order], b[order], c[order]
    for k in np.unique(bid):                                    # the properties the kernel relies on
        rs = np.nonzero(bid == k)[0]
        syms = [s for r in rs for s in (int(pa[r]), int(pb[r]), int(pc[r]))]
        assert len(rs) == 1 or (len(set(syms)) == len(syms)), f"batch {k} is not symbol-disjoint"
        assert len(rs) == 1 or all(pa[r] != pb[r] for r in rs)
    pos = np.empty(len(a), int); pos[order] = np.arange(len(a))  # interacting ranks keep their relative order
    for r in range(len(a)):
        sr = {int(a[r]), int(b[r]), int(c[r])}
        for r2 in range(r + 1, min(len(a), r + 60)):
            if sr & {int(a[r2]), int(b[r2]), int(c[r2])}:
                assert bid[pos[r]] < bid[pos[r2]], (r, r2)
    rng = np.random.default_rng(17)
    words = [str(s) for s in G["seqs"]][:60]
    words += ["".join(rng.choice(list("ACGT"), n, p=[0.4, 0.1, 0.1, 0.4])) for n in (500, 4000, 8000)]
    words += ["A" * 700 + "ACGT" * 300 + "T" * 513 + "CA" * 400, "".join(rng.choice(list("ACGTRYN"), 3000))]
    long_words = ["".join(rng.choice(list("ACGT"), 301_000)), "AC" * 60_000 + "".join(rng.choice(list("ACGTN"), 150_000)) + "T" * 31_000]
    bpe = O.OracleBPE()
    for group in (words, long_words):
        buf, lens, mx = _pack(group)
        cap = mx
        t1, c1 = ops.bpe_tokenize(buf, lens, mx, tk.merges, cap, cap)
        t0, c0 = ops.bpe_tokenize(buf, lens, mx, tk.merges[:3], cap, cap)
        torch.cuda.synchronize()
        assert torch.equal(c0, c1) and torch.equal(t0, t1)
        for i, w in enumerate(group):
            want = bpe.encode(w)
            assert int(c1[i]) == len(want) and (t1[i, :len(want)].cpu().numpy() == want).all()


def test_bpe_fixed_and_chunked_shapes():
    tk = WindowTokenizer(DEV)
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACGT"), n)) for n in (30, 350, 450, 1000, 3000)]
    buf, lens, mx = _pack(seqs)
    tok, mask, cnt = tk.tokenize_fixed(buf, lens, mx)
    bpe = O.OracleBPE()
    for i, s in enumerate(seqs):
        o, m = O.adjust_length(bpe.encode(s), 200)
        assert (tok[i].cpu().numpy() == o).all() and (mask[i].cpu().numpy() == m).all()
    gene = "".join(rng.choice(list("ACGT"), 20000)) + "NNNN" + "".join(rng.choice(list("acgt"), 9000))
    buf, lens, mx = _pack([gene, gene[:5000]])
    chunks, cnt = tk.tokenize_chunked(buf, lens, mx)
    for (t, m), s in zip(chunks, [gene, gene[:5000]]):
        o, om = O.chunkify(bpe.encode(s), 200, 200)
        assert t.shape == o.shape and (t.cpu().numpy() == o).all() and (m.cpu().numpy() == om).all()
    tk2 = WindowTokenizer(DEV, max_chunks=3)
    chunks, _ = tk2.tokenize_chunked(buf, lens, mx)
    o, om = O.chunkify(bpe.encode(gene), 200, 3)
    assert chunks[0][0].shape == (3, 200) and (chunks[0][0].cpu().numpy() == o).all()


@pytest.mark.parametrize("snp_only", [False, True])
def test_encode_windows_matches_oracle(snp_only):
    rng = np.random.default_rng(1234)
    chroms = {"chr1": synth.make_chromosome(rng, 2_000_000), "chr2": synth.make_chromosome(rng, 1_500_000)}
    var = {c: synth.make_variants(rng, s) for c, s in chroms.items()}
    genome = Genome.from_arrays(chroms, DEV)
    sv = SampleVariants(var, DEV)
    tk = WindowTokenizer(DEV)
    names, w0, w1, rc = [], [], [], []
    for c, s in chroms.items():
        for _ in range(150):
            a = int(rng.integers(0, len(s) - 400)); a0, a1 = cre_window(a, a + int(rng.integers(150, 351)), 50)
            names.append(c); w0.append(a0); w1.append(a1); rc.append(int(rng.integers(0, 2)))
    for strand, (gs, ge) in (("+", (400_000, 460_000)), ("-", (700_000, 1_250_000)), ("+", (5_000, 390_000))):
        a0, a1 = gene_window(gs, ge, strand, 1000, 300000)
        names.append("chr1"); w0.append(a0); w1.append(a1); rc.append(int(strand == "-"))
    for use_var in (True, False):
        out, out_len, err = tk.sequences(genome, names, w0, w1, rc, sv if use_var else None, snp_only=snp_only)
        out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
        assert int(err.item()) == 0
        for i, c in enumerate(names):
            v = var[c]
            lens = np.array([len(a) for a in v["alt"]], np.int32)
            pool = np.frombuffer(b"".join(v["alt"]), np.uint8)
            if use_var:
                want = O.apply_variants(chroms[c], w0[i], w1[i], v["pos"], v["ref_len"], np.cumsum(lens) - lens, lens,
                                        v["gt"], pool, snp_only=snp_only)
            else:
                want = chroms[c][w0[i]:w1[i]].tobytes()
            if rc[i]:
                want = O.reverse_complement(want)
            got = out[i, :out_len[i]].tobytes()
            assert got == want, f"window {i} {c}:{w0[i]}-{w1[i]} rc={rc[i]} var={use_var}"


def test_empty_and_all_n_windows():
    tk = WindowTokenizer(DEV)
    buf, lens, mx = _pack(["NNNNNNNN", "A", "nnnnACGTnnnn"])
    lens[1] = 0                                              # empty window
    tok, mask, cnt = tk.tokenize_fixed(buf, lens, mx)
    want = O.OracleBPE().encode("nnnnACGTnnnn").tolist()            # [ACG, T]
    assert cnt.cpu().tolist() == [0, 0, len(want)] and mask[0].all() and mask[1].all()
    assert tok[2, :len(want)].cpu().tolist() == want
    assert (tok[:2] == 0).all()
