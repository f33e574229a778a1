"""Per-kernel parity tests (GPU box).  Each CUDA kernel is called through the C ABI
(variantformer_b200.ops -> libvf_b200.so) and compared with a plain torch fp32
reference of the same op on the same bf16-rounded operands."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from variantformer_b200 import ops  # noqa: E402
from variantformer_b200._lib import (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GEGLU_BF16, EPI_BIAS_GELU_BF16,  # noqa: E402
                                     EPI_BIAS_RESID_F32)
from variantformer_b200.engine import interleave_geglu  # noqa: E402

DEV = "cuda"


def _describe_mismatch(got, want, tol):
    err = (got - want).abs()
    bad = err > tol
    rows = bad.any(1).nonzero().flatten().tolist()
    cols = bad.any(0).nonzero().flatten().tolist()
    return (f"max err {err.max().item():.4g} (tol {tol:.3g}), {int(bad.sum())} bad of {bad.numel()}; "
            f"bad rows[:16]={rows[:16]} (n={len(rows)}) bad cols[:16]={cols[:16]} (n={len(cols)}); "
            f"got[0,:4]={got[0, :4].tolist()} want[0,:4]={want[0, :4].tolist()}")


GEMM_SHAPES = [(128, 256, 64), (128, 256, 512), (300, 576, 192), (9, 3072, 1536), (1000, 512, 1024),
               (4096, 4608, 1536), (12663, 1536, 1536), (777, 1536, 512),
               # M >= 4096: the residual epilogue runs as the split (reader / writer warp pairs) kernel — single CTA (K < 1024)
               # and CTA pair, a ragged last row block, N that leaves slabs of the last column tile empty
               (5001, 512, 512), (4500, 576, 192), (20000, 1536, 1024), (4097, 328, 1024)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_RESID_F32, EPI_BIAS_GELU_BF16])
def test_gemm(M, N, K, epi):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    resid = torch.randn(M, N, generator=g).to(DEV) if epi == EPI_BIAS_RESID_F32 else None
    want = a.float() @ w.float().t() + bias
    if resid is not None:
        want = want + resid
    if epi == EPI_BIAS_GELU_BF16:
        want = F.gelu(want)
    out2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV) if epi in (EPI_BIAS_F32, EPI_BIAS_RESID_F32) else None
    got = ops.gemm(a, w, epi, bias=bias, resid=resid, out2=out2)
    torch.cuda.synchronize()
    tol = 2e-2 if got.dtype == torch.bfloat16 else 2e-3
    assert torch.allclose(got.float(), want, atol=tol, rtol=tol), _describe_mismatch(got.float(), want, tol)
    if out2 is not None:
        assert torch.allclose(out2.float(), want, atol=2e-2, rtol=2e-2), _describe_mismatch(out2.float(), want, 2e-2)


@pytest.mark.parametrize("M,K", [(128, 512), (1000, 1536), (333, 192), (2500, 1536), (1300, 1024)])
def test_gemm_geglu(M, K):
    N = 2048
    g = torch.Generator(device="cpu").manual_seed(M + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    y = a.float() @ w.float().t() + bias
    u, gate = y.chunk(2, -1)
    want = u * F.gelu(gate)
    got = ops.gemm(a, interleave_geglu(w), EPI_BIAS_GEGLU_BF16, bias=interleave_geglu(bias))
    torch.cuda.synchronize()
    assert got.shape == (M, N // 2)
    assert torch.allclose(got.float(), want, atol=3e-2, rtol=2e-2), _describe_mismatch(got.float(), want, 3e-2)


def test_gemm_inplace_residual_and_strided_output():
    M, N, K = 515, 512, 1024
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / 32).to(DEV).bfloat16()
    x = torch.randn(M, N, generator=g).to(DEV)
    want = a.float() @ w.float().t() + x
    ops.gemm(a, w, EPI_BIAS_RESID_F32, resid=x, out=x)
    torch.cuda.synchronize()
    assert torch.allclose(x, want, atol=2e-3, rtol=2e-3), _describe_mismatch(x, want, 2e-3)
    big = torch.zeros(M, 3 * N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, EPI_BIAS_BF16, out=big[:, N:2 * N])
    torch.cuda.synchronize()
    assert torch.allclose(big[:, N:2 * N].float(), a.float() @ w.float().t(), atol=3e-2, rtol=2e-2)
    assert big[:, :N].abs().max() == 0 and big[:, 2 * N:].abs().max() == 0


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 1536, 1536), (4099, 4608, 1536), (77, 256, 192), (9000, 512, 512),
                                   (4200, 1536, 1024)])
def test_gemm_stats_out_and_layernorm_fold(M, N, K):
    """The epilogue of an fp32 GEMM accumulates each output row's (sum, sum of squares); a following GEMM on the bf16
    mirror with gamma/beta folded into W/bias reproduces Linear(LayerNorm(x)) (layers.py:116-163)."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    resid = (torch.randn(M, N, generator=g) * 0.7 + 0.3).to(DEV)
    x = torch.empty(M, N, device=DEV); xb = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    st = torch.full((M, ops.stats_parts(N), 2), 7.0, device=DEV)           # every partial is overwritten
    ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=resid, out=x, out2=xb, stats_out=st)
    torch.cuda.synchronize()
    want_x = a.float() @ w.float().t() + bias + resid
    assert torch.allclose(x, want_x, atol=2e-3, rtol=2e-3)
    tot = st.sum(1)
    assert torch.allclose(tot[:, 0], x.sum(1), atol=2e-2, rtol=1e-4), (tot[:4, 0], x.sum(1)[:4])
    assert torch.allclose(tot[:, 1], (x * x).sum(1), atol=2e-2, rtol=1e-4)
    st2 = ops.rowstats(x)                                                   # the stand-alone statistics kernel agrees
    torch.cuda.synchronize()
    assert torch.allclose(st2[:, 0], tot, atol=2e-2, rtol=1e-4)
    st_again = torch.empty_like(st)                                         # plain stores in a fixed order: bit-reproducible
    ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=resid, out=torch.empty_like(x), stats_out=st_again)
    torch.cuda.synchronize()
    assert torch.equal(st, st_again)
    # consumer: Linear(LayerNorm(x)) with a [N2, N] weight
    N2 = 512
    gamma = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV); beta = (0.1 * torch.randn(N, generator=g)).to(DEV)
    w2 = (torch.randn(N2, N, generator=g) / math.sqrt(N)).to(DEV); b2 = torch.randn(N2, generator=g).to(DEV)
    want = F.linear(F.layer_norm(x, (N,), gamma, beta, 1e-5), w2, b2)
    wf = (w2 * gamma[None, :]).bfloat16(); cs = wf.float().sum(1).contiguous(); bf = (b2 + w2 @ beta).contiguous()
    got = ops.gemm(xb, wf.contiguous(), EPI_BIAS_BF16, bias=bf, ln=(st, cs, N, 1e-5))
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=4e-2, rtol=2e-2), _describe_mismatch(got.float(), want, 4e-2)
    # GeGLU consumer (tile-interleaved weight / bias / column sums)
    w3 = (torch.randn(2048, N, generator=g) / math.sqrt(N)).to(DEV); b3 = torch.randn(2048, generator=g).to(DEV)
    y = F.linear(F.layer_norm(x, (N,), gamma, beta, 1e-5), w3, b3)
    u, gate = y.chunk(2, -1)
    want3 = u * F.gelu(gate)
    wf3 = (w3 * gamma[None, :]).bfloat16(); cs3 = wf3.float().sum(1); bf3 = b3 + w3 @ beta
    got3 = ops.gemm(xb, interleave_geglu(wf3), EPI_BIAS_GEGLU_BF16, bias=interleave_geglu(bf3),
                    ln=(st, interleave_geglu(cs3), N, 1e-5))
    torch.cuda.synchronize()
    assert torch.allclose(got3.float(), want3, atol=5e-2, rtol=3e-2), _describe_mismatch(got3.float(), want3, 5e-2)


@pytest.mark.parametrize("M,N,K", [(515, 512, 512), (3000, 1536, 1024), (1301, 768, 512), (2049, 512, 1536), (1100, 1280, 512)])
def test_gemm_bf16_residual_and_mirror_only(M, N, K):
    """Intra-layer temporaries: x1' = x1 + Linear(a) with x1 held in bf16, in place, only mirror + statistics written."""
    g = torch.Generator(device="cpu").manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    x1 = torch.randn(M, N, generator=g).to(DEV).bfloat16()
    want = a.float() @ w.float().t() + bias + x1.float()
    st = torch.empty(M, ops.stats_parts(N), 2, device=DEV)
    got = ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x1, out2=x1, stats_out=st, mirror_only=True)
    torch.cuda.synchronize()
    assert got.data_ptr() == x1.data_ptr()
    assert torch.allclose(x1.float(), want, atol=3e-2, rtol=2e-2), _describe_mismatch(x1.float(), want, 3e-2)
    tot = st.sum(1)                                        # statistics are taken from the fp32 values before rounding
    assert torch.allclose(tot[:, 0], want.sum(1), atol=5e-2, rtol=1e-3) and torch.allclose(tot[:, 1], (want * want).sum(1), rtol=2e-3)


@pytest.mark.parametrize("form", ["ffn2", "out_proj"])
def test_gemm_residual_rows_do_not_depend_on_how_many_rows_are_batched(form):
    """Which kernel a residual GEMM runs through depends on M (single CTA below 1024 rows, CTA pairs with the TMA-staged
    epilogues above): values, mirror and row statistics of a row must be bit-identical either way, or results would depend
    on how genes are batched (and the multi-GPU gather check could not be bit-exact)."""
    g = torch.Generator(device="cpu").manual_seed(77)
    M, N, K = 2304, 768, 512
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    x0 = (torch.randn(M, N, generator=g) * 3).to(DEV)

    def run(rows):
        st = torch.empty(rows, ops.stats_parts(N), 2, device=DEV)
        if form == "ffn2":
            x = x0[:rows].clone(); xb = torch.empty(rows, N, device=DEV, dtype=torch.bfloat16)
            ops.gemm(a[:rows].contiguous(), w, EPI_BIAS_RESID_F32, bias=bias, resid=x, out=x, out2=xb, stats_out=st)
            return x, xb, st
        xb = x0[:rows].bfloat16().contiguous()
        ops.gemm(a[:rows].contiguous(), w, EPI_BIAS_RESID_F32, bias=bias, resid=xb, out2=xb, stats_out=st, mirror_only=True)
        return xb, xb, st
    big = run(M)                                          # CTA pairs (TMA-staged epilogue)
    small = run(640)                                      # single-CTA kernel (register-prefetch epilogue)
    torch.cuda.synchronize()
    for b, s_ in zip(big, small):
        assert torch.equal(b[:640], s_)


def test_rowstats_mirror():
    x = torch.randn(1001, 1536, device=DEV) * 2 + 0.5
    xb = torch.empty(1001, 1536, dtype=torch.bfloat16, device=DEV)
    st = ops.rowstats(x, None, xb)
    torch.cuda.synchronize()
    assert torch.equal(xb, x.bfloat16())
    assert st.shape == (1001, 1, 2)
    assert torch.allclose(st[:, 0, 0], x.sum(1), rtol=1e-5, atol=1e-2) and torch.allclose(st[:, 0, 1], (x * x).sum(1), rtol=1e-5)


def test_center_uncenter_and_pivoted_meanpool():
    M, d = 1001, 1536
    g = torch.Generator(device="cpu").manual_seed(4)
    x0 = (torch.randn(M, d, generator=g) * 0.7 + torch.randn(M, 1, generator=g) * 9).to(DEV)
    x = x0.clone(); xb = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    piv, st = ops.center_rows(x, out_bf16=xb)
    torch.cuda.synchronize()
    assert torch.allclose(piv, x0.mean(1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(x, x0 - x0.mean(1, keepdim=True), atol=2e-5, rtol=1e-5)
    assert torch.equal(xb, x.bfloat16())
    assert torch.allclose(st[:, 0, 0], x.sum(1), atol=1e-2) and torch.allclose(st[:, 0, 1], (x * x).sum(1), rtol=1e-4)
    back, back_bf = ops.uncenter_rows(x, piv, want_bf16=True)
    assert torch.allclose(back, x0, atol=2e-5, rtol=1e-5) and torch.equal(back_bf, back.bfloat16())
    idx = torch.tensor([5, 0, 1000, 17], dtype=torch.int32, device=DEV)
    rows, _ = ops.gather_rows(x, None, idx)
    raw, _ = ops.uncenter_rows(rows, piv, idx=idx)
    assert torch.allclose(raw, x0[idx.long()], atol=2e-5, rtol=1e-5)
    lens = [3, 500, 1, 497]
    cu = ops.cu_seqlens(lens, DEV)
    _, pooled = ops.masked_meanpool(x, cu, len(lens), want_f32=True, pivot=piv)
    want = torch.stack([x0[a:b].mean(0) for a, b in zip(np.cumsum([0] + lens[:-1]), np.cumsum(lens))])
    torch.cuda.synchronize()
    assert torch.allclose(pooled, want, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("mean_mult,outlier", [(20.0, 0.0), (-35.0, 0.0), (20.0, 50.0), (0.0, 50.0)])
def test_layernorm_fold_on_adversarial_streams(mean_mult, outlier):
    """Trained checkpoints have rows with |mean| >> std and a few outlier channels.  The LayerNorm fold reads the bf16
    mirror of the UN-normalised row, so the stream is kept row-centred (vf_center_rows): Linear(LayerNorm(x)) through
    center_rows + vf_gemm_bf16_ln must stay as close to the fp32 reference as the reference's own bf16 path
    (fp32 LayerNorm -> bf16 -> Linear), here and after a residual GEMM has produced the next row statistics
    (two stacked LayerNorm-fold consumers).  Tolerance: max |err| / max |want| <= 6e-3 (the reference's bf16 path sits
    at 2-4e-3 on these streams; the un-centred mirror is at 2.6e-2 for mean = 20 std)."""
    M, d, N = 2048, 1536, 512
    g = torch.Generator(device="cpu").manual_seed(int(abs(mean_mult)) + int(outlier))
    x0 = torch.randn(M, d, generator=g) + mean_mult * torch.where(torch.rand(M, 1, generator=g) < 0.5, -1.0, 1.0)
    if outlier:
        x0[:, torch.randint(0, d, (6,), generator=g)] *= outlier
    x0 = x0.to(DEV)
    gamma = (1 + 0.2 * torch.randn(d, generator=g)).to(DEV); beta = (0.1 * torch.randn(d, generator=g)).to(DEV)
    w = (torch.randn(N, d, generator=g) / math.sqrt(d)).to(DEV); b = torch.randn(N, generator=g).to(DEV)

    def folded(w_, b_, gam, bet):
        wf = (w_ * gam[None, :]).bfloat16()
        return wf.contiguous(), wf.float().sum(1).contiguous(), (b_ + w_ @ bet).contiguous()

    def err(got, want):
        return float((got.float() - want).abs().max() / want.abs().max())
    wf, cs, bf = folded(w, b, gamma, beta)
    want = F.linear(F.layer_norm(x0.double(), (d,), gamma.double(), beta.double(), 1e-5), w.double(), b.double()).float()
    x = x0.clone(); xb = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    piv, st = ops.center_rows(x, out_bf16=xb)
    got = ops.gemm(xb, wf, EPI_BIAS_BF16, bias=bf, ln=(st, cs, d, 1e-5))
    torch.cuda.synchronize()
    e1 = err(got, want)
    assert e1 <= 6e-3, e1
    if outlier == 0.0:                 # the control: the raw mirror loses the normalised signal's bits (the test has teeth)
        xb_raw = x0.bfloat16(); st_raw = ops.rowstats(x0)
        e_raw = err(ops.gemm(xb_raw, wf, EPI_BIAS_BF16, bias=bf, ln=(st_raw, cs, d, 1e-5)), want)
        assert e_raw > 3 * e1, (e_raw, e1)
    # second layer: x2 = x + f W2^T (+ b2) from a residual epilogue that also writes the mirror and the row statistics
    K2 = 1024
    f = (torch.randn(M, K2, generator=g) * 0.5).to(DEV).bfloat16()
    w2 = (torch.randn(d, K2, generator=g) / math.sqrt(K2)).to(DEV).bfloat16(); b2 = torch.randn(d, generator=g).to(DEV)
    st2 = torch.empty(M, ops.stats_parts(d), 2, device=DEV)
    ops.gemm(f, w2, EPI_BIAS_RESID_F32, bias=b2, resid=x, out=x, out2=xb, stats_out=st2)
    x2_raw = x0.double() + f.double() @ w2.double().t() + b2.double()
    gamma2 = (1 + 0.2 * torch.randn(d, generator=g)).to(DEV); beta2 = (0.1 * torch.randn(d, generator=g)).to(DEV)
    wf2, cs2, bf2 = folded(w, b, gamma2, beta2)
    want2 = F.linear(F.layer_norm(x2_raw, (d,), gamma2.double(), beta2.double(), 1e-5), w.double(), b.double()).float()
    got2 = ops.gemm(xb, wf2, EPI_BIAS_BF16, bias=bf2, ln=(st2, cs2, d, 1e-5))
    raw2, _ = ops.uncenter_rows(x, piv)
    torch.cuda.synchronize()
    assert torch.allclose(raw2, x2_raw.float(), atol=2e-3 * max(1.0, abs(mean_mult), outlier), rtol=1e-5)
    e2 = err(got2, want2)
    assert e2 <= 6e-3, e2


def _ref_attention(q, k, v, lens_q, lens_k, H, hd, slopes):
    out = torch.empty(q.shape[0], H * hd, device=q.device)
    qs = ks = 0
    for sq, sk in zip(lens_q, lens_k):
        qq = q[qs:qs + sq].float().view(sq, H, hd); kk = k[ks:ks + sk].float().view(sk, H, hd)
        vv = v[ks:ks + sk].float().view(sk, H, hd)
        s = torch.einsum("thd,shd->hts", qq, kk) / math.sqrt(hd)
        if slopes is not None:
            i = torch.arange(sq, device=q.device)[:, None]; j = torch.arange(sk, device=q.device)[None, :]
            s = s - slopes[:, None, None] * (i + sk - sq - j).abs()[None]
        out[qs:qs + sq] = torch.einsum("hts,shd->thd", s.softmax(-1), vv).reshape(sq, H * hd)
        qs += sq; ks += sk
    return out


@pytest.mark.parametrize("H,hd,alibi,lens", [
    (4, 48, True, [201] * 5), (4, 48, False, [1, 97, 128, 129, 64, 65, 7, 200, 33]), (32, 48, True, [201, 640, 33, 1024]),
    (8, 64, False, [97, 130, 200, 12, 128, 77, 200]), (2, 64, True, [300, 5, 257]), (4, 48, True, [1000]),
])
def test_mc_self_attention(H, hd, alibi, lens):
    n, d = sum(lens), H * hd
    g = torch.Generator(device="cpu").manual_seed(H * hd + len(lens))
    qkv = torch.randn(n, 3 * d, generator=g).to(DEV).bfloat16()
    slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
    slots = ops.SlotMap(lens, DEV)
    got = ops.attention_mc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], slots, H, hd, slopes)
    want = _ref_attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], lens, lens, H, hd, slopes)
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=2e-2, rtol=2e-2), _describe_mismatch(got.float(), want, 2e-2)


@pytest.mark.parametrize("H,hd", [(4, 48), (32, 48), (4, 64)])
def test_mc_cross_attention_stacked_queries(H, hd):
    lens_q, lens_k = [603, 201, 1500, 128], [300, 64, 1024, 1]
    d = H * hd
    g = torch.Generator(device="cpu").manual_seed(17 + H)
    q = torch.randn(sum(lens_q), d, generator=g).to(DEV).bfloat16()
    kv = torch.randn(sum(lens_k), 2 * d, generator=g).to(DEV).bfloat16()
    slots = ops.SlotMap(lens_q, DEV, k_lens=lens_k)
    got = ops.attention_mc(q, kv[:, :d], kv[:, d:], slots, H, hd, None)
    want = _ref_attention(q, kv[:, :d], kv[:, d:], lens_q, lens_k, H, hd, None)
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=2e-2, rtol=2e-2), _describe_mismatch(got.float(), want, 2e-2)


@pytest.mark.parametrize("H,hd,alibi,n_seq,lo,hi", [(8, 64, False, 1200, 60, 128), (8, 64, True, 700, 1, 200),
                                                    (32, 48, True, 60, 150, 420), (32, 48, False, 40, 1, 700)])
def test_mc_many_items_per_cta(H, hd, alibi, n_seq, lo, hi):
    """More work items than resident CTAs: every CTA walks a long stream of items (ring / phase bookkeeping across
    items, shared and split K/V streams mixed)."""
    rng = np.random.default_rng(n_seq + hi)
    lens = rng.integers(lo, hi + 1, n_seq).tolist()
    n, d = sum(lens), H * hd
    g = torch.Generator(device="cpu").manual_seed(n_seq)
    qkv = torch.randn(n, 3 * d, generator=g).to(DEV).bfloat16()
    slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
    slots = ops.SlotMap(lens, DEV)
    assert slots.n_items * H > 2 * 148 * 3
    got = ops.attention_mc(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], slots, H, hd, slopes)
    torch.cuda.synchronize()
    want = _ref_attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], lens, lens, H, hd, slopes)
    assert torch.allclose(got.float(), want, atol=2e-2, rtol=2e-2), _describe_mismatch(got.float(), want, 2e-2)


_RAGGED = np.random.default_rng(3).integers(1, 201, 3000).tolist()


@pytest.mark.parametrize("lens,H,hd,alibi", [([97] * 2048, 8, 64, False), ([201] * 300 + [97] * 301, 32, 48, True),
                                             (_RAGGED, 8, 64, False)])
def test_mc_repeated_launches_are_bit_identical(lens, H, hd, alibi):
    """The kernel is deterministic: the same launch repeated next to unrelated traffic on another stream must give the
    same bits every time, and those bits must be right.  This is the check that caught round 1's cross-slot parity alias
    on split-key work items (two tiles with different key ranges in one item); such items are built again now that the
    issuers order their waits behind the producer's issue counters.  Long form: tools/stress_attention.py."""
    n, d = sum(lens), H * hd
    g = torch.Generator(device="cpu").manual_seed(11)
    q, k, v = (torch.randn(n, d, generator=g).to(DEV).bfloat16() for _ in range(3))
    slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
    slots = ops.SlotMap(lens, DEV)
    tab = slots.table.cpu().numpy()
    both = (tab[:, 0, 1] > 0) & (tab[:, 1, 1] > 0)
    if ops.PAIR_UNRELATED_TILES:
        assert np.any(tab[both, 0, 2] != tab[both, 1, 2]), "the case must contain split-key items"
    ref = ops.attention_mc(q, k, v, slots, H, hd, slopes).clone()
    want = _ref_attention(q, k, v, lens, lens, H, hd, slopes)
    assert torch.allclose(ref.float(), want, atol=2e-2, rtol=2e-2), _describe_mismatch(ref.float(), want, 2e-2)
    out = torch.empty_like(ref)
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    side = torch.cuda.Stream()
    a = torch.randn(2048, 2048, device=DEV)
    for i in range(400):
        if i % 5 == 0:
            with torch.cuda.stream(side):
                (a @ a).sum()
        ops.attention_mc(q, k, v, slots, H, hd, slopes, out=out)
        bad += (out.view(torch.int16) != ref.view(torch.int16)).any().to(torch.int64)
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_mc_tail_rows_and_keyless_sequences_do_not_leak():
    """A sequence's last K/V block over-fetches rows of its neighbour: NaN there (an all-N window upstream) must not
    reach this sequence's output; a sequence without keys gets zeros, never stale buffer contents."""
    H, hd = 4, 48
    d = H * hd
    lens_q, lens_k = [130, 40, 64, 9], [70, 0, 130, 3]
    g = torch.Generator(device="cpu").manual_seed(8)
    q = torch.randn(sum(lens_q), d, generator=g).to(DEV).bfloat16()
    k = torch.randn(sum(lens_k), d, generator=g).to(DEV).bfloat16()
    v = torch.randn(sum(lens_k), d, generator=g).to(DEV).bfloat16()
    want = _ref_attention(q, k, v, lens_q, lens_k, H, hd, None)                      # computed without the poison
    k[70:80] = float("nan"); v[70:80] = float("nan")                                 # first rows of the third sequence
    out = torch.full((sum(lens_q), d), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.attention_mc(q, k, v, ops.SlotMap(lens_q, DEV, k_lens=lens_k), H, hd, None, out=out)
    torch.cuda.synchronize()
    assert torch.allclose(out[:130].float(), want[:130], atol=2e-2, rtol=2e-2)       # neighbour's NaN rows not seen
    assert bool((out[130:170] == 0).all())                                           # keyless sequence
    assert bool(torch.isnan(out[170:234]).all())                                     # its own NaN keys do poison it
    assert torch.allclose(out[234:].float(), want[234:], atol=2e-2, rtol=2e-2)


def test_mc_output_of_a_sequence_does_not_depend_on_its_neighbours():
    """Batching invariance down to the bit: the rows of a tile past a sequence's last query row belong to whatever
    follows it in the q tensor; they must not influence the live rows (they once voted on the lazy raise of the
    reference maximum, which changes the bf16 rounding of the probabilities — found by the 8-rank gather check)."""
    H, hd = 4, 48
    d = H * hd
    g = torch.Generator(device="cpu").manual_seed(31)
    la, lb = 201, 333
    qa = (torch.randn(la, d, generator=g) * 2).to(DEV).bfloat16()
    ka = (torch.randn(la, d, generator=g) * 2 * torch.linspace(0.3, 2.5, la)[:, None]).to(DEV).bfloat16()   # scores grow
    va = torch.randn(la, d, generator=g).to(DEV).bfloat16()
    slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV)
    outs = []
    for scale in (0.1, 30.0):                         # a tame and a wild neighbour behind sequence A
        qb = (torch.randn(lb, d, generator=g) * scale).to(DEV).bfloat16()
        kb = (torch.randn(lb, d, generator=g) * scale).to(DEV).bfloat16()
        vb = torch.randn(lb, d, generator=g).to(DEV).bfloat16()
        q, k, v = torch.cat([qa, qb]), torch.cat([ka, kb]), torch.cat([va, vb])
        outs.append(ops.attention_mc(q, k, v, ops.SlotMap([la, lb], DEV), H, hd, slopes)[:la].clone())
    alone = ops.attention_mc(qa, ka, va, ops.SlotMap([la], DEV), H, hd, slopes)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], alone)
    want = _ref_attention(qa, ka, va, [la], [la], H, hd, slopes)
    assert torch.allclose(alone.float(), want, atol=3e-2, rtol=3e-2)


def test_mc_large_scores_raise_the_lazy_maximum():
    """Scores that grow along the key axis force the reference maximum to be raised (O rescaled in TMEM) many times."""
    H, hd, lens = 4, 48, [700, 130]
    n, d = sum(lens), H * hd
    g = torch.Generator(device="cpu").manual_seed(3)
    q = (torch.randn(n, d, generator=g) * 3).to(DEV).bfloat16()
    k = (torch.randn(n, d, generator=g) * 3 * torch.linspace(0.2, 3.0, n)[:, None]).to(DEV).bfloat16()
    v = torch.randn(n, d, generator=g).to(DEV).bfloat16()
    slots = ops.SlotMap(lens, DEV)
    got = ops.attention_mc(q, k, v, slots, H, hd, None)
    want = _ref_attention(q, k, v, lens, lens, H, hd, None)
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=3e-2, rtol=3e-2), _describe_mismatch(got.float(), want, 3e-2)


def test_label_attention_equals_full_attention_over_labels():
    H, hd, C = 4, 48, [37, 120]
    D = H * hd
    g = torch.Generator(device="cpu").manual_seed(3)
    q = torch.randn(sum(C), D, generator=g).to(DEV).bfloat16()
    kv9 = torch.randn(9, 2 * D, generator=g).to(DEV)
    labels = [torch.randint(0, 9, (c,), generator=g) for c in C]
    labels[0][labels[0] == 4] = 5                                      # make one class absent in gene 0
    counts = torch.stack([torch.bincount(l, minlength=9) for l in labels]).float()
    logc = counts.log().to(DEV)
    row_seq = torch.repeat_interleave(torch.arange(2), torch.tensor(C)).int().to(DEV)
    got = ops.label_attention(q, kv9, logc, row_seq, H, hd)
    kfull = torch.cat([kv9[l.to(DEV)] for l in labels])               # what the reference materialises
    want = _ref_attention(q, kfull[:, :D], kfull[:, D:], C, C, H, hd, None)
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=1e-2, rtol=1e-2), _describe_mismatch(got.float(), want, 1e-2)


@pytest.mark.parametrize("M,d,gelu", [(1000, 1536, False), (77, 512, False), (50, 192, True), (63, 1536, True)])
def test_layernorm(M, d, gelu):
    g = torch.Generator(device="cpu").manual_seed(d)
    x = (torch.randn(M, d, generator=g) * 3 + 1).to(DEV)
    gam = torch.randn(d, generator=g).to(DEV); bet = torch.randn(d, generator=g).to(DEV)
    want = F.layer_norm(x, (d,), gam, bet, 1e-5)
    if gelu:
        want = F.gelu(want)
    got = ops.layernorm(x, gam, bet, gelu=gelu)
    torch.cuda.synchronize()
    assert torch.allclose(got.float(), want, atol=2e-2, rtol=1e-2), _describe_mismatch(got.float(), want, 2e-2)


def test_unpad_embed_meanpool():
    n, L, d, V = 50, 200, 512, 500
    g = torch.Generator(device="cpu").manual_seed(1)
    tok = torch.randint(4, V, (n, L), generator=g, dtype=torch.int32)
    mask = torch.rand(n, L, generator=g) < 0.4                      # arbitrary (non-prefix) masks
    mask[3] = False; mask[4, 1:] = True
    lens = (~mask).sum(1).numpy()
    emb = torch.randn(V, d, generator=g).to(DEV); pe = torch.randn(L, d, generator=g).to(DEV)
    tok_d, mask_d = tok.to(DEV), mask.to(torch.uint8).to(DEV)
    got_lens = ops.window_lengths(mask_d)
    assert got_lens.cpu().tolist() == lens.tolist()
    cu = ops.cu_seqlens(lens, DEV)
    ids, pos = ops.compact_tokens(tok_d, mask_d, cu, int(lens.sum()))
    keep = ~mask
    assert ids.cpu().tolist() == tok[keep].tolist()
    assert pos.cpu().tolist() == torch.arange(L).expand(n, L)[keep].tolist()
    x = ops.embed_tokens(ids, pos, emb, pe)
    want = emb[tok_d[keep.to(DEV)].long()] + pe[pos.long()]
    assert torch.equal(x, want)
    pooled_bf, pooled = ops.masked_meanpool(x, cu, n, want_f32=True)
    seg = torch.repeat_interleave(torch.arange(n), torch.from_numpy(lens)).to(DEV)
    wantp = torch.zeros(n, d, device=DEV).index_add_(0, seg, want) / torch.from_numpy(lens).to(DEV)[:, None]
    torch.cuda.synchronize()
    assert torch.allclose(pooled, wantp, atol=1e-5, rtol=1e-5)
    assert torch.allclose(pooled_bf.float(), wantp, atol=2e-2, rtol=1e-2)


def test_gather_and_head_out():
    g = torch.Generator(device="cpu").manual_seed(2)
    ta = torch.randn(40, 192, generator=g).to(DEV); tb = torch.randn(63, 192, generator=g).to(DEV)
    idx = torch.tensor([-1, 0, 1, 2, -63, 39, 5, -7], dtype=torch.int32)
    of, ob = ops.gather_rows(ta, tb, idx.to(DEV), want_f32=True, want_bf16=True)
    want = torch.stack([ta[i] if i >= 0 else tb[-i - 1] for i in idx.tolist()])
    assert torch.equal(of, want) and torch.equal(ob, want.bfloat16())
    h = torch.randn(21, 1536, generator=g).to(DEV).bfloat16()
    w = (torch.randn(1536, generator=g) / 40).to(DEV); b = torch.randn(1, generator=g).to(DEV)
    got = ops.head_out(h, w, b)
    want = F.softplus(h.float() @ w + b)
    torch.cuda.synchronize()
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)
