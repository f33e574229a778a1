// vf_elementwise.cu — the memory-bound glue of stages 2-4 (everything the reference
// runs as separate ATen eager kernels): LayerNorm, token embedding + positional
// encoding, window unpadding, masked mean-pool, stream assembly / row gathers, the
// 9-class CRE x label cross-attention collapse and the final head dot + Softplus.
// All kernels are warp-per-row with 16-byte accesses where the shape allows.
#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, affine), fp32 in -> bf16 out, optional exact-erf GELU.
// nn.LayerNorm sites: seq2reg/modules.py:143-144, seq2gene/modules/layers.py:74-76, :1080.
// Exact two-pass statistics in fp32 (mean, then centred variance).
// ---------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int M, int d, float eps, __nv_bfloat16* __restrict__ out, int ldo,
                 int act_gelu) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)row * ldx;
    __nv_bfloat16* orow = out + (size_t)row * ldo;
    if constexpr (VEC) {
        // d % 128 == 0 and d <= 2048: the whole row lives in registers (<= 16 float4 per lane)
        float4 v[16];
        const int nv = d >> 7;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) {
                v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
                q += (a * a + b * b) + (c * c + e * e);
            }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) {
                const int c0 = (i * 32 + lane) * 4;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0));
                float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
                float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
                if (act_gelu) { y0 = gelu_erf(y0); y1 = gelu_erf(y1); y2 = gelu_erf(y2); y3 = gelu_erf(y3); }
                uint2 pk; pk.x = pack_bf16x2(y0, y1); pk.y = pack_bf16x2(y2, y3);
                *reinterpret_cast<uint2*>(orow + c0) = pk;
            }
    } else {
        float s = 0.f;
        for (int c = lane; c < d; c += 32) s += xr[c];
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
        for (int c = lane; c < d; c += 32) { const float a = xr[c] - mean; q += a * a; }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + eps);
        for (int c = lane; c < d; c += 32) {
            float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
            if (act_gelu) y = gelu_erf(y);
            orow[c] = __float2bfloat16_rn(y);
        }
    }
}

int layernorm(const float* x, int ldx, const float* gamma, const float* beta, int M, int d, float eps, void* out,
              int ldo, int act_gelu, cudaStream_t s) {
    if (M == 0) return 0;
    const int rows_per_block = 8;
    const int grid = (M + rows_per_block - 1) / rows_per_block;
    const bool vec = (d % 128 == 0) && d <= 2048 && (ldx % 4 == 0) && (ldo % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    if (vec)
        layernorm_kernel<true><<<grid, 256, 0, s>>>(x, ldx, gamma, beta, M, d, eps, (__nv_bfloat16*)out, ldo, act_gelu);
    else
        layernorm_kernel<false><<<grid, 256, 0, s>>>(x, ldx, gamma, beta, M, d, eps, (__nv_bfloat16*)out, ldo, act_gelu);
    VF_LAUNCH_OK("layernorm_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Row statistics of a freshly assembled fp32 stream: (sum, sum of squares) per row and
// the bf16 mirror that the LayerNorm-folded GEMMs read as their A operand (vf_gemm.cu).
// Used once where a stream is born (token embedding, registry-token assembly); every
// later LayerNorm input gets its statistics from the GEMM epilogue that produced it.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowstats_kernel(const float* __restrict__ x, int ldx, int M, int d, float* __restrict__ stats,
                __nv_bfloat16* __restrict__ out, int ldo) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)row * ldx;
    float s = 0.f, q = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        s += (v.x + v.y) + (v.z + v.w);
        q += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
        if (out) {
            uint2 pk; pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(out + (size_t)row * ldo + c) = pk;
        }
    }
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) { stats[2 * (size_t)row] = s; stats[2 * (size_t)row + 1] = q; }
}

int rowstats(const float* x, int ldx, int M, int d, float* stats, void* out_bf16, int ldo, cudaStream_t s) {
    if (M == 0) return 0;
    VF_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && (!out_bf16 || ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
               "rowstats: width and strides must be multiples of 4 elements, 16-byte aligned input");
    rowstats_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, ldx, M, d, stats, (__nv_bfloat16*)out_bf16, ldo);
    VF_LAUNCH_OK("rowstats_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Row-centred streams.  Every consumer of a residual stream is a LayerNorm (invariant under a per-row shift) or a
// residual add (which carries a shift along), so the engine stores x' = x - c_r with c_r = the row's mean at the point
// where the stream is assembled, and keeps c_r in a side vector.  The LayerNorm fold reads the bf16 mirror of the
// UN-normalised row: without the shift a row with |mean| >> std loses log2(|mean|/std) bits of the normalised signal
// to bf16 rounding that the reference's fp32 LayerNorm keeps (measured: 2.6e-2 vs 1.8e-3 at mean = 20 std).
// center_rows: x <- x - mean_r in place, pivot[r] = mean_r, bf16 mirror and (sum, sum of squares) of the shifted row.
// uncenter_rows: out[r] = x[r] + pivot[idx ? idx[r] : r] for the places that need raw values (cross-attention
// context, pooled windows, the returned embeddings).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
center_rows_kernel(float* __restrict__ x, int ldx, int M, int d, float* __restrict__ pivot, float* __restrict__ stats,
                   __nv_bfloat16* __restrict__ out, int ldo) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float* xr = x + (size_t)row * ldx;
    float s = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / (float)d;
    s = 0.f;
    float q = 0.f;
    for (int c = lane * 4; c < d; c += 128) {                 // second read of the row hits L1
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
        *reinterpret_cast<float4*>(xr + c) = v;
        s += (v.x + v.y) + (v.z + v.w);
        q += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
        if (out) {
            uint2 pk; pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(out + (size_t)row * ldo + c) = pk;
        }
    }
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) { pivot[row] = mean; stats[2 * (size_t)row] = s; stats[2 * (size_t)row + 1] = q; }
}

int center_rows(float* x, int ldx, int M, int d, float* pivot, float* stats, void* out_bf16, int ldo, cudaStream_t s) {
    if (M == 0) return 0;
    VF_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && (!out_bf16 || ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
               "center_rows: width and strides must be multiples of 4 elements, 16-byte aligned input");
    center_rows_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, ldx, M, d, pivot, stats, (__nv_bfloat16*)out_bf16, ldo);
    VF_LAUNCH_OK("center_rows_kernel launch");
    return 0;
}

__global__ void __launch_bounds__(256)
uncenter_rows_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ pivot, const int* __restrict__ idx,
                     int M, int d, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int ldo) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const float c0 = pivot[idx ? idx[row] : row];
    for (int c = lane * 4; c < d; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * ldx + c);
        v.x += c0; v.y += c0; v.z += c0; v.w += c0;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * ldo + c) = v;
        if (out_bf16) {
            uint2 pk; pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * ldo + c) = pk;
        }
    }
}

int uncenter_rows(const float* x, int ldx, const float* pivot, const int* idx, int M, int d, float* out_f32,
                  void* out_bf16, int ldo, cudaStream_t s) {
    if (M == 0) return 0;
    VF_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
               "uncenter_rows: width and strides must be multiples of 4 elements, 16-byte aligned input");
    uncenter_rows_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, ldx, pivot, idx, M, d, out_f32, (__nv_bfloat16*)out_bf16, ldo);
    VF_LAUNCH_OK("uncenter_rows_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Window unpadding (flash_attn.bert_padding.unpad_input at seq2reg/modules.py:156-161):
// valid tokens of each [L]-token window are compacted in order; `pos` keeps the
// original in-window position for the positional encoding.  One warp per window.
// ---------------------------------------------------------------------------------
__global__ void window_lengths_kernel(const uint8_t* __restrict__ pad_mask, int n_win, int L, int* __restrict__ lens) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_win) return;
    const int lane = threadIdx.x & 31;
    int c = 0;
    for (int i = lane; i < L; i += 32) c += pad_mask[(size_t)w * L + i] ? 0 : 1;
    c = (int)warp_sum((float)c);
    if (lane == 0) lens[w] = c;
}

__global__ void compact_tokens_kernel(const int* __restrict__ tokens, const uint8_t* __restrict__ pad_mask,
                                      const int* __restrict__ cu, int n_win, int L, int* __restrict__ out_ids,
                                      int* __restrict__ out_pos) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_win) return;
    const int lane = threadIdx.x & 31;
    int base = cu[w];
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        const bool keep = i < L && pad_mask[(size_t)w * L + i] == 0;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int o = base + __popc(b & ((1u << lane) - 1));
            out_ids[o] = tokens[(size_t)w * L + i];
            out_pos[o] = i;
        }
        base += __popc(b);
    }
}

int window_lengths(const uint8_t* pad_mask, int n_win, int L, int* lens, cudaStream_t s) {
    if (n_win == 0) return 0;
    window_lengths_kernel<<<(n_win + 7) / 8, 256, 0, s>>>(pad_mask, n_win, L, lens);
    VF_LAUNCH_OK("window_lengths_kernel launch");
    return 0;
}
int compact_tokens(const int* tokens, const uint8_t* pad_mask, const int* cu, int n_win, int L, int* out_ids,
                   int* out_pos, cudaStream_t s) {
    if (n_win == 0) return 0;
    compact_tokens_kernel<<<(n_win + 7) / 8, 256, 0, s>>>(tokens, pad_mask, cu, n_win, L, out_ids, out_pos);
    VF_LAUNCH_OK("compact_tokens_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Token embedding + sinusoidal positional encoding (seq2reg/model.py:214-220):
//   x[t] = E[id[t]] + PE[pos[t]]           (PE table precomputed exactly like :15-37)
// ---------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const int* __restrict__ ids, const int* __restrict__ pos,
                                    const float* __restrict__ emb, const float* __restrict__ pe, int n_tok, int d,
                                    float* __restrict__ out) {
    const int dv = d >> 2;
    const size_t total = (size_t)n_tok * dv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / dv), c = (int)(i % dv) * 4;
        float4 e = __ldg(reinterpret_cast<const float4*>(emb + (size_t)ids[t] * d + c));
        if (pe) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (size_t)pos[t] * d + c));
            e.x += p.x; e.y += p.y; e.z += p.z; e.w += p.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)t * d + c) = e;
    }
}

int embed_tokens(const int* ids, const int* pos, const float* emb, const float* pe, int n_tok, int d, float* out,
                 cudaStream_t s) {
    VF_REQUIRE(d % 4 == 0, "embed_tokens: d must be a multiple of 4");
    if (n_tok == 0) return 0;
    const size_t total = (size_t)n_tok * (d / 4);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    embed_tokens_kernel<<<grid, 256, 0, s>>>(ids, pos, emb, pe, n_tok, d, out);
    VF_LAUNCH_OK("embed_tokens_kernel launch");
    return 0;
}

// embed_tokens + center_rows in one pass (the coarse seq2reg entry point): a warp builds its token's row in registers,
// takes the row mean, and writes the centred fp32 row, its bf16 mirror, the pivot and the first row statistics — the
// uncentred rows (2.3 GB per benchmark slab) are never written and re-read.  Same arithmetic, same summation order as the
// two kernels: bit-identical results.  NV = float4 per lane (d <= 128 * NV).
template <int NV>
__global__ void __launch_bounds__(256)
embed_center_kernel(const int* __restrict__ ids, const int* __restrict__ pos, const float* __restrict__ emb,
                    const float* __restrict__ pe, int n_tok, int d, float* __restrict__ x, float* __restrict__ pivot,
                    float* __restrict__ stats, __nv_bfloat16* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_tok) return;
    const int lane = threadIdx.x & 31;
    const float* er = emb + (size_t)ids[row] * d;
    const float* pr = pe ? pe + (size_t)pos[row] * d : nullptr;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + 128 * k;
        if (c < d) {
            float4 e = __ldg(reinterpret_cast<const float4*>(er + c));
            if (pr) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c));
                e.x += q.x; e.y += q.y; e.z += q.z; e.w += q.w;
            }
            v[k] = e;
            s += (e.x + e.y) + (e.z + e.w);
        }
    }
    const float mean = warp_sum(s) / (float)d;
    s = 0.f;
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + 128 * k;
        if (c < d) {
            float4 e = v[k];
            e.x -= mean; e.y -= mean; e.z -= mean; e.w -= mean;
            *reinterpret_cast<float4*>(x + (size_t)row * d + c) = e;
            s += (e.x + e.y) + (e.z + e.w);
            q2 += fmaf(e.x, e.x, e.y * e.y) + fmaf(e.z, e.z, e.w * e.w);
            uint2 pk; pk.x = pack_bf16x2(e.x, e.y); pk.y = pack_bf16x2(e.z, e.w);
            *reinterpret_cast<uint2*>(out + (size_t)row * d + c) = pk;
        }
    }
    s = warp_sum(s); q2 = warp_sum(q2);
    if (lane == 0) { pivot[row] = mean; stats[2 * (size_t)row] = s; stats[2 * (size_t)row + 1] = q2; }
}

int embed_center(const int* ids, const int* pos, const float* emb, const float* pe, int n_tok, int d, float* x,
                 float* pivot, float* stats, void* out_bf16, cudaStream_t s) {
    if (n_tok == 0) return 0;
    if (d % 4 != 0 || d > 1024 || !out_bf16) {               // shapes the register form does not cover: the two kernels
        if (int rc = embed_tokens(ids, pos, emb, pe, n_tok, d, x, s)) return rc;
        return center_rows(x, d, n_tok, d, pivot, stats, out_bf16, d, s);
    }
    const int grid = (n_tok + 7) / 8;
    if (d <= 512) embed_center_kernel<4><<<grid, 256, 0, s>>>(ids, pos, emb, pe, n_tok, d, x, pivot, stats, (__nv_bfloat16*)out_bf16);
    else embed_center_kernel<8><<<grid, 256, 0, s>>>(ids, pos, emb, pe, n_tok, d, x, pivot, stats, (__nv_bfloat16*)out_bf16);
    VF_LAUNCH_OK("embed_center_kernel launch");
    return 0;
}


// ---------------------------------------------------------------------------------
// Masked mean over each window's valid tokens (seq2reg/model.py:263-267).
// One CTA per window; 0/0 -> NaN for an empty window, as upstream.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
meanpool_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ cu, int d, const float* __restrict__ pivot,
                __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int ldo) {
    const int w = blockIdx.x;
    const int b = cu[w], e = cu[w + 1];
    const float inv = 1.0f / (float)(e - b);        // inf for an empty window -> 0*inf = NaN
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = b; t < e; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)t * ldx + c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (pivot) {                                  // rows of a centred stream: add the mean of their pivots back
            float c0 = 0.f;
            for (int t = b; t < e; ++t) c0 += pivot[t];
            a.x += c0; a.y += c0; a.z += c0; a.w += c0;
        }
        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)w * ldo + c) = a;
        if (out_bf16) {
            uint2 pk; pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
            *reinterpret_cast<uint2*>(out_bf16 + (size_t)w * ldo + c) = pk;
        }
    }
}

int masked_meanpool(const float* x, int ldx, const int* cu, int n_win, int d, const float* pivot, void* out_bf16,
                    float* out_f32, int ldo, cudaStream_t s) {
    VF_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "meanpool: d/ld must be multiples of 4");
    if (n_win == 0) return 0;
    meanpool_kernel<<<n_win, 128, 0, s>>>(x, ldx, cu, d, pivot, (__nv_bfloat16*)out_bf16, out_f32, ldo);
    VF_LAUNCH_OK("meanpool_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Row gather with two source tables: idx >= 0 -> table_a[idx], idx < 0 -> table_b[-idx-1].
// Builds the gene stream (registry token of the tissue + gene-chunk embeddings shared by all
// tissue copies: layers.py:508-521 + model_combined_modulator.py:622-649) and extracts the
// registry rows at the end (pool_outputs :391-392).
// ---------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ ta, int lda, const float* __restrict__ tb, int ldb,
                                   const int* __restrict__ idx, int n_rows, int d, float* __restrict__ out_f32,
                                   __nv_bfloat16* __restrict__ out_bf16, int ldo) {
    const int dv = d >> 2;
    const size_t total = (size_t)n_rows * dv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / dv), c = (int)(i % dv) * 4;
        const int j = idx[r];
        const float4 v = (j >= 0) ? *reinterpret_cast<const float4*>(ta + (size_t)j * lda + c)
                                  : *reinterpret_cast<const float4*>(tb + (size_t)(-j - 1) * ldb + c);
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)r * ldo + c) = v;
        if (out_bf16) {
            uint2 pk; pk.x = pack_bf16x2(v.x, v.y); pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(out_bf16 + (size_t)r * ldo + c) = pk;
        }
    }
}

int gather_rows(const float* ta, int lda, const float* tb, int ldb, const int* idx, int n_rows, int d, float* out_f32,
                void* out_bf16, int ldo, cudaStream_t s) {
    VF_REQUIRE(d % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0 && (tb == nullptr || ldb % 4 == 0),
               "gather_rows: d/ld must be multiples of 4");
    if (n_rows == 0) return 0;
    const size_t total = (size_t)n_rows * (d / 4);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    gather_rows_kernel<<<grid, 256, 0, s>>>(ta, lda, tb, ldb, idx, n_rows, d, out_f32, (__nv_bfloat16*)out_bf16, ldo);
    VF_LAUNCH_OK("gather_rows_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// CRE x label cross-attention, collapsed over the 9 reference-cCRE classes.
// The reference attends every CRE query over K/V = Wkv·Emb9[label_j] for all C keys of
// its gene (layers.py:142-150, model_combined_modulator.py:168,261-267); keys of one class
// are identical, so softmax over C keys == softmax over the 9 class logits + log(count_c).
// Exact identity (SURVEY App. D.12, fp32 check 4.8e-7).  fp32 CUDA-core math.
//   q     bf16 [n_rows, H*HD]
//   kv9   fp32 [9, 2*H*HD]   ((two, h, d) order like flash_attn's Wkv)
//   logc  fp32 [n_seq, 9]    log(count of class c among the sequence's CREs), -inf if absent
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
label_attention_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const float* __restrict__ kv9,
                       const float* __restrict__ logc, const int* __restrict__ row_seq, int n_rows, int H, int HD,
                       float scale, __nv_bfloat16* __restrict__ out, int ldo) {
    const int row = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int D = H * HD;
    const float* lc = logc + (size_t)row_seq[row] * 9;
    for (int h = warp; h < H; h += nw) {
        const int c0 = h * HD;
        const float q0 = lane < HD ? __bfloat162float(q[(size_t)row * ldq + c0 + lane]) : 0.f;
        const float q1 = lane + 32 < HD ? __bfloat162float(q[(size_t)row * ldq + c0 + lane + 32]) : 0.f;
        float logit[9];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const float* kr = kv9 + (size_t)c * 2 * D + c0;
            float pdot = (lane < HD ? q0 * __ldg(kr + lane) : 0.f) + (lane + 32 < HD ? q1 * __ldg(kr + lane + 32) : 0.f);
            pdot = warp_sum(pdot);
            logit[c] = pdot * scale + lc[c];
            mx = fmaxf(mx, logit[c]);
        }
        float den = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const float pc = __expf(logit[c] - mx);      // -inf logit -> 0
            den += pc;
            const float* vr = kv9 + (size_t)c * 2 * D + D + c0;
            if (lane < HD) o0 += pc * __ldg(vr + lane);
            if (lane + 32 < HD) o1 += pc * __ldg(vr + lane + 32);
        }
        const float inv = 1.0f / den;
        if (lane < HD) out[(size_t)row * ldo + c0 + lane] = __float2bfloat16_rn(o0 * inv);
        if (lane + 32 < HD) out[(size_t)row * ldo + c0 + lane + 32] = __float2bfloat16_rn(o1 * inv);
    }
}

// Same computation, ONE THREAD PER (row, head): the 9 class keys / values of 8 heads sit in shared memory (every lane of
// a warp works on the same head, so each read is a broadcast), the query row and the output row stay in registers;
// no shuffles.  ~20x fewer instructions than the warp-per-(row, head) kernel above.
template <int HD>
__global__ void __launch_bounds__(256)
label_attention_rows_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const float* __restrict__ kv9,
                            const float* __restrict__ logc, const int* __restrict__ row_seq, int n_rows, int H,
                            float scale, __nv_bfloat16* __restrict__ out, int ldo) {
    __shared__ float sk[8][9][HD], sv[8][9][HD];
    const int D = H * HD, h0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < 8 * 9 * HD; i += 256) {
        const int hh = i / (9 * HD), c = (i / HD) % 9, d = i % HD;
        const bool ok = h0 + hh < H;
        sk[hh][c][d] = ok ? __ldg(kv9 + (size_t)c * 2 * D + (h0 + hh) * HD + d) : 0.f;
        sv[hh][c][d] = ok ? __ldg(kv9 + (size_t)c * 2 * D + D + (h0 + hh) * HD + d) : 0.f;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, h = h0 + w, row = blockIdx.x * 32 + (threadIdx.x & 31);
    if (h >= H || row >= n_rows) return;
    float qv[HD];
    const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)row * ldq + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        const uint4 u = __ldg(qp + i);
        const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            qv[i * 8 + 2 * j] = __uint_as_float(ww[j] << 16);
            qv[i * 8 + 2 * j + 1] = __uint_as_float(ww[j] & 0xffff0000u);
        }
    }
    const float* lc = logc + (size_t)row_seq[row] * 9;
    float logit[9], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 2) { a0 = fmaf(qv[d], sk[w][c][d], a0); a1 = fmaf(qv[d + 1], sk[w][c][d + 1], a1); }
        logit[c] = (a0 + a1) * scale + __ldg(lc + c);
        mx = fmaxf(mx, logit[c]);
    }
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) { logit[c] = __expf(logit[c] - mx); den += logit[c]; }     // -inf logit -> 0
    const float inv = 1.0f / den;
    uint4* op = reinterpret_cast<uint4*>(out + (size_t)row * ldo + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < 9; ++c) a = fmaf(logit[c], sv[w][c][i * 8 + j], a);
            o[j] = a * inv;
        }
        op[i] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

int label_attention(const void* q, int ldq, const float* kv9, const float* logc, const int* row_seq, int n_rows, int H,
                    int HD, void* out, int ldo, cudaStream_t s) {
    VF_REQUIRE(HD <= 64, "label_attention: head_dim must be <= 64");
    if (n_rows == 0) return 0;
    const bool vec = (ldq % 8 == 0) && (ldo % 8 == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec && (HD == 48 || HD == 64)) {
        const dim3 grid((n_rows + 31) / 32, (H + 7) / 8);
        const float scale = 1.0f / sqrtf((float)HD);
        if (HD == 48)
            label_attention_rows_kernel<48><<<grid, 256, 0, s>>>((const __nv_bfloat16*)q, ldq, kv9, logc, row_seq, n_rows, H,
                                                                 scale, (__nv_bfloat16*)out, ldo);
        else
            label_attention_rows_kernel<64><<<grid, 256, 0, s>>>((const __nv_bfloat16*)q, ldq, kv9, logc, row_seq, n_rows, H,
                                                                 scale, (__nv_bfloat16*)out, ldo);
        VF_LAUNCH_OK("label_attention_rows_kernel launch");
        return 0;
    }
    label_attention_kernel<<<n_rows, 256, 0, s>>>((const __nv_bfloat16*)q, ldq, kv9, logc, row_seq, n_rows, H, HD,
                                                   1.0f / sqrtf((float)HD), (__nv_bfloat16*)out, ldo);
    VF_LAUNCH_OK("label_attention_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Last head layer: y = softplus(w·h + b)  (layers.py:1084-1087; Softplus beta=1, threshold=20).
// ---------------------------------------------------------------------------------
__global__ void head_out_kernel(const __nv_bfloat16* __restrict__ h, int ldh, const float* __restrict__ w,
                                const float* __restrict__ b, int n_rows, int d, int softplus,
                                float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc += __bfloat162float(h[(size_t)row * ldh + c]) * __ldg(w + c);
    acc = warp_sum(acc) + b[0];
    if (lane == 0) out[row] = softplus ? (acc > 20.f ? acc : log1pf(expf(acc))) : acc;
}

int head_out(const void* h, int ldh, const float* w, const float* b, int n_rows, int d, int softplus, float* out,
             cudaStream_t s) {
    if (n_rows == 0) return 0;
    head_out_kernel<<<(n_rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)h, ldh, w, b, n_rows, d, softplus, out);
    VF_LAUNCH_OK("head_out_kernel launch");
    return 0;
}

// fp32 -> bf16 cast (weights at load time, contexts)
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16_rn(x[i]);
}
int cast_f32_to_bf16(const float* x, void* y, size_t n, cudaStream_t s) {
    if (n == 0) return 0;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    cast_bf16_kernel<<<grid, 256, 0, s>>>(x, (__nv_bfloat16*)y, n);
    VF_LAUNCH_OK("cast_bf16_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// Gradient-boosted forest inference on embeddings: the AD-risk head (processors/ad_risk.py:41-52, 157-176 evaluates one
// treelite GBDT per (gene, tissue) on the 1536-d registry embedding with treelite.gtil.predict, one Python call per row).
// Here every row picks its forest; one warp per row: the row is staged in shared memory, lane l walks trees l, l + 32,
// ... of the row's forest (nodes: split feature or -1 for a leaf, threshold, children, leaf value), the leaf values are
// summed over the warp and squashed: p = sigmoid(base + sum).  Split rule x[feature] <= threshold goes left (sklearn /
// treelite "<="; op_lt != 0: strict "<", the xgboost convention); NaN goes to `default_left`'s side (bit 31 of feat).
// ---------------------------------------------------------------------------------
constexpr int kForestWarps = 8;
__global__ void __launch_bounds__(kForestWarps * 32)
forest_predict_kernel(const float* __restrict__ x, int ldx, int n_rows, int d, const int* __restrict__ row_forest,
                      const int* __restrict__ forest_tree_off, const float* __restrict__ forest_base,
                      const int* __restrict__ tree_root, const int* __restrict__ node_feat,
                      const float* __restrict__ node_thr, const int* __restrict__ node_left,
                      const int* __restrict__ node_right, const float* __restrict__ node_value, int op_lt,
                      float* __restrict__ out) {
    extern __shared__ float s_row[];                          // kForestWarps x d
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kForestWarps + warp;
    if (row >= n_rows) return;
    float* xr = s_row + (size_t)warp * d;
    for (int c = lane; c < d; c += 32) xr[c] = x[(size_t)row * ldx + c];
    __syncwarp();
    const int f = row_forest[row];
    const int t0 = forest_tree_off[f], t1 = forest_tree_off[f + 1];
    float acc = 0.f;
    for (int t = t0 + lane; t < t1; t += 32) {
        int n = tree_root[t];
        for (int depth = 0; depth < 64; ++depth) {            // (bounded: a malformed table cannot hang the GPU)
            const int ft = node_feat[n];
            if (ft == -1) break;
            const float v = xr[ft & 0x7fffffff], thr = node_thr[n];
            const bool left = (v != v) ? (ft < 0) : (op_lt ? v < thr : v <= thr);
            n = left ? node_left[n] : node_right[n];
        }
        acc += node_value[n];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = 1.0f / (1.0f + __expf(-(forest_base[f] + acc)));
}

int forest_predict(const float* x, int ldx, int n_rows, int d, const int* row_forest, const int* forest_tree_off,
                   const float* forest_base, const int* tree_root, const int* node_feat, const float* node_thr,
                   const int* node_left, const int* node_right, const float* node_value, int op_lt, float* out,
                   cudaStream_t s) {
    if (n_rows == 0) return 0;
    const size_t smem = (size_t)kForestWarps * d * sizeof(float);
    VF_REQUIRE(d > 0 && smem <= 200 * 1024, "forest_predict: %d features do not fit the row staging buffer", d);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        VF_CUDA_OK(cudaFuncSetAttribute(forest_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    forest_predict_kernel<<<(n_rows + kForestWarps - 1) / kForestWarps, kForestWarps * 32, smem, s>>>(
        x, ldx, n_rows, d, row_forest, forest_tree_off, forest_base, tree_root, node_feat, node_thr, node_left, node_right,
        node_value, op_lt, out);
    VF_LAUNCH_OK("forest_predict_kernel launch");
    return 0;
}

}  // namespace vf
