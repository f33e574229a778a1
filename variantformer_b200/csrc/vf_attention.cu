// vf_attention.cu — variable-length, non-causal multi-head attention with optional
// ALiBi bias, FlashAttention-style online softmax (fp32 statistics, bf16 operands).
//
// Replaces flash_attn's varlen kernels on the reference's hot path:
//   flash_attn_varlen_qkvpacked_func  (seq2reg/modules.py:159-171, seq2gene/modules/layers.py:454-467)
//   flash_attn_varlen_kvpacked_func   (seq2gene/modules/layers.py:372-439)
// Semantics (flash_attn.modules.mha): scores = (q·k)/sqrt(hd) - slope_h*|i + Sk - Sq - j|,
// softmax over the sequence's own keys, no dropout in eval.
//
// Round-1 implementation: warp-level mma.sync m16n8k16 (bf16 -> fp32) with
// cp.async double-buffered K/V tiles and ldmatrix fragments; one CTA owns BM query
// rows of one (sequence, head).  The tcgen05/TMEM version replaces this file's
// mainloop in a later round; results are identical by construction
// of the parity tests, which only reference the C-ABI.
#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {

struct AttnParams {
    const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; __nv_bfloat16* o;
    int ldq, ldk, ldv, ldo;            // row strides in elements; head h lives at column h*HD
    const int* cu_q; const int* cu_k;  // [n_seq+1] prefix sums of query / key rows
    const int* tile_seq; const int* tile_q0;   // per query tile: sequence index, first query row inside it
    const float* slopes;               // [H] ALiBi slopes or nullptr
    float scale_log2;                  // (1/sqrt(hd)) * log2(e)
};

template <int HD, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attention_kernel(const AttnParams p) {
    constexpr int BM = NWARPS * 16, BN = 64;
    constexpr int PITCH = HD + 8;                 // elements; keeps ldmatrix rows 16B aligned and conflict-free
    constexpr int KSTEPS = HD / 16;               // k-steps of Q·K^T
    constexpr int DT = HD / 8;                    // n-tiles of the output
    constexpr int CHUNKS = HD / 8;                // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t attn_smem[];
    __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(attn_smem);                 // [BM][PITCH]
    __nv_bfloat16 (*sk)[BN * PITCH] = reinterpret_cast<__nv_bfloat16 (*)[BN * PITCH]>(sq + BM * PITCH);   // [2][BN][PITCH]
    __nv_bfloat16 (*sv)[BN * PITCH] = sk + 2;                                                          // [2][BN][PITCH]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int head = blockIdx.y;
    const int seq = p.tile_seq[blockIdx.x], q0 = p.tile_q0[blockIdx.x];
    const int qbeg = p.cu_q[seq], Sq = p.cu_q[seq + 1] - qbeg;
    const int kbeg = p.cu_k[seq], Sk = p.cu_k[seq + 1] - kbeg;
    const int hoff = head * HD;

    // ---- stage Q tile (rows beyond the sequence are zero filled) ----
    for (int i = tid; i < BM * CHUNKS; i += NWARPS * 32) {
        const int r = i / CHUNKS, c = i % CHUNKS;
        const bool ok = q0 + r < Sq;
        const __nv_bfloat16* src = p.q + (size_t)(qbeg + (ok ? q0 + r : 0)) * p.ldq + hoff + c * 8;
        cp_async_16(smem_u32(&sq[r * PITCH + c * 8]), src, ok);
    }
    auto load_kv = [&](int buf, int j0) {
        for (int i = tid; i < BN * CHUNKS; i += NWARPS * 32) {
            const int r = i / CHUNKS, c = i % CHUNKS;
            const bool ok = j0 + r < Sk;
            const size_t row = (size_t)(kbeg + (ok ? j0 + r : 0));
            cp_async_16(smem_u32(&sk[buf][r * PITCH + c * 8]), p.k + row * p.ldk + hoff + c * 8, ok);
            cp_async_16(smem_u32(&sv[buf][r * PITCH + c * 8]), p.v + row * p.ldv + hoff + c * 8, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();

    const int n_kv = (Sk + BN - 1) / BN;
    float o_acc[DT][4];
#pragma unroll
    for (int d = 0; d < DT; ++d) { o_acc[d][0] = o_acc[d][1] = o_acc[d][2] = o_acc[d][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[KSTEPS][4];
    const float slope_l2 = p.slopes ? p.slopes[head] * 1.4426950408889634f : 0.f;
    const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;          // this thread's two query rows (in-sequence index)
    const int shift = Sk - Sq;                                   // flash_attn ALiBi: |i + Sk - Sq - j|

    for (int it = 0; it < n_kv; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_kv) load_kv(buf ^ 1, (it + 1) * BN);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (it == 0) {
            // Q fragments (A operand) for this warp's 16 rows, loaded once
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int r = warp * 16 + (lane & 15), c = ks * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qf[ks], smem_u32(&sq[r * PITCH + c]));
            }
        }
        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[BN / 8][4];
#pragma unroll
        for (int n = 0; n < BN / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < BN / 16; ++np) {
                // matrices: (keys 16np..+7, k lo), (same keys, k hi), (keys +8..+15, k lo), (keys +8.., k hi)
                uint32_t kf[4];
                const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3), c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(kf, smem_u32(&sk[buf][r * PITCH + c]));
                mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- scale, ALiBi, key masking; online softmax in the log2 domain ----
        const int j_base = it * BN;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < BN / 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j_base + n * 8 + t4 * 2 + (e & 1);
                const int qi = (e < 2) ? qi0 : qi1;
                float x = s[n][e] * p.scale_log2;
                if (p.slopes) x -= slope_l2 * fabsf((float)(qi + shift - j));
                if (j >= Sk) x = -INFINITY;
                s[n][e] = x;
                mx[e >> 1] = fmaxf(mx[e >> 1], x);
            }
        }
        float corr[2], m_new[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            m_new[h] = fmaxf(m_run[h], mx[h]);
            // every tile has >= 1 valid key, so m_new is finite
            corr[h] = exp2f(m_run[h] - m_new[h]);
            m_run[h] = m_new[h];
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[BN / 16][4];
#pragma unroll
        for (int n = 0; n < BN / 8; ++n) {
            const float p0 = exp2f(s[n][0] - m_new[0]), p1 = exp2f(s[n][1] - m_new[0]);
            const float p2 = exp2f(s[n][2] - m_new[1]), p3 = exp2f(s[n][3] - m_new[1]);
            rs[0] += p0 + p1; rs[1] += p2 + p3;
            pf[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
        for (int d = 0; d < DT; ++d) {
            o_acc[d][0] *= corr[0]; o_acc[d][1] *= corr[0]; o_acc[d][2] *= corr[1]; o_acc[d][3] *= corr[1];
        }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
#pragma unroll
            for (int dp = 0; dp < DT / 2; ++dp) {
                // matrices: (keys 16kk..+7, d 16dp..+7), (keys +8..+15, same d), (keys lo, d +8..+15), (keys hi, d +8..)
                uint32_t vf4[4];
                const int r = kk * 16 + (lane & 15), c = dp * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(vf4, smem_u32(&sv[buf][r * PITCH + c]));
                mma_bf16_16816(o_acc[2 * dp], pf[kk], vf4[0], vf4[1]);
                mma_bf16_16816(o_acc[2 * dp + 1], pf[kk], vf4[2], vf4[3]);
            }
        }
        __syncthreads();     // everyone done with buf before it is refilled two iterations later
    }
    cp_async_wait<0>();

    // ---- normalise and store ----
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
    for (int d = 0; d < DT; ++d) {
        const int col = hoff + d * 8 + t4 * 2;
        if (qi0 < Sq)
            *reinterpret_cast<uint32_t*>(p.o + (size_t)(qbeg + qi0) * p.ldo + col) =
                pack_bf16x2(o_acc[d][0] * inv0, o_acc[d][1] * inv0);
        if (qi1 < Sq)
            *reinterpret_cast<uint32_t*>(p.o + (size_t)(qbeg + qi1) * p.ldo + col) =
                pack_bf16x2(o_acc[d][2] * inv1, o_acc[d][3] * inv1);
    }
}

template <int HD, int NWARPS>
static int launch_attn(const AttnParams& p, int n_tiles, int heads, cudaStream_t s) {
    constexpr size_t smem = (size_t)(NWARPS * 16 + 4 * 64) * (HD + 8) * sizeof(__nv_bfloat16);
    static bool attr_set = false;
    if (!attr_set) {
        VF_CUDA_OK(cudaFuncSetAttribute(attention_kernel<HD, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        attr_set = true;
    }
    dim3 grid(n_tiles, heads);
    attention_kernel<HD, NWARPS><<<grid, NWARPS * 32, smem, s>>>(p);
    VF_LAUNCH_OK("attention_kernel launch");
    return 0;
}

int attention_varlen(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                     const int* cu_q, const int* cu_k, const int* tile_seq, const int* tile_q0, int n_tiles,
                     int block_m, int heads, int head_dim, const float* slopes, cudaStream_t stream) {
    VF_REQUIRE(head_dim == 48 || head_dim == 64 || head_dim == 32, "attention: head_dim %d not supported (32/48/64)",
               head_dim);
    VF_REQUIRE(block_m == 64 || block_m == 128, "attention: block_m must be 64 or 128");
    VF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: strides must keep 16B alignment");
    if (n_tiles == 0) return 0;
    AttnParams p;
    p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.o = (__nv_bfloat16*)o;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.cu_q = cu_q; p.cu_k = cu_k; p.tile_seq = tile_seq;
    p.tile_q0 = tile_q0; p.slopes = slopes;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    if (block_m == 64) {
        if (head_dim == 48) return launch_attn<48, 4>(p, n_tiles, heads, stream);
        if (head_dim == 64) return launch_attn<64, 4>(p, n_tiles, heads, stream);
        return launch_attn<32, 4>(p, n_tiles, heads, stream);
    }
    if (head_dim == 48) return launch_attn<48, 8>(p, n_tiles, heads, stream);
    if (head_dim == 64) return launch_attn<64, 8>(p, n_tiles, heads, stream);
    return launch_attn<32, 8>(p, n_tiles, heads, stream);
}

}  // namespace vf
