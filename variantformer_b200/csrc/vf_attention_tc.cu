// vf_attention_tc.cu — variable-length non-causal attention on the 5th-gen tensor cores.
//
// Replaces flash_attn's varlen kernels (see vf_attention.cu for the reference call sites) for every problem with
// at least two 128-row query tiles per (sequence, head) work item; the legacy mma.sync kernel keeps the tiny
// seq2reg windows.  Semantics: softmax((q·k)/sqrt(hd) - slope_h*|i + Sk - Sq - j|) v, fp32 statistics.
//
// One persistent CTA per SM walks work items = (sequence, head, block of up to 4 query tiles of 128 rows).
//   warp 0      TMA producer: Q tiles once per item, K and V blocks (64 keys) through 4-stage rings; all tiles are
//               [rows, 64 cols] bf16, SWIZZLE_128B (hd=48 over-fetches 16 columns that no MMA ever reads).
//   warp 1      tcgen05.mma issuer.  S = Q K^T : M128 x N64 x K(hd) into one of the owning warpgroup's TWO TMEM
//               S buffers; O_t += P V : M128 x N(hd) x K64 with P from one of the warpgroup's two shared-memory
//               buffers (K-major) and V as an MN-major operand.  QK^T runs up to two steps per warpgroup ahead.
//   warps 2-9   two softmax warpgroups (query tile t is owned by warpgroup t&1): ONE THREAD PER QUERY ROW, the row
//               arrives straight from TMEM (tcgen05.ld 32x32b), so max / sum need no cross-thread reduction.
// Softmax is single-pass with a LAZY reference maximum (exact arithmetic, FA4-style): every row keeps a reference
// m_ref; P = exp2(x - m_ref) may reach 2^8 before m_ref is raised.  m_ref is only raised (and l / O_t rescaled by
// exp2(m_old - m_new), O_t through tcgen05.ld/st) when a 32-key slab exceeds it by more than 8 in the log2 domain:
// for free on the first slab of a tile, otherwise by a rare restart of the tile with its exact maximum.
// TMEM: columns [128 wg + 64 buf, +64) S buffers, [256 + 64 t, +64) O_t  (512 columns).
#include <cuda.h>

#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {

constexpr int kQT = 128;            // query rows per tile
constexpr int kKB = 64;             // keys per block
constexpr int kMaxQT = 4;           // query tiles per work item
constexpr int kKStages = 4, kVStages = 4;
constexpr uint32_t kTileBytes = 128 * 64 * 2;                 // one [128 x 64] bf16 tile (Q tile, P buffer)
constexpr uint32_t kKvBytes = kKB * 64 * 2;                   // one [64 keys x 64] bf16 tile
constexpr int kAttnThreads = 64 + 256;
constexpr int kNumBars = 4 + 2 * kKStages + 2 * kVStages + 16 + 4;
constexpr size_t kAttnSmem = 1024 + (size_t)kMaxQT * kTileBytes + (size_t)(kKStages + kVStages) * kKvBytes +
                             4 * kTileBytes + kNumBars * 8 + 64;

struct AttnTcParams {
    const int* cu_q; const int* cu_k;
    const int* item_seq; const int* item_q0;
    int n_items, heads;
    __nv_bfloat16* o; int ldo;
    const float* slopes;
    float scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


constexpr float kLazyThreshold = 8.0f;      // log2 units: P stays below 2^8 between reference-max updates

// scaled (and biased / masked) score of element e of a 32-key slab
template <bool ALIBI, bool TAIL>
__device__ __forceinline__ float score(uint32_t raw, int e, float scale, float slope, float d0, int nvalid) {
    float x = __uint_as_float(raw) * scale;
    if constexpr (ALIBI) x = fmaf(__uint_as_float(raw), scale, -slope * fabsf(d0 - (float)e));
    if constexpr (TAIL) { if (e >= nvalid) x = -INFINITY; }
    return x;
}
template <bool ALIBI, bool TAIL>
__device__ __forceinline__ float slab_max(const uint32_t (&r)[32], float scale, float slope, float d0, int nvalid) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // 4 independent chains (latency, not issue, bound)
    if constexpr (!ALIBI && !TAIL) {            // scale > 0 commutes with max: one FMNMX per element
#pragma unroll
        for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(r[e]));
        return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale;
    } else {                                    // r already holds the finished scores (see softmax_tile)
#pragma unroll
        for (int e = 0; e < 32; ++e) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(r[e]));
        return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    }
}
// p = exp2(x - m_ref) for one slab: row-sum, bf16 pack, swizzled store into the P tile (K-major SWIZZLE_128B:
// one 64-key tile, 16-byte chunk index XOR (row & 7); slab c covers chunks 4c .. 4c+3)
template <bool ALIBI, bool TAIL>
__device__ __forceinline__ void slab_exp_store(const uint32_t (&r)[32], float scale, float slope, float d0, int nvalid,
                                               float m_ref, float& sum, uint8_t* dst, int c, int row) {
    uint32_t pk[16];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
        float x0, x1;
        if constexpr (!ALIBI && !TAIL) {
            x0 = fmaf(__uint_as_float(r[e]), scale, -m_ref); x1 = fmaf(__uint_as_float(r[e + 1]), scale, -m_ref);
        } else {
            x0 = __uint_as_float(r[e]) - m_ref; x1 = __uint_as_float(r[e + 1]) - m_ref;
        }
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        acc[(e >> 1) & 3] += p0 + p1;
        pk[e >> 1] = pack_bf16x2(p0, p1);
    }
    sum += (acc[0] + acc[1]) + (acc[2] + acc[3]);
    uint8_t* base = dst + row * 128;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (c * 4 + q4) ^ (row & 7);
        *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
    }
}
// O_t row *= corr, in tensor memory (warp-collective; lanes that need no change pass corr = 1)
__device__ __forceinline__ void rescale_o(uint32_t o_addr, float corr) {
    uint32_t v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        tmem_ld_32x32(o_addr + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * corr);
        tmem_st_32x32(o_addr + h * 32, v);
    }
    tmem_st_wait();
}

// One 128 x 64 score tile for this thread's row.  Returns through m_ref / l / (O_t in TMEM) / the P buffer.
template <bool ALIBI, bool TAIL>
__device__ __forceinline__ void softmax_tile(uint32_t s_addr, uint32_t o_addr, bool have_o, uint64_t* s_empty_bar,
                                             uint64_t* p_empty_bar, uint32_t p_empty_parity, uint64_t* prev_pv_bar,
                                             uint32_t prev_pv_parity, float scale, float slope, float qpos, int key0,
                                             int Sk, float& m_ref, float& l, uint8_t* dst, int row, int lane) {
    constexpr int kSlabs = kKB / 32;
    uint32_t r[kSlabs][32];
    float sum = 0.f;
    bool restart = false;
    float seen = -INFINITY;
#pragma unroll
    for (int c = 0; c < kSlabs; ++c) tmem_ld_32x32(s_addr + c * 32, r[c]);
    tmem_ld_wait();
    tc_fence_before();                                       // whole tile is in registers: release the S buffer
    __syncwarp();
    if (lane == 0) mbar_arrive(s_empty_bar);
    float cmax[kSlabs];
#pragma unroll
    for (int c = 0; c < kSlabs; ++c) {
        if constexpr (ALIBI || TAIL) {                       // biased / masked scores are computed once, in place
#pragma unroll
            for (int e = 0; e < 32; ++e)
                r[c][e] = __float_as_uint(score<ALIBI, TAIL>(r[c][e], e, scale, slope, qpos - (float)(key0 + c * 32),
                                                             Sk - (key0 + c * 32)));
        }
        cmax[c] = slab_max<ALIBI, TAIL>(r[c], scale, slope, 0.f, 0);
        seen = fmaxf(seen, cmax[c]);
    }
    // P buffer free <=> the PV that read it two of this warpgroup's steps ago has retired; every earlier PV (in
    // particular the last one that wrote O_t) has then retired too, so O_t may be rescaled below
    mbar_wait(p_empty_bar, p_empty_parity);
    tc_fence_after();
    if (__any_sync(0xffffffffu, seen > m_ref + kLazyThreshold)) {
        const float m_new = fmaxf(m_ref, seen);              // exact tile maximum: no restart is ever needed
        const float corr = ex2_approx(m_ref - m_new);        // m_ref = -inf -> 0
        l *= corr;
        if (have_o) {
            // the warpgroup's previous step used the OTHER P buffer and may target the same O_t: its PV must have
            // retired (old-reference contributions fully accumulated) before O_t is rescaled
            if (prev_pv_bar) { mbar_wait(prev_pv_bar, prev_pv_parity); tc_fence_after(); }
            rescale_o(o_addr, corr);
        }
        m_ref = m_new;
    }
    (void)restart;
#pragma unroll
    for (int c = 0; c < kSlabs; ++c)
        slab_exp_store<ALIBI, TAIL>(r[c], scale, slope, 0.f, 0, m_ref, sum, dst, c, row);
    l += sum;
}

// SHORT: every work item has <= 2 query tiles (e.g. the 201-token gene self-attention).  Q slots and O accumulators
// are then double-buffered across consecutive items (slot / accumulator base 2*(item&1)), so the TMA producer and
// the MMA warp run into the next item while the softmax warps still drain the previous one.
template <int HD, bool ALIBI, bool SHORT>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sm_q = smem;                                         // kMaxQT tiles of 16 KB
    uint8_t* sm_k = sm_q + kMaxQT * kTileBytes;                   // kKStages tiles of 8 KB
    uint8_t* sm_v = sm_k + kKStages * kKvBytes;                   // kVStages tiles of 8 KB
    uint8_t* sm_p = sm_v + kVStages * kKvBytes;                   // [warpgroup][buffer] tiles of 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 4 * kTileBytes);
    uint64_t* q_full = bars + 0;  uint64_t* q_empty = bars + 2;                        // [2] each (index item&1 if SHORT)
    uint64_t* k_full = bars + 4;                 uint64_t* k_empty = k_full + kKStages;
    uint64_t* v_full = k_empty + kKStages;       uint64_t* v_empty = v_full + kVStages;
    uint64_t* s_full = v_empty + kVStages;       uint64_t* s_empty = s_full + 4;       // [wg * 2 + buf]
    uint64_t* p_full = s_empty + 4;              uint64_t* p_empty = p_full + 4;       // [wg * 2 + buf]
    uint64_t* o_full = p_empty + 4;              uint64_t* o_empty = o_full + 2;       // [2] each
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = p.n_items * p.heads;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 8);
        }
        for (int i = 0; i < kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work decomposition shared by all roles: head-major so concurrently running CTAs share one head's K/V in L2
    auto decode = [&](int w, int& seq, int& head, int& q0, int& qbeg, int& Sq, int& kbeg, int& Sk, int& nq, int& nk) {
        head = w / p.n_items;
        const int item = w - head * p.n_items;
        seq = p.item_seq[item]; q0 = p.item_q0[item];
        qbeg = p.cu_q[seq]; Sq = p.cu_q[seq + 1] - qbeg;
        kbeg = p.cu_k[seq]; Sk = p.cu_k[seq + 1] - kbeg;
        nq = min(kMaxQT, (Sq - q0 + kQT - 1) / kQT);
        nk = (Sk + kKB - 1) / kKB;
    };

    if (warp == 0) {
        // ============================ TMA producer ============================
        if (elect_one()) {
            int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk;
                decode(w, seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk);
                const int col = head * HD;
                const int ib = SHORT ? (it & 1) : 0;                       // Q slot pair / barrier index of this item
                const uint32_t ipar = SHORT ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&q_empty[ib], ipar ^ 1);
                mbar_arrive_expect_tx(&q_full[ib], nq * kTileBytes);
                for (int t = 0; t < nq; ++t)
                    tma_load_2d(sm_q + (ib * 2 + t) * kTileBytes, &tmQ, &q_full[ib], col, qbeg + q0 + t * kQT);
                for (int j = 0; j < nk; ++j) {
                    mbar_wait(&k_empty[ks], kph ^ 1);
                    mbar_arrive_expect_tx(&k_full[ks], kKvBytes);
                    tma_load_2d(sm_k + ks * kKvBytes, &tmK, &k_full[ks], col, kbeg + j * kKB);
                    if (++ks == kKStages) { ks = 0; kph ^= 1; }
                    mbar_wait(&v_empty[vs], vph ^ 1);
                    mbar_arrive_expect_tx(&v_full[vs], kKvBytes);
                    tma_load_2d(sm_v + vs * kKvBytes, &tmV, &v_full[vs], col, kbeg + j * kKB);
                    if (++vs == kVStages) { vs = 0; vph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        if (elect_one()) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kKB);                         // A, B K-major
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD) | (1u << 16);            // B (= V) MN-major
            int ks = 0, vs = 0; uint32_t kph = 0, vph = 0; int it = 0;
            uint32_t m_s[2] = {0, 0}, m_p[2] = {0, 0};              // per-warpgroup local step counters (QK / PV issued)
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk;
                decode(w, seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk);
                const int ib = SHORT ? (it & 1) : 0;
                const uint32_t ipar = SHORT ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&q_full[ib], ipar);
                mbar_wait(&o_empty[ib], ipar ^ 1);                  // the O accumulators of this slot pair have been read out
                tc_fence_after();
                // linearised steps n = j*nq + t (key block j, query tile t, warpgroup t&1)
                const int N = nk * nq;
                int k_seen = 0, v_seen = 0;                         // key blocks whose k_full / v_full were awaited
                int ks_of[4], vs_of[4];                             // smem stage of key block j (indexed j & 3)
                auto qk_step = [&](int n) {
                    const int j = n / nq, t = n - j * nq, wg = t & 1;
                    while (k_seen <= j) {
                        mbar_wait(&k_full[ks], kph);
                        ks_of[k_seen & 3] = ks;
                        if (++ks == kKStages) { ks = 0; kph ^= 1; }
                        ++k_seen;
                    }
                    const uint32_t m = m_s[wg]++;
                    const int buf = m & 1;
                    mbar_wait(&s_empty[wg * 2 + buf], ((m >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sm_q + (ib * 2 + t) * kTileBytes));
                    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sm_k + ks_of[j & 3] * kKvBytes));
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(tmem_base + wg * 128 + buf * kKB, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                    umma_commit(&s_full[wg * 2 + buf]);
                    if (t == nq - 1) umma_commit(&k_empty[ks_of[j & 3]]);
                };
                auto pv_step = [&](int n) {
                    const int j = n / nq, t = n - j * nq, wg = t & 1;
                    while (v_seen <= j) {
                        mbar_wait(&v_full[vs], vph);
                        vs_of[v_seen & 3] = vs;
                        if (++vs == kVStages) { vs = 0; vph ^= 1; }
                        ++v_seen;
                    }
                    const uint32_t m = m_p[wg]++;
                    const int buf = m & 1;
                    mbar_wait(&p_full[wg * 2 + buf], (m >> 1) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + 256 + (ib * 2 + t) * 64;
                    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sm_p + (wg * 2 + buf) * kTileBytes));
                    const uint32_t vb = smem_u32(sm_v + vs_of[j & 3] * kKvBytes);
#pragma unroll
                    for (int kk = 0; kk < kKB / 16; ++kk)
                        umma_bf16(d, da + 2 * kk, umma_desc_kmajor_sw128(vb + kk * 2048), idesc_pv, (j | kk) != 0);
                    umma_commit(&p_empty[wg * 2 + buf]);
                    if (t == nq - 1) umma_commit(&v_empty[vs_of[j & 3]]);
                };
                // S = Q K^T runs ahead of O += P V by two steps per active warpgroup (each owns two S buffers);
                // with a single query tile only warpgroup 0 works and the look-ahead is its two buffers
                const int LA = nq >= 2 ? 4 : 2;
                for (int n = 0; n < LA && n < N; ++n) qk_step(n);
                for (int n = 0; n < N; ++n) {
                    if (n + LA < N) qk_step(n + LA);
                    pv_step(n);
                }
                umma_commit(&o_full[ib]);
                umma_commit(&q_empty[ib]);
            }
        }
        __syncwarp();
    } else {
        // ============================ softmax warpgroups ============================
        const int wg = (warp - 2) >> 2;                       // owns query tiles t with (t & 1) == wg
        const int quad = warp & 3;                            // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;                     // row inside a 128-row query tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t m_loc = 0; int it = 0;                       // this warpgroup's local step counter
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk;
            decode(w, seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk);
            const float slope = ALIBI ? p.slopes[head] * 1.4426950408889634f : 0.f;
            const int shift = Sk - Sq;
            const int ib = SHORT ? (it & 1) : 0;
            const uint32_t ipar = SHORT ? ((it >> 1) & 1) : (it & 1);
            float m_ref[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
            for (int j = 0; j < nk; ++j) {
                const int key0 = j * kKB;
                const bool tail = key0 + kKB > Sk;                // block holds keys beyond the sequence
                for (int tt = 0; wg + 2 * tt < nq; ++tt) {
                    const int t = wg + 2 * tt;
                    const float qpos = (float)(q0 + t * kQT + row + shift);
                    const uint32_t m = m_loc++;
                    const int buf = m & 1, bar = wg * 2 + buf;
                    const uint32_t par = (m >> 1) & 1;
                    mbar_wait(&s_full[bar], par);
                    tc_fence_after();
                    const uint32_t s_addr = t_lane + wg * 128 + buf * kKB, o_addr = t_lane + 256 + (ib * 2 + t) * 64;
                    uint8_t* my_p = sm_p + bar * kTileBytes;
                    uint64_t* prev_bar = m > 0 ? &p_empty[bar ^ 1] : nullptr;       // PV of this warpgroup's step m-1
                    const uint32_t prev_par = ((m - 1) >> 1) & 1;
                    if (tail) softmax_tile<ALIBI, true>(s_addr, o_addr, j > 0, &s_empty[bar], &p_empty[bar], par ^ 1, prev_bar, prev_par, p.scale_log2, slope, qpos, key0, Sk, m_ref[tt], l_run[tt], my_p, row, lane);
                    else      softmax_tile<ALIBI, false>(s_addr, o_addr, j > 0, &s_empty[bar], &p_empty[bar], par ^ 1, prev_bar, prev_par, p.scale_log2, slope, qpos, key0, Sk, m_ref[tt], l_run[tt], my_p, row, lane);
                    fence_proxy_async_smem();                     // generic-proxy writes -> visible to the UMMA (async proxy)
                    tc_fence_before();                            // orders a possible tcgen05.st rescale before the PV
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[bar]);
                }
            }
            // ---- epilogue: O_t / l -> bf16 -> global ----
            mbar_wait(&o_full[ib], ipar);
            tc_fence_after();
            for (int tt = 0; wg + 2 * tt < nq; ++tt) {
                const int t = wg + 2 * tt;
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(t_lane + 256 + (ib * 2 + t) * 64, o0);
                tmem_ld_32x32(t_lane + 256 + (ib * 2 + t) * 64 + 32, o1);
                tmem_ld_wait();
                const int qi = q0 + t * kQT + row;
                if (qi < Sq) {
                    const float inv = 1.0f / l_run[tt];
                    __nv_bfloat16* dst = p.o + (size_t)(qbeg + qi) * p.ldo + head * HD;
#pragma unroll
                    for (int e = 0; e < 32; e += 8)
                        *reinterpret_cast<uint4*>(dst + e) = make_uint4(
                            pack_bf16x2(__uint_as_float(o0[e]) * inv, __uint_as_float(o0[e + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o0[e + 2]) * inv, __uint_as_float(o0[e + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o0[e + 4]) * inv, __uint_as_float(o0[e + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o0[e + 6]) * inv, __uint_as_float(o0[e + 7]) * inv));
#pragma unroll
                    for (int e = 0; e < HD - 32; e += 8)
                        *reinterpret_cast<uint4*>(dst + 32 + e) = make_uint4(
                            pack_bf16x2(__uint_as_float(o1[e]) * inv, __uint_as_float(o1[e + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o1[e + 2]) * inv, __uint_as_float(o1[e + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o1[e + 4]) * inv, __uint_as_float(o1[e + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o1[e + 6]) * inv, __uint_as_float(o1[e + 7]) * inv));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[ib]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap_rows64(CUtensorMap* tm, const void* base, long rows, int cols, int ld, int box_rows) {
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_encodeTiled>(fp);
    }
    VF_REQUIRE(enc, "cuTensorMapEncodeTiled entry point not available");
    VF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, "attention operands must be 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for an attention operand", (int)r);
    return 0;
}

template <int HD, bool ALIBI, bool SHORT>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p,
                          cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        VF_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<HD, ALIBI, SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kAttnSmem));
        attr_set = true;
    }
    int dev = 0, sms = 0;
    VF_CUDA_OK(cudaGetDevice(&dev));
    VF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long work = (long)p.n_items * p.heads;
    const int grid = (int)(work < sms ? work : sms);
    attention_tc_kernel<HD, ALIBI, SHORT><<<grid, kAttnThreads, kAttnSmem, s>>>(tq, tk, tv, p);
    VF_LAUNCH_OK("attention_tc_kernel launch");
    return 0;
}

int attention_tc_varlen(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                        long rows_q, long rows_k, const int* cu_q, const int* cu_k, const int* item_seq,
                        const int* item_q0, int n_items, int heads, int head_dim, const float* slopes,
                        int short_items, cudaStream_t stream) {
    VF_REQUIRE(head_dim == 48 || head_dim == 64, "attention_tc: head_dim %d not supported (48/64)", head_dim);
    VF_REQUIRE(ldo % 8 == 0, "attention_tc: output stride must keep 16-byte alignment");
    if (n_items == 0) return 0;
    const int d = heads * head_dim;
    CUtensorMap tq, tk, tv;
    if (make_tmap_rows64(&tq, q, rows_q, d, ldq, kQT)) return -1;
    if (make_tmap_rows64(&tk, k, rows_k, d, ldk, kKB)) return -1;
    if (make_tmap_rows64(&tv, v, rows_k, d, ldv, kKB)) return -1;
    AttnTcParams p;
    p.cu_q = cu_q; p.cu_k = cu_k; p.item_seq = item_seq; p.item_q0 = item_q0; p.n_items = n_items; p.heads = heads;
    p.o = (__nv_bfloat16*)o; p.ldo = ldo; p.slopes = slopes;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
#define VF_ATTN_DISPATCH(HD_)                                                                                     \
    if (short_items) return slopes ? launch_attn_tc<HD_, true, true>(tq, tk, tv, p, stream)                      \
                                   : launch_attn_tc<HD_, false, true>(tq, tk, tv, p, stream);                    \
    return slopes ? launch_attn_tc<HD_, true, false>(tq, tk, tv, p, stream)                                      \
                  : launch_attn_tc<HD_, false, false>(tq, tk, tv, p, stream);
    if (head_dim == 48) { VF_ATTN_DISPATCH(48) }
    VF_ATTN_DISPATCH(64)
#undef VF_ATTN_DISPATCH
}

}  // namespace vf
