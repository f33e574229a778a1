// vf_attention_tc128.cu — the 128-key-block variant of vf_attention_tc.cu (same algorithm and contract; one S and
// one P buffer per softmax warpgroup, 3-stage K / 2-stage V rings).  Per-tile fixed costs are amortised over twice
// as many keys, which wins for long key ranges (the stacked gene->CRE cross-attention, 1024 keys); the 64-key,
// double-buffered variant wins for short sequences.  The two files are unified in a later round.
// --- original header follows ---
// vf_attention_tc.cu — variable-length non-causal attention on the 5th-gen tensor cores.
//
// Replaces flash_attn's varlen kernels (see vf_attention.cu for the reference call sites) for every problem with
// at least two 128-row query tiles per (sequence, head) work item; the legacy mma.sync kernel keeps the tiny
// seq2reg windows.  Semantics: softmax((q·k)/sqrt(hd) - slope_h*|i + Sk - Sq - j|) v, fp32 statistics.
//
// One persistent CTA per SM walks work items = (sequence, head, block of up to 4 query tiles of 128 rows).
//   warp 0      TMA producer: Q tiles once per item, K blocks (128 keys) through a 3-stage ring, V blocks through
//               a 2-stage ring; all tiles are [rows, 64 cols] bf16, SWIZZLE_128B (hd=48 over-fetches 16 columns
//               that no MMA ever reads).
//   warp 1      tcgen05.mma issuer.  S = Q K^T : M128 x N128 x K(hd) into one of two TMEM S buffers;
//               O_t += P V : M128 x N(hd) x K128 with P from shared memory (K-major) and V as an MN-major operand.
//   warps 2-9   two softmax warpgroups (query tile t is owned by warpgroup t&1): ONE THREAD PER QUERY ROW, the row
//               arrives straight from TMEM (tcgen05.ld 32x32b), so max / sum need no cross-thread reduction.
// Softmax is single-pass with a LAZY reference maximum (exact arithmetic, FA4-style): every row keeps a reference
// m_ref; P = exp2(x - m_ref) may reach 2^8 before m_ref is raised.  m_ref is only raised (and l / O_t rescaled by
// exp2(m_old - m_new), O_t through tcgen05.ld/st) when a 32-key slab exceeds it by more than 8 in the log2 domain:
// for free on the first slab of a tile, otherwise by a rare restart of the tile with its exact maximum.
// TMEM: columns [0,128) S0, [128,256) S1, [256 + 64 t, +64) O_t  (512 columns).
#include <cuda.h>

#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {
namespace tc128 {

constexpr int kQT = 128;            // query rows per tile
constexpr int kKB = 128;            // keys per block
constexpr int kMaxQT = 4;           // query tiles per work item
constexpr int kKStages = 3, kVStages = 2;
constexpr uint32_t kTileBytes = 128 * 64 * 2;                 // one [128 x 64] bf16 tile
constexpr int kAttnThreads = 64 + 256;
constexpr size_t kAttnSmem = 1024 + (size_t)(kMaxQT + kKStages + kVStages) * kTileBytes + 2 * 2 * kTileBytes + 512;

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
// sub-tile = 64 keys, 16-byte chunk index XOR (row & 7); slab c covers chunks (c&1)*4 .. +3 of sub-tile c>>1)
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
    uint8_t* base = dst + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = ((c & 1) * 4 + q4) ^ (row & 7);
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

// One 128 x 128 score tile for this thread's row.  Returns through m_ref / l / (O_t in TMEM) / the P tile.
template <bool ALIBI, bool TAIL>
__device__ __forceinline__ void softmax_tile(uint32_t s_addr, uint32_t o_addr, bool have_o, uint64_t* s_empty_bar,
                                             float scale, float slope, float qpos, int key0, int Sk, float& m_ref,
                                             float& l, uint8_t* dst, int row, int lane) {
    uint32_t r[2][32];
    float sum = 0.f;
    bool restart = false;
    float seen = -INFINITY;
    tmem_ld_32x32(s_addr, r[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c + 1 < 4) tmem_ld_32x32(s_addr + (c + 1) * 32, r[(c + 1) & 1]);
        const float d0 = qpos - (float)(key0 + c * 32);
        const int nvalid = Sk - (key0 + c * 32);
        if constexpr (ALIBI || TAIL) {                       // biased / masked scores are computed once, in place
#pragma unroll
            for (int e = 0; e < 32; ++e)
                r[c & 1][e] = __float_as_uint(score<ALIBI, TAIL>(r[c & 1][e], e, scale, slope, d0, nvalid));
        }
        const float cmax = slab_max<ALIBI, TAIL>(r[c & 1], scale, slope, d0, nvalid);
        seen = fmaxf(seen, cmax);
        if (__any_sync(0xffffffffu, cmax > m_ref + kLazyThreshold)) {
            if (c == 0) {                                    // nothing of this tile is written yet: raise in place
                const float m_new = fmaxf(m_ref, cmax);
                const float corr = ex2_approx(m_ref - m_new);             // m_ref = -inf -> 0
                l *= corr;
                if (have_o) rescale_o(o_addr, corr);
                m_ref = m_new;
            } else {
                restart = true;                              // warp-uniform
            }
        }
        if (restart) break;
        if (c == 3) {                                        // whole tile is in registers and accepted: release S
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty_bar);
        }
        slab_exp_store<ALIBI, TAIL>(r[c & 1], scale, slope, d0, nvalid, m_ref, sum, dst, c, row);
    }
    if (restart) {
        // rare: exact maximum of the whole tile first, then one clean pass (no further raise can trigger)
        tmem_ld_wait();                                      // drain the in-flight prefetch
        for (int c = 0; c < 4; ++c) {
            tmem_ld_32x32(s_addr + c * 32, r[0]);
            tmem_ld_wait();
            if constexpr (ALIBI || TAIL) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    r[0][e] = __float_as_uint(score<ALIBI, TAIL>(r[0][e], e, scale, slope, qpos - (float)(key0 + c * 32), Sk - (key0 + c * 32)));
            }
            seen = fmaxf(seen, slab_max<ALIBI, TAIL>(r[0], scale, slope, qpos - (float)(key0 + c * 32), Sk - (key0 + c * 32)));
        }
        const float m_new = fmaxf(m_ref, seen);
        const float corr = ex2_approx(m_ref - m_new);
        l *= corr;
        if (have_o) rescale_o(o_addr, corr);
        m_ref = m_new;
        sum = 0.f;
        for (int c = 0; c < 4; ++c) {
            tmem_ld_32x32(s_addr + c * 32, r[0]);
            tmem_ld_wait();
            if constexpr (ALIBI || TAIL) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    r[0][e] = __float_as_uint(score<ALIBI, TAIL>(r[0][e], e, scale, slope, qpos - (float)(key0 + c * 32), Sk - (key0 + c * 32)));
            }
            if (c == 3) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty_bar);
            }
            slab_exp_store<ALIBI, TAIL>(r[0], scale, slope, qpos - (float)(key0 + c * 32), Sk - (key0 + c * 32), m_ref,
                                        sum, dst, c, row);
        }
    }
    l += sum;
}

template <int HD, bool ALIBI>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sm_q = smem;                                         // kMaxQT tiles
    uint8_t* sm_k = sm_q + kMaxQT * kTileBytes;                   // kKStages tiles
    uint8_t* sm_v = sm_k + kKStages * kTileBytes;                 // kVStages tiles
    uint8_t* sm_p = sm_v + kVStages * kTileBytes;                 // 2 buffers x (2 sub-tiles of 64 keys)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 4 * kTileBytes);
    uint64_t* q_full = bars + 0;  uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;  uint64_t* k_empty = bars + 5;   // [3] each
    uint64_t* v_full = bars + 8;  uint64_t* v_empty = bars + 10;  // [2] each
    uint64_t* s_full = bars + 12; uint64_t* s_empty = bars + 14;  // [2] each
    uint64_t* p_full = bars + 16; uint64_t* p_empty = bars + 18;  // [2] each
    uint64_t* o_full = bars + 20; uint64_t* o_empty = bars + 21;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = p.n_items * p.heads;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int i = 0; i < kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1); mbar_init(o_empty, 8);
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
                mbar_wait(q_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, nq * kTileBytes);
                for (int t = 0; t < nq; ++t) tma_load_2d(sm_q + t * kTileBytes, &tmQ, q_full, col, qbeg + q0 + t * kQT);
                for (int j = 0; j < nk; ++j) {
                    mbar_wait(&k_empty[ks], kph ^ 1);
                    mbar_arrive_expect_tx(&k_full[ks], kTileBytes);
                    tma_load_2d(sm_k + ks * kTileBytes, &tmK, &k_full[ks], col, kbeg + j * kKB);
                    if (++ks == kKStages) { ks = 0; kph ^= 1; }
                    mbar_wait(&v_empty[vs], vph ^ 1);
                    mbar_arrive_expect_tx(&v_full[vs], kTileBytes);
                    tma_load_2d(sm_v + vs * kTileBytes, &tmV, &v_full[vs], col, kbeg + j * kKB);
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
            uint32_t n_s[2] = {0, 0}, n_p[2] = {0, 0};                                      // uses of S / P buffers
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk;
                decode(w, seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk);
                mbar_wait(q_full, it & 1);
                tc_fence_after();
                auto issue_qk = [&](int t, int kstage) {
                    const int b = t & 1;
                    mbar_wait(&s_empty[b], (n_s[b] & 1) ^ 1); ++n_s[b];
                    tc_fence_after();
                    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sm_q + t * kTileBytes));
                    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sm_k + kstage * kTileBytes));
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + b * kKB, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                    umma_commit(&s_full[b]);
                };
                // ---- S = Q K^T two steps ahead of O_t += P V ----
                mbar_wait(o_empty, (it & 1) ^ 1);
                tc_fence_after();
                // linearised steps n = j*nq + t.  S(n+2) is issued BEFORE waiting for P(n): the warpgroup that owns
                // step n finds its next score tile ready the moment it finishes writing P(n).
                const int N = nk * nq;
                int kj_issued = 0;                                  // key blocks whose k_full has been awaited
                int ks_of[2] = {0, 0};                              // smem stage of key block j (parity-indexed, <=2 live)
                int vs_of[2] = {0, 0};
                auto ensure_k = [&](int j) {                        // wait for K_j / V_j exactly once, in order
                    while (kj_issued <= j) {
                        mbar_wait(&k_full[ks], kph);
                        mbar_wait(&v_full[vs], vph);
                        tc_fence_after();
                        ks_of[kj_issued & 1] = ks; vs_of[kj_issued & 1] = vs;
                        if (++ks == kKStages) { ks = 0; kph ^= 1; }
                        if (++vs == kVStages) { vs = 0; vph ^= 1; }
                        ++kj_issued;
                    }
                };
                auto qk_step = [&](int n) {
                    const int j = n / nq, t = n - j * nq;
                    ensure_k(j);
                    issue_qk(t, ks_of[j & 1]);
                    if (t == nq - 1) umma_commit(&k_empty[ks_of[j & 1]]);
                };
                // look-ahead of 2 steps needs two live key blocks at most (nq >= 2); a single query tile per item
                // looks ahead one step so that the 2-stage V ring and the parity-indexed stage tables stay valid
                const int LA = nq >= 2 ? 2 : 1;
                for (int n = 0; n < LA && n < N; ++n) qk_step(n);
                for (int n = 0; n < N; ++n) {
                    if (n + LA < N) qk_step(n + LA);
                    const int j = n / nq, t = n - j * nq, b = t & 1;
                    mbar_wait(&p_full[b], n_p[b] & 1); ++n_p[b];
                    tc_fence_after();
                    const uint32_t d = tmem_base + 256 + t * 64;
                    const int vstage = vs_of[j & 1];
#pragma unroll
                    for (int kk = 0; kk < kKB / 16; ++kk) {
                        const uint64_t da = umma_desc_kmajor_sw128(
                            smem_u32(sm_p + (b * 2 + (kk >> 2)) * kTileBytes)) + 2 * (kk & 3);
                        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sm_v + vstage * kTileBytes + kk * 2048));
                        umma_bf16(d, da, db, idesc_pv, (j | kk) != 0);
                    }
                    umma_commit(&p_empty[b]);
                    if (t == nq - 1) umma_commit(&v_empty[vstage]);
                }
                umma_commit(o_full);
                umma_commit(q_empty);
            }
        }
        __syncwarp();
    } else {
        // ============================ softmax warpgroups ============================
        const int wg = (warp - 2) >> 2;                       // owns query tiles t with (t & 1) == wg
        const int quad = warp & 3;                            // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;                     // row inside a 128-row query tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint8_t* my_p = sm_p + wg * 2 * kTileBytes;
        uint32_t n_s = 0, n_p = 0; int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk;
            decode(w, seq, head, q0, qbeg, Sq, kbeg, Sk, nq, nk);
            const float slope = ALIBI ? p.slopes[head] * 1.4426950408889634f : 0.f;
            const int shift = Sk - Sq;
            float m_ref[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
            for (int j = 0; j < nk; ++j) {
                const int key0 = j * kKB;
                const bool tail = key0 + kKB > Sk;                // block holds keys beyond the sequence
                for (int tt = 0; wg + 2 * tt < nq; ++tt) {
                    const int t = wg + 2 * tt;
                    const float qpos = (float)(q0 + t * kQT + row + shift);
                    mbar_wait(&s_full[wg], n_s & 1); ++n_s;
                    // P buffer free <=> the PV two of this warpgroup's tiles back has retired; every earlier PV
                    // (in particular the last one that wrote O_t) has then retired too, so O_t may be rescaled
                    mbar_wait(&p_empty[wg], (n_p & 1) ^ 1); ++n_p;
                    tc_fence_after();
                    const uint32_t s_addr = t_lane + wg * kKB, o_addr = t_lane + 256 + t * 64;
                    if (tail) softmax_tile<ALIBI, true>(s_addr, o_addr, j > 0, &s_empty[wg], p.scale_log2, slope, qpos, key0, Sk, m_ref[tt], l_run[tt], my_p, row, lane);
                    else      softmax_tile<ALIBI, false>(s_addr, o_addr, j > 0, &s_empty[wg], p.scale_log2, slope, qpos, key0, Sk, m_ref[tt], l_run[tt], my_p, row, lane);
                    fence_proxy_async_smem();                     // generic-proxy writes -> visible to the UMMA (async proxy)
                    tc_fence_before();                            // orders a possible tcgen05.st rescale before the PV
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[wg]);
                }
            }
            // ---- epilogue: O_t / l -> bf16 -> global ----
            mbar_wait(o_full, it & 1);
            tc_fence_after();
            for (int tt = 0; wg + 2 * tt < nq; ++tt) {
                const int t = wg + 2 * tt;
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(t_lane + 256 + t * 64, o0);
                tmem_ld_32x32(t_lane + 256 + t * 64 + 32, o1);
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
            if (lane == 0) mbar_arrive(o_empty);
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

static int make_tmap_rows64(CUtensorMap* tm, const void* base, long rows, int cols, int ld) {
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
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for an attention operand", (int)r);
    return 0;
}

template <int HD, bool ALIBI>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p,
                          cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        VF_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<HD, ALIBI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kAttnSmem));
        attr_set = true;
    }
    int dev = 0, sms = 0;
    VF_CUDA_OK(cudaGetDevice(&dev));
    VF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long work = (long)p.n_items * p.heads;
    const int grid = (int)(work < sms ? work : sms);
    attention_tc_kernel<HD, ALIBI><<<grid, kAttnThreads, kAttnSmem, s>>>(tq, tk, tv, p);
    VF_LAUNCH_OK("attention_tc_kernel launch");
    return 0;
}

int attention_tc128_varlen(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                        long rows_q, long rows_k, const int* cu_q, const int* cu_k, const int* item_seq,
                        const int* item_q0, int n_items, int heads, int head_dim, const float* slopes,
                        cudaStream_t stream) {
    VF_REQUIRE(head_dim == 48 || head_dim == 64, "attention_tc: head_dim %d not supported (48/64)", head_dim);
    VF_REQUIRE(ldo % 8 == 0, "attention_tc: output stride must keep 16-byte alignment");
    if (n_items == 0) return 0;
    const int d = heads * head_dim;
    CUtensorMap tq, tk, tv;
    if (make_tmap_rows64(&tq, q, rows_q, d, ldq)) return -1;
    if (make_tmap_rows64(&tk, k, rows_k, d, ldk)) return -1;
    if (make_tmap_rows64(&tv, v, rows_k, d, ldv)) return -1;
    AttnTcParams p;
    p.cu_q = cu_q; p.cu_k = cu_k; p.item_seq = item_seq; p.item_q0 = item_q0; p.n_items = n_items; p.heads = heads;
    p.o = (__nv_bfloat16*)o; p.ldo = ldo; p.slopes = slopes;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    if (head_dim == 48) return slopes ? launch_attn_tc<48, true>(tq, tk, tv, p, stream) : launch_attn_tc<48, false>(tq, tk, tv, p, stream);
    return slopes ? launch_attn_tc<64, true>(tq, tk, tv, p, stream) : launch_attn_tc<64, false>(tq, tk, tv, p, stream);
}

}  // namespace tc128

int attention_tc128_varlen(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                           long rows_q, long rows_k, const int* cu_q, const int* cu_k, const int* item_seq,
                           const int* item_q0, int n_items, int heads, int head_dim, const float* slopes,
                           cudaStream_t stream) {
    return tc128::attention_tc128_varlen(q, ldq, k, ldk, v, ldv, o, ldo, rows_q, rows_k, cu_q, cu_k, item_seq, item_q0,
                                         n_items, heads, head_dim, slopes, stream);
}
}  // namespace vf
