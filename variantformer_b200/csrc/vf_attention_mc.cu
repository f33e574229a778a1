// vf_attention_mc.cu — variable-length non-causal attention on the 5th-gen tensor cores, TWO CTAs PER SM.
//
// Same contract as vf_attention_tc.cu (flash_attn varlen call sites: seq2reg/modules.py:159-171,
// seq2gene/modules/layers.py:372-467): softmax((q·k)/sqrt(hd) - slope_h*|i + Sk - Sq - j|) v, fp32 statistics.
//
// Why a second kernel: at head_dim 48/64 attention is bound by the softmax (one ex2 per score, 16/clk/SM), not by
// the tensor pipe, and with one CTA per SM the 8 softmax warps (2 per scheduler) leave most issue slots and ~60 % of
// the MUFU pipe idle behind TMEM-load, barrier and MUFU latencies (ncu: profiles/r01_attention_tc_*).  This kernel
// halves every per-CTA resource (256 TMEM columns, ~106 KB shared memory, 2 query tiles per work item) so that two
// CTAs are co-resident: 16 softmax warps per SM in four independent pipelines whose bubbles overlap.
//
// One CTA = 384 threads = 3 warpgroups:
//   warp 0      TMA producer: Q tiles of the item's (up to) two SLOTS, K / V blocks of 64 keys through 3- / 2-stage
//               rings; all tiles are [rows, 64 cols] bf16, SWIZZLE_128B (hd=48 over-fetches 16 columns never read)
//   warps 1-3   idle: they exist so that warpgroup 0 can hand its registers to the softmax warps (setmaxnreg)
//   warps 4-11  two softmax warpgroups, warpgroup s owns slot s: ONE THREAD PER QUERY ROW, the 64 scores of a row
//               come straight from TMEM (tcgen05.ld 32x32b) into registers.  There is NO separate MMA warp: one
//               thread of each warpgroup issues its slot's tcgen05.mma the moment the warpgroup (named barrier) has
//               the previous scores in registers / has written P, so no hand-shake with another warp sits on the
//               critical path:  S_s = Q_s K^T (M128 x N64 x K=hd) into slot s's TMEM score buffer;
//               O_s += P_s V (M128 x N=hd x K64), P from shared memory (K-major), V as an MN-major operand.
//               The first Q K^T of the next item is issued during the last tile of the current one.
// A work item is (head, two slots); a slot is a 128-row query tile of some sequence, described by a host-built record
// {first query row, valid rows, first key row, keys, ALiBi position of row 0} (absolute row numbers: no cu_seqlens
// lookups on the device).  Two slots with the same key range share one K/V stream (long sequences: every K/V block
// feeds two query tiles); slots of different sequences (short sequences, <= 128 rows: seq2reg windows) stream their
// own K/V, so both warpgroups stay busy either way.
//
// Softmax: single pass with a LAZY reference maximum (exact arithmetic): P = exp2(x - m_ref) may reach 2^8 before
// m_ref is raised; raising rescales l and (through tcgen05.ld/st) O_s by exp2(m_old - m_new).  The whole 64-key
// score row is in registers before anything is written, so a raise never needs a restart.  The affine part of the
// exponent (scale, -m_ref, and away from the diagonal the ALiBi term, which is linear in the key index there) is
// folded into one or two FFMA per score.
// TMEM (256 columns per CTA): [64 s, +64) scores of slot s, [128 + 64 s, +64) O_s.
#include <cuda.h>

#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {
namespace mc {

constexpr int kQT = 128;            // query rows per slot
constexpr int kKB = 64;             // keys per block
constexpr int kKStages = 3, kVStages = 2;
constexpr uint32_t kQBytes = 128 * 64 * 2;                    // one [128 x 64] bf16 tile (Q tile, P buffer)
constexpr uint32_t kKvBytes = kKB * 64 * 2;                   // one [64 keys x 64] bf16 tile
constexpr int kThreads = 384;
constexpr int kFirstSoftmaxWarp = 4;
constexpr uint32_t kTmemCols = 256;
constexpr size_t kSmem = 1024 + 2 * kQBytes + (size_t)(kKStages + kVStages) * kKvBytes + 2 * kQBytes + 256;
constexpr float kLazyThreshold = 8.0f;      // log2 units: P stays below 2^8 between reference-max updates

// one slot of a work item (32 bytes): nrows == 0 marks an empty slot
struct SlotRec { int qrow, nrows, krow, Sk, qpos0, pad0, pad1, pad2; };

struct Params {
    const SlotRec* slots;                         // [n_items][2]
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

// O_s row *= corr, in tensor memory, 8 columns at a time (few live registers: the caller holds a whole score row).
// Warp-collective; lanes that need no change pass corr = 1.
template <int HD>
__device__ __forceinline__ void rescale_o(uint32_t o_addr, float corr) {
#pragma unroll 1
    for (int c = 0; c < HD; c += 8) {
        uint32_t u[8];
        tmem_ld_32x8(o_addr + c, u);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = __float_as_uint(__uint_as_float(u[e]) * corr);
        tmem_st_32x8(o_addr + c, u);
    }
    tmem_st_wait();
}

// exponent of one score: x = raw * scale + (bias - m_ref), written back in place; returns nothing.
// MODE 0: no positional term.  MODE 1: ALiBi, the whole 64-key block on one side of the diagonal for every row of the
// warp: bias = c + sl * e (sl = +-slope, c folded by the caller).  MODE 2: general (diagonal block and / or tail mask).
template <int MODE>
__device__ __forceinline__ void exponents(uint32_t (&r)[64], float scale, float c, float sl, float slope, float d0,
                                          int nvalid, float& mx_out) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int e = 0; e < 64; ++e) {
        float x;
        if constexpr (MODE == 0) {
            x = fmaf(__uint_as_float(r[e]), scale, c);
        } else if constexpr (MODE == 1) {
            x = fmaf(__uint_as_float(r[e]), scale, fmaf(sl, (float)e, c));
        } else {
            x = fmaf(__uint_as_float(r[e]), scale, fmaf(-slope, fabsf(d0 - (float)e), c));
            if (e >= nvalid) x = -INFINITY;
        }
        r[e] = __float_as_uint(x);
        mx[e & 3] = fmaxf(mx[e & 3], x);
    }
    mx_out = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// One 128 x 64 score tile for this thread's row (r = the 64 raw scores): -> P (shared memory, bf16, K-major
// SWIZZLE_128B: 16-byte chunk index XOR (row & 7)), running m_ref / l, O_s rescaled in TMEM when the reference
// maximum is raised.
template <int HD, bool ALIBI>
__device__ __forceinline__ void softmax_tile(uint32_t (&r)[64], uint32_t o_addr, bool first, uint64_t* p_empty_bar,
                                             uint32_t p_empty_parity, float scale, float slope, float qpos, int key0,
                                             int Sk, float& m_ref, float& l, uint8_t* dst, int row, int lane) {
    const float base = first ? 0.f : m_ref;                  // exponents are first taken against `base`
    const float d0 = qpos - (float)key0;                     // query position minus the block's first key
    const int nvalid = Sk - key0;
    float mx;
    bool general = nvalid < kKB;
    if constexpr (ALIBI) general = general || !(__all_sync(0xffffffffu, d0 >= 63.f) || __all_sync(0xffffffffu, d0 <= 0.f));
    if (general) {
        exponents<2>(r, scale, -base, 0.f, ALIBI ? slope : 0.f, d0, nvalid, mx);
    } else if constexpr (ALIBI) {
        const float sl = d0 > 0.f ? slope : -slope;          // |d0 - e| = +-(d0 - e) for the whole block
        exponents<1>(r, scale, -slope * fabsf(d0) - base, sl, slope, d0, nvalid, mx);
    } else {
        exponents<0>(r, scale, -base, 0.f, 0.f, d0, nvalid, mx);
    }
    // P buffer free <=> the PV that read it (this slot's previous step) has retired; that PV is also the last writer
    // of O_s, so O_s may be rescaled below
    mbar_wait(p_empty_bar, p_empty_parity);
    tc_fence_after();
    if (first || __any_sync(0xffffffffu, mx > kLazyThreshold)) {
        // first tile: the exact row maximum becomes the reference (may be negative).  later: raise by max(mx, 0).
        const float delta = first ? mx : fmaxf(mx, 0.f);
        const float corr = first ? 0.f : ex2_approx(-delta);
        l *= corr;
        if (!first) rescale_o<HD>(o_addr, corr);
        m_ref = base + delta;
#pragma unroll
        for (int e = 0; e < 64; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) - delta);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t* prow = dst + row * 128;
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) {                          // 8 scores -> one 16-byte chunk of the P row
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float p0 = ex2_approx(__uint_as_float(r[q8 * 8 + 2 * h]));
            const float p1 = ex2_approx(__uint_as_float(r[q8 * 8 + 2 * h + 1]));
            acc[h] += p0 + p1;
            pk[h] = pack_bf16x2(p0, p1);
        }
        *reinterpret_cast<uint4*>(prow + ((q8 ^ (row & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    l += (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// global position of K/V block (j, slot) of an item inside the CTA's K / V rings, relative to the item's first load.
// One stream when the slots share their keys; otherwise the blocks of the two slots alternate while both have keys.
__device__ __forceinline__ int ring_offset(int j, int s, bool same, int nk0, int nk1) {
    if (same) return j;
    return min(j, nk0) + min(j, nk1) + ((s == 1 && j < nk0) ? 1 : 0);
}

template <int HD, bool ALIBI>
__global__ void __launch_bounds__(kThreads, 2)
attention_mc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sm_q = smem;                                         // 2 slots x 16 KB
    uint8_t* sm_k = sm_q + 2 * kQBytes;                           // kKStages x 8 KB
    uint8_t* sm_v = sm_k + kKStages * kKvBytes;                   // kVStages x 8 KB
    uint8_t* sm_p = sm_v + kVStages * kKvBytes;                   // 2 slots x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 2 * kQBytes);
    uint64_t* q_full = bars + 0;   uint64_t* q_empty = bars + 2;                   // [slot]
    uint64_t* s_full = bars + 4;   uint64_t* p_empty = bars + 6;                   // [slot]
    uint64_t* o_full = bars + 8;                                                   // [slot]
    uint64_t* k_full = bars + 10;  uint64_t* k_empty = k_full + kKStages;
    uint64_t* v_full = k_empty + kKStages;   uint64_t* v_empty = v_full + kVStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + kVStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = p.n_items * p.heads;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1);
        }
        // a K / V stage is released by TWO tcgen05.commit arrivals: one per slot when the slots share the block, both
        // from its only user otherwise
        for (int i = 0; i < kKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }
        for (int i = 0; i < kVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work index w -> (head, item): head-major so concurrently running CTAs share one head's K/V in L2.  Every role
    // reads only the fields it needs from the item's two slot records.
    const int4* recs = reinterpret_cast<const int4*>(p.slots);         // 4 int4 per item: {s0.lo, s0.hi, s1.lo, s1.hi}

    if (warp < kFirstSoftmaxWarp) {
        reg_dealloc<32>();
        if (warp == 0 && elect_one()) {
            // ============================ TMA producer ============================
            uint32_t ring = 0;                               // K / V blocks loaded so far (same count for both rings)
            uint32_t nq[2] = {0, 0};                         // Q tiles loaded so far per slot
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int head = w / p.n_items, item = w - head * p.n_items;
                const int4 r0 = __ldg(recs + 4 * item), r1 = __ldg(recs + 4 * item + 2);   // {qrow, nrows, krow, Sk}
                const int qrow[2] = {r0.x, r1.x}, krow[2] = {r0.z, r1.z};
                const int nk[2] = {r0.y > 0 ? (r0.w + kKB - 1) / kKB : 0, r1.y > 0 ? (r1.w + kKB - 1) / kKB : 0};
                const bool same = nk[0] > 0 && nk[1] > 0 && r0.z == r1.z;
                const int col = head * HD;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (nk[s] == 0) continue;
                    const uint32_t m = nq[s]++;
                    mbar_wait(&q_empty[s], (m & 1) ^ 1);
                    mbar_arrive_expect_tx(&q_full[s], kQBytes);
                    tma_load_2d(sm_q + s * kQBytes, &tmQ, &q_full[s], col, qrow[s]);
                }
                // K / V blocks in exactly the order the warpgroups consume them: K of block j+1 before V of block j
                const int nkmax = max(nk[0], nk[1]);
                const uint32_t base = ring;
                auto load_k = [&](int j, int s) {
                    const uint32_t idx = base + ring_offset(j, s, same, nk[0], nk[1]);
                    const uint32_t st = idx % kKStages;
                    mbar_wait(&k_empty[st], ((idx / kKStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&k_full[st], kKvBytes);
                    tma_load_2d(sm_k + st * kKvBytes, &tmK, &k_full[st], col, krow[s] + j * kKB);
                };
                auto load_v = [&](int j, int s) {
                    const uint32_t idx = base + ring_offset(j, s, same, nk[0], nk[1]);
                    const uint32_t st = idx % kVStages;
                    mbar_wait(&v_empty[st], ((idx / kVStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&v_full[st], kKvBytes);
                    tma_load_2d(sm_v + st * kKvBytes, &tmV, &v_full[st], col, krow[s] + j * kKB);
                };
#pragma unroll
                for (int s = 0; s < 2; ++s) if (nk[s] > 0 && !(s == 1 && same)) load_k(0, s);
                for (int j = 0; j < nkmax; ++j) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) if (j + 1 < nk[s] && !(s == 1 && same)) load_k(j + 1, s);
#pragma unroll
                    for (int s = 0; s < 2; ++s) if (j < nk[s] && !(s == 1 && same)) load_v(j, s);
                }
                ring += same ? nk[0] : nk[0] + nk[1];
            }
        }
    } else {
        // ============================ softmax warpgroups (they also issue their slot's MMAs) ============================
        reg_alloc<104>();
        const int s = (warp - kFirstSoftmaxWarp) >> 2;        // slot owned by this warpgroup
        const int quad = warp & 3;                            // TMEM lane quadrant of this warp
        const int row = quad * 32 + lane;                     // row inside the 128-row query tile
        const bool issuer = (threadIdx.x & 127) == 0;         // the one thread of the warpgroup that issues tcgen05.mma
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t s_addr = t_lane + s * kKB, o_addr = t_lane + 128 + s * 64;
        uint8_t* my_p = sm_p + s * kQBytes;
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kKB);                         // A, B K-major
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD) | (1u << 16);            // B (= V) MN-major
        uint32_t ring = 0;                                    // K / V blocks the CTA has consumed before this item
        uint32_t n_tiles = 0, n_items_mine = 0;               // score tiles / items this slot has processed so far

        auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + s) : "memory"); };
        // S = Q K^T for key block j of an item (issuer thread only)
        auto issue_qk = [&](uint32_t base, int j, bool same, int nk0, int nk1, bool first_of_item, bool last_of_item) {
            const uint32_t idx = base + ring_offset(j, s, same, nk0, nk1);
            const uint32_t st = idx % kKStages;
            if (first_of_item) mbar_wait(&q_full[s], n_items_mine & 1);
            mbar_wait(&k_full[st], (idx / kKStages) & 1);
            tc_fence_after();
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sm_q + s * kQBytes));
            const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sm_k + st * kKvBytes));
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + s * kKB, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
            umma_commit(&s_full[s]);
            umma_commit(&k_empty[st]);
            if (!same) umma_commit(&k_empty[st]);             // sole user of the block: both release arrivals
            if (last_of_item) umma_commit(&q_empty[s]);       // the Q slot may be refilled
        };
        auto issue_pv = [&](uint32_t base, int j, bool same, int nk0, int nk1, bool last_of_item) {
            const uint32_t idx = base + ring_offset(j, s, same, nk0, nk1);
            const uint32_t st = idx % kVStages;
            mbar_wait(&v_full[st], (idx / kVStages) & 1);
            tc_fence_after();
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(my_p));
            const uint32_t vb = smem_u32(sm_v + st * kKvBytes);
#pragma unroll
            for (int kk = 0; kk < kKB / 16; ++kk)
                umma_bf16(tmem_base + 128 + s * 64, da + 2 * kk, umma_desc_kmajor_sw128(vb + kk * 2048), idesc_pv,
                          (j | kk) != 0);
            umma_commit(&p_empty[s]);
            umma_commit(&v_empty[st]);
            if (!same) umma_commit(&v_empty[st]);
            if (last_of_item) umma_commit(&o_full[s]);
        };
        struct Item { int qrow, nrows, Sk, qpos0, nk0, nk1, head; bool same; };
        auto fetch = [&](int w, Item& it) {
            it.head = w / p.n_items;
            const int item = w - it.head * p.n_items;
            const int4 r0 = __ldg(recs + 4 * item), r1 = __ldg(recs + 4 * item + 2);
            it.nk0 = r0.y > 0 ? (r0.w + kKB - 1) / kKB : 0;
            it.nk1 = r1.y > 0 ? (r1.w + kKB - 1) / kKB : 0;
            it.same = it.nk0 > 0 && it.nk1 > 0 && r0.z == r1.z;
            const int4 me = s == 0 ? r0 : r1;
            it.qrow = me.x; it.nrows = me.y; it.Sk = me.w;
            it.qpos0 = __ldg(reinterpret_cast<const int*>(recs + 4 * item + 2 * s + 1));
        };

        Item cur;
        int w = blockIdx.x;
        bool have = w < n_work;
        if (have) {
            fetch(w, cur);
            const int my_nk = s == 0 ? cur.nk0 : cur.nk1;
            if (issuer && my_nk > 0) issue_qk(ring, 0, cur.same, cur.nk0, cur.nk1, true, my_nk == 1);
        }
        while (have) {
            const int my_nk = s == 0 ? cur.nk0 : cur.nk1;
            const uint32_t base = ring;
            const uint32_t next_base = base + (cur.same ? cur.nk0 : cur.nk0 + cur.nk1);
            Item nxt;
            const int w_next = w + gridDim.x;
            const bool have_next = w_next < n_work;
            if (have_next) fetch(w_next, nxt);
            const int next_nk = have_next ? (s == 0 ? nxt.nk0 : nxt.nk1) : 0;
            const float slope = ALIBI ? p.slopes[cur.head] * 1.4426950408889634f : 0.f;
            const float qpos = (float)(cur.qpos0 + row);
            float m_ref = 0.f, l_run = 0.f;
            bool next_issued = false;                         // issuer only: next item's first Q K^T already issued
            for (int j = 0; j < my_nk; ++j) {
                const uint32_t m = n_tiles++;
                mbar_wait(&s_full[s], m & 1);
                tc_fence_after();
                // ---- scores of this row -> registers; then the score buffer is free for the next Q K^T ----
                uint32_t r[64];
                tmem_ld_32x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                tmem_ld_32x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                tmem_ld_wait();
                tc_fence_before();
                wg_sync();
                if (issuer) {
                    if (j + 1 < my_nk) {
                        issue_qk(base, j + 1, cur.same, cur.nk0, cur.nk1, false, j + 2 == my_nk);
                    } else if (next_nk > 0) {
                        // run into the next item: its first score tile — but only if its Q tile and K block have
                        // already landed.  Blocking here could deadlock: the producer may need the V block this
                        // warpgroup is about to consume before it can reach the next item's loads.
                        const uint32_t idx = next_base + ring_offset(0, s, nxt.same, nxt.nk0, nxt.nk1);
                        if (mbar_try_wait(&q_full[s], (n_items_mine + 1) & 1) &&
                            mbar_try_wait(&k_full[idx % kKStages], (idx / kKStages) & 1)) {
                            ++n_items_mine;
                            issue_qk(next_base, 0, nxt.same, nxt.nk0, nxt.nk1, true, next_nk == 1);
                            --n_items_mine;
                            next_issued = true;
                        }
                    }
                }
                softmax_tile<HD, ALIBI>(r, o_addr, j == 0, &p_empty[s], (m & 1) ^ 1, p.scale_log2, slope, qpos, j * kKB,
                                        cur.Sk, m_ref, l_run, my_p, row, lane);
                fence_proxy_async_smem();                     // generic-proxy P writes -> visible to the UMMA (async proxy)
                tc_fence_before();                            // orders a possible tcgen05.st rescale before the PV
                wg_sync();
                if (issuer) issue_pv(base, j, cur.same, cur.nk0, cur.nk1, j + 1 == my_nk);
            }
            if (my_nk > 0) {
                if (issuer && next_nk > 0 && !next_issued) {  // everything of this item is consumed: blocking is safe now
                    ++n_items_mine;
                    issue_qk(next_base, 0, nxt.same, nxt.nk0, nxt.nk1, true, next_nk == 1);
                    --n_items_mine;
                }
                // ---- epilogue: O_s / l -> bf16 -> global ----
                mbar_wait(&o_full[s], n_items_mine & 1);
                tc_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(o_addr, o0);
                if constexpr (HD > 48) tmem_ld_32x32(o_addr + 32, o1);
                else tmem_ld_32x16(o_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(&o1[0]));
                tmem_ld_wait();
                tc_fence_before();                            // the next item's first PV (issued after a wg_sync) overwrites O_s
                if (row < cur.nrows) {
                    const float inv = 1.0f / l_run;
                    __nv_bfloat16* dst = p.o + (size_t)(cur.qrow + row) * p.ldo + cur.head * HD;
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
                ++n_items_mine;
            } else if (issuer && next_nk > 0) {
                // this slot was empty in the current item: nobody issued the next item's first Q K^T yet
                issue_qk(next_base, 0, nxt.same, nxt.nk0, nxt.nk1, true, next_nk == 1);
            }
            ring = next_base;
            cur = nxt; w = w_next; have = have_next;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
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

template <int HD, bool ALIBI>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        VF_CUDA_OK(cudaFuncSetAttribute(attention_mc_kernel<HD, ALIBI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kSmem));
        VF_CUDA_OK(cudaFuncSetAttribute(attention_mc_kernel<HD, ALIBI>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
        // (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for every kernel that allocates tensor memory;
        //  tools/ubench/occ.cu shows two such CTAs with 256 columns and 108 KB each are in fact co-resident.)
        attr_set = true;
    }
    int dev = 0, sms = 0;
    VF_CUDA_OK(cudaGetDevice(&dev));
    VF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long work = (long)p.n_items * p.heads;
    const int grid = (int)(work < 2L * sms ? work : 2L * sms);
    attention_mc_kernel<HD, ALIBI><<<grid, kThreads, kSmem, s>>>(tq, tk, tv, p);
    VF_LAUNCH_OK("attention_mc_kernel launch");
    return 0;
}

}  // namespace mc

int attention_mc_varlen(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                        long rows_q, long rows_k, const int* slots, int n_items, int heads, int head_dim,
                        const float* slopes, cudaStream_t stream) {
    using namespace mc;
    VF_REQUIRE(head_dim == 48 || head_dim == 64, "attention_mc: head_dim %d not supported (48/64)", head_dim);
    VF_REQUIRE(ldo % 8 == 0, "attention_mc: output stride must keep 16-byte alignment");
    if (n_items == 0) return 0;
    const int d = heads * head_dim;
    CUtensorMap tq, tk, tv;
    if (make_tmap_rows64(&tq, q, rows_q, d, ldq, kQT)) return -1;
    if (make_tmap_rows64(&tk, k, rows_k, d, ldk, kKB)) return -1;
    if (make_tmap_rows64(&tv, v, rows_k, d, ldv, kKB)) return -1;
    Params p;
    VF_REQUIRE((reinterpret_cast<uintptr_t>(slots) & 15) == 0, "attention_mc: slot table must be 16-byte aligned");
    p.slots = reinterpret_cast<const SlotRec*>(slots); p.n_items = n_items; p.heads = heads;
    p.o = (__nv_bfloat16*)o; p.ldo = ldo; p.slopes = slopes;
    p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
    if (head_dim == 48) return slopes ? launch<48, true>(tq, tk, tv, p, stream) : launch<48, false>(tq, tk, tv, p, stream);
    return slopes ? launch<64, true>(tq, tk, tv, p, stream) : launch<64, false>(tq, tk, tv, p, stream);
}

}  // namespace vf
