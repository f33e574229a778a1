// vf_encode.cu — stage 1 on the GPU: genotype application (IUPAC diploid encoding)
// and BPE-500 tokenisation of every cis-regulatory / gene window.  Integer, byte
// and index work only: HBM/L2-bound, no tensor cores.  Bit-exact against
// oracle/vf_oracle.c (which is pinned to the reference's utils/seq.py + tokenizers).
//
// Replaces, per window:
//   `samtools faidx | bcftools consensus -H I` subprocess pairs  (utils/data_process.py:17-101, :367-467)
//   reverse_complement                                          (utils/functions.py:129-172)
//   BPEEncoder.normalize / encode (HF tokenizers BPE)            (utils/seq.py:32-62)
//   __adjust_length / chunkify_data                              (datasets/vcfdataset.py:198-217, :338-394)
#include <cooperative_groups.h>

#include "vf_common.cuh"
#include "vf_internal.h"

namespace vf {

// ---------------------------------------------------------------------------------
// encode_windows: reference slice + sample genotypes -> (optionally reverse-complemented) sequence
// ---------------------------------------------------------------------------------
constexpr int kMaxApplied = 2048;     // applied variants per window kept in shared memory
constexpr int kEncSlice = 16384;      // reference bytes per CTA

__device__ __forceinline__ uint8_t comp_base(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'a': return 't'; case 'C': return 'G'; case 'c': return 'g';
        case 'G': return 'C'; case 'g': return 'c'; case 'T': return 'A'; case 't': return 'a';
        case 'R': return 'Y'; case 'r': return 'y'; case 'Y': return 'R'; case 'y': return 'r';
        case 'K': return 'M'; case 'k': return 'm'; case 'M': return 'K'; case 'm': return 'k';
        case 'B': return 'V'; case 'b': return 'v'; case 'V': return 'B'; case 'v': return 'b';
        case 'D': return 'H'; case 'd': return 'h'; case 'H': return 'D'; case 'h': return 'd';
        default: return c;
    }
}
__device__ __forceinline__ uint8_t iupac_het(uint8_t ref, uint8_t alt) {   // vepdataset.py:75-104
    const int r = ref == 'A' ? 0 : ref == 'C' ? 1 : ref == 'G' ? 2 : ref == 'T' ? 3 : -1;
    const int a = alt == 'A' ? 0 : alt == 'C' ? 1 : alt == 'G' ? 2 : alt == 'T' ? 3 : -1;
    if (r < 0 || a < 0) return 'N';
    const char tbl[17] = "AMRWMCSYRSGKWYKT";
    return (uint8_t)tbl[r * 4 + a];
}

struct EncodeParams {
    const uint8_t* genome;            // concatenated chromosomes, 1 byte/base as in the FASTA (case kept)
    const int64_t* win_base;          // [n_win] offset of the window's chromosome in `genome`
    const int32_t* w0; const int32_t* w1;         // [n_win] window [w0, w1) inside the chromosome
    const int32_t* var_lo; const int32_t* var_hi; // [n_win] range of candidate variants (pos in [w0, w1))
    const uint8_t* flags;             // [n_win] bit0 = reverse-complement, bit1 = SNP-only filter
    const int32_t* v_pos; const int32_t* v_ref_len; const int32_t* v_alt_off; const int32_t* v_alt_len;
    const uint8_t* v_gt;              // 0 skip, 1 het, 2 hom-alt
    const uint8_t* alt_pool;
    uint8_t* out; int64_t pitch;      // [n_win, pitch]
    int32_t* out_len;                 // [n_win]
    int32_t* err;                     // device error flag (1 = pitch overflow, 2 = too many variants)
};

__global__ void __launch_bounds__(256)
encode_windows_kernel(const EncodeParams p) {
    __shared__ int32_t s_vidx[kMaxApplied];     // applied variant index
    __shared__ int32_t s_shift[kMaxApplied];    // cumulative (alt_len - ref_len) BEFORE this variant
    __shared__ int s_n, s_total;
    const int w = blockIdx.y;
    const int w0 = p.w0[w], w1 = p.w1[w];
    const int slice0 = w0 + blockIdx.x * kEncSlice;
    if (slice0 >= w1 && !(blockIdx.x == 0)) return;
    const uint8_t fl = p.flags[w];
    const bool rc = fl & 1, snp_only = fl & 2;
    if (threadIdx.x == 0) {
        // sequential scan: which records are applied (overlap rule needs order) and where they land
        int n = 0, cur = w0, shift = 0, bad = 0;
        for (int i = p.var_lo[w]; i < p.var_hi[w]; ++i) {
            if (p.v_gt[i] == 0) continue;
            const int pos = p.v_pos[i], rl = p.v_ref_len[i], al = p.v_alt_len[i];
            if (pos < cur || pos + rl > w1) continue;
            if (snp_only && !(rl == 1 && al == 1)) continue;
            if (n == kMaxApplied) { bad = 2; break; }
            s_vidx[n] = i; s_shift[n] = shift; ++n;
            shift += al - rl;
            cur = pos + rl;
        }
        s_n = n; s_total = (w1 - w0) + shift;
        if (bad) atomicMax(p.err, bad);
        if (s_total > p.pitch) atomicMax(p.err, 1);
        // an overflowing window writes nothing: report length 0 (plus err = 1) so the tokenizer never reads past the row
        if (blockIdx.x == 0) p.out_len[w] = s_total > p.pitch ? 0 : s_total;
    }
    __syncthreads();
    const int n = s_n, total = s_total;
    if (total > p.pitch) return;
    const uint8_t* ref = p.genome + p.win_base[w];
    uint8_t* out = p.out + (size_t)w * p.pitch;
    const int slice1 = min(w1, slice0 + kEncSlice);
    // (1) reference bytes of this slice that survive (not inside an applied record's REF span)
    for (int x = slice0 + threadIdx.x; x < slice1; x += blockDim.x) {
        // last applied variant with pos <= x
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.v_pos[s_vidx[mid]] <= x) lo = mid + 1; else hi = mid; }
        int shift = 0;
        if (lo > 0) {
            const int v = s_vidx[lo - 1];
            if (x < p.v_pos[v] + p.v_ref_len[v]) continue;              // replaced by the record's ALT
            shift = s_shift[lo - 1] + p.v_alt_len[v] - p.v_ref_len[v];
        }
        const int o = x - w0 + shift;
        const uint8_t b = ref[x];
        if (rc) out[total - 1 - o] = comp_base(b); else out[o] = b;
    }
    // (2) ALT bytes of the applied records that start inside this slice
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int v = s_vidx[k];
        const int pos = p.v_pos[v];
        if (pos < slice0 || pos >= slice1) continue;
        const int al = p.v_alt_len[v], rl = p.v_ref_len[v];
        const int o0 = pos - w0 + s_shift[k];
        if (al == 1 && rl == 1) {
            uint8_t a = p.alt_pool[p.v_alt_off[v]];
            if (p.v_gt[v] == 1) {
                uint8_t r = ref[pos];
                if (r >= 'a' && r <= 'z') r -= 32;
                a = iupac_het(r, a);
            }
            if (rc) out[total - 1 - o0] = comp_base(a); else out[o0] = a;
        } else {
            for (int j = 0; j < al; ++j) {
                const uint8_t a = p.alt_pool[p.v_alt_off[v] + j];
                if (rc) out[total - 1 - (o0 + j)] = comp_base(a); else out[o0 + j] = a;
            }
        }
    }
}

int encode_windows(const uint8_t* genome, const int64_t* win_base, const int32_t* w0, const int32_t* w1,
                   const int32_t* var_lo, const int32_t* var_hi, const uint8_t* flags, const int32_t* v_pos,
                   const int32_t* v_ref_len, const int32_t* v_alt_off, const int32_t* v_alt_len, const uint8_t* v_gt,
                   const uint8_t* alt_pool, int n_win, int max_window, uint8_t* out, int64_t pitch, int32_t* out_len,
                   int32_t* err, cudaStream_t s) {
    if (n_win == 0) return 0;
    VF_REQUIRE(max_window > 0 && pitch >= max_window, "encode_windows: pitch %lld < max window %d", (long long)pitch,
               max_window);
    EncodeParams p{genome, win_base, w0, w1, var_lo, var_hi, flags, v_pos, v_ref_len, v_alt_off, v_alt_len, v_gt,
                   alt_pool, out, pitch, out_len, err};
    dim3 grid((max_window + kEncSlice - 1) / kEncSlice, n_win);
    encode_windows_kernel<<<grid, 256, 0, s>>>(p);
    VF_LAUNCH_OK("encode_windows_kernel launch");
    return 0;
}

// ---------------------------------------------------------------------------------
// BPE-500 tokenisation by ascending-rank sweeps in the base-position domain.
//
// HF tokenizers merges the globally lowest-ranked adjacent pair first (leftmost first among
// equals).  Because the token created by rank r can only be consumed by a merge of rank > r
// (verified on the 482-entry table at load), processing ranks 0..R-1 in order and, within a
// rank, all matches left to right (non-overlapping for self pairs) yields the identical
// segmentation.  State: one uint16 per base position — token id at the
// token's first base, DEAD on its remaining bases, SEP on non-IUPAC characters (word
// boundaries, utils/seq.py:36-38).  Tokens are <= 8 bases long, so "next alive symbol" is a
// scan over <= 7 DEAD slots and no compaction is needed between sweeps.
// One CTA per window; symbols in shared memory when the window fits, else in a global
// scratch row (gene windows: ~301 k symbols).  Ranks whose operands are absent from the
// window (per-CTA occupancy counters) are skipped without touching the symbols.
//
// Rank BATCHES.  Consecutive ranks whose symbol sets {left, right, new} are pairwise disjoint
// (and none of which is a self pair) commute: a merge only creates adjacencies that involve its
// own new token, the symbol it kills always sits right behind a symbol of its own left type, so
// no match of another rank of the batch appears, disappears or moves.  Such a run is applied in
// ONE sweep with ONE barrier: every position looks its symbol up in a small table (symbol ->
// rank of the batch whose left operand it is).  The host supplies the batch id of every rank
// (merge_batch, nondecreasing; stage1.merge_batches builds it: 482 ranks -> 162 batches of up
// to 12); null = every rank alone.  Self-pair ranks are always alone (detect / apply phases).
// ---------------------------------------------------------------------------------
constexpr int kMaxBatch = 16;
constexpr uint16_t kDead = 0xFFFF, kSep = 0xFFFE;
constexpr int kBpeSmemSyms = 8192;    // windows up to this many symbols stay in shared memory
constexpr int kMaxVocab = 512;

struct BpeParams {
    const uint8_t* seq; int64_t pitch; const int32_t* len;   // [n_win, pitch] bytes
    const uint16_t* merge_a; const uint16_t* merge_b; const uint16_t* merge_new; int n_merges;
    const uint16_t* merge_batch;                              // [n_merges] nondecreasing batch id per rank, or null (see below)
    uint16_t* scratch; int64_t scratch_pitch;                 // [n_win, scratch_pitch] for long windows (may be null)
    int32_t* out_tokens; int out_pitch; int out_cap;          // [n_win, out_pitch]; first out_cap tokens kept, rest of the row zeroed
    int32_t* out_count;                                       // [n_win] total token count (before truncation)
    int32_t* out_start; int64_t start_pitch;                  // optional [n_win, start_pitch]: first base of every token
    int max_len;                                              // lengths are clamped to this (and to pitch): a corrupt or
                                                              // overflowed length can never index past a row
};

__device__ __forceinline__ uint16_t base_symbol(uint8_t c) {
    // upper-case + 14-letter alphabet A,B,C,D,G,H,K,M,R,S,T,V,W,Y -> ids 4..17; everything else separates words
    if (c >= 'a' && c <= 'z') c -= 32;
    switch (c) {
        case 'A': return 4; case 'B': return 5; case 'C': return 6; case 'D': return 7; case 'G': return 8;
        case 'H': return 9; case 'K': return 10; case 'M': return 11; case 'R': return 12; case 'S': return 13;
        case 'T': return 14; case 'V': return 15; case 'W': return 16; case 'Y': return 17;
        default: return kSep;
    }
}

__global__ void bpe_tokenize_kernel(const BpeParams p) {
    extern __shared__ uint16_t s_sym[];                 // kBpeSmemSyms (only used when the window fits)
    __shared__ int s_cnt[kMaxVocab];                    // alive symbols per token id
    __shared__ int s_warp_tot[32];
    __shared__ int s_any;
    const int w = blockIdx.x;
    const int n = max(0, min(p.len[w], p.max_len));
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool in_smem = n <= kBpeSmemSyms;
    uint16_t* sym = in_smem ? s_sym : p.scratch + (size_t)w * p.scratch_pitch;
    const uint8_t* src = p.seq + (size_t)w * p.pitch;

    for (int i = tid; i < kMaxVocab; i += nt) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const uint16_t b = base_symbol(src[i]);
        sym[i] = b;
        if (b != kSep) atomicAdd(&s_cnt[b], 1);
    }
    __syncthreads();

    __shared__ short s_rk[kMaxVocab];                   // symbol -> index (within the batch) of the rank it is the left operand of
    __shared__ int s_mcnt[kMaxBatch];
    for (int i = tid; i < kMaxVocab; i += nt) s_rk[i] = -1;
    __syncthreads();
    for (int r0 = 0; r0 < p.n_merges;) {
        int r1 = r0 + 1;
        if (p.merge_batch) {
            const uint16_t bid = p.merge_batch[r0];
            while (r1 < p.n_merges && r1 - r0 < kMaxBatch && p.merge_batch[r1] == bid) ++r1;
        }
        const uint16_t a = p.merge_a[r0], b = p.merge_b[r0], c = p.merge_new[r0];
        const bool self_pair = a == b;                    // (always a batch of its own)
        // uniform skip: all threads read the same counters (stable since the last barrier)
        bool any = false;
        for (int r = r0; r < r1; ++r) {
            const uint16_t ra = p.merge_a[r], rb = p.merge_b[r];
            any |= s_cnt[ra] != 0 && s_cnt[rb] != 0 && !(ra == rb && s_cnt[ra] < 2);
        }
        if (!any) { r0 = r1; continue; }
        if (!self_pair) {
            // every thread must have taken the skip decision above before any occupancy counter moves
            __syncthreads();
            if (tid < r1 - r0) {
                const uint16_t ra = p.merge_a[r0 + tid], rb = p.merge_b[r0 + tid];
                if (s_cnt[ra] != 0 && s_cnt[rb] != 0) s_rk[ra] = (short)tid;
                s_mcnt[tid] = 0;
            }
            __syncthreads();
            // matches of the batch's ranks can never share a symbol: apply immediately
            for (int i = tid; i < n; i += nt) {
                const uint16_t sy = sym[i];
                if (sy >= kMaxVocab) continue;            // DEAD / SEP
                const int k = s_rk[sy];
                if (k < 0) continue;
                int j = i + 1;
                while (j < n && sym[j] == kDead) ++j;
                if (j < n && sym[j] == p.merge_b[r0 + k]) {
                    sym[i] = p.merge_new[r0 + k]; sym[j] = kDead;
                    atomicAdd(&s_mcnt[k], 1);
                }
            }
            __syncthreads();
            if (tid < r1 - r0) {                          // symbol sets are disjoint: no two threads touch one counter
                const uint16_t ra = p.merge_a[r0 + tid], rb = p.merge_b[r0 + tid], rc = p.merge_new[r0 + tid];
                const int m = s_mcnt[tid];
                if (m) { s_cnt[rc] += m; s_cnt[ra] -= m; s_cnt[rb] -= m; }
                s_rk[ra] = -1;
            }
        } else {
            int merged = 0;
            // self pair: within a run of consecutive a's merge (1st,2nd), (3rd,4th), ... -> decide from the
            // parity of the number of a's that precede i in its run; detect first, apply after a barrier
            // (the tentative mark c|0x8000 is treated as `a` by concurrent scans, so marking is race-free)
            const uint16_t mark = (uint16_t)(c | 0x8000);
            if (tid == 0) s_any = 0;
            __syncthreads();
            for (int i = tid; i < n; i += nt) {
                if (sym[i] != a) continue;
                int k = 0, q = i - 1;
                for (;;) {
                    while (q >= 0 && sym[q] == kDead) --q;
                    if (q < 0) break;
                    const uint16_t sq = sym[q];
                    if (sq != a && sq != mark) break;
                    ++k; --q;
                }
                if (k & 1) continue;                     // i is the right half of the previous pair
                int j = i + 1;
                while (j < n && sym[j] == kDead) ++j;
                if (j < n && (sym[j] == a || sym[j] == mark)) { sym[i] = mark; s_any = 1; }
            }
            __syncthreads();
            if (s_any) {
                for (int i = tid; i < n; i += nt) {
                    if (sym[i] != mark) continue;
                    int j = i + 1;
                    while (j < n && sym[j] == kDead) ++j;
                    sym[j] = kDead; sym[i] = c; ++merged;
                }
            }
            if (merged) { atomicAdd(&s_cnt[c], merged); atomicSub(&s_cnt[a], 2 * merged); }
        }
        __syncthreads();
        r0 = r1;
    }

    // ---- ordered compaction: each thread owns a contiguous segment ----
    const int per = (n + nt - 1) / nt;
    const int b0 = min(n, tid * per), b1 = min(n, b0 + per);
    int mine = 0;
    for (int i = b0; i < b1; ++i) mine += (sym[i] < kSep);
    int incl = mine;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = lane < (nt >> 5) ? s_warp_tot[lane] : 0;
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
        s_warp_tot[lane] = inc2 - v;                       // exclusive warp offsets
        if (lane == 31) s_any = inc2;                      // grand total
    }
    __syncthreads();
    int o = s_warp_tot[wid] + incl - mine;
    const int total = s_any;
    int32_t* out = p.out_tokens + (size_t)w * p.out_pitch;
    int32_t* ost = p.out_start ? p.out_start + (size_t)w * p.start_pitch : nullptr;
    for (int i = b0; i < b1; ++i) {
        const uint16_t sy = sym[i];
        if (sy < kSep) {
            if (o < p.out_cap) out[o] = sy;
            if (ost) ost[o] = i;
            ++o;
        }
    }
    for (int i = min(total, p.out_cap) + tid; i < p.out_pitch; i += nt) out[i] = 0;    // <pad> = 0
    if (tid == 0) p.out_count[w] = total;
}

// ---------------------------------------------------------------------------------
// Long windows (gene windows, ~301 k symbols): one thread-block CLUSTER per window.  Each CTA keeps a contiguous
// segment of the symbol array in its own shared memory; neighbours are reached through distributed shared memory
// (forward "next alive symbol" scans cross a segment boundary by at most 7 slots, self-pair run scans by the run
// length).  One cluster barrier per applied rank (two for self pairs).  The per-token-id occupancy counters that
// drive the uniform skip decision are replicated in every CTA and kept identical by gathering the per-CTA merge
// counts of every applied sweep through DSMEM right after that sweep's cluster barrier.
// ---------------------------------------------------------------------------------
constexpr int kBpeCluster = 8;

__global__ void __launch_bounds__(1024, 1)
bpe_tokenize_cluster_kernel(const BpeParams p, const int seg_cap) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint16_t s_sym[];                  // seg_cap symbols of this CTA's segment
    __shared__ int s_cnt[2][kMaxVocab];                  // replicated occupancy counters (double buffered)
    __shared__ uint16_t* s_peer[kBpeCluster];
    __shared__ int* s_peer_cnt[kBpeCluster];
    __shared__ int s_segcnt[kBpeCluster];
    __shared__ int s_warp_tot[32];
    const int rank = (int)cluster.block_rank();
    const int w = blockIdx.x / kBpeCluster;
    const int n = max(0, min(p.len[w], p.max_len));
    const int tid = threadIdx.x, nt = blockDim.x;
    const int seg = max(8, (n + kBpeCluster - 1) / kBpeCluster);      // <= seg_cap (host guarantees)
    const int lo = min(n, rank * seg), hi = min(n, lo + seg);
    const uint8_t* src = p.seq + (size_t)w * p.pitch;

    if (tid < kBpeCluster) {
        s_peer[tid] = cluster.map_shared_rank(s_sym, tid);
        s_peer_cnt[tid] = cluster.map_shared_rank(&s_cnt[0][0], tid);
    }
    for (int i = tid; i < 2 * kMaxVocab; i += nt) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    // local histogram in replica 0 slots of a scratch copy: count into registers-free smem then publish
    __shared__ int s_hist[32];                            // base ids 4..17 only
    if (tid < 32) s_hist[tid] = 0;
    cluster.sync();                                       // every CTA's counters are zeroed before anyone publishes
    for (int i = lo + tid; i < hi; i += nt) {
        const uint16_t b = base_symbol(src[i]);
        s_sym[i - lo] = b;
        if (b != kSep) atomicAdd(&s_hist[b], 1);
    }
    __syncthreads();
    if (tid < 32 && s_hist[tid] != 0) {
        for (int c = 0; c < kBpeCluster; ++c) {
            atomicAdd(s_peer_cnt[c] + tid, s_hist[tid]);
        }
    }
    cluster.sync();

    auto LD = [&](int pos) -> uint16_t {
        if (pos >= lo && pos < hi) return s_sym[pos - lo];
        const int c = pos / seg;
        return s_peer[c][pos - c * seg];
    };
    auto ST = [&](int pos, uint16_t v) {
        if (pos >= lo && pos < hi) { s_sym[pos - lo] = v; return; }
        const int c = pos / seg;
        s_peer[c][pos - c * seg] = v;
    };
    // Per applied batch every CTA posts the merge counts of the batch's ranks in its own shared memory (slot k&1); after
    // the cluster barrier each CTA gathers the 8 count vectors through DSMEM and updates its private replica of the
    // counters, so all replicas stay identical without remote atomics.
    __shared__ int s_pub[2][kMaxBatch];
    __shared__ int s_mcnt[2][kMaxBatch];
    __shared__ short s_rk[kMaxVocab];                     // symbol -> index (within the batch) of the rank it is the left operand of
    for (int i = tid; i < kMaxVocab; i += nt) s_rk[i] = -1;
    if (tid < 2 * kMaxBatch) (&s_mcnt[0][0])[tid] = 0;
    __syncthreads();
    int k = 0;                                            // applied batches so far (cluster-uniform)
    for (int r0 = 0; r0 < p.n_merges;) {
        int r1 = r0 + 1;
        if (p.merge_batch) {
            const uint16_t bid = p.merge_batch[r0];
            while (r1 < p.n_merges && r1 - r0 < kMaxBatch && p.merge_batch[r1] == bid) ++r1;
        }
        const int nb = r1 - r0;
        const uint16_t a = p.merge_a[r0], b = p.merge_b[r0], c = p.merge_new[r0];
        const bool self_pair = a == b;                    // (always a batch of its own)
        const int* cnt = s_cnt[0];
        bool any = false;                                 // uniform over the whole cluster (identical replicas)
        for (int r = r0; r < r1; ++r) {
            const uint16_t ra = p.merge_a[r], rb = p.merge_b[r];
            any |= cnt[ra] != 0 && cnt[rb] != 0 && !(ra == rb && cnt[ra] < 2);
        }
        if (!any) { r0 = r1; continue; }
        int* mc = s_mcnt[k & 1];
        if (!self_pair) {
            if (tid < nb) {
                const uint16_t ra = p.merge_a[r0 + tid], rb = p.merge_b[r0 + tid];
                if (cnt[ra] != 0 && cnt[rb] != 0) s_rk[ra] = (short)tid;
            }
            __syncthreads();
            for (int i = lo + tid; i < hi; i += nt) {
                const uint16_t sy = s_sym[i - lo];
                if (sy >= kMaxVocab) continue;            // DEAD / SEP
                const int kk = s_rk[sy];
                if (kk < 0) continue;
                int j = i + 1;
                while (j < n && LD(j) == kDead) ++j;
                if (j < n && LD(j) == p.merge_b[r0 + kk]) {
                    s_sym[i - lo] = p.merge_new[r0 + kk]; ST(j, kDead);
                    atomicAdd(&mc[kk], 1);
                }
            }
        } else {
            int merged = 0;
            const uint16_t mark = (uint16_t)(c | 0x8000);
            for (int i = lo + tid; i < hi; i += nt) {
                if (s_sym[i - lo] != a) continue;
                int kk = 0, q = i - 1;
                for (;;) {
                    while (q >= 0 && LD(q) == kDead) --q;
                    if (q < 0) break;
                    const uint16_t sq = LD(q);
                    if (sq != a && sq != mark) break;
                    ++kk; --q;
                }
                if (kk & 1) continue;
                int j = i + 1;
                while (j < n && LD(j) == kDead) ++j;
                if (j < n) { const uint16_t sj = LD(j); if (sj == a || sj == mark) s_sym[i - lo] = mark; }
            }
            cluster.sync();
            for (int i = lo + tid; i < hi; i += nt) {
                if (s_sym[i - lo] != mark) continue;
                int j = i + 1;
                while (j < n && LD(j) == kDead) ++j;
                ST(j, kDead); s_sym[i - lo] = c; ++merged;
            }
            if (merged) atomicAdd(&mc[0], merged);
        }
        __syncthreads();
        if (tid < kMaxBatch) {
            s_pub[k & 1][tid] = tid < nb ? mc[tid] : 0;
            s_mcnt[(k + 1) & 1][tid] = 0;
            if (tid < nb && !self_pair) s_rk[p.merge_a[r0 + tid]] = -1;
        }
        cluster.sync();                                   // all merges of this batch + every CTA's counts are visible
        if (tid < 32 * kMaxBatch / 4) {                   // 4 ranks per warp: lanes (8 CTAs x 4 ranks)
            const int q = tid >> 5, lane = tid & 31;      // warp q handles ranks 4q .. 4q+3
            const int rr = 4 * q + (lane >> 3), cta = lane & 7;
            int v = rr < nb ? cluster.map_shared_rank(&s_pub[0][0], cta)[(k & 1) * kMaxBatch + rr] : 0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (cta == 0 && rr < nb && v != 0) {          // symbol sets of a batch are disjoint: one writer per counter
                const uint16_t ra = p.merge_a[r0 + rr], rb = p.merge_b[r0 + rr], rc = p.merge_new[r0 + rr];
                s_cnt[0][rc] += v;
                if (ra == rb) s_cnt[0][ra] -= 2 * v; else { s_cnt[0][ra] -= v; s_cnt[0][rb] -= v; }
            }
        }
        __syncthreads();
        ++k;
        r0 = r1;
    }
    cluster.sync();                                       // no CTA may exit (or reuse s_pub) while peers still read it

    // ---- ordered compaction of this CTA's segment; token offset = alive symbols of the lower-ranked segments ----
    const int len_seg = hi - lo;
    const int per = (len_seg + nt - 1) / nt;
    const int b0 = min(len_seg, tid * per), b1 = min(len_seg, b0 + per);
    int mine = 0;
    for (int i = b0; i < b1; ++i) mine += (s_sym[i] < kSep);
    int incl = mine;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int v = lane < (nt >> 5) ? s_warp_tot[lane] : 0;
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, o); if (lane >= o) inc2 += u; }
        s_warp_tot[lane] = inc2 - v;
        if (lane == 31) {
            for (int c2 = 0; c2 < kBpeCluster; ++c2) cluster.map_shared_rank(&s_segcnt[0], c2)[rank] = inc2;
        }
    }
    cluster.sync();
    int base = 0, total = 0;
    for (int c2 = 0; c2 < kBpeCluster; ++c2) { if (c2 < rank) base += s_segcnt[c2]; total += s_segcnt[c2]; }
    int o = base + s_warp_tot[wid] + incl - mine;
    int32_t* out = p.out_tokens + (size_t)w * p.out_pitch;
    int32_t* ost = p.out_start ? p.out_start + (size_t)w * p.start_pitch : nullptr;
    for (int i = b0; i < b1; ++i) {
        const uint16_t sy = s_sym[i];
        if (sy < kSep) {
            if (o < p.out_cap) out[o] = sy;
            if (ost) ost[o] = lo + i;
            ++o;
        }
    }
    for (int i = min(total, p.out_cap) + rank * nt + tid; i < p.out_pitch; i += nt * kBpeCluster) out[i] = 0;
    if (rank == 0 && tid == 0) p.out_count[w] = total;
    cluster.sync();                                       // keep every CTA's shared memory alive until all peers are done
}

int bpe_tokenize(const uint8_t* seq, int64_t pitch, const int32_t* len, int n_win, int max_len,
                 const uint16_t* merge_a, const uint16_t* merge_b, const uint16_t* merge_new,
                 const uint16_t* merge_batch, int n_merges, uint16_t* scratch, int64_t scratch_pitch, int32_t* out_tokens, int out_pitch, int out_cap,
                 int32_t* out_count, int32_t* out_start, int64_t start_pitch, int block_threads, cudaStream_t s) {
    if (n_win == 0) return 0;
    VF_REQUIRE(out_cap <= out_pitch, "bpe_tokenize: out_cap %d > out_pitch %d", out_cap, out_pitch);
    VF_REQUIRE(out_start == nullptr || start_pitch >= max_len, "bpe_tokenize: start_pitch too small");
    BpeParams p{seq, pitch, len, merge_a, merge_b, merge_new, n_merges, merge_batch, scratch, scratch_pitch,
                out_tokens, out_pitch, out_cap, out_count, out_start, start_pitch,
                (int)(pitch < (int64_t)max_len ? pitch : (int64_t)max_len)};
    if (max_len > kBpeSmemSyms) {
        // cluster path: 8 CTAs per window, segment of the symbol array per CTA in shared memory
        const int seg_cap = (max_len + kBpeCluster - 1) / kBpeCluster + 8;
        const size_t smem = (size_t)seg_cap * sizeof(uint16_t);
        VF_REQUIRE(smem <= 200 * 1024, "bpe_tokenize: window of %d symbols exceeds the cluster kernel's capacity", max_len);
        static size_t attr_smem = 0;
        if (smem > attr_smem) {
            VF_CUDA_OK(cudaFuncSetAttribute(bpe_tokenize_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
            attr_smem = smem;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_win * kBpeCluster); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kBpeCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        VF_CUDA_OK(cudaLaunchKernelEx(&cfg, bpe_tokenize_cluster_kernel, p, seg_cap));
        return 0;
    }
    // block size is a pure performance hint (typical window length); any value works for any window
    const int threads = (block_threads == 128 || block_threads == 256 || block_threads == 512 || block_threads == 1024)
                            ? block_threads : (max_len <= 1024 ? 128 : 1024);
    const size_t smem = (size_t)kBpeSmemSyms * sizeof(uint16_t);
    bpe_tokenize_kernel<<<n_win, threads, smem, s>>>(p);
    VF_LAUNCH_OK("bpe_tokenize_kernel launch");
    return 0;
}

}  // namespace vf
