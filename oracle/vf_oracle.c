/*
 * vf_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of stage 1 of VariantFormer's inference hot path
 * (genotype -> IUPAC sequence -> BPE-500 tokens -> fixed 200-token windows).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product path
 * (variantformer_b200/) never does and fails loudly without its CUDA library.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference checkout).  The BPE algorithm itself lives in a
 * third-party dependency that is not vendored in the reference:
 *   HuggingFace `tokenizers` (pyproject.toml:41 pins ~=0.21.1; 0.22.2 is what
 *   the build container has).  Its published word-merge algorithm
 *   (models/bpe/word.rs `Word::merge_all`) is restated in vfo_bpe_word();
 *   parity is pinned by tests/golden/stage1_golden.npz, produced by
 *   tests/golden/make_stage1_golden.py from the reference's own
 *   utils/seq.py::BPEEncoder running on that wheel.
 * Genotype application: `bcftools consensus -H I` (htslib/bcftools 1.21,
 *   Dockerfile:24-52) is NOT available offline -> the SNP rule is pinned by the
 *   reference's in-repo datasets/vepdataset.py:75-131; indel handling is
 *   "parity unpinned" (policy documented at vfo_apply_variants()).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VFO_MAX_VOCAB 512

typedef struct {
    int n_vocab;                 /* 500 */
    int n_merges;                /* 482 */
    int16_t pair_rank[VFO_MAX_VOCAB * VFO_MAX_VOCAB]; /* (a,b) -> rank or -1 */
    int16_t new_id[VFO_MAX_VOCAB];                    /* rank -> merged token id */
    int8_t  base_id[256];        /* upper-cased byte -> token id (4..17) or -1 */
} vfo_bpe_t;

/* utils/constants.py:2-17 — the 14-letter IUPAC alphabet (N is NOT a member);
 * vocab ids 4..17 follow the order A,B,C,D,G,H,K,M,R,S,T,V,W,Y
 * (vocabs/bpe_vocabulary_500.json model.vocab). */
static const char VFO_ALPHABET[] = "ABCDGHKMRSTVWY";

vfo_bpe_t *vfo_bpe_create(const int16_t *merge_left, const int16_t *merge_right,
                          const int16_t *merge_new, int n_merges, int n_vocab)
{
    if (n_vocab > VFO_MAX_VOCAB || n_merges > VFO_MAX_VOCAB) return NULL;
    vfo_bpe_t *t = (vfo_bpe_t *)malloc(sizeof(vfo_bpe_t));
    if (!t) return NULL;
    t->n_vocab = n_vocab;
    t->n_merges = n_merges;
    for (int i = 0; i < VFO_MAX_VOCAB * VFO_MAX_VOCAB; ++i) t->pair_rank[i] = -1;
    for (int r = 0; r < n_merges; ++r) {
        t->pair_rank[merge_left[r] * VFO_MAX_VOCAB + merge_right[r]] = (int16_t)r;
        t->new_id[r] = merge_new[r];
    }
    for (int i = 0; i < 256; ++i) t->base_id[i] = -1;
    for (int i = 0; VFO_ALPHABET[i]; ++i) {
        t->base_id[(unsigned char)VFO_ALPHABET[i]] = (int8_t)(4 + i);
        t->base_id[(unsigned char)(VFO_ALPHABET[i] + 32)] = (int8_t)(4 + i); /* .upper(), utils/seq.py:35 */
    }
    return t;
}

void vfo_bpe_destroy(vfo_bpe_t *t) { free(t); }

/* ---- min-heap of candidate merges ordered by (rank, pos) ------------------ */
typedef struct { int32_t rank; int32_t pos; int32_t new_id; } vfo_cand_t;
typedef struct { vfo_cand_t *a; int n, cap; } vfo_heap_t;

static int cand_less(const vfo_cand_t *x, const vfo_cand_t *y)
{
    if (x->rank != y->rank) return x->rank < y->rank;
    return x->pos < y->pos;
}
static int heap_push(vfo_heap_t *h, vfo_cand_t c)
{
    if (h->n == h->cap) {
        int nc = h->cap ? h->cap * 2 : 1024;
        vfo_cand_t *na = (vfo_cand_t *)realloc(h->a, (size_t)nc * sizeof(vfo_cand_t));
        if (!na) return -1;
        h->a = na; h->cap = nc;
    }
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (!cand_less(&c, &h->a[p])) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i] = c;
    return 0;
}
static vfo_cand_t heap_pop(vfo_heap_t *h)
{
    vfo_cand_t top = h->a[0];
    vfo_cand_t last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        const vfo_cand_t *best = &last;
        if (l < h->n && cand_less(&h->a[l], best)) { m = l; best = &h->a[l]; }
        if (r < h->n && cand_less(&h->a[r], best)) { m = r; best = &h->a[r]; }
        if (m == i) break;
        h->a[i] = h->a[m]; i = m;
    }
    if (h->n > 0) h->a[i] = last;
    return top;
}

/*
 * One "word" (maximal run of IUPAC letters) -> token ids.
 * Restates tokenizers' Word::merge_all with dropout=None: seed a priority queue
 * with every adjacent pair that has a merge rule; pop lowest (rank, position);
 * skip stale entries (left symbol consumed, no right neighbour, or the pair now
 * at that position no longer maps to the recorded new id); merge right into
 * left; enqueue the pairs formed with the previous and next symbols.
 * ids[] holds the base ids on entry (length n); returns the token count and
 * compacts ids[] in place.  start[] (optional) receives each token's first
 * character offset inside the word (== Encoding.offsets[i][0], used by
 * utils/seq.py:149-154).
 */
static int vfo_bpe_word(const vfo_bpe_t *t, int32_t *ids, int n, int32_t *start,
                        int32_t *prev, int32_t *next, int32_t *len, vfo_heap_t *h)
{
    if (n <= 0) return 0;
    h->n = 0;
    for (int i = 0; i < n; ++i) { prev[i] = i - 1; next[i] = (i + 1 < n) ? i + 1 : -1; len[i] = 1; }
    for (int i = 0; i + 1 < n; ++i) {
        int r = t->pair_rank[ids[i] * VFO_MAX_VOCAB + ids[i + 1]];
        if (r >= 0) { vfo_cand_t c = { r, i, t->new_id[r] }; if (heap_push(h, c)) return -1; }
    }
    while (h->n > 0) {
        vfo_cand_t top = heap_pop(h);
        int p = top.pos;
        if (len[p] == 0) continue;
        if (next[p] < 0) continue;
        int q = next[p];
        int r = t->pair_rank[ids[p] * VFO_MAX_VOCAB + ids[q]];
        if (r < 0 || t->new_id[r] != top.new_id) continue;
        ids[p] = top.new_id;
        len[p] += len[q];
        len[q] = 0;
        next[p] = next[q];
        if (next[q] >= 0) prev[next[q]] = p;
        if (prev[p] >= 0) {
            int pp = prev[p];
            int rr = t->pair_rank[ids[pp] * VFO_MAX_VOCAB + ids[p]];
            if (rr >= 0) { vfo_cand_t c = { rr, pp, t->new_id[rr] }; if (heap_push(h, c)) return -1; }
        }
        if (next[p] >= 0) {
            int nn = next[p];
            int rr = t->pair_rank[ids[p] * VFO_MAX_VOCAB + ids[nn]];
            if (rr >= 0) { vfo_cand_t c = { rr, p, t->new_id[rr] }; if (heap_push(h, c)) return -1; }
        }
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if (len[i]) { if (start) start[m] = i; ids[m++] = ids[i]; }
    return m;
}

/*
 * BPEEncoder.encode forward strand (utils/seq.py:32-62): upper-case, every
 * character outside the 14-letter alphabet becomes a separator, each maximal
 * run is tokenised independently and the ids are concatenated.
 * out_ids must hold n entries.  tok_char_start (optional, n entries) receives,
 * for every token, the index in `seq` of its first character.
 * Returns the number of tokens, or -1 on allocation failure.
 */
int vfo_bpe_encode(const vfo_bpe_t *t, const uint8_t *seq, int n, int32_t *out_ids,
                   int32_t *tok_char_start)
{
    if (n <= 0) return 0;
    int32_t *buf = (int32_t *)malloc((size_t)n * 5 * sizeof(int32_t));
    if (!buf) return -1;
    int32_t *ids = buf, *prev = buf + n, *next = buf + 2 * n, *len = buf + 3 * n, *st = buf + 4 * n;
    vfo_heap_t h = { NULL, 0, 0 };
    int total = 0, i = 0;
    while (i < n) {
        while (i < n && t->base_id[seq[i]] < 0) ++i;
        int w0 = i;
        while (i < n && t->base_id[seq[i]] >= 0) { ids[i - w0] = t->base_id[seq[i]]; ++i; }
        int wl = i - w0;
        if (wl == 0) break;
        int m = vfo_bpe_word(t, ids, wl, st, prev, next, len, &h);
        if (m < 0) { free(buf); free(h.a); return -1; }
        for (int k = 0; k < m; ++k) {
            out_ids[total] = ids[k];
            if (tok_char_start) tok_char_start[total] = w0 + st[k];
            ++total;
        }
    }
    free(buf); free(h.a);
    return total;
}

/*
 * BPEEncoder.encode_with_position (utils/seq.py:68-174): index (over the
 * concatenated token list) of the token covering character `position`.
 * Returns -1 when the position is out of range (ValueError :84-87) and -2 when
 * it points at a non-IUPAC character (ValueError :93-97).
 */
int vfo_bpe_token_at(const vfo_bpe_t *t, const uint8_t *seq, int n, int position)
{
    if (position < 0 || position >= n) return -1;
    if (t->base_id[seq[position]] < 0) return -2;
    int32_t *ids = (int32_t *)malloc((size_t)n * 2 * sizeof(int32_t));
    if (!ids) return -3;
    int m = vfo_bpe_encode(t, seq, n, ids, ids + n);
    int ans = -3;
    for (int k = 0; k < m; ++k) if (ids[n + k] <= position) ans = k; else break;
    free(ids);
    return ans;
}

/* utils/functions.py:129-172 (dup. datasets/vepdataset.py:40-73,94-98):
 * reverse, then complement over IUPAC incl. lower case; unknown bytes pass. */
static uint8_t vfo_comp(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'a': return 't'; case 'C': return 'G'; case 'c': return 'g';
    case 'G': return 'C'; case 'g': return 'c'; case 'T': return 'A'; case 't': return 'a';
    case 'R': return 'Y'; case 'r': return 'y'; case 'Y': return 'R'; case 'y': return 'r';
    case 'K': return 'M'; case 'k': return 'm'; case 'M': return 'K'; case 'm': return 'k';
    case 'B': return 'V'; case 'b': return 'v'; case 'V': return 'B'; case 'v': return 'b';
    case 'D': return 'H'; case 'd': return 'h'; case 'H': return 'D'; case 'h': return 'd';
    default: return c; /* S, W, N, '-', '.' and anything else map to themselves */
    }
}
void vfo_reverse_complement(const uint8_t *in, int n, uint8_t *out)
{
    for (int i = 0; i < n; ++i) out[i] = vfo_comp(in[n - 1 - i]);
}

/* datasets/vepdataset.py:75-104 — het genotype code; anything outside the
 * 16-pair ACGT table is 'N'.  Case-sensitive exactly like the dict lookup. */
uint8_t vfo_iupac_het(uint8_t ref, uint8_t alt)
{
    static const char tbl[4][4] = { /* rows/cols A C G T */
        { 'A', 'M', 'R', 'W' }, { 'M', 'C', 'S', 'Y' }, { 'R', 'S', 'G', 'K' }, { 'W', 'Y', 'K', 'T' } };
    int r = ref == 'A' ? 0 : ref == 'C' ? 1 : ref == 'G' ? 2 : ref == 'T' ? 3 : -1;
    int a = alt == 'A' ? 0 : alt == 'C' ? 1 : alt == 'G' ? 2 : alt == 'T' ? 3 : -1;
    if (r < 0 || a < 0) return 'N';
    return (uint8_t)tbl[r][a];
}

/*
 * Genotype application for one window [w_start, w_end) (0-based, half open) of
 * a chromosome; restates what `samtools faidx chrom:start+1-end | bcftools
 * consensus -H I -e 'ALT~"<.*>"'` yields for utils/data_process.py:17-101.
 *
 * Variants: sorted by pos (0-based), one ALT allele each:
 *   pos[i], ref_len[i], alt_off[i]/alt_len[i] into alt_pool, gt[i]
 *   gt: 0 = hom-ref / missing (skip), 1 = heterozygous, 2 = homozygous ALT.
 * Rules (SNP rule pinned by datasets/vepdataset.py:94-131; rest is the
 * documented policy of SURVEY Appendix D.4b, "parity unpinned"):
 *   - ref_len==1 && alt_len==1, het  -> IUPAC code of (REF base as in the
 *     FASTA upper-cased, ALT)       ; hom -> ALT base.
 *   - otherwise (indel / MNP), het or hom -> the REF span is replaced by ALT.
 *   - a record starting before the end of the previously APPLIED record is
 *     skipped (bcftools skips overlapping records); records that do not lie
 *     completely inside the window are skipped.
 *   - snp_only != 0 restates the VEP path's `|| TYPE!="snp"` filter
 *     (data_process.py:40-49): non-SNP records are ignored.
 * out must hold (w_end - w_start) + sum(alt_len) bytes.  Returns bytes written.
 */
int vfo_apply_variants(const uint8_t *chrom_seq, int64_t w_start, int64_t w_end,
                       const int64_t *pos, const int32_t *ref_len, const int32_t *alt_off,
                       const int32_t *alt_len, const uint8_t *gt, int n_var,
                       const uint8_t *alt_pool, int snp_only, uint8_t *out)
{
    int64_t cur = w_start;
    int o = 0;
    for (int i = 0; i < n_var; ++i) {
        if (gt[i] == 0) continue;
        int64_t p = pos[i];
        if (p < cur) continue;                       /* before window or overlapping an applied record */
        if (p + ref_len[i] > w_end) continue;        /* sticks out of the window */
        int is_snp = (ref_len[i] == 1 && alt_len[i] == 1);
        if (snp_only && !is_snp) continue;
        while (cur < p) out[o++] = chrom_seq[cur++];
        if (is_snp) {
            uint8_t a = alt_pool[alt_off[i]];
            if (gt[i] == 1) {
                uint8_t r = chrom_seq[p];
                if (r >= 'a' && r <= 'z') r = (uint8_t)(r - 32);
                out[o++] = vfo_iupac_het(r, a);
            } else {
                out[o++] = a;
            }
        } else {
            for (int k = 0; k < alt_len[i]; ++k) out[o++] = alt_pool[alt_off[i] + k];
        }
        cur = p + ref_len[i];
    }
    while (cur < w_end) out[o++] = chrom_seq[cur++];
    return o;
}

/* datasets/vcfdataset.py:198-217 (== vepdataset.py:495-507): pad with id 0 /
 * truncate to max_length; mask 1 = padding. */
void vfo_adjust_length(const int32_t *ids, int n, int max_length, int32_t *out_ids, uint8_t *out_mask)
{
    for (int i = 0; i < max_length; ++i) {
        if (i < n) { out_ids[i] = ids[i]; out_mask[i] = 0; }
        else       { out_ids[i] = 0;      out_mask[i] = 1; }
    }
}

/* datasets/vcfdataset.py:338-394: consecutive max_length-token chunks, last one
 * padded with id 0 (mask 1), at most max_chunks chunks.  Returns chunk count. */
int vfo_chunkify(const int32_t *ids, int n, int max_length, int max_chunks,
                 int32_t *out_ids, uint8_t *out_mask)
{
    int g = 0;
    for (int s = 0; s < n && g < max_chunks; s += max_length, ++g) {
        for (int j = 0; j < max_length; ++j) {
            int k = s + j;
            out_ids[g * max_length + j] = k < n ? ids[k] : 0;
            out_mask[g * max_length + j] = k < n ? 0 : 1;
        }
    }
    return g;
}

/* Window arithmetic.  CRE: utils/data_process.py:21-24 -> [max(0,start-nb), end+nb).
 * Gene: :387-400 -> '-' : [max(start, end-down), end+up) ;
 *                  '+' : s=max(0,start-up); [s, min(end, s+down))  (uses the shifted start). */
void vfo_cre_window(int64_t start, int64_t end, int64_t nb, int64_t *w0, int64_t *w1)
{
    *w0 = start - nb > 0 ? start - nb : 0;
    *w1 = end + nb;
}
void vfo_gene_window(int64_t start, int64_t end, int minus_strand, int64_t up, int64_t down,
                     int64_t *w0, int64_t *w1)
{
    if (minus_strand) {
        *w0 = start > end - down ? start : end - down;
        *w1 = end + up;
    } else {
        int64_t s = start - up > 0 ? start - up : 0;
        *w0 = s;
        *w1 = end < s + down ? end : s + down;
    }
}
