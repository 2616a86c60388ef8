/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (subphaser_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the CPU baseline.
 *
 * CPU restatement of the k-mer counting step the reference delegates to the third-party binary
 * jellyfish 2.2.10 (pinned in /root/reference/SubPhaser.yaml:66; NOT vendored, NOT installed here):
 *
 *   cat CHR.fasta | jellyfish count -t T -m K -s 100000000 --canonical /dev/stdin -o P.jf
 *   jellyfish dump -c -o P.fa P.jf -L lower_count          (reference: subphaser/Jellyfish.py:697-700)
 *
 * Published semantics restated here (jellyfish 2.2 manual, "Counting k-mers in sequencing reads",
 * option -C/--canonical; dump -c -L):
 *   - every window of K consecutive A/C/G/T characters (either case) of a record is one k-mer;
 *   - any other character (N, IUPAC codes, '-', ...) cannot be part of a k-mer: windows covering it
 *     are skipped; line breaks inside a record are ignored; k-mers never span two records;
 *   - with --canonical a k-mer and its reverse complement are the same key, represented by the
 *     lexicographically smaller of the two (A < C < G < T);
 *   - counts are exact; dump -c prints "KMER COUNT" for every key with COUNT >= L, in arbitrary order.
 * PARITY UNPINNED against jellyfish itself: neither the binary nor any reference golden vector exists
 * in this environment.  The restatement is pinned instead against hand-computed
 * known-answer vectors and an independent string-level Python counter (tests/test_oracle.py).
 *
 * Implementation: parse to a compact code array, then T pthreads roll forward/reverse-complement
 * words over disjoint position ranges and insert into one shared lock-free open-addressed table
 * (CAS on the key, atomic add on the count) — the same scheme jellyfish uses, so this doubles as the
 * multi-threaded "port" CPU baseline of bench.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EMPTY_KEY (~0ull)

typedef struct {
    uint64_t* keys;
    uint32_t* counts;
    uint64_t slots; /* power of two */
} table_t;

static inline uint64_t hash64(uint64_t x) {
    /* splitmix64 finaliser (deliberately different from the GPU's murmur finaliser) */
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

static int table_add(table_t* t, uint64_t key, uint32_t add) {
    uint64_t mask = t->slots - 1;
    uint64_t s = hash64(key) & mask;
    for (uint64_t probes = 0; probes < t->slots; probes++) {
        uint64_t cur = __atomic_load_n(&t->keys[s], __ATOMIC_RELAXED);
        if (cur == EMPTY_KEY) {
            uint64_t expected = EMPTY_KEY;
            if (__atomic_compare_exchange_n(&t->keys[s], &expected, key, 0, __ATOMIC_RELAXED,
                                            __ATOMIC_RELAXED))
                cur = key;
            else
                cur = expected;
        }
        if (cur == key) {
            __atomic_fetch_add(&t->counts[s], add, __ATOMIC_RELAXED);
            return 0;
        }
        s = (s + 1) & mask;
    }
    return -1;
}

/* FASTA bytes -> codes: 0..3 = A,C,G,T (case-insensitive), 4 = anything else / record separator.
 * Returns the number of codes written (<= n). */
uint64_t orc_fasta_to_codes(const uint8_t* buf, uint64_t n, uint8_t* codes, uint64_t* n_records) {
    uint64_t out = 0, recs = 0;
    int at_line_start = 1, in_header = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t c = buf[i];
        if (c == '\n') {
            at_line_start = 1;
            in_header = 0;
            continue;
        }
        if (at_line_start && c == '>') {
            in_header = 1;
            recs++;
            if (i != 0) codes[out++] = 4; /* k-mers never span records */
        }
        at_line_start = 0;
        if (in_header || c == '\r') continue;
        switch (c) {
            case 'A': case 'a': codes[out++] = 0; break;
            case 'C': case 'c': codes[out++] = 1; break;
            case 'G': case 'g': codes[out++] = 2; break;
            case 'T': case 't': codes[out++] = 3; break;
            default: codes[out++] = 4;
        }
    }
    if (n_records) *n_records = recs;
    return out;
}

typedef struct {
    const uint8_t* codes;
    uint64_t n, beg, end; /* k-mer START positions [beg, end) */
    int k;
    table_t* tab;
    uint64_t n_valid;
    int failed;
} job_t;

static void* count_range(void* arg) {
    job_t* j = (job_t*)arg;
    const int k = j->k;
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fwd = 0, rc = 0;
    int run = 0; /* number of consecutive valid bases ending at the current one */
    uint64_t last = j->end + (uint64_t)k - 1;
    if (last > j->n) last = j->n;
    for (uint64_t p = j->beg; p < last; p++) {
        uint8_t c = j->codes[p];
        if (c > 3) {
            run = 0;
            fwd = rc = 0;
            continue;
        }
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
        if (++run >= k) {
            uint64_t canon = fwd < rc ? fwd : rc;
            if (table_add(j->tab, canon, 1) != 0) j->failed = 1;
            j->n_valid++;
        }
    }
    return NULL;
}

typedef struct {
    uint64_t n_bases, n_records, n_valid_kmers, n_distinct, n_dumped, sum_dumped;
} orc_stats_t;

/* Count the canonical k-mers of FASTA bytes.  On success *keys_out / *counts_out are malloc'ed arrays
 * (free with orc_free) holding every k-mer with count >= lower_count (unordered), *n_out their
 * number.  Key encoding: 2 bits per base, first base most significant, A=0 C=1 G=2 T=3. */
int orc_count_fasta(const uint8_t* buf, uint64_t nbytes, int k, uint32_t lower_count, int nthreads,
                    uint64_t** keys_out, uint32_t** counts_out, uint64_t* n_out, orc_stats_t* st) {
    if (k < 1 || k > 32 || nthreads < 1) return -1;
    uint8_t* codes = (uint8_t*)malloc(nbytes + 1);
    if (!codes) return -2;
    uint64_t recs = 0;
    uint64_t n = orc_fasta_to_codes(buf, nbytes, codes, &recs);
    table_t tab;
    uint64_t want = (uint64_t)((double)n / 0.6) + 1024;
    if (k < 31 && (1ull << (2 * k)) * 2 < want) want = (1ull << (2 * k)) * 2;
    tab.slots = 1024;
    while (tab.slots < want) tab.slots <<= 1;
    tab.keys = (uint64_t*)malloc(tab.slots * sizeof(uint64_t));
    tab.counts = (uint32_t*)calloc(tab.slots, sizeof(uint32_t));
    if (!tab.keys || !tab.counts) {
        free(codes); free(tab.keys); free(tab.counts);
        return -2;
    }
    memset(tab.keys, 0xFF, tab.slots * sizeof(uint64_t));

    uint64_t n_starts = (n >= (uint64_t)k) ? n - k + 1 : 0;
    job_t* jobs = (job_t*)calloc(nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
    uint64_t per = (n_starts + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        jobs[t].codes = codes;
        jobs[t].n = n;
        jobs[t].k = k;
        jobs[t].tab = &tab;
        jobs[t].beg = (uint64_t)t * per < n_starts ? (uint64_t)t * per : n_starts;
        jobs[t].end = jobs[t].beg + per < n_starts ? jobs[t].beg + per : n_starts;
        if (nthreads == 1) count_range(&jobs[t]);
        else pthread_create(&th[t], NULL, count_range, &jobs[t]);
    }
    uint64_t n_valid = 0;
    int failed = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        n_valid += jobs[t].n_valid;
        failed |= jobs[t].failed;
    }
    free(jobs); free(th); free(codes);
    if (failed) {
        free(tab.keys); free(tab.counts);
        return -3;
    }
    uint64_t distinct = 0, dumped = 0, sum = 0;
    for (uint64_t s = 0; s < tab.slots; s++)
        if (tab.keys[s] != EMPTY_KEY) {
            distinct++;
            if (tab.counts[s] >= lower_count) {
                dumped++;
                sum += tab.counts[s];
            }
        }
    uint64_t* ok = (uint64_t*)malloc((dumped + 1) * sizeof(uint64_t));
    uint32_t* oc = (uint32_t*)malloc((dumped + 1) * sizeof(uint32_t));
    uint64_t w = 0;
    for (uint64_t s = 0; s < tab.slots; s++)
        if (tab.keys[s] != EMPTY_KEY && tab.counts[s] >= lower_count) {
            ok[w] = tab.keys[s];
            oc[w] = tab.counts[s];
            w++;
        }
    free(tab.keys); free(tab.counts);
    *keys_out = ok;
    *counts_out = oc;
    *n_out = dumped;
    if (st) {
        st->n_bases = n;
        st->n_records = recs;
        st->n_valid_kmers = n_valid;
        st->n_distinct = distinct;
        st->n_dumped = dumped;
        st->sum_dumped = sum;
    }
    return 0;
}

void orc_free(void* p) { free(p); }

/* Per-position lookup restating Seqs.map_kmer_each4 (/root/reference/subphaser/Seqs.py:209-237) on a
 * code array: sorted canonical keys + subgenome ids, binary search; counts[line * S + sg] += 1 with
 * line = i / bin_size + (chunk ? (i + k - 1) / chunk : 0). Used as the CPU baseline of the map stage. */
uint64_t orc_map_bins(const uint8_t* codes, uint64_t n, int k, const uint64_t* keys, const uint8_t* sgs,
                      uint64_t nkeys, int S, uint64_t bin_size, uint64_t chunk, uint32_t* counts) {
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fwd = 0, rc = 0, hits = 0;
    int run = 0;
    for (uint64_t p = 0; p < n; p++) {
        uint8_t c = codes[p];
        if (c > 3) { run = 0; fwd = rc = 0; continue; }
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
        if (++run >= k) {
            uint64_t canon = fwd < rc ? fwd : rc;
            uint64_t lo = 0, hi = nkeys;
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if (keys[mid] < canon) lo = mid + 1; else hi = mid;
            }
            if (lo < nkeys && keys[lo] == canon) {
                uint64_t i = p + 1 - (uint64_t)k;
                uint64_t line = i / bin_size + (chunk ? (i + (uint64_t)k - 1) / chunk : 0);
                counts[line * (uint64_t)S + sgs[lo]]++;
                hits++;
            }
        }
    }
    return hits;
}
