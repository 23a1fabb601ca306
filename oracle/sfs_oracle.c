/*
 * sfs_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (not part of the product).
 *
 * Plain-C restatement of the array-level 2DSFS-scan hot path, following the reference's algorithm step by step
 * (paths relative to uricchio/2DSFS-scan):
 *   oracle_decode      scripts/src/twoDSFS_class.py:118-130  per-sample ref/alt counting (on the 2-bit codes)
 *   spectra_of         :140-232 (2D: joint fold :199-206, skip :212, bin :216), :398-444 (1D raw alt), :446-463 (fold)
 *   multinomial_logpmf scipy 1.18.1 stats/_multivariate.py _logpmf: gammaln(n+1) + sum(xlogy(x,p) - gammaln(x+1))
 *   clr                :478-537 / :625-684  interior bins = sorted keys [1:-1], None when N == 0 or B == 0
 *   oracle_scan        :843-949 fixed-bp walk, :1515-1535 fixed-SNP walk, backgrounds :809-825 / whole genome :1970-1981
 * Like the reference it builds the FULL dense (2n1+1)(2n2+1) spectrum for every window and evaluates the likelihood
 * over every interior bin, so its cost profile (a fixed cost per window proportional to the bin count) is the
 * reference's.  Parity status: PINNED by tests/test_oracle_c.py against the reference's chr1 golden outputs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* minimal pthread parallel-for over chunks (the image's default CC has no OpenMP runtime) */
typedef void (*chunk_fn)(void* ctx, int64_t i);
typedef struct { chunk_fn fn; void* ctx; int64_t n; atomic_llong next; } pf_t;
static void* pf_worker(void* a) {
  pf_t* p = (pf_t*)a;
  for (;;) {
    long long i = atomic_fetch_add(&p->next, 1);
    if (i >= p->n) break;
    p->fn(p->ctx, i);
  }
  return NULL;
}
int oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
static void parallel_for(int nthreads, int64_t n, chunk_fn fn, void* ctx) {
  if (nthreads <= 0) nthreads = oracle_max_threads();
  if (nthreads > n) nthreads = (int)(n > 0 ? n : 1);
  pf_t p;
  p.fn = fn; p.ctx = ctx; p.n = n;
  atomic_init(&p.next, 0);
  if (nthreads <= 1) { pf_worker(&p); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, pf_worker, &p);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
}

typedef struct { const uint32_t* G; int64_t S; int W1, W2, ns1, ns2; uint16_t* cnt; } dec_t;
#define DEC_CHUNK 4096
static void decode_chunk(void* a, int64_t ci) {
  const dec_t* d = (const dec_t*)a;
  const uint32_t* G = d->G;
  const int W1 = d->W1, W2 = d->W2, ns1 = d->ns1, ns2 = d->ns2;
  uint16_t* cnt = d->cnt;
  const int RW = W1 + W2;
  const int64_t s_end = (ci + 1) * DEC_CHUNK < d->S ? (ci + 1) * DEC_CHUNK : d->S;
  for (int64_t s = ci * DEC_CHUNK; s < s_end; ++s) {
    /* B32 layout: word w of SNP s at ((s / 32) * RW + w) * 32 + s % 32 */
    const uint32_t* row = G + (s >> 5) * (int64_t)RW * 32 + (s & 31);
    for (int pop = 0; pop < 2; ++pop) {
      const uint32_t* blk = pop ? row + (int64_t)W1 * 32 : row;
      const int ns = pop ? ns2 : ns1;
      int ref = 0, alt = 0;
      for (int i = 0; i < ns; ++i) { /* one sample at a time, as the reference does */
        /* samples 32g..32g+31: word 2g holds the lo bits of their codes, word 2g+1 the hi bits */
        const unsigned lo = (blk[(2 * (i >> 5)) * 32] >> (i & 31)) & 1u, hi = (blk[(2 * (i >> 5) + 1) * 32] >> (i & 31)) & 1u;
        const unsigned code = lo | (hi << 1);
        if (code == 0) ref += 2;
        else if (code == 1) { ref += 1; alt += 1; }
        else if (code == 3) alt += 2; /* code 2 = missing: counts nothing */
      }
      cnt[s * 4 + 2 * pop] = (uint16_t)ref;
      cnt[s * 4 + 2 * pop + 1] = (uint16_t)alt;
    }
  }
}
void oracle_decode_mt(const uint32_t* G, int64_t S, int W1, int W2, int ns1, int ns2, uint16_t* cnt, int nthreads) {
  dec_t d = {G, S, W1, W2, ns1, ns2, cnt};
  parallel_for(nthreads, (S + DEC_CHUNK - 1) / DEC_CHUNK, decode_chunk, &d);
}
void oracle_decode(const uint32_t* G, int64_t S, int W1, int W2, int ns1, int ns2, uint16_t* cnt) {
  oracle_decode_mt(G, S, W1, W2, ns1, ns2, cnt, 1);
}

typedef struct {
  int64_t* h2;  /* (2n1+1)(2n2+1) */
  int64_t* h1a; /* 2n1+1 raw */
  int64_t* h1b; /* 2n2+1 raw */
} spectra_t;

static void spectra_alloc(spectra_t* s, int n1, int n2) {
  s->h2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(2 * n1 + 1) * (2 * n2 + 1));
  s->h1a = (int64_t*)malloc(sizeof(int64_t) * (2 * n1 + 1));
  s->h1b = (int64_t*)malloc(sizeof(int64_t) * (2 * n2 + 1));
}
static void spectra_free(spectra_t* s) { free(s->h2); free(s->h1a); free(s->h1b); }

/* returns -1 when a count exceeds 2n (the reference raises KeyError) */
static int spectra_of(const uint16_t* cnt, const uint8_t* inc, int64_t lo, int64_t hi, int n1, int n2, int fold, spectra_t* out) {
  const int R1 = 2 * n1 + 1, R2 = 2 * n2 + 1;
  memset(out->h2, 0, sizeof(int64_t) * (size_t)R1 * R2); /* the reference initialises every (i,j) key (:161-163) */
  memset(out->h1a, 0, sizeof(int64_t) * R1);
  memset(out->h1b, 0, sizeof(int64_t) * R2);
  for (int64_t s = lo; s < hi; ++s) {
    if (inc && !(inc[s] & 1)) continue;
    const int r1 = cnt[4 * s], a1 = cnt[4 * s + 1], r2 = cnt[4 * s + 2], a2 = cnt[4 * s + 3];
    int k1 = a1, k2 = a2;
    if (fold && a1 + a2 > n1 + n2) { k1 = r1; k2 = r2; }
    if (k1 >= R1 || k2 >= R2 || a1 >= R1 || a2 >= R2) return -1;
    if (!(k1 == 0 && k2 == 0)) out->h2[(int64_t)k1 * R2 + k2] += 1;
    if (a1) out->h1a[a1] += 1;
    if (a2) out->h1b[a2] += 1;
  }
  return 0;
}

static void fold1d(const int64_t* raw, int n, int64_t* folded /* n+1 */) {
  for (int k = 0; k <= n; ++k) folded[k] = 0;
  for (int f = 0; f <= 2 * n; ++f) {
    int m = f < 2 * n - f ? f : 2 * n - f;
    folded[m] += raw[f];
  }
}

static double multinomial_logpmf(const int64_t* x, int64_t nb, int64_t n, const double* p) {
  double acc = 0.0;
  for (int64_t i = 0; i < nb; ++i) {
    const int64_t xi = x[i];
    const double xl = xi == 0 ? 0.0 : (double)xi * log(p[i]); /* xlogy */
    acc += xl - lgamma((double)xi + 1.0);
  }
  return lgamma((double)n + 1.0) + acc;
}

/* 2*(ll_fg - ll_bg) over interior vectors x (counts) and b (background); *none = 1 when the reference returns None */
static double clr(const int64_t* x, const int64_t* b, int64_t nb, double* pf, double* pb, int* none) {
  int64_t N = 0, B = 0;
  for (int64_t i = 0; i < nb; ++i) { N += x[i]; B += b[i]; }
  *none = (N == 0 || B == 0);
  if (*none) return NAN;
  for (int64_t i = 0; i < nb; ++i) { pf[i] = (double)x[i] / (double)N; pb[i] = (double)b[i] / (double)B; }
  return 2.0 * (multinomial_logpmf(x, nb, N, pf) - multinomial_logpmf(x, nb, N, pb));
}

typedef struct { int32_t chrom; int64_t start, end, lo, hi; } win_t;

typedef struct {
  const uint16_t* cnt; const uint8_t* inc; const win_t* wins; int64_t nw; int n1, n2, fold, snp_mode, bg_genome; int64_t bins;
  spectra_t* bg; int64_t** bgf1; int64_t** bgf2; uint8_t* keep; double *T2, *T1a, *T1b; uint8_t* nn; int nthreads;
  int nslab; atomic_int bad; atomic_llong next;
} scan_t;

/* each worker owns its dense scratch and pulls windows from a shared counter (dynamic schedule) */
static void scan_slab(void* a, int64_t slab) {
  (void)slab;
  scan_t* q = (scan_t*)a;
  const int n1 = q->n1, n2 = q->n2;
  const int64_t bins = q->bins;
  spectra_t fg;
  spectra_alloc(&fg, n1, n2);
  int64_t* f1 = (int64_t*)malloc(sizeof(int64_t) * (n1 + 1));
  int64_t* f2 = (int64_t*)malloc(sizeof(int64_t) * (n2 + 1));
  double* pf = (double*)malloc(sizeof(double) * (size_t)bins);
  double* pb = (double*)malloc(sizeof(double) * (size_t)bins);
  for (;;) {
    const long long w = atomic_fetch_add(&q->next, 1);
    if (w >= q->nw) break;
    const int g = q->bg_genome ? 0 : q->wins[w].chrom;
    if (spectra_of(q->cnt, q->inc, q->wins[w].lo, q->wins[w].hi, n1, n2, q->fold, &fg)) { atomic_store(&q->bad, 1); continue; }
    if (q->snp_mode) { /* :1496 skip windows whose 2D spectrum sums to zero */
      int64_t tot = 0;
      for (int64_t k = 0; k < bins; ++k) tot += fg.h2[k];
      if (tot == 0) continue;
    }
    q->keep[w] = 1;
    int none;
    q->T2[w] = clr(fg.h2 + 1, q->bg[g].h2 + 1, bins - 2, pf, pb, &none);
    q->nn[w] |= none ? 1 : 0;
    fold1d(fg.h1a, n1, f1);
    fold1d(fg.h1b, n2, f2);
    q->T1a[w] = clr(f1 + 1, q->bgf1[g] + 1, n1 - 1 > 0 ? n1 - 1 : 0, pf, pb, &none);
    q->nn[w] |= none ? 2 : 0;
    q->T1b[w] = clr(f2 + 1, q->bgf2[g] + 1, n2 - 1 > 0 ? n2 - 1 : 0, pf, pb, &none);
    q->nn[w] |= none ? 4 : 0;
  }
  spectra_free(&fg);
  free(f1); free(f2); free(pf); free(pb);
}

int64_t oracle_scan(const uint16_t* cnt, const int32_t* pos, const int64_t* off, int C, int n1, int n2, int fold,
                    int64_t W, int snp_mode, int bg_genome, int nthreads, const uint8_t* inc, int64_t cap,
                    int32_t* o_chrom, int64_t* o_start, int64_t* o_end, int32_t* o_count, double* o_T2, double* o_T1a,
                    double* o_T1b, uint8_t* o_none) {
  const int R1 = 2 * n1 + 1, R2 = 2 * n2 + 1;
  const int64_t bins = (int64_t)R1 * R2;
  const int64_t S = off[C];
  /* ---- windows, by the sequential walk ---- */
  win_t* wins = (win_t*)malloc(sizeof(win_t) * (size_t)(S + 1));
  int64_t nw = 0;
  for (int c = 0; c < C; ++c) {
    const int64_t clo = off[c], chi = off[c + 1];
    if (snp_mode) {
      for (int64_t j = 0; (j + 1) * W <= chi - clo; ++j) {
        win_t w = {c, j == 0 ? pos[clo] : (int64_t)pos[clo + j * W - 1] + 1, pos[clo + (j + 1) * W - 1], clo + j * W, clo + (j + 1) * W};
        wins[nw++] = w;
      }
    } else {
      int64_t start = 1, wlo = clo;
      for (int64_t s = clo; s < chi; ++s) {
        if (!(pos[s] < start + W)) {
          if (s > wlo) { win_t w = {c, start, start + W - 1, wlo, s}; wins[nw++] = w; }
          start += W * ((pos[s] - start) / W);
          wlo = s;
        }
      }
      if (chi > wlo) { win_t w = {c, start, start + W - 1, wlo, chi}; wins[nw++] = w; }
    }
  }
  /* ---- backgrounds ---- */
  const int NG = bg_genome ? 1 : C;
  spectra_t* bg = (spectra_t*)malloc(sizeof(spectra_t) * NG);
  int64_t** bgf1 = (int64_t**)malloc(sizeof(int64_t*) * NG);
  int64_t** bgf2 = (int64_t**)malloc(sizeof(int64_t*) * NG);
  int bad = 0;
  for (int g = 0; g < NG; ++g) {
    spectra_alloc(&bg[g], n1, n2);
    bgf1[g] = (int64_t*)malloc(sizeof(int64_t) * (n1 + 1));
    bgf2[g] = (int64_t*)malloc(sizeof(int64_t) * (n2 + 1));
    if (spectra_of(cnt, inc, bg_genome ? 0 : off[g], bg_genome ? S : off[g + 1], n1, n2, fold, &bg[g])) bad = 1;
    fold1d(bg[g].h1a, n1, bgf1[g]);
    fold1d(bg[g].h1b, n2, bgf2[g]);
  }
  /* ---- windows in parallel (the reference's only admissible parallelism is over independent windows / chromosomes) ---- */
  int64_t emitted = 0;
  uint8_t* keep = (uint8_t*)calloc((size_t)nw + 1, 1);
  double *T2 = (double*)malloc(sizeof(double) * (nw + 1)), *T1a = (double*)malloc(sizeof(double) * (nw + 1)),
         *T1b = (double*)malloc(sizeof(double) * (nw + 1));
  uint8_t* nn = (uint8_t*)calloc((size_t)nw + 1, 1);
  scan_t sc;
  memset(&sc, 0, sizeof sc);
  sc.cnt = cnt; sc.inc = inc; sc.wins = wins; sc.nw = nw; sc.n1 = n1; sc.n2 = n2; sc.fold = fold; sc.snp_mode = snp_mode;
  sc.bg_genome = bg_genome; sc.bins = bins; sc.bg = bg; sc.bgf1 = bgf1; sc.bgf2 = bgf2; sc.keep = keep; sc.T2 = T2; sc.T1a = T1a;
  sc.T1b = T1b; sc.nn = nn; sc.nthreads = nthreads;
  atomic_init(&sc.bad, 0);
  atomic_init(&sc.next, 0);
  {
    int nt = nthreads > 0 ? nthreads : oracle_max_threads();
    /* one task per thread-sized slab so the per-thread dense scratch is allocated once */
    sc.nslab = nt;
    parallel_for(nt, nt, scan_slab, &sc);
  }
  if (atomic_load(&sc.bad)) bad = 1;
  for (int64_t w = 0; w < nw; ++w) {
    if (!keep[w]) continue;
    if (emitted < cap) {
      o_chrom[emitted] = wins[w].chrom; o_start[emitted] = wins[w].start; o_end[emitted] = wins[w].end;
      int64_t n = 0;
      for (int64_t s = wins[w].lo; s < wins[w].hi; ++s) n += inc ? ((inc[s] >> 1) & 1) : 1;
      o_count[emitted] = (int32_t)n;
      o_T2[emitted] = T2[w]; o_T1a[emitted] = T1a[w]; o_T1b[emitted] = T1b[w]; o_none[emitted] = nn[w];
    }
    ++emitted;
  }
  for (int g = 0; g < NG; ++g) { spectra_free(&bg[g]); free(bgf1[g]); free(bgf2[g]); }
  free(bg); free(bgf1); free(bgf2); free(wins); free(keep); free(T2); free(T1a); free(T1b); free(nn);
  return bad ? -1 : emitted;
}

