/* tools/synth.c -- synthetic workload generator for bench.py and the full-size property tests
 * (SURVEY.md section 8(d), configs 2-5).  Bench infrastructure, not part of the product library.
 *
 * PRNG: xoshiro256** seeded through splitmix64, so every workload is reproducible from its seed on
 * any box.  All sequences are lower-case a/c/g/t.
 *
 *   synth_contig   config 2: genes and spacers alternate on random strands until `target` bases.
 *                  gene = atg + n codons (n = 100 + Geom(mean 230), drawn from a stop-free codon table)
 *                  + a stop codon from {taa,tag,tga}; spacer = 30 + Geom(mean 90) i.i.d. bases at `gc`.
 *   synth_reads    configs 3/5: fixed-length reads at uniform positions / random strand of a contig; with
 *                  indel != 0, 454-like homopolymer errors: every run of length r >= 3 gains or loses one
 *                  copy with probability min(0.3, 0.02 r).
 *   synth_coding   config 4: n stop-free coding sequences of `codons` codons each.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t s[4];
} rng_t;

static uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) {
  for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
  uint64_t* s = r->s;
  const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_u01(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t* r, uint64_t n) { return (uint64_t)(rng_u01(r) * (double)n); }
/* number of failures before the first success, success probability 1/(mean+1): E = mean */
static int rng_geom(rng_t* r, double mean) {
  double u = rng_u01(r);
  return (int)floor(log(1.0 - u) / log(mean / (mean + 1.0)));
}

static const char ACGT[4] = {'a', 'c', 'g', 't'};
static int is_stop(int c) { return c == 48 /*taa*/ || c == 50 /*tag*/ || c == 56 /*tga*/; }

typedef struct {
  double cdf[64];
} codon_table;

static void make_table(const double* freq, codon_table* t) {
  double sum = 0.0, acc = 0.0;
  for (int c = 0; c < 64; c++) sum += is_stop(c) ? 0.0 : freq[c];
  for (int c = 0; c < 64; c++) {
    acc += is_stop(c) ? 0.0 : freq[c] / sum;
    t->cdf[c] = acc;
  }
  t->cdf[63] = 1.0;
}
static int draw_codon(rng_t* r, const codon_table* t) {
  double u = rng_u01(r);
  int lo = 0, hi = 63;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (t->cdf[mid] > u) hi = mid;
    else lo = mid + 1;
  }
  while (is_stop(lo)) lo--; /* cdf is flat over stops; step back to the codon that owns the mass */
  return lo;
}
static void put_codon(char* out, int c) {
  out[0] = ACGT[(c >> 4) & 3];
  out[1] = ACGT[(c >> 2) & 3];
  out[2] = ACGT[c & 3];
}
static void revcomp(char* s, int64_t n) {
  for (int64_t i = 0, j = n - 1; i <= j; i++, j--) {
    char a = s[i], b = s[j];
#define COMP(ch) ((ch) == 'a' ? 't' : (ch) == 'c' ? 'g' : (ch) == 'g' ? 'c' : 'a')
    s[i] = COMP(b);
    s[j] = COMP(a);
#undef COMP
  }
}

/* out must hold target + 64 bytes; returns target */
int64_t synth_contig(uint64_t seed, int64_t target, const double* codon_freq, double gc, char* out) {
  rng_t r;
  rng_seed(&r, seed);
  codon_table t;
  make_table(codon_freq, &t);
  static const int stops[3] = {48, 50, 56};
  int64_t n = 0;
  int cap = 1 << 16;
  char* gene = (char*)malloc((size_t)cap);
  while (n < target) {
    /* spacer */
    int sp = 30 + rng_geom(&r, 90.0);
    for (int i = 0; i < sp && n < target; i++) {
      double u = rng_u01(&r);
      int b = (u < gc) ? (rng_u01(&r) < 0.5 ? 1 : 2) : (rng_u01(&r) < 0.5 ? 0 : 3);
      out[n++] = ACGT[b];
    }
    if (n >= target) break;
    /* gene */
    int nc = 100 + rng_geom(&r, 230.0);
    int len = 3 * (nc + 2);
    if (len > cap) {
      cap = len * 2;
      gene = (char*)realloc(gene, (size_t)cap);
    }
    put_codon(gene, 14 /* atg */);
    for (int i = 0; i < nc; i++) put_codon(gene + 3 + 3 * i, draw_codon(&r, &t));
    put_codon(gene + 3 + 3 * nc, stops[rng_below(&r, 3)]);
    if (rng_next(&r) & 1) revcomp(gene, len);
    int64_t take = (n + len <= target) ? len : target - n;
    memcpy(out + n, gene, (size_t)take);
    n += take;
  }
  free(gene);
  return n;
}

/* out must hold n_reads * (rlen + rlen / 3 + 2) bytes; off has n_reads + 1 entries; returns total bases */
int64_t synth_reads(uint64_t seed, const char* contig, int64_t clen, int64_t n_reads, int rlen, int indel, char* out,
                    int64_t* off) {
  rng_t r;
  rng_seed(&r, seed);
  char* buf = (char*)malloc((size_t)rlen + 8);
  int64_t n = 0;
  off[0] = 0;
  for (int64_t i = 0; i < n_reads; i++) {
    int64_t a = (int64_t)rng_below(&r, (uint64_t)(clen - rlen + 1));
    memcpy(buf, contig + a, (size_t)rlen);
    if (rng_next(&r) & 1) revcomp(buf, rlen);
    if (!indel) {
      memcpy(out + n, buf, (size_t)rlen);
      n += rlen;
    } else {
      int p = 0;
      while (p < rlen) {
        int e = p + 1;
        while (e < rlen && buf[e] == buf[p]) e++;
        int run = e - p, emit = run;
        if (run >= 3) {
          double pr = 0.02 * run;
          if (pr > 0.3) pr = 0.3;
          if (rng_u01(&r) < pr) emit += (rng_next(&r) & 1) ? 1 : -1;
        }
        for (int k = 0; k < emit; k++) out[n++] = buf[p];
        p = e;
      }
    }
    off[i + 1] = n;
  }
  free(buf);
  return n;
}

/* out must hold n_seqs * 3 * codons bytes; returns total bases */
int64_t synth_coding(uint64_t seed, int64_t n_seqs, int codons, const double* codon_freq, char* out) {
  rng_t r;
  rng_seed(&r, seed);
  codon_table t;
  make_table(codon_freq, &t);
  int64_t n = 0;
  for (int64_t i = 0; i < n_seqs; i++)
    for (int k = 0; k < codons; k++, n += 3) put_codon(out + n, draw_codon(&r, &t));
  return n;
}
