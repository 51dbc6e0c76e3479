/* oracle/icm_oracle.c -- TEST INFRASTRUCTURE ONLY (see icm_oracle.h).
 *
 * Plain-C CPU restatement of the reference's ICM scoring/training hot path, written
 * from the algorithm and citing the reference file:line each
 * function follows.  Parity status: PINNED (goldens + compiled reference, see
 * tests/test_oracle_*.py).  The product never links this file.
 *
 * Compile: gcc -O2 -fno-fast-math -ffp-contract=off (no FMA contraction: the
 * reference is plain x86-64 SSE2 double arithmetic).
 */
#define _GNU_SOURCE
#include "icm_oracle.h"

#include <ctype.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ALPHA 4
#define ICM_VERSION_ID 200
#define ID_STRING_LEN 150
#define PARENT(x) (((x)-1) / ALPHA)

/* ------------------------------------------------------------------------- */
/* alphabet: Filter (Common/gene.cc:1139-1175), Subscript (ICM/icm.cc:2008-2027),
 * Complement (gene.cc:1084, COMPLEMENT_TABLE :15-19)                          */

static int orc_filter(int ch) {
  switch (tolower(ch)) {
    case 'a': case 'c': case 'g': case 't': return ch;
    case 'r': return 'g';
    case 'y': return 'c';
    case 's': return 'c';
    case 'w': return 't';
    case 'm': return 'c';
    case 'k': return 't';
    case 'b': return 'c';
    case 'd': return 'g';
    case 'h': return 'c';
    case 'v': return 'c';
    default: return 'c';
  }
}

static int sub_of(int ch) {
  switch (tolower(orc_filter(ch))) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    default: return 3; /* 't' */
  }
}

/* only ever applied to filtered lower-case acgt in the paths restated here */
static int comp_of(int ch) {
  switch (ch) {
    case 'a': return 't';
    case 'c': return 'g';
    case 'g': return 'c';
    case 't': return 'a';
    default: return 'n';
  }
}

/* ------------------------------------------------------------------------- */
/* model container (ICM/icm.cc:24-61) */

static int int_power(int b, int e) {
  int r = 1;
  while (e-- > 0) r *= b;
  return r;
}

orc_icm* orc_icm_new(int w, int d, int p) {
  orc_icm* m = (orc_icm*)calloc(1, sizeof(orc_icm));
  m->model_len = w;
  m->model_depth = d;
  m->periodicity = p;
  m->num_nodes = (int_power(ALPHA, d + 1) - 1) / (ALPHA - 1);
  m->mip = (short*)calloc((size_t)p * m->num_nodes, sizeof(short));
  m->prob = (float*)calloc((size_t)p * m->num_nodes * 4, sizeof(float));
  return m;
}

void orc_icm_free(orc_icm* m) {
  if (!m) return;
  free(m->mip);
  free(m->prob);
  free(m);
}

#define MIP(m, f, n) ((m)->mip[(size_t)(f) * (m)->num_nodes + (n)])
#define PROB(m, f, n) ((m)->prob + 4 * ((size_t)(f) * (m)->num_nodes + (n)))

/* ICM_t::Input (ICM/icm.cc:614-726): binary model format */
orc_icm* orc_icm_read(const char* path) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return NULL;
  char line[ID_STRING_LEN];
  int param[6];
  if (fread(line, 1, ID_STRING_LEN, fp) != ID_STRING_LEN || fread(param, sizeof(int), 6, fp) != 6 ||
      param[0] != ICM_VERSION_ID || param[1] != ID_STRING_LEN) {
    fclose(fp);
    return NULL;
  }
  orc_icm* m = (orc_icm*)calloc(1, sizeof(orc_icm));
  m->model_len = param[2];
  m->model_depth = param[3];
  m->periodicity = param[4];
  m->num_nodes = param[5];
  m->mip = (short*)calloc((size_t)m->periodicity * m->num_nodes, sizeof(short));
  m->prob = (float*)calloc((size_t)m->periodicity * m->num_nodes * 4, sizeof(float));
  int period = -1, prev_node = 0, node_id;
  while (fread(&node_id, sizeof(int), 1, fp) == 1) {
    if (node_id < 0) break;
    if (node_id == 0) period++;
    if (period < 0 || period >= m->periodicity || node_id >= m->num_nodes ||
        fread(PROB(m, period, node_id), sizeof(float), 4, fp) != 4 ||
        fread(&MIP(m, period, node_id), sizeof(short), 1, fp) != 1) {
      fclose(fp);
      orc_icm_free(m);
      return NULL;
    }
    /* nodes missing from the file were cut: mark -2 (icm.cc:699-721) */
    if (node_id != 0 && prev_node != node_id - 1)
      for (int i = prev_node + 1; i < node_id; i++) MIP(m, period, i) = -2;
    if (node_id == 0 && period > 0)
      for (int i = prev_node + 1; i < m->num_nodes; i++) MIP(m, period - 1, i) = -2;
    prev_node = node_id;
  }
  fclose(fp);
  if (period != m->periodicity - 1) {
    orc_icm_free(m);
    return NULL;
  }
  if (prev_node != m->num_nodes - 1)
    for (int i = prev_node + 1; i < m->num_nodes; i++) MIP(m, period, i) = -2;
  return m;
}

/* ICM_t::Output / Output_Node / Write_Header, binary form (icm.cc:729-803, 961-998) */
int orc_icm_write(const orc_icm* m, const char* path) {
  FILE* fp = fopen(path, "wb");
  if (!fp) return -1;
  char line[ID_STRING_LEN];
  memset(line, 0, sizeof line);
  snprintf(line, sizeof line, ">ver = %.2f  len = %d  depth = %d  periodicity = %d  nodes = %d\n",
           ICM_VERSION_ID / 100.0, m->model_len, m->model_depth, m->periodicity, m->num_nodes);
  fwrite(line, 1, ID_STRING_LEN, fp);
  int param[6] = {ICM_VERSION_ID, ID_STRING_LEN, m->model_len, m->model_depth, m->periodicity, m->num_nodes};
  fwrite(param, sizeof(int), 6, fp);
  for (int f = 0; f < m->periodicity; f++)
    for (int i = 0; i < m->num_nodes; i++)
      if (i == 0 || MIP(m, f, i) >= -1) {
        fwrite(&i, sizeof(int), 1, fp);
        fwrite(PROB(m, f, i), sizeof(float), 4, fp);
        fwrite(&MIP(m, f, i), sizeof(short), 1, fp);
      }
  int end_marker = -1;
  fwrite(&end_marker, sizeof(int), 1, fp);
  fclose(fp);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* ICM_t::Build_Indep_WO_Stops (ICM/icm.cc:65-216) */

orc_icm* orc_build_indep_wo_stops(double gc, const char* const* stops, int n_stops) {
  orc_icm* m = orc_icm_new(3, 2, 3);
  double codon_prob[64], base_prob[4], acc[3][21][4];
  int pattern[3] = {0, 0, 0};
  base_prob[1] = base_prob[2] = gc / 2.0;
  base_prob[0] = base_prob[3] = 0.5 - base_prob[1];
  for (int i = 0; i < 64; i++) {
    codon_prob[i] = base_prob[pattern[0]] * base_prob[pattern[1]] * base_prob[pattern[2]];
    for (int j = 2; j >= 0; j--) {
      pattern[j]++;
      if (pattern[j] == 4) pattern[j] = 0;
      else break;
    }
  }
  /* stop codons entered reversed: scoring runs 3'->5' (icm.cc:113-125) */
  for (int i = 0; i < n_stops; i++) {
    int j = sub_of(stops[i][0]) + 4 * sub_of(stops[i][1]) + 16 * sub_of(stops[i][2]);
    codon_prob[j] = 1e-20;
  }
  double sum = 0.0;
  for (int i = 0; i < 64; i++) sum += codon_prob[i];
  for (int i = 0; i < 64; i++) codon_prob[i] /= sum;

  /* NOTE: the reference accumulates into the float prob[] fields (icm.cc:155,173,192);
   * float += double rounds to float at every step, mirrored here. */
  float facc[3][21][4];
  memset(facc, 0, sizeof facc);
  (void)acc;
  for (int i = 0; i < 3; i++) {
    int d1 = int_power(4, (3 - i) % 3);
    MIP(m, i, 0) = (i == 1) ? -1 : 1;
    for (int j = 0; j < 64; j++) facc[i][0][(j / d1) % 4] += codon_prob[j];
  }
  for (int i = 0; i < 3; i++) {
    int d1 = int_power(4, (3 - i) % 3), d2 = int_power(4, (4 - i) % 3);
    for (int j = 0; j < 4; j++) MIP(m, i, 1 + j) = (i == 2) ? -1 : 0;
    if (i != 1)
      for (int j = 0; j < 64; j++) facc[i][1 + (j / d2) % 4][(j / d1) % 4] += codon_prob[j];
  }
  {
    int i = 0;
    int d1 = int_power(4, (3 - i) % 3), d2 = int_power(4, (4 - i) % 3), d3 = int_power(4, (5 - i) % 3);
    for (int j = 0; j < 16; j++) MIP(m, i, 5 + j) = -1;
    for (int j = 0; j < 64; j++) {
      int k = 4 * ((j / d2) % 4) + (j / d3) % 4;
      facc[i][5 + k][(j / d1) % 4] += codon_prob[j];
    }
  }
  /* normalise + logs; sum is a double accumulating floats (icm.cc:203-211) */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 21; j++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += facc[i][j][k];
      for (int k = 0; k < 4; k++) PROB(m, i, j)[k] = (float)(s == 0.0 ? 0.0 : log(facc[i][j][k] / s));
    }
  return m;
}

/* ------------------------------------------------------------------------- */
/* scalar scoring */

/* ICM_t::Full_Window_Prob (ICM/icm.cc:557-610) */
double orc_full_window_prob(const orc_icm* m, const char* w, int frame) {
  int node = 0, pos;
  for (int i = 0; i < m->model_depth; i++) {
    pos = MIP(m, frame, node);
    if (pos == -1) break;
    if (pos < -1) {
      node = PARENT(node);
      break;
    }
    node = node * ALPHA + sub_of(w[pos]) + 1;
  }
  if (MIP(m, frame, node) < -1) node = PARENT(node);
  return (double)PROB(m, frame, node)[sub_of(w[m->model_len - 1])];
}

/* ICM_t::Partial_Window_Prob (ICM/icm.cc:807-842) */
double orc_partial_window_prob(const orc_icm* m, int predict_pos, const char* s, int frame) {
  int start = predict_pos - (m->model_len - 1);
  int node = 0;
  for (int i = 0; i < m->model_depth; i++) {
    int pos = start + MIP(m, frame, node);
    if (pos < 0) break;
    node = node * ALPHA + sub_of(s[pos]) + 1;
  }
  if (MIP(m, frame, node) == -2) node = PARENT(node);
  return (double)PROB(m, frame, node)[sub_of(s[predict_pos])];
}

/* ICM_t::Score_String (ICM/icm.cc:864-903) */
double orc_score_string(const orc_icm* m, const char* s, int len, int frame) {
  double result = 0.0;
  int i, start, stop = m->model_len - 1;
  if (m->periodicity == 1) frame = 0;
  for (i = 0; i < len && i < m->model_len - 1; i++) {
    result += orc_partial_window_prob(m, i, s, frame);
    frame = (frame + 1) % m->periodicity;
  }
  for (start = 0; stop < len; start++, stop++) {
    result += orc_full_window_prob(m, s + start, frame);
    frame = (frame + 1) % m->periodicity;
  }
  return result;
}

/* ICM_t::Cumulative_Score (ICM/icm.cc:354-405) */
void orc_cumulative_score(const orc_icm* m, const char* s, int len, int frame, double* out) {
  double result = 0.0;
  int i, start, stop;
  if (m->periodicity == 1) frame = 0;
  stop = (m->model_len - 1 < len) ? m->model_len - 1 : len;
  for (i = 0; i < stop; i++) {
    result += orc_partial_window_prob(m, i, s, frame);
    frame = (frame == m->periodicity - 1) ? 0 : frame + 1;
    out[i] = result;
  }
  for (start = 0; i < len; start++, i++) {
    result += orc_full_window_prob(m, s + start, frame);
    frame = (frame == m->periodicity - 1) ? 0 : frame + 1;
    out[i] = result;
  }
}

/* ICM_t::Frame_Score (ICM/icm.cc:485-509): fixed period for every position */
void orc_frame_score(const orc_icm* m, const char* s, int len, int frame, double* out) {
  int i, start;
  int stop = (m->model_len - 1 < len) ? m->model_len - 1 : len;
  for (i = 0; i < stop; i++) out[i] = orc_partial_window_prob(m, i, s, frame);
  for (start = 0; i < len; start++, i++) out[i] = orc_full_window_prob(m, s + start, frame);
}

/* ------------------------------------------------------------------------- */
/* glimmer-mg scoring half */

void orc_default_params(orc_params* p, int metagenomic) {
  memset(p, 0, sizeof *p);
  p->min_gene_len = 75;                    /* glimmer_base.hh:24 */
  p->allow_truncated = metagenomic ? 1 : 0; /* glimmer-mg.cc:12 / glimmer3.cc:23 */
  p->min_indel_orf_len = 15;               /* glimmer_base.cc:40 */
  p->indel_quality_threshold = 18;         /* glimmer-mg.cc:134 */
  p->indel_max = 2;                        /* glimmer-mg.cc:136 */
  p->indel_suffix_score_threshold = -12;   /* glimmer-mg.cc:132 */
  p->ignore_score_len = INT_MAX;
  p->n_start = 3;
  strcpy(p->start_codon[0], "atg");
  strcpy(p->start_codon[1], "gtg");
  strcpy(p->start_codon[2], "ttg");
  p->n_stop = 3;
  strcpy(p->stop_codon[0], "taa");
  strcpy(p->stop_codon[1], "tag");
  strcpy(p->stop_codon[2], "tga");
}

/* Score_All_Frames (glimmer-mg.cc:1468-1510) */
void orc_score_all_frames(const orc_icm* gene, const orc_icm* indep, const char* seq, int len, double* fs) {
  char* buff = (char*)malloc((size_t)len + 1);
  double* g = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  double* n = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  for (int i = 0; i < len; i++) buff[i] = seq[len - 1 - i]; /* Reverse_Transfer */
  buff[len] = 0;
  for (int f = 0; f < 3; f++) {
    orc_frame_score(gene, buff, len, f, g);
    orc_frame_score(indep, buff, len, f, n);
    for (int i = 0; i < len; i++) fs[(size_t)f * len + i] = g[len - 1 - i] - n[len - 1 - i];
  }
  for (int i = 0; i < len; i++) buff[i] = (char)comp_of(seq[i]); /* Complement_Transfer */
  for (int f = 0; f < 3; f++) {
    orc_frame_score(gene, buff, len, f, g);
    orc_frame_score(indep, buff, len, f, n);
    for (int i = 0; i < len; i++) fs[(size_t)(3 + f) * len + i] = g[i] - n[i];
  }
  free(buff);
  free(g);
  free(n);
}

/* exact 3-base codon tests.  After Filter every base is one of acgt, so the 12-bit
 * IUPAC masks of Codon_t::Can_Be / Must_Be (gene.cc:39-95) reduce to string equality
 * once three bases have been shifted in; `nb` counts shifted-in bases. */
static int codon_index(const char c[3], int nb, char list[][4], int n) {
  if (nb < 3) return -1;
  for (int i = 0; i < n; i++)
    if (c[0] == list[i][0] && c[1] == list[i][1] && c[2] == list[i][2]) return i;
  return -1;
}

/* Save_Prev_Stops (glimmer-mg.cc:675-729) */
void orc_save_prev_stops(const char* seq, int len, const orc_params* p, int* fwd_prev, int* rev_next) {
  int last_stops[3] = {0, 1, -1};
  int frame = 0;
  char c[3] = {0, 0, 0};
  orc_params* pp = (orc_params*)p;
  for (int i = 0; i < len; i++) {
    c[0] = c[1]; c[1] = c[2]; c[2] = seq[i];
    if (i >= 2 && codon_index(c, 3, pp->stop_codon, p->n_stop) >= 0) last_stops[frame] = i;
    fwd_prev[i] = last_stops[frame];
    frame = (frame + 1) % 3;
  }
  last_stops[0] = len - 1;
  last_stops[1] = len - 2;
  last_stops[2] = len;
  frame = 0;
  for (int i = len - 1; i >= 0; i--) {
    c[0] = c[1]; c[1] = c[2]; c[2] = (char)comp_of(seq[i]);
    if (i <= len - 3 && codon_index(c, 3, pp->stop_codon, p->n_stop) >= 0) last_stops[frame] = i;
    rev_next[i] = last_stops[frame];
    frame = (frame + 1) % 3;
  }
}

/* Set_Quality_454 (glimmer-mg.cc:1865-1906) */
void orc_set_quality_454(const char* seq, int len, int* q) {
  static const int run_q[6] = {31, 26, 21, 16, 11, 6};
  int run = 0, i;
  char last = ' ';
  if (len <= 0) return;
  for (i = 0; i < len; i++) {
    if (seq[i] != last) {
      if (i > 0) q[i - 1] = run < 6 ? run_q[run] : run_q[5];
      run = 1;
    } else {
      q[i - 1] = 31;
      run++;
    }
    last = seq[i];
  }
  q[i - 1] = run < 6 ? run_q[run] : run_q[5];
}

/* Clean_Quality_454 (glimmer-mg.cc:519-546) */
void orc_clean_quality_454(const char* seq, int len, int* q, int threshold) {
  for (int i = 0; i < len; i++)
    if (q[i] <= 0) q[i] = 1;
  for (int i = 1; i < len; i++)
    if (seq[i] == seq[i - 1] && q[i - 1] < threshold + 1) q[i - 1] = threshold + 1;
}

/* Set_GC_Fraction (glimmer_base.cc:2564-2595) over already-read sequences */
double orc_gc_fraction(const char* const* seqs, const int* lens, int n) {
  unsigned int ct = 0, total = 0;
  for (int s = 0; s < n; s++) {
    total += (unsigned)lens[s];
    for (int j = 0; j < lens[s]; j++) {
      int ch = orc_filter(tolower(seqs[s][j]));
      if (ch == 'g' || ch == 'c') ct++;
    }
  }
  return (double)ct / total;
}

/* Set_Ignore_Score_Len (glimmer_base.cc:2597-2633) */
int orc_ignore_score_len(double gc, const orc_params* p) {
  double lambda = 0.0;
  for (int i = 0; i < p->n_stop; i++) {
    double x = 1.0;
    for (int j = 0; j < 3; j++)
      if (p->stop_codon[i][j] == 'c' || p->stop_codon[i][j] == 'g') x *= gc / 2.0;
      else x *= (1.0 - gc) / 2.0;
    lambda += x;
  }
  return (int)(long)floor(3.0 * log(2.0 * 1000000 * lambda) / lambda);
}

/* ------------------------------------------------------------------------- */
/* Find_Orfs (glimmer_base.cc:638-817) with Do_Fwd_Stop_Codon (:461-503),
 * Do_Rev_Stop_Codon (:505-540), Handle_First_Forward_Stop (:946-981),
 * Handle_First_Reverse_Stop (:985-1008), Handle_Last_Reverse_Stop (:1012-1062),
 * Finish_Orfs (:783-817).  Linear sequences, no ignore regions. */

typedef struct {
  orc_orf* v;
  int n, cap;
} orf_vec;

static void orf_push(orf_vec* L, int stop, int frame, int gene_len, int orf_len) {
  if (L->n == L->cap) {
    L->cap = L->cap ? 2 * L->cap : 64;
    L->v = (orc_orf*)realloc(L->v, sizeof(orc_orf) * (size_t)L->cap);
  }
  orc_orf* o = &L->v[L->n++];
  o->stop_position = stop;
  o->frame = frame;
  o->gene_len = gene_len;
  o->orf_len = orf_len;
}

static int keep_orf(const orc_params* p, int gene_len, int orf_len) {
  return gene_len >= p->min_gene_len ||
         ((p->allow_indels || p->allow_subs) && orf_len >= p->min_indel_orf_len);
}

static void do_fwd_stop(int i, int frame, int prev_fwd_stop[3], int first_fwd_start[3], int first_base,
                        const orc_params* p, orf_vec* L) {
  int gene_len, orf_len;
  if (prev_fwd_stop[frame] == 0) {
    int pos = i - 1, start_pos = first_fwd_start[frame];
    orf_len = pos - first_base;
    orf_len -= orf_len % 3;
    gene_len = (start_pos == INT_MAX) ? 0 : pos - start_pos;
    if (p->allow_truncated && gene_len < p->min_gene_len) gene_len = orf_len;
  } else {
    gene_len = i - first_fwd_start[frame] - 1;
    orf_len = i - prev_fwd_stop[frame] - 4;
  }
  if (keep_orf(p, gene_len, orf_len)) orf_push(L, i - 1, 1 + (frame + 1) % 3, gene_len, orf_len);
  first_fwd_start[frame] = INT_MAX;
  prev_fwd_stop[frame] = i - 1;
}

static void do_rev_stop(int i, int frame, int prev_rev_stop[3], int last_rev_start[3],
                        const orc_params* p, orf_vec* L) {
  int gene_len, orf_len, orf_stop = 0;
  if (prev_rev_stop[frame] == 0) {
    if (!p->allow_truncated) gene_len = 0; /* orf_stop stays 0 (glimmer_base.cc:513,999-1003) */
    else {
      orf_stop = (i - 1) % 3;
      if (orf_stop > 0) orf_stop -= 3;
      gene_len = last_rev_start[frame] - orf_stop;
    }
  } else {
    orf_stop = prev_rev_stop[frame];
    gene_len = last_rev_start[frame] - orf_stop;
  }
  orf_len = i - orf_stop - 4;
  if (keep_orf(p, gene_len, orf_len)) orf_push(L, orf_stop, -1 - (frame + 1) % 3, gene_len, orf_len);
  last_rev_start[frame] = 0;
  prev_rev_stop[frame] = i - 1;
}

int orc_find_orfs(const char* seq, int len, const orc_params* p, orc_orf** out) {
  orf_vec L = {NULL, 0, 0};
  orc_params* pp = (orc_params*)p;
  int first_fwd_start[3] = {INT_MAX, INT_MAX, INT_MAX};
  int last_rev_start[3] = {0, 0, 0}, prev_fwd_stop[3] = {0, 0, 0}, prev_rev_stop[3] = {0, 0, 0};
  int frame = 0, i, n = len, nb = 0;
  char c[3] = {0, 0, 0}, rc[3];
  *out = NULL;
  if (n < p->min_gene_len) return 0;
  for (i = 0; i < n; i++) {
    c[0] = c[1]; c[1] = c[2]; c[2] = seq[i];
    if (nb < 3) nb++;
    /* reverse-strand patterns are the reverse complements of the forward ones */
    rc[0] = (char)comp_of(c[2]); rc[1] = (char)comp_of(c[1]); rc[2] = (char)comp_of(c[0]);
    if (codon_index(c, nb, pp->start_codon, p->n_start) >= 0 && first_fwd_start[frame] == INT_MAX)
      first_fwd_start[frame] = i - 1;
    if (codon_index(rc, nb, pp->start_codon, p->n_start) >= 0) last_rev_start[frame] = i - 1;
    if (codon_index(c, nb, pp->stop_codon, p->n_stop) >= 0)
      do_fwd_stop(i, frame, prev_fwd_stop, first_fwd_start, 1, p, &L);
    if (codon_index(rc, nb, pp->stop_codon, p->n_stop) >= 0)
      do_rev_stop(i, frame, prev_rev_stop, last_rev_start, p, &L);
    frame = (frame == 2) ? 0 : frame + 1;
  }
  /* Finish_Orfs(false, ..., Sequence_Len) */
  for (int fr = 0; fr < 3; fr++) {
    int orf_stop, orf_len, gene_len;
    if (prev_rev_stop[fr] == 0) orf_stop = (fr == 0) ? -1 : (fr == 1 ? 0 : -2);
    else orf_stop = prev_rev_stop[fr];
    orf_len = len - orf_stop - 2;
    orf_len -= orf_len % 3;
    gene_len = (last_rev_start[fr] == 0) ? 0 : last_rev_start[fr] - orf_stop;
    if (p->allow_truncated && gene_len < p->min_gene_len) gene_len = orf_len;
    if (keep_orf(p, gene_len, orf_len)) orf_push(&L, orf_stop, -1 - (fr + 1) % 3, gene_len, orf_len);
  }
  if (p->allow_truncated)
    for (; i < n + 3; i++) {
      do_fwd_stop(i, frame, prev_fwd_stop, first_fwd_start, 1, p, &L);
      frame = (frame == 2) ? 0 : frame + 1;
    }
  *out = L.v;
  return L.n;
}

/* ------------------------------------------------------------------------- */
/* Score_Orf_Starts / Score_Indels / Pass_Stop_Penalty recursion
 * (glimmer-mg.cc:1693-1862, 1513-1602, 961-995) */

typedef struct {
  orc_start* v;
  int n, cap;
} start_vec;

static void start_push(start_vec* S, const orc_start* s) {
  if (S->n == S->cap) {
    S->cap = S->cap ? 2 * S->cap : 256;
    S->v = (orc_start*)realloc(S->v, sizeof(orc_start) * (size_t)S->cap);
  }
  S->v[S->n++] = *s;
}

typedef struct {
  const char* seq;
  int len;
  const double* fs; /* [6][len] */
  const int* fwd_prev;
  const int* rev_next;
  const int* qual;
  const orc_params* p;
} mg_ctx;

typedef struct {
  int n;
  int pos[2], type[2];
} err_list;

static double pass_stop_penalty(const mg_ctx* c, int frame, int lo, int hi) {
  double default_p = 0.999;
  double codon_p[3] = {default_p, default_p, default_p};
  int stop_i[3] = {lo - 3, lo - 2, lo - 1};
  if (frame < 0) {
    stop_i[0] = hi + 1;
    stop_i[1] = hi;
    stop_i[2] = hi - 1;
  }
  if (c->p->have_quality_file)
    for (int k = 0; k < 3; k++) codon_p[k] = 1.0 - pow(10.0, -(double)c->qual[stop_i[k]] / 10.0);
  double p_stop = codon_p[0];
  /* the reference indexes Sequence without bounds checks (glimmer-mg.cc:984-991);
   * out-of-range reads are treated as "not a/t" here. */
  char s1 = (stop_i[1] >= 0 && stop_i[1] < c->len) ? c->seq[stop_i[1]] : 0;
  char s2 = (stop_i[2] >= 0 && stop_i[2] < c->len) ? c->seq[stop_i[2]] : 0;
  if ((frame > 0 && s1 == 'a') || (frame < 0 && s1 == 't')) p_stop *= 2.0 / 3.0 * codon_p[1] + 1.0 / 3.0;
  else p_stop *= codon_p[1];
  if ((frame > 0 && s2 == 'a') || (frame < 0 && s2 == 't')) p_stop *= 2.0 / 3.0 * codon_p[2] + 1.0 / 3.0;
  else p_stop *= codon_p[2];
  return log(1.0 - p_stop) - log(p_stop);
}

static void score_orf_starts(const mg_ctx* c, int frame, start_vec* out, int end_point, double suffix_score,
                             int suffix_j, err_list errors) {
  const orc_params* p = c->p;
  int lo, hi, len, k, orf_is_truncated;
  int L = c->len;

  if (frame > 0) {
    hi = end_point;
    int e = end_point - 1;
    lo = ((e >= 0 && e < L) ? c->fwd_prev[e] : e) + 1; /* Fwd_Prev_Stop (:642) */
    len = hi - lo;
    orf_is_truncated = (lo < 3 && p->allow_truncated);
    k = lo - 1;
  } else {
    lo = end_point;
    int e = end_point - 1;
    hi = ((e >= 0 && e < L) ? c->rev_next[e] : e) + 1; /* Rev_Next_Stop (:1436) */
    len = hi - lo;
    orf_is_truncated = (L - (hi - 1) < 3 && p->allow_truncated);
    k = hi + 1;
  }
  if (len < 0) len = 0; /* cannot happen (prev-stop tables are monotone); guard only */

  /* Cumulative_Frame_Score (glimmer-mg.cc:561-604); indep_score is identically 0 */
  double* score = (double*)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
  {
    double cum = 0;
    int f = 1;
    int si = (frame > 0) ? hi - 1 : lo - 1;
    for (int i = 0; i < len; i++) {
      score[i] = cum + c->fs[(size_t)((frame > 0 ? 0 : 3) + f) * L + si];
      cum = score[i];
      si += (frame > 0) ? -1 : 1;
      f = (f == 2) ? 0 : f + 1;
    }
  }

  /* substitution through the previous stop (glimmer-mg.cc:1771-1806) */
  if (p->allow_subs && errors.n < 1) {
    int error_end_point, error_pos;
    if (frame > 0) {
      error_end_point = lo - 3;
      error_pos = lo - 2;
    } else {
      error_end_point = hi + 3;
      error_pos = hi + 2;
    }
    if (error_end_point >= 0 && error_end_point - 2 < L) {
      int error_suffix_j = suffix_j + len;
      double error_suffix_score = suffix_score + pass_stop_penalty(c, frame, lo, hi);
      if (len > 0) error_suffix_score += score[len - 1] - 0.0;
      err_list e2 = errors;
      e2.pos[e2.n] = error_pos;
      e2.type[e2.n] = 2;
      e2.n++;
      score_orf_starts(c, frame, out, error_end_point, error_suffix_score, error_suffix_j, e2);
    }
  }

  /* find starts (glimmer-mg.cc:1811-1861) */
  int m = len;
  int lowest_j = (3 < p->min_gene_len - 3) ? 3 : p->min_gene_len - 3;
  int first_pos = 0, nb = 0;
  char cod[3] = {0, 0, 0};
  orc_params* pp = (orc_params*)p;
  for (int j = m - 1; j >= lowest_j; j--) {
    int bidx = (frame > 0) ? hi - 1 - j : lo - 1 + j;
    if (p->allow_indels && c->qual[bidx] <= p->indel_quality_threshold && errors.n < p->indel_max) {
      /* Score_Indels (glimmer-mg.cc:1513-1602) */
      int q = c->qual[bidx];
      double prob_err = pow(10.0, -(double)q / 10.0);
      double score_penalty = log(prob_err / 2.0) - log(1.0 - prob_err);
      int esj = suffix_j + j + 2 - (j % 3);
      double ess = suffix_score + score[j] - 0.0 + score_penalty;
      if (ess > p->indel_suffix_score_threshold) { /* deletion */
        err_list e2 = errors;
        e2.pos[e2.n] = (frame > 0) ? k + 3 : k - 1;
        e2.type[e2.n] = 1;
        e2.n++;
        score_orf_starts(c, frame, out, (frame > 0) ? k + (j % 3) : k - (j % 3), ess, esj, e2);
      }
      ess = suffix_score + score[j - 1] - 0.0 + score_penalty;
      if (ess > p->indel_suffix_score_threshold) { /* insertion */
        err_list e2 = errors;
        e2.pos[e2.n] = (frame > 0) ? k + 2 : k - 2;
        e2.type[e2.n] = 0;
        e2.n++;
        score_orf_starts(c, frame, out, (frame > 0) ? k - (2 - (j % 3)) : k + 2 - (j % 3), ess, esj, e2);
      }
    }
    cod[0] = cod[1]; cod[1] = cod[2];
    cod[2] = (frame > 0) ? c->seq[bidx] : (char)comp_of(c->seq[bidx]);
    if (nb < 3) nb++;
    if (j % 3 == 0) {
      int which = codon_index(cod, nb, pp->start_codon, p->n_start);
      if ((which >= 0 || (first_pos == 0 && orf_is_truncated)) && j + 3 + suffix_j >= p->min_gene_len) {
        orc_start st;
        memset(&st, 0, sizeof st);
        double next_s = score[j - 1] - 0.0;
        st.j = j + 2 + suffix_j;
        st.pos = k;
        st.score = next_s + suffix_score;
        st.first = (first_pos == 0);
        st.n_err = errors.n;
        for (int e = 0; e < errors.n; e++) {
          st.err_pos[e] = errors.pos[e];
          st.err_type[e] = errors.type[e];
        }
        if (which >= 0 && first_pos == 0 && orf_is_truncated) {
          st.which = -1;
          st.truncated = 1;
          start_push(out, &st);
          st.first = 0;
        }
        st.which = which;
        st.truncated = (which < 0);
        start_push(out, &st);
        if (first_pos == 0) first_pos = k;
      }
    }
    if (frame > 0) k++;
    else k--;
  }
  free(score);
}

/* Score_Orfs_Errors up to the boost (glimmer-mg.cc:1605-1651) */
int orc_mg_score_orfs(const orc_icm* gene, const orc_icm* indep, const char* seq, int len, const int* qual,
                      const orc_params* p, const orc_orf* orfs, int n_orf, int* start_off, orc_start** starts) {
  mg_ctx c;
  start_vec S = {NULL, 0, 0};
  double* fs = (double*)malloc(sizeof(double) * 6 * (size_t)(len > 0 ? len : 1));
  int* fwd_prev = (int*)malloc(sizeof(int) * (size_t)(len > 0 ? len : 1));
  int* rev_next = (int*)malloc(sizeof(int) * (size_t)(len > 0 ? len : 1));
  int* q = (int*)malloc(sizeof(int) * (size_t)(len > 0 ? len : 1));
  orc_score_all_frames(gene, indep, seq, len, fs);
  orc_save_prev_stops(seq, len, p, fwd_prev, rev_next);
  if (qual) {
    memcpy(q, qual, sizeof(int) * (size_t)len);
    if (p->allow_indels) orc_clean_quality_454(seq, len, q, p->indel_quality_threshold);
  } else if (p->allow_indels) {
    orc_set_quality_454(seq, len, q);
  } else {
    for (int i = 0; i < len; i++) q[i] = 31;
  }
  c.seq = seq; c.len = len; c.fs = fs; c.fwd_prev = fwd_prev; c.rev_next = rev_next; c.qual = q; c.p = p;
  for (int i = 0; i < n_orf; i++) {
    err_list none;
    memset(&none, 0, sizeof none);
    start_off[i] = S.n;
    int end_point = (orfs[i].frame > 0) ? orfs[i].stop_position - 1 : orfs[i].stop_position + 3;
    score_orf_starts(&c, orfs[i].frame, &S, end_point, 0, 0, none);
    for (int s = start_off[i]; s < S.n; s++) /* boost long ORFs (:1649-1651) */
      if (S.v[s].j > p->ignore_score_len && 0.0 > S.v[s].score) S.v[s].score = 0.0; /* Max(0.0, score) */
  }
  start_off[n_orf] = S.n;
  *starts = S.v;
  free(fs); free(fwd_prev); free(rev_next); free(q);
  return S.n;
}

/* ------------------------------------------------------------------------- */
/* glimmer3 Score_Orfs start enumeration (glimmer3.cc:1275-1466) */

int orc_g3_score_orfs(const orc_icm* gene, const orc_icm* indep, const char* seq, int len, const orc_params* p,
                      const orc_orf* orfs, int n_orf, int* start_off, orc_start** starts) {
  start_vec S = {NULL, 0, 0};
  orc_params* pp = (orc_params*)p;
  int cap = 1024;
  char* buff = (char*)malloc((size_t)cap + 1);
  double* score = (double*)malloc(sizeof(double) * (size_t)cap);
  double* indep_score = (double*)malloc(sizeof(double) * (size_t)cap);
  for (int i = 0; i < n_orf; i++) {
    int frame = orfs[i].frame, olen = orfs[i].orf_len, lo, hi, k, orf_is_truncated;
    start_off[i] = S.n;
    if (olen > cap) {
      cap = 2 * olen;
      buff = (char*)realloc(buff, (size_t)cap + 1);
      score = (double*)realloc(score, sizeof(double) * (size_t)cap);
      indep_score = (double*)realloc(indep_score, sizeof(double) * (size_t)cap);
    }
    if (frame > 0) {
      hi = orfs[i].stop_position - 1;
      if (hi <= 0) hi += len;
      lo = hi - olen;
      /* Reverse_Transfer (glimmer_base.cc:2505-2530), with its wraparound */
      int start = hi - 1;
      for (int j = 0; j < olen; j++, start--) {
        buff[j] = seq[start];
        if (start <= 0) start += len;
      }
      orf_is_truncated = (lo < 3 && p->allow_truncated);
      k = orfs[i].stop_position - olen - 2;
    } else {
      lo = orfs[i].stop_position + 2;
      if (lo >= len) lo -= len;
      hi = lo + olen;
      /* Complement_Transfer (glimmer_base.cc:410-432) */
      int start = lo;
      for (int j = 0; j < olen; j++, start++) {
        if (start >= len) start -= len;
        buff[j] = (char)comp_of(seq[start]);
      }
      orf_is_truncated = (len - hi < 3 && p->allow_truncated);
      k = orfs[i].stop_position + olen + 4;
    }
    buff[olen > 0 ? olen : 0] = 0;
    orc_cumulative_score(gene, buff, olen, 1, score);
    orc_cumulative_score(indep, buff, olen, 1, indep_score);
    int m = olen, first_pos = 0, nb = 0;
    int lowest_j = (3 < p->min_gene_len - 3) ? 3 : p->min_gene_len - 3;
    char cod[3] = {0, 0, 0};
    for (int j = m - 1; j >= lowest_j; j--) {
      cod[0] = cod[1]; cod[1] = cod[2]; cod[2] = buff[j];
      if (nb < 3) nb++;
      if (j % 3 == 0) {
        int which = codon_index(cod, nb, pp->start_codon, p->n_start);
        if ((which >= 0 || (first_pos == 0 && orf_is_truncated)) && j + 3 >= p->min_gene_len) {
          orc_start st;
          memset(&st, 0, sizeof st);
          st.j = j + 2;
          st.pos = k;
          st.score = score[j - 1] - indep_score[j - 1];
          st.first = (first_pos == 0);
          if (which >= 0 && first_pos == 0 && orf_is_truncated) {
            st.which = -1;
            st.truncated = 1;
            start_push(&S, &st);
            st.first = 0;
          }
          st.which = which;
          st.truncated = (which < 0);
          start_push(&S, &st);
          if (first_pos == 0) first_pos = k;
        }
      }
      if (frame > 0) k++;
      else k--;
    }
    for (int s = start_off[i]; s < S.n; s++)
      if (S.v[s].j > p->ignore_score_len && 0.0 > S.v[s].score) S.v[s].score = 0.0; /* Max(0.0, score) */
  }
  start_off[n_orf] = S.n;
  *starts = S.v;
  free(buff); free(score); free(indep_score);
  return S.n;
}

/* ------------------------------------------------------------------------- */
/* training (ICM/icm.cc:1010-1463, 1841-1954) */

static const float CHI2_VAL[7] = {2.37f, 4.11f, 6.25f, 7.81f, 9.35f, 11.3f, 12.8f};           /* icm.hh:36-37 */
static const float CHI2_SIGNIFICANCE[7] = {0.50f, 0.75f, 0.90f, 0.95f, 0.975f, 0.99f, 0.995f}; /* icm.hh:39-40 */
#define MUT_INFO_BIAS 0.03
#define MUT_INFO_EPSILON 1e-4
#define PSEUDO_COUNT 0.001
#define SAMPLE_SIZE_BOUND 400

/* Get_Mutual_Info (ICM/icm.cc:1900-1954), n = 4 */
static double mutual_info(const int ct[16], int sum) {
  double mi = 0.0, left[4] = {0, 0, 0, 0}, right[4] = {0, 0, 0, 0};
  if (sum == 0) return 0.0;
  for (int i = 0, k = 0; i < 4; i++)
    for (int j = 0; j < 4; j++, k++) {
      left[i] += ct[k];
      right[j] += ct[k];
    }
  for (int i = 0; i < 4; i++) {
    left[i] /= sum;
    right[i] /= sum;
  }
  for (int i = 0, k = 0; i < 4; i++)
    for (int j = 0; j < 4; j++, k++) {
      double prob = (double)ct[k] / sum;
      if (prob != 0.0 && left[i] != 0.0 && right[j] != 0.0) mi += prob * log(prob / (left[i] * right[j]));
    }
  return mi;
}

/* counts[(f*num_nodes + node)*(w-1) + pos][16] */
#define CNT(m, cnt, f, node, pos) ((cnt) + 16 * ((((size_t)(f) * (m)->num_nodes + (node)) * ((m)->model_len - 1)) + (pos)))

/* Count_Char_Pairs (icm.cc:1841-1870) for the root of period `frame` (icm.cc:1377-1399) */
static void count_root(const orc_icm* m, int* counts, const char* s, int frame) {
  int w = m->model_len, period = m->periodicity;
  int offset = frame - (w % period);
  if (offset < 0) offset += period;
  int end = (int)strlen(s);
  if (offset > end) return; /* data[i] + offset past the NUL would be UB in the reference */
  s += offset;
  end -= offset;
  for (int start = 0, stop = w - 1; stop < end; start += period, stop += period) {
    int last = sub_of(s[stop]);
    for (int i = 0; i < w - 1; i++) CNT(m, counts, frame, 0, i)[ALPHA * sub_of(s[start + i]) + last]++;
  }
}

/* Count_Char_Pairs_Restricted + Get_Training_Node (icm.cc:1190-1256) */
static void count_restricted(const orc_icm* m, int* counts, const char* s, int level) {
  int w = m->model_len, period = m->periodicity;
  int end = (int)strlen(s), frame = w % period;
  for (int start = 0, stop = w - 1; stop < end; start++, stop++) {
    int node = 0, ok = 1;
    for (int i = 0; i < level; i++) {
      int j = MIP(m, frame, node);
      if (j < 0) {
        ok = 0;
        break;
      }
      node = node * ALPHA + sub_of(s[start + j]) + 1;
    }
    if (ok) {
      int last = sub_of(s[stop]);
      for (int i = 0; i < w - 1; i++) CNT(m, counts, frame, node, i)[ALPHA * sub_of(s[start + i]) + last]++;
    }
    frame++;
    if (frame == period) frame = 0;
  }
}

void orc_count_level(const orc_icm* m, const char* const* strings, int n, int level, int* counts) {
  if (level == 0) {
    for (int f = 0; f < m->periodicity; f++)
      for (int i = 0; i < n; i++) count_root(m, counts, strings[i], f);
  } else {
    for (int i = 0; i < n; i++) count_restricted(m, counts, strings[i], level);
  }
}

/* Interpolate_Probs (icm.cc:1260-1330).  prob[] are floats; every store rounds. */
static void interpolate_probs(orc_icm* m, int frame, int sub, const int ct[4]) {
  int parent = PARENT(sub);
  float* pr = PROB(m, frame, sub);
  const float* pp = PROB(m, frame, parent);
  double total = 0.0;
  for (int i = 0; i < 4; i++) total += ct[i];
  for (int i = 0; i < 4; i++) pr[i] = (float)((ct[i] + PSEUDO_COUNT * pp[i]) / (total + PSEUDO_COUNT));
  if (total >= SAMPLE_SIZE_BOUND) return;
  double chi2 = 0.0;
  for (int i = 0; i < 4; i++) {
    double expected = total * pp[i];
    if (expected > 0.0) chi2 += pow(ct[i] - expected, 2.0) / expected;
  }
  int i;
  for (i = 0; i < 7 && CHI2_VAL[i] < chi2; i++)
    ;
  double lambda;
  if (i == 0) lambda = 0.0;
  else if (i == 7) lambda = 1.0;
  else
    lambda = CHI2_SIGNIFICANCE[i - 1] +
             ((chi2 - CHI2_VAL[i - 1]) / (CHI2_VAL[i] - CHI2_VAL[i - 1])) *
                 (CHI2_SIGNIFICANCE[i] - CHI2_SIGNIFICANCE[i - 1]);
  lambda *= total / SAMPLE_SIZE_BOUND;
  if (lambda > 1.0) lambda = 1.0;
  for (int k = 0; k < 4; k++) {
    pr[k] = (float)(pr[k] * lambda);
    pr[k] = (float)(pr[k] + (1.0 - lambda) * pp[k]);
  }
}

/* Train_Model + Complete_Tree + Take_Logs (icm.cc:1356-1463, 1061-1186, 1334-1352) */
orc_icm* orc_icm_train(const char* const* strings, int n, int w, int d, int p) {
  orc_icm* m = orc_icm_new(w, d, p);
  size_t ncnt = (size_t)p * m->num_nodes * (w - 1) * 16;
  int* counts = (int*)calloc(ncnt ? ncnt : 1, sizeof(int));
  /* model_depth == 0 (Count_Single_Chars) is not restated: no driver uses it */
  for (int frame = 0; frame < p; frame++) {
    int final_ct[4] = {0, 0, 0, 0}, sum = 0;
    for (int i = 0; i < n; i++) count_root(m, counts, strings[i], frame);
    const int* c0 = CNT(m, counts, frame, 0, 0);
    for (int i = 0, k = 0; i < 4; i++)
      for (int j = 0; j < 4; j++, k++) {
        sum += c0[k];
        final_ct[j] += c0[k];
      }
    /* float arithmetic: (int + float) / float (icm.cc:1411-1413) */
    for (int j = 0; j < 4; j++)
      PROB(m, frame, 0)[j] = ((float)final_ct[j] + (float)(PSEUDO_COUNT / ALPHA)) / (float)(sum + PSEUDO_COUNT);
    int max_pos = 0;
    double best = mutual_info(CNT(m, counts, frame, 0, 0), sum);
    for (int i = 1; i < w - 1; i++) {
      double next = mutual_info(CNT(m, counts, frame, 0, i), sum);
      if (next >= best) {
        best = next;
        max_pos = i;
      } else if (next >= best / (1.0 + MUT_INFO_BIAS))
        max_pos = i;
    }
    MIP(m, frame, 0) = (short)max_pos;
  }
  /* Complete_Tree */
  int first_node = 1, nodes_on_level = ALPHA;
  for (int level = 1; level <= d; level++) {
    for (int i = 0; i < n; i++) count_restricted(m, counts, strings[i], level);
    int last_node = first_node + nodes_on_level - 1;
    for (int frame = 0; frame < p; frame++)
      for (int sub = first_node; sub <= last_node; sub++) {
        int final_ct[4] = {0, 0, 0, 0}, sum = 0;
        if (MIP(m, frame, PARENT(sub)) < 0) {
          MIP(m, frame, sub) = -2;
          continue;
        }
        const int* c0 = CNT(m, counts, frame, sub, 0);
        for (int i = 0, k = 0; i < 4; i++)
          for (int j = 0; j < 4; j++, k++) {
            sum += c0[k];
            final_ct[j] += c0[k];
          }
        int max_pos = 0;
        double best = mutual_info(c0, sum);
        for (int i = 1; i < w - 1; i++) {
          double next = mutual_info(CNT(m, counts, frame, sub, i), sum);
          if (next >= best) {
            best = next;
            max_pos = i;
          } else if (next >= best / (1.0 + MUT_INFO_BIAS))
            max_pos = i;
        }
        if (best <= MUT_INFO_EPSILON && sum < SAMPLE_SIZE_BOUND) max_pos = -1;
        MIP(m, frame, sub) = (short)max_pos;
        interpolate_probs(m, frame, sub, final_ct);
      }
    first_node = last_node + 1;
    nodes_on_level *= ALPHA;
  }
  /* Take_Logs: float argument binds to logf under libstdc++ (icm.cc:1345-1347) */
  for (size_t i = 0; i < (size_t)p * m->num_nodes * 4; i++)
    m->prob[i] = (m->prob[i] > 0.0f) ? logf(m->prob[i]) : -FLT_MAX;
  free(counts);
  return m;
}
