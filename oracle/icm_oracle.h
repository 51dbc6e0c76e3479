/* oracle/icm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's (davek44/Glimmer-MG) ICM scoring and
 * training hot path.  It exists to CHECK the CUDA implementation: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (glimmer_mg_b200/) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement against
 *   - the reference's own golden vectors (5 994 Score_String values,
 *     NC_000915.run1 models / predictions), and
 *   - the unmodified reference compiled into oracle/_ref (libref_icm.so,
 *     glimmer-mg-dump), bit for bit.
 *
 * All file:line citations are relative to /root/reference/src/.
 */
#ifndef ICM_ORACLE_H
#define ICM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ICM_t (ICM/icm.hh:116-129) with the per-node tables flattened [period][node]. */
typedef struct {
  int model_len, model_depth, periodicity, num_nodes;
  short* mip;  /* mut_info_pos, [periodicity][num_nodes] */
  float* prob; /* natural-log probs, [periodicity][num_nodes][4] */
} orc_icm;

/* Orf_t (Common/gene.hh:101-136) */
typedef struct {
  int frame;         /* +-(1..3) */
  int stop_position; /* 1-based lowest coordinate of the stop codon */
  int orf_len, gene_len;
} orc_orf;

/* Start_t (Glimmer/glimmer_base.hh:80-88) with the error vector flattened (<= 2). */
typedef struct {
  int j, pos;
  double score;
  int which, truncated, first;
  int n_err;
  int err_pos[2];
  int err_type[2]; /* 0 insertion, 1 deletion, 2 substitution */
} orc_start;

/* scoring-half parameters (defaults of glimmer-mg.cc / glimmer_base.hh) */
typedef struct {
  int min_gene_len;        /* 75 */
  int allow_truncated;     /* mg: 1, glimmer3: 0 */
  int allow_indels;        /* -i */
  int allow_subs;          /* -s */
  int min_indel_orf_len;   /* 15 */
  int indel_quality_threshold; /* 18 */
  int indel_max;           /* 2 */
  double indel_suffix_score_threshold; /* -12 */
  int ignore_score_len;    /* Set_Ignore_Score_Len or INT_MAX */
  int have_quality_file;   /* Quality_File_Name != NULL */
  int n_start, n_stop;
  char start_codon[8][4];
  char stop_codon[8][4];
} orc_params;

void orc_default_params(orc_params* p, int metagenomic);

/* model container + I/O (ICM/icm.cc:614-803, 961-998) */
orc_icm* orc_icm_new(int w, int d, int p);
orc_icm* orc_icm_read(const char* path);
int orc_icm_write(const orc_icm* m, const char* path);
void orc_icm_free(orc_icm* m);
orc_icm* orc_build_indep_wo_stops(double gc, const char* const* stops, int n_stops);

/* scalar scoring ops (ICM/icm.cc:354-405, 485-509, 557-610, 807-842, 864-903) */
double orc_full_window_prob(const orc_icm* m, const char* w, int frame);
double orc_partial_window_prob(const orc_icm* m, int predict_pos, const char* s, int frame);
double orc_score_string(const orc_icm* m, const char* s, int len, int frame);
void orc_cumulative_score(const orc_icm* m, const char* s, int len, int frame, double* out);
void orc_frame_score(const orc_icm* m, const char* s, int len, int frame, double* out);

/* glimmer-mg scoring half (Glimmer/glimmer-mg.cc) */
void orc_score_all_frames(const orc_icm* gene, const orc_icm* indep, const char* seq, int len,
                          double* fs /* [6][len] */);
void orc_save_prev_stops(const char* seq, int len, const orc_params* p, int* fwd_prev, int* rev_next);
void orc_set_quality_454(const char* seq, int len, int* q);
void orc_clean_quality_454(const char* seq, int len, int* q, int threshold);
double orc_gc_fraction(const char* const* seqs, const int* lens, int n);
int orc_ignore_score_len(double gc, const orc_params* p);

/* Find_Orfs (Glimmer/glimmer_base.cc:638-817), linear genomes, no ignore regions.
 * Returns number of ORFs; *out is malloc'd. */
int orc_find_orfs(const char* seq, int len, const orc_params* p, orc_orf** out);

/* Score_Orfs_Errors up to (not including) the filter/Add_Events stage
 * (glimmer-mg.cc:1605-1651): for every ORF the raw start_list in generation order,
 * long-ORF boost applied.  start_off[n_orf+1] are offsets into *starts (malloc'd).
 * qual may be NULL (synthesised with Set_Quality_454 when allow_indels). */
int orc_mg_score_orfs(const orc_icm* gene, const orc_icm* indep, const char* seq, int len,
                      const int* qual, const orc_params* p, const orc_orf* orfs, int n_orf,
                      int* start_off, orc_start** starts);

/* glimmer3 Score_Orfs start enumeration (glimmer3.cc:1275-1466), boost applied. */
int orc_g3_score_orfs(const orc_icm* gene, const orc_icm* indep, const char* seq, int len,
                      const orc_params* p, const orc_orf* orfs, int n_orf,
                      int* start_off, orc_start** starts);

/* training (ICM/icm.cc:1010-1463, 1841-1954); strings as given to Train_Model
 * (already lower-cased and, for build-icm -r, reversed). */
orc_icm* orc_icm_train(const char* const* strings, int n, int w, int d, int p);

/* context counts only: count[period][node][w-1][16] for the nodes of `level`
 * given the mut_info_pos chosen so far (icm.cc:1190-1256, 1841-1870).  level 0 =
 * root counts.  Used to check the K4 histogram kernel in isolation. */
void orc_count_level(const orc_icm* m, const char* const* strings, int n, int level, int* counts);

#ifdef __cplusplus
}
#endif
#endif
