// glimmer_mg_b200/csrc/gmg_mg_flat.cuh -- K3 (glimmer-mg), FLAT form: the Score_Orf_Starts / Score_Indels /
// Pass_Stop_Penalty recursion (/root/reference/src/Glimmer/glimmer-mg.cc:1693-1862, 1513-1602, 961-995) without a
// recursion.
//
// The reference's call tree has at most three levels (errors so far = 0, 1, 2: Indel_Max = 2, one substitution at
// most and only from the root).  Every call is a pure function of (end point, suffix score, suffix length, error
// list), its children sit at the call's "gate" positions (quality <= Indel_Quality_Threshold; two candidates each:
// deletion, insertion) plus one substitution candidate in front, and its own start records are start-codon bits of
// one codon-bitmap stream.  So the tree is enumerated level by level with one thread per CANDIDATE:
//
//   pass A  per ORF            root call, its candidate count n1, its own record count
//   pass B  per level-1 cand.  exists? (suffix score > Indel_Suffix_Score_Threshold), geometry, n2, own records
//   pass C  per level-2 cand.  exists? -> record count of the leaf call
//   pass D  per level-1 cand.  records of the whole subtree
//   pass E  per ORF            records of the ORF (CSR offsets after a scan)
//   write   per ORF / level-1 / level-2 candidate: own records at their place in the reference's generation order
//
// with exclusive scans in between.  The reference's order -- children of position j (deletion, insertion) before
// the position's own start, positions descending, a call's substitution child first -- makes every subtree one
// contiguous range of the output, so all places are differences of the scans plus popcounts.
//
// Everything here is __host__ __device__: gmg_score.cu wraps each pass in a kernel, and tests/mgflat_host_check.cu
// (test infrastructure) runs the very same functions on the host against the CPU checker.
#pragma once

#include <stdint.h>

#include "../../include/gmg_icm.h"

#if defined(__CUDACC__)
#define MGF_HD __host__ __device__ __forceinline__
#else
#define MGF_HD inline
#endif

#if !defined(__CUDACC__)
struct uint2 {
  unsigned x, y;
};
#endif

struct CodonSets {
  unsigned long long start_mask;  // bit c set: 6-bit codon c is a start codon
  unsigned long long stop_mask;
  // the same sets over the RAW packed code  b(c) | b(c+1) << 2 | b(c+2) << 4  of the bases at c, c+1, c+2:
  // [0] forward start, [1] forward stop, [2] reverse-strand start, [3] reverse-strand stop (k_codon_bits)
  unsigned long long raw_mask[4];
  unsigned char which[64];        // index of the first matching start codon (Can_Be order)
};

struct DevParams {
  int min_gene_len, allow_truncated, allow_indels, allow_subs, min_indel_orf_len, indel_q_thresh, indel_max,
      ignore_score_len, have_quality_file;
  double indel_suffix_thresh;
};

template <class T>
MGF_HD T mgf_ld(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
MGF_HD int mgf_popc(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
MGF_HD int mgf_ffs(unsigned x) {  // 1-based index of the lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}
MGF_HD int mgf_clz(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
MGF_HD int mgf_mod3(int x) {
  int r = x % 3;
  return r < 0 ? r + 3 : r;
}

// One batch: the packed sequences and what K2 made of them.
struct MgfBatch {
  const uint64_t* words;      // 2-bit bases, base p at bits 2 (p % 32) of word p / 32 (zero padding either side)
  const int64_t* off;         // [n_seq + 1]
  int64_t total;
  const double* cum;          // [6][total]: forward classes suffix sums, reverse classes prefix sums (K2)
  const int32_t* fwd_prev;    // [total]
  const int32_t* rev_next;    // [total]
  const uint8_t* qual;        // [total] (NULL when neither indels nor a quality file are in play)
  const uint8_t* cert;        // [n_seq] K2's exactness certificate; sequences with 0 take the ordered kernel
  const uint2* cb;            // codon bitmaps [6][nwc] {start bits, stop bits}
  int64_t nwc;
  const uint32_t* gate_bits;  // [total / 32 + 1] bit i of word w: quality of base 32 w + i <= Indel_Quality_Threshold
  const uint32_t* gate_rank;  // [total / 32 + 2] gates in all earlier words
  const uint32_t* gate_pos;   // [n_gates] global base index of every gate, ascending
  const double* tables;       // [516]: 0..255 indel penalty by quality, 256..259 stop penalty, 260..515 1 - 10^(-q/10)
  const double* sub_pen;      // [n_orfs] Pass_Stop_Penalty of every root when a quality file is in use (host log), else NULL
};

struct MgfSeq {
  int64_t a;
  int L;
};

// A call of Score_Orf_Starts that exists.  64 bytes.
struct MgfCall {
  int32_t lo, hi;        // the reference's lo / hi of this call
  int32_t suffix_j;
  int32_t n_gates;       // gates in the call's j range if the call may branch, else 0
  uint32_t gate0;        // index into gate_pos of the call's first gate in j-descending order
  int32_t err_pos;       // the error that opened this call (levels 1, 2 keep their parent's in the parent)
  int8_t err_type, n_err, ok, fwd;
  int32_t orf;           // ORF this call belongs to
  int32_t jpar;          // position j of the parent call at which this call was opened (-1: substitution child / root)
  double suffix_score;
  double cbase;          // value subtracted from the prefix-sum row to get score[] of this call
  int64_t base;          // place of the call's first record in the output (filled by the write passes)
};

MGF_HD int mgf_base_at(const uint64_t* words, int64_t p) { return (int)((mgf_ld(words + (p >> 5)) >> ((p & 31) * 2)) & 3); }

// 6-bit code (b0*16 + b1*4 + b2) of the forward-strand codon whose three bases start at global index g
MGF_HD int mgf_codon6_at(const uint64_t* words, int64_t g) {
  const int64_t w = g >> 5;
  const int sh = (int)(g & 31) * 2;
  uint64_t v = mgf_ld(words + w) >> sh;
  if (sh > 58) v |= mgf_ld(words + w + 1) << (64 - sh);
  const int raw = (int)(v & 63);
  return ((raw & 3) << 4) | (raw & 12) | (raw >> 4);
}

// ---- geometry of a call (glimmer-mg.cc:1719-1758) -------------------------------------------------------------
MGF_HD void mgf_open(const MgfBatch& B, const MgfSeq& S, bool fwd, int end_point, int* lo, int* hi) {
  const int e = end_point - 1;
  const bool in = e >= 0 && e < S.L;
  if (fwd) {
    *hi = end_point;
    *lo = (in ? mgf_ld(B.fwd_prev + S.a + e) : e) + 1;
  } else {
    *lo = end_point;
    *hi = (in ? mgf_ld(B.rev_next + S.a + e) : e) + 1;
  }
}
MGF_HD const double* mgf_row(const MgfBatch& B, const MgfSeq& S, bool fwd, int lo, int hi) {
  return fwd ? B.cum + (size_t)mgf_mod3(hi) * (size_t)B.total + S.a : B.cum + (size_t)(3 + mgf_mod3(lo - 1)) * (size_t)B.total + S.a;
}
MGF_HD double mgf_cbase(const MgfBatch& B, const MgfSeq& S, bool fwd, int lo, int hi) {
  const double* row = mgf_row(B, S, fwd, lo, hi);
  const int q = fwd ? hi : lo - 2;
  return ((unsigned)q < (unsigned)S.L) ? mgf_ld(row + q) : 0.0;
}
// score[j] of a call (Cumulative_Frame_Score, glimmer-mg.cc:561-604) as a difference of two prefix-sum entries
MGF_HD double mgf_score(const MgfSeq& S, bool fwd, const double* row, int lo, int hi, double cbase, int j) {
  const int q = fwd ? hi - 1 - j : lo - 1 + j;
  return (((unsigned)q < (unsigned)S.L) ? mgf_ld(row + q) : 0.0) - cbase;
}

// ---- gates: positions where Score_Indels may branch -----------------------------------------------------------
MGF_HD uint32_t mgf_gate_rank(const MgfBatch& B, int64_t p) {  // gates at global positions < p
  const int64_t w = p >> 5;
  const unsigned m = mgf_ld(B.gate_bits + w) & ((1u << (int)(p & 31)) - 1u);
  return mgf_ld(B.gate_rank + w) + (uint32_t)mgf_popc(m);
}
// the call's gate range: sequence positions [b_lo, b_hi] of j = m-1 .. lowest_j, clamped to the sequence
MGF_HD void mgf_gates(const MgfBatch& B, const MgfSeq& S, bool fwd, int lo, int hi, int lowest_j, int* n_gates,
                      uint32_t* gate0) {
  int m = hi - lo;
  if (m < 0) m = 0;
  *n_gates = 0;
  *gate0 = 0;
  if (m - 1 < lowest_j) return;
  int b_lo = fwd ? hi - 1 - (m - 1) : lo - 1 + lowest_j;
  int b_hi = fwd ? hi - 1 - lowest_j : lo - 1 + (m - 1);
  if (b_lo < 0) b_lo = 0;
  if (b_hi > S.L - 1) b_hi = S.L - 1;
  if (b_hi < b_lo) return;
  const uint32_t r_lo = mgf_gate_rank(B, S.a + b_lo), r_hi = mgf_gate_rank(B, S.a + b_hi + 1);
  *n_gates = (int)(r_hi - r_lo);
  *gate0 = fwd ? r_lo : r_hi - 1u;
}
// number of the call's gates at positions j' >= j
MGF_HD int mgf_gates_at_or_above(const MgfBatch& B, const MgfSeq& S, const MgfCall& c, int j) {
  if (c.n_gates == 0) return 0;
  const bool fwd = c.fwd != 0;
  int q = fwd ? c.hi - 1 - j : c.lo - 1 + j;  // sequence position of j
  int n;
  if (fwd) {
    if (q < 0) return 0;
    if (q > S.L - 1) q = S.L - 1;
    n = (int)(mgf_gate_rank(B, S.a + q + 1) - c.gate0);
  } else {
    if (q > S.L - 1) return 0;
    if (q < 0) q = 0;
    n = (int)(c.gate0 + 1u - mgf_gate_rank(B, S.a + q));
  }
  return n < 0 ? 0 : (n > c.n_gates ? c.n_gates : n);
}

// ---- own start records of a call (glimmer-mg.cc:1809-1861) ----------------------------------------------------
struct MgfOwn {
  int lo, hi, m, trunc;
  int j_lo, j_hi, j_hs;  // eligible j (multiples of 3): j_lo..j_hi; start-codon test only for j <= j_hs
  const uint2* st;       // bitmap stream
  uint32_t cpos;         // forward: slot(j) = (cpos - j) / 3; reverse: slot(j) = (cpos + j) / 3 (batches < 2^32 bases)
};
MGF_HD void mgf_own_open(const MgfBatch& B, const MgfSeq& S, const DevParams& P, bool fwd, int lo, int hi, int suffix_j,
                         MgfOwn& f) {
  const int lowest_j = P.min_gene_len - 3 < 3 ? P.min_gene_len - 3 : 3;
  f.lo = lo;
  f.hi = hi;
  f.m = hi - lo;
  if (f.m < 0) f.m = 0;
  f.trunc = fwd ? (lo < 3 && P.allow_truncated) : (S.L - (hi - 1) < 3 && P.allow_truncated);
  int jl = lowest_j > P.min_gene_len - 3 - suffix_j ? lowest_j : P.min_gene_len - 3 - suffix_j;
  if (jl < 0) jl = 0;
  f.j_lo = jl + (3 - jl % 3) % 3;
  f.j_hi = (f.m - 1) >= 0 ? (f.m - 1) - (f.m - 1) % 3 : -3;
  f.j_hs = (f.m - 3) >= 0 ? (f.m - 3) - (f.m - 3) % 3 : -3;
  if (fwd) {
    f.cpos = (uint32_t)(S.a + hi - 3);  // first base of the codon ending at hi-1-j is cpos - j
    f.st = B.cb + (size_t)(f.cpos % 3u) * (size_t)B.nwc;
  } else {
    f.cpos = (uint32_t)(S.a + lo - 1);  // first base of the reverse codon starting at lo-1+j is cpos + j
    f.st = B.cb + (size_t)(3u + f.cpos % 3u) * (size_t)B.nwc;
  }
}
MGF_HD bool mgf_own_bit(const MgfOwn& f, bool fwd, int j) {
  const uint32_t sl = (fwd ? f.cpos - (uint32_t)j : f.cpos + (uint32_t)j) / 3u;
  return (mgf_ld(&f.st[sl >> 5].x) >> (sl & 31u)) & 1u;
}
// number of start bits at eligible j in [ja, jb] (multiples of 3, ja <= jb)
MGF_HD int mgf_own_popc(const MgfOwn& f, bool fwd, int ja, int jb) {
  const uint32_t s1 = (fwd ? f.cpos - (uint32_t)jb : f.cpos + (uint32_t)ja) / 3u;
  const uint32_t s2 = (fwd ? f.cpos - (uint32_t)ja : f.cpos + (uint32_t)jb) / 3u;
  const uint32_t w1 = s1 >> 5, w2 = s2 >> 5;
  const unsigned m1 = ~0u << (s1 & 31u), m2 = (2u << (s2 & 31u)) - 1u;
  int cnt = 0;
  for (uint32_t w = w1; w <= w2; w++)
    cnt += mgf_popc(mgf_ld(&f.st[w].x) & (w == w1 ? m1 : ~0u) & (w == w2 ? m2 : ~0u));
  return cnt;
}
// own records at positions j' > j_above (-1: all of them)
MGF_HD int mgf_own_count(const MgfOwn& f, bool fwd, int j_above) {
  if (f.j_hi < f.j_lo) return 0;
  int cnt = 0, jt = f.j_hi;
  bool state = true;  // first_pos == 0
  if (f.trunc) {      // every eligible position emits while first_pos is still 0 (glimmer-mg.cc:1836-1853)
    while (state && jt >= f.j_lo) {
      if (jt <= j_above) return cnt;
      cnt += (jt <= f.j_hs && mgf_own_bit(f, fwd, jt)) ? 2 : 1;
      const int k = fwd ? f.lo + f.m - 2 - jt : f.lo + jt + 2;
      if (k != 0) state = false;
      jt -= 3;
    }
  }
  const int jb = jt < f.j_hs ? jt : f.j_hs;
  int ja = f.j_lo;
  if (j_above + 1 > ja) {
    ja = j_above + 1;
    ja += (3 - ja % 3) % 3;
  }
  if (jb < ja) return cnt;
  return cnt + mgf_own_popc(f, fwd, ja, jb);
}

MGF_HD void mgf_put(gmg_start* dst, const DevParams& P, int jj, int k, double sc, int which, int truncated, int first, int n_err,
                    const int* err_pos, const int* err_type) {
  gmg_start st;
  st.j = jj;
  st.pos = k;
  st.score = (jj > P.ignore_score_len && 0.0 > sc) ? 0.0 : sc;  // long-ORF boost (glimmer-mg.cc:1649-1651)
  st.which = which;
  st.truncated = truncated;
  st.first = first;
  st.n_err = n_err;
  st.err_pos[0] = n_err > 0 ? err_pos[0] : 0;
  st.err_pos[1] = n_err > 1 ? err_pos[1] : 0;
  st.err_type[0] = n_err > 0 ? err_type[0] : 0;
  st.err_type[1] = n_err > 1 ? err_type[1] : 0;
  *dst = st;
}

// Write the call's own records in the reference's order (j descending).  `extra(j)` = records of the call's
// children at positions >= j (they precede the record of j); the i-th own record goes to out[extra(j) + i].
// `score_before(j)` = score[j - 1] of the call.
template <class Extra, class ScoreBefore>
MGF_HD int mgf_own_write_with(const MgfBatch& B, const MgfSeq& S, const DevParams& P, const CodonSets& cs, const MgfOwn& f, bool fwd,
                              double suffix_score, int suffix_j, int n_err, const int* err_pos, const int* err_type,
                              gmg_start* out, Extra extra, ScoreBefore score_before) {
  if (f.j_hi < f.j_lo) return 0;
  int cnt = 0, jt = f.j_hi;
  bool state = true;
  auto emit = [&](int j, int which, int truncated, int first, int64_t ex) {
    const int k = fwd ? f.lo + f.m - 2 - j : f.lo + j + 2;
    const double sc = (score_before(j) - 0.0) + suffix_score;
    mgf_put(out + ex + cnt, P, j + 2 + suffix_j, k, sc, which, truncated, first, n_err, err_pos, err_type);
    cnt++;
  };
  auto which_at = [&](int j) -> int {
    const int bidx = fwd ? f.hi - 1 - j : f.lo - 1 + j;
    int cd = mgf_codon6_at(B.words, S.a + (fwd ? bidx - 2 : bidx));
    if (!fwd) {
      cd = 63 - cd;
      cd = ((cd & 3) << 4) | (cd & 12) | (cd >> 4);
    }
    return (int)cs.which[cd];
  };
  if (f.trunc) {
    while (state && jt >= f.j_lo) {
      const int64_t ex = extra(jt);
      if (jt <= f.j_hs && mgf_own_bit(f, fwd, jt)) {
        emit(jt, -1, 1, 1, ex);
        emit(jt, which_at(jt), 0, 0, ex);
      } else {
        emit(jt, -1, 1, 1, ex);
      }
      const int k = fwd ? f.lo + f.m - 2 - jt : f.lo + jt + 2;
      if (k != 0) state = false;
      jt -= 3;
    }
  }
  const int jb = jt < f.j_hs ? jt : f.j_hs;
  if (jb < f.j_lo) return cnt;
  const uint32_t s1 = (fwd ? f.cpos - (uint32_t)jb : f.cpos + (uint32_t)f.j_lo) / 3u;
  const uint32_t s2 = (fwd ? f.cpos - (uint32_t)f.j_lo : f.cpos + (uint32_t)jb) / 3u;
  const uint32_t w1 = s1 >> 5, w2 = s2 >> 5;
  const unsigned m1 = ~0u << (s1 & 31u), m2 = (2u << (s2 & 31u)) - 1u;
  const uint32_t rr = f.cpos % 3u;  // (cpos -/+ j) = 3 sl + rr
  for (uint32_t wi = 0; wi <= w2 - w1; wi++) {
    const uint32_t w = fwd ? w1 + wi : w2 - wi;
    unsigned x = mgf_ld(&f.st[w].x) & (w == w1 ? m1 : ~0u) & (w == w2 ? m2 : ~0u);
    while (x) {
      const int b = fwd ? mgf_ffs(x) - 1 : 31 - mgf_clz(x);
      x &= ~(1u << b);
      const uint32_t sl = (w << 5) + (uint32_t)b;
      const int j = (int)(fwd ? f.cpos - rr - 3u * sl : 3u * sl + rr - f.cpos);
      const int k = fwd ? f.lo + f.m - 2 - j : f.lo + j + 2;
      emit(j, which_at(j), 0, state ? 1 : 0, extra(j));
      if (k != 0) state = false;
    }
  }
  return cnt;
}

// ---- the same records position by position (plain glimmer-mg: the fused K2 + K3 kernel gives every lane one position) --
// What the serial loop above does at position j only depends on j, the call's geometry and on how many records with a
// non-zero start coordinate came before it (`first_pos` of the reference, glimmer-mg.cc:1836-1853): the truncated records
// sit at j_hi (and at j_hi - 3 when the coordinate at j_hi is 0); start-codon records follow at the eligible j <= jb.
struct MgfPlan {
  int nT;           // positions j_hi, j_hi - 3, ... that emit a truncated record
  int state_after;  // first_pos is still 0 after them
  int jb;           // start-codon records of the chain: eligible j <= jb with a start bit
};
MGF_HD int mgf_kpos(const MgfOwn& f, bool fwd, int j) { return fwd ? f.lo + f.m - 2 - j : f.lo + j + 2; }
MGF_HD void mgf_plan(const MgfOwn& f, bool fwd, MgfPlan& pl) {
  pl.nT = 0;
  pl.state_after = 1;
  int jt = f.j_hi;
  if (f.trunc) {
    bool state = true;
    while (state && jt >= f.j_lo) {
      pl.nT++;
      if (mgf_kpos(f, fwd, jt) != 0) state = false;
      jt -= 3;
    }
    pl.state_after = state ? 1 : 0;
  }
  pl.jb = jt < f.j_hs ? jt : f.j_hs;
}
// records at the eligible position j (a multiple of 3 in [j_lo, j_hi]): 0, 1 or 2.  *trunc_rec: the first of them is the
// truncated record (which = -1, truncated, first); *chain: the position's only record is a start-codon record of the
// chain, whose `first` flag is state_after && "no chain record with a non-zero coordinate at a higher j".
MGF_HD int mgf_recs_at(const MgfOwn& f, const MgfPlan& pl, bool fwd, int j, bool* trunc_rec, bool* chain) {
  const bool tpos = j > f.j_hi - 3 * pl.nT;
  const bool bit = j <= f.j_hs && mgf_own_bit(f, fwd, j);
  *trunc_rec = tpos;
  *chain = !tpos && j <= pl.jb && bit;
  if (tpos) return bit ? 2 : 1;
  return *chain ? 1 : 0;
}
// 6-bit code of the (sense-strand) codon at position j, and the index of that start codon (Can_Be order), as which_at above
MGF_HD int mgf_codon_at(const MgfBatch& B, const MgfSeq& S, const MgfOwn& f, bool fwd, int j) {
  const int bidx = fwd ? f.hi - 1 - j : f.lo - 1 + j;
  int cd = mgf_codon6_at(B.words, S.a + (fwd ? bidx - 2 : bidx));
  if (!fwd) {
    cd = 63 - cd;
    cd = ((cd & 3) << 4) | (cd & 12) | (cd >> 4);
  }
  return cd;
}
MGF_HD int mgf_which_at(const MgfBatch& B, const MgfSeq& S, const unsigned char* which, const MgfOwn& f, bool fwd, int j) {
  return (int)which[mgf_codon_at(B, S, f, fwd, j)];
}

// the same with score[] read off K2's prefix-sum rows (Cumulative_Frame_Score as a difference of two entries)
template <class Extra>
MGF_HD int mgf_own_write(const MgfBatch& B, const MgfSeq& S, const DevParams& P, const CodonSets& cs, const MgfOwn& f, bool fwd,
                         double suffix_score, int suffix_j, double cbase, int n_err, const int* err_pos, const int* err_type,
                         gmg_start* out, Extra extra) {
  const double* row = mgf_row(B, S, fwd, f.lo, f.hi);
  return mgf_own_write_with(B, S, P, cs, f, fwd, suffix_score, suffix_j, n_err, err_pos, err_type, out, extra,
                            [&](int j) { return mgf_score(S, fwd, row, f.lo, f.hi, cbase, j - 1); });
}

// ---- children ---------------------------------------------------------------------------------------------------
MGF_HD bool mgf_may_branch(const DevParams& P, int n_err) { return P.allow_indels && n_err < P.indel_max; }
MGF_HD bool mgf_may_sub(const DevParams& P, int n_err) { return P.allow_subs && n_err < 1; }

// Pass_Stop_Penalty without a quality file (glimmer-mg.cc:961-995): table index 2 a1 + a2
MGF_HD double mgf_stop_penalty(const MgfBatch& B, const MgfSeq& S, bool fwd, int lo, int hi) {
  const int i1 = fwd ? lo - 2 : hi, i2 = fwd ? lo - 1 : hi - 1;
  const int want = fwd ? 0 : 3;  // 'a' forward, 't' reverse
  const bool a1 = (i1 >= 0 && i1 < S.L) && mgf_base_at(B.words, S.a + i1) == want;
  const bool a2 = (i2 >= 0 && i2 < S.L) && mgf_base_at(B.words, S.a + i2) == want;
  return mgf_ld(B.tables + 256 + 2 * (int)a1 + (int)a2);
}
// with a quality file: the probability whose log-odds is the penalty (the log itself is taken on the host, glibc)
MGF_HD double mgf_stop_pstop(const MgfBatch& B, const MgfSeq& S, bool fwd, int lo, int hi) {
  int idx[3];
  if (fwd) { idx[0] = lo - 3; idx[1] = lo - 2; idx[2] = lo - 1; }
  else { idx[0] = hi + 1; idx[1] = hi; idx[2] = hi - 1; }
  const int want = fwd ? 0 : 3;
  const bool a1 = (idx[1] >= 0 && idx[1] < S.L) && mgf_base_at(B.words, S.a + idx[1]) == want;
  const bool a2 = (idx[2] >= 0 && idx[2] < S.L) && mgf_base_at(B.words, S.a + idx[2]) == want;
  double cp[3];
  for (int t = 0; t < 3; t++)
    cp[t] = (idx[t] >= 0 && idx[t] < S.L) ? mgf_ld(B.tables + 260 + (int)mgf_ld(B.qual + S.a + idx[t])) : 0.999;
  double p_stop = cp[0];
  p_stop *= a1 ? (2.0 / 3.0 * cp[1] + 1.0 / 3.0) : cp[1];
  p_stop *= a2 ? (2.0 / 3.0 * cp[2] + 1.0 / 3.0) : cp[2];
  return p_stop;
}

// A candidate child of call `c`: local index `ci` among the call's candidates ([substitution] then, per gate in
// j-descending order, deletion and insertion).  Returns false if the child does not exist; else fills `ch`
// (geometry, suffix, error) -- n_gates / gate0 / base are left to the caller.
MGF_HD bool mgf_child(const MgfBatch& B, const MgfSeq& S, const DevParams& P, const MgfCall& c, int ci, MgfCall& ch) {
  const bool fwd = c.fwd != 0;
  const int sub_slots = mgf_may_sub(P, c.n_err) ? 1 : 0;
  int m = c.hi - c.lo;
  if (m < 0) m = 0;
  const double* row = mgf_row(B, S, fwd, c.lo, c.hi);
  int eep;
  ch.fwd = c.fwd;
  ch.orf = c.orf;
  ch.n_err = (int8_t)(c.n_err + 1);
  ch.n_gates = 0;
  ch.gate0 = 0;
  ch.base = 0;
  if (ci < sub_slots) {  // substitution through the previous stop (glimmer-mg.cc:1771-1806)
    eep = fwd ? c.lo - 3 : c.hi + 3;
    if (!(eep >= 0 && eep - 2 < S.L)) return false;
    double ess = c.suffix_score + (B.sub_pen ? mgf_ld(B.sub_pen + c.orf) : mgf_stop_penalty(B, S, fwd, c.lo, c.hi));
    if (m > 0) ess += mgf_score(S, fwd, row, c.lo, c.hi, c.cbase, m - 1) - 0.0;
    ch.suffix_score = ess;
    ch.suffix_j = c.suffix_j + m;
    ch.err_pos = fwd ? c.lo - 2 : c.hi + 2;
    ch.err_type = 2;
    ch.jpar = -1;
  } else {
    const int t = (ci - sub_slots) >> 1, ph = (ci - sub_slots) & 1;
    const int64_t g = (int64_t)mgf_ld(B.gate_pos + (fwd ? c.gate0 + (uint32_t)t : c.gate0 - (uint32_t)t));
    const int q = (int)(g - S.a);
    const int j = fwd ? c.hi - 1 - q : q - (c.lo - 1);
    const int k = fwd ? c.lo + m - 2 - j : c.lo + j + 2;
    const int j3 = j % 3;
    const double pen = mgf_ld(B.tables + (int)mgf_ld(B.qual + g));
    // Score_Indels (glimmer-mg.cc:1513-1602): deletion uses score[j], insertion score[j-1]
    const double ess = c.suffix_score + mgf_score(S, fwd, row, c.lo, c.hi, c.cbase, ph == 0 ? j : j - 1) - 0.0 + pen;
    if (!(ess > P.indel_suffix_thresh)) return false;
    if (ph == 0) {
      eep = fwd ? k + j3 : k - j3;
      ch.err_pos = fwd ? k + 3 : k - 1;
      ch.err_type = 1;
    } else {
      eep = fwd ? k - (2 - j3) : k + 2 - j3;
      ch.err_pos = fwd ? k + 2 : k - 2;
      ch.err_type = 0;
    }
    ch.suffix_score = ess;
    ch.suffix_j = c.suffix_j + j + 2 - j3;
    ch.jpar = j;
  }
  int lo, hi;
  mgf_open(B, S, fwd, eep, &lo, &hi);
  ch.lo = lo;
  ch.hi = hi;
  ch.cbase = mgf_cbase(B, S, fwd, lo, hi);
  ch.ok = 1;
  return true;
}

MGF_HD int mgf_n_candidates(const DevParams& P, const MgfCall& c) {
  return (mgf_may_sub(P, c.n_err) ? 1 : 0) + 2 * c.n_gates;
}

// parent of every candidate: segment p of the scan `off` owns candidates [off[p], off[p + 1]).  Written once per level
// (a binary search per candidate in each of the passes that need the parent cost a third of their time).
MGF_HD void mgf_fill_parent(const uint32_t* off, uint32_t p, uint32_t* par) {
  const uint32_t a = mgf_ld(off + p), b = mgf_ld(off + p + 1);
  for (uint32_t i = a; i < b; i++) par[i] = p;
}

// ---- the passes ---------------------------------------------------------------------------------------------------
struct MgfWork {
  const gmg_orf* orfs;
  const int32_t* orf_seq;
  uint32_t n_orfs;
  // level 0
  MgfCall* root;    // [n_orfs]
  uint32_t* n1;     // [n_orfs + 1] candidates of each root; off1 = its exclusive scan
  uint32_t* off1;
  uint32_t* own0;   // [n_orfs]
  // level 1
  uint32_t c1;      // number of level-1 candidates (= off1[n_orfs]; known to the host before pass B)
  uint32_t* par1;   // [c1] ORF of every level-1 candidate
  MgfCall* call1;   // [c1]
  uint32_t* n2;     // [c1 + 1]; off2 = its exclusive scan
  uint32_t* off2;
  uint32_t* own1;   // [c1]
  uint32_t* t1;     // [c1 + 1] records of the subtree of every level-1 candidate; s1 = its exclusive scan
  uint32_t* s1;
  // level 2
  uint32_t c2;      // number of level-2 candidates (= off2[c1])
  uint32_t* par2;   // [c2] level-1 candidate (parent call) of every level-2 candidate
  uint32_t* cnt3;   // [c2 + 1] records of every level-2 candidate; s3 = its exclusive scan
  uint32_t* s3;
  // output
  int64_t* counts;           // [n_orfs + 1] records per ORF
  const int64_t* start_off;  // exclusive scan of counts
  gmg_start* starts;
};

MGF_HD MgfSeq mgf_seq_of(const MgfBatch& B, int32_t s) {
  MgfSeq S;
  S.a = mgf_ld(B.off + s);
  S.L = (int)(mgf_ld(B.off + s + 1) - S.a);
  return S;
}

MGF_HD void mgf_pass_a(const MgfBatch& B, const DevParams& P, const MgfWork& W, uint32_t o) {
  const gmg_orf orf = W.orfs[o];
  const int32_t s = W.orf_seq[o];
  const MgfSeq S = mgf_seq_of(B, s);
  MgfCall c;
  c.fwd = orf.frame > 0;
  c.orf = (int32_t)o;
  c.n_err = 0;
  c.err_pos = 0;
  c.err_type = 0;
  c.jpar = -1;
  c.suffix_j = 0;
  c.suffix_score = 0.0;
  c.base = 0;
  c.n_gates = 0;
  c.gate0 = 0;
  c.lo = c.hi = 0;
  c.cbase = 0.0;
  c.ok = (B.cert == 0 || B.cert[s] != 0) ? 1 : 0;  // uncertified sequences: the ordered kernel scores their ORFs
  uint32_t n1 = 0, own = 0;
  if (c.ok) {
    const bool fwd = c.fwd != 0;
    int lo, hi;
    mgf_open(B, S, fwd, fwd ? orf.stop_position - 1 : orf.stop_position + 3, &lo, &hi);
    c.lo = lo;
    c.hi = hi;
    c.cbase = mgf_cbase(B, S, fwd, lo, hi);
    if (mgf_may_branch(P, 0)) {
      const int lowest_j = P.min_gene_len - 3 < 3 ? P.min_gene_len - 3 : 3;
      int ng;
      uint32_t g0;
      mgf_gates(B, S, fwd, lo, hi, lowest_j, &ng, &g0);
      c.n_gates = ng;
      c.gate0 = g0;
    }
    n1 = (uint32_t)mgf_n_candidates(P, c);
    MgfOwn f;
    mgf_own_open(B, S, P, fwd, lo, hi, 0, f);
    own = (uint32_t)mgf_own_count(f, fwd, -1);
  }
  W.root[o] = c;
  W.n1[o] = n1;
  W.own0[o] = own;
}

MGF_HD void mgf_pass_b(const MgfBatch& B, const DevParams& P, const MgfWork& W, uint32_t i) {
  const uint32_t o = mgf_ld(W.par1 + i);
  const MgfCall r = W.root[o];
  const MgfSeq S = mgf_seq_of(B, W.orf_seq[o]);
  MgfCall ch;
  uint32_t n2 = 0, own = 0;
  if (mgf_child(B, S, P, r, (int)(i - mgf_ld(W.off1 + o)), ch)) {
    const bool fwd = ch.fwd != 0;
    if (mgf_may_branch(P, ch.n_err)) {
      const int lowest_j = P.min_gene_len - 3 < 3 ? P.min_gene_len - 3 : 3;
      int ng;
      uint32_t g0;
      mgf_gates(B, S, fwd, ch.lo, ch.hi, lowest_j, &ng, &g0);
      ch.n_gates = ng;
      ch.gate0 = g0;
    }
    n2 = (uint32_t)mgf_n_candidates(P, ch);
    MgfOwn f;
    mgf_own_open(B, S, P, fwd, ch.lo, ch.hi, ch.suffix_j, f);
    own = (uint32_t)mgf_own_count(f, fwd, -1);
    W.call1[i] = ch;
  } else {
    W.call1[i].ok = 0;
  }
  W.n2[i] = n2;
  W.own1[i] = own;
}

MGF_HD void mgf_pass_c(const MgfBatch& B, const DevParams& P, const MgfWork& W, uint32_t i) {
  const uint32_t p = mgf_ld(W.par2 + i);
  const MgfCall c = W.call1[p];
  const MgfSeq S = mgf_seq_of(B, W.orf_seq[c.orf]);
  MgfCall ch;
  uint32_t cnt = 0;
  if (mgf_child(B, S, P, c, (int)(i - mgf_ld(W.off2 + p)), ch)) {
    MgfOwn f;
    mgf_own_open(B, S, P, ch.fwd != 0, ch.lo, ch.hi, ch.suffix_j, f);
    cnt = (uint32_t)mgf_own_count(f, ch.fwd != 0, -1);
  }
  W.cnt3[i] = cnt;
}

MGF_HD void mgf_pass_d(const MgfWork& W, uint32_t i) {
  W.t1[i] = W.own1[i] + (mgf_ld(W.s3 + mgf_ld(W.off2 + i + 1)) - mgf_ld(W.s3 + mgf_ld(W.off2 + i)));
}

MGF_HD void mgf_pass_e(const MgfWork& W, uint32_t o) {
  if (!W.root[o].ok) return;  // left to the ordered kernel
  W.counts[o] = (int64_t)W.own0[o] + (int64_t)(mgf_ld(W.s1 + mgf_ld(W.off1 + o + 1)) - mgf_ld(W.s1 + mgf_ld(W.off1 + o)));
}

// records of the children of call c (candidates [first, ...) with sizes scanned in `scan`) at positions >= j
MGF_HD int64_t mgf_children_before(const MgfBatch& B, const MgfSeq& S, const DevParams& P, const MgfCall& c, uint32_t first,
                                   const uint32_t* scan, int j) {
  const uint32_t idx = first + (mgf_may_sub(P, c.n_err) ? 1u : 0u) + 2u * (uint32_t)mgf_gates_at_or_above(B, S, c, j);
  return (int64_t)(mgf_ld(scan + idx) - mgf_ld(scan + first));
}

// own records of the root
MGF_HD void mgf_write_0(const MgfBatch& B, const DevParams& P, const CodonSets& cs, const MgfWork& W, uint32_t o) {
  const MgfCall c = W.root[o];
  if (!c.ok || W.own0[o] == 0) return;
  const MgfSeq S = mgf_seq_of(B, W.orf_seq[o]);
  const bool fwd = c.fwd != 0;
  MgfOwn f;
  mgf_own_open(B, S, P, fwd, c.lo, c.hi, 0, f);
  const uint32_t first = mgf_ld(W.off1 + o);
  const int ep[2] = {0, 0}, et[2] = {0, 0};
  mgf_own_write(B, S, P, cs, f, fwd, 0.0, 0, c.cbase, 0, ep, et, W.starts + W.start_off[o],
                [&](int j) { return mgf_children_before(B, S, P, c, first, W.s1, j); });
}

// place of a level-1 call's subtree, then its own records
MGF_HD void mgf_write_1(const MgfBatch& B, const DevParams& P, const CodonSets& cs, const MgfWork& W, uint32_t i) {
  MgfCall c = W.call1[i];
  if (!c.ok) return;
  const uint32_t o = (uint32_t)c.orf;
  const MgfCall r = W.root[o];
  const MgfSeq S = mgf_seq_of(B, W.orf_seq[o]);
  const bool fwd = c.fwd != 0;
  int64_t base = W.start_off[o] + (int64_t)(mgf_ld(W.s1 + i) - mgf_ld(W.s1 + mgf_ld(W.off1 + o)));
  if (c.jpar >= 0) {  // the root's own records above the branching position precede this subtree
    MgfOwn fr;
    mgf_own_open(B, S, P, fwd, r.lo, r.hi, 0, fr);
    base += mgf_own_count(fr, fwd, c.jpar);
  }
  W.call1[i].base = base;
  if (W.own1[i] == 0) return;
  MgfOwn f;
  mgf_own_open(B, S, P, fwd, c.lo, c.hi, c.suffix_j, f);
  const uint32_t first = mgf_ld(W.off2 + i);
  const int ep[2] = {c.err_pos, 0}, et[2] = {c.err_type, 0};
  mgf_own_write(B, S, P, cs, f, fwd, c.suffix_score, c.suffix_j, c.cbase, 1, ep, et, W.starts + base,
                [&](int j) { return mgf_children_before(B, S, P, c, first, W.s3, j); });
}

// records of a level-2 (leaf) call
MGF_HD void mgf_write_2(const MgfBatch& B, const DevParams& P, const CodonSets& cs, const MgfWork& W, uint32_t i) {
  if (mgf_ld(W.s3 + i + 1) == mgf_ld(W.s3 + i)) return;  // no records (or the call does not exist)
  const uint32_t p = mgf_ld(W.par2 + i);
  const MgfCall c = W.call1[p];
  const MgfSeq S = mgf_seq_of(B, W.orf_seq[c.orf]);
  const bool fwd = c.fwd != 0;
  MgfCall ch;
  const uint32_t first = mgf_ld(W.off2 + p);
  if (!mgf_child(B, S, P, c, (int)(i - first), ch)) return;
  MgfOwn fp;
  mgf_own_open(B, S, P, fwd, c.lo, c.hi, c.suffix_j, fp);
  const int64_t base = c.base + (int64_t)(mgf_ld(W.s3 + i) - mgf_ld(W.s3 + first)) + mgf_own_count(fp, fwd, ch.jpar);
  MgfOwn f;
  mgf_own_open(B, S, P, fwd, ch.lo, ch.hi, ch.suffix_j, f);
  const int ep[2] = {c.err_pos, ch.err_pos}, et[2] = {c.err_type, ch.err_type};
  mgf_own_write(B, S, P, cs, f, fwd, ch.suffix_score, ch.suffix_j, ch.cbase, 2, ep, et, W.starts + base,
                [](int) { return (int64_t)0; });
}
