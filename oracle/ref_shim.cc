// oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim (our code) around the UNMODIFIED reference ICM library
// (/root/reference/src/ICM/icm.{hh,cc}, compiled by oracle/Makefile into
// oracle/_ref/obj/icm.o).  It lets the python tests call the real ICM_t /
// ICM_Training_t through ctypes to pin oracle/icm_oracle.c and the CUDA path.
// Nothing here is part of the product.

#include "icm.hh"
#include "fasta.hh"
#include <string>
#include <vector>
#include <cstring>

namespace {
// ICM_t keeps its tables protected (icm.hh:118-129); a derived peeker exposes them.
struct Peek : public ICM_t {
  int nodes() const { return num_nodes; }
  int len() const { return model_len; }
  int depth() const { return model_depth; }
  int period() const { return periodicity; }
  const ICM_Score_Node_t* row(int p) const { return score[p]; }
};
}  // namespace

extern "C" {

void* ref_icm_read(const char* path) {
  ICM_t* m = new ICM_t();
  m->Read(const_cast<char*>(path));
  return m;
}

void* ref_icm_build_indep(double gc, const char** stops, int n_stops) {
  ICM_t* m = new ICM_t(3, 2, 3);
  std::vector<const char*> v(stops, stops + n_stops);
  m->Build_Indep_WO_Stops(gc, v);
  return m;
}

void ref_icm_free(void* h) { delete static_cast<ICM_t*>(h); }

void ref_icm_dims(void* h, int* dims /*len, depth, period, nodes*/) {
  const Peek* p = static_cast<const Peek*>(static_cast<ICM_t*>(h));
  dims[0] = p->len(); dims[1] = p->depth(); dims[2] = p->period(); dims[3] = p->nodes();
}

// mip: int16[period*nodes], prob: float[period*nodes*4]
void ref_icm_tables(void* h, short* mip, float* prob) {
  const Peek* p = static_cast<const Peek*>(static_cast<ICM_t*>(h));
  for (int f = 0; f < p->period(); f++)
    for (int i = 0; i < p->nodes(); i++) {
      mip[f * p->nodes() + i] = p->row(f)[i].mut_info_pos;
      memcpy(prob + 4 * (size_t)(f * p->nodes() + i), p->row(f)[i].prob, 4 * sizeof(float));
    }
}

double ref_full_window_prob(void* h, const char* w, int frame) {
  return static_cast<ICM_t*>(h)->Full_Window_Prob(w, frame);
}
double ref_partial_window_prob(void* h, int predict_pos, const char* s, int frame) {
  return static_cast<ICM_t*>(h)->Partial_Window_Prob(predict_pos, s, frame);
}
double ref_score_string(void* h, const char* s, int len, int frame) {
  return static_cast<ICM_t*>(h)->Score_String(s, len, frame);
}
void ref_cumulative_score(void* h, const char* s, int len, int frame, double* out) {
  std::string str(s, len);
  std::vector<double> v;
  static_cast<ICM_t*>(h)->Cumulative_Score(str, v, frame);
  for (int i = 0; i < len; i++) out[i] = v[i];
}
void ref_frame_score(void* h, const char* s, int len, int frame, double* out) {
  std::string str(s, len);
  std::vector<double> v;
  static_cast<ICM_t*>(h)->Frame_Score(str, v, frame);
  for (int i = 0; i < len; i++) out[i] = v[i];
}

// Train an ICM on n NUL-terminated strings (already lower-cased / reversed by the
// caller, as build-icm.cc:111-118 does) and write it in binary form to `path`.
void* ref_icm_train(const char** strings, int n, int w, int d, int p) {
  ICM_Training_t* m = new ICM_Training_t(w, d, p);
  std::vector<char*> data;
  for (int i = 0; i < n; i++) data.push_back(strdup(strings[i]));
  m->Train_Model(data);
  for (int i = 0; i < n; i++) free(data[i]);
  return static_cast<ICM_t*>(m);
}
void ref_icm_train_free(void* h) { delete static_cast<ICM_Training_t*>(static_cast<ICM_t*>(h)); }

int ref_icm_write(void* h, const char* path) {
  FILE* fp = fopen(path, "wb");
  if (!fp) return -1;
  static_cast<ICM_t*>(h)->Output(fp, true);
  fclose(fp);
  return 0;
}

// Fasta_Read (Common/fasta.cc:236-283) over a whole file: the records the reference's reader sees.  Returns the number
// of records; *seqs / *hdrs are malloc'd buffers of NUL-separated strings (sequence characters as read, header text).
int ref_fasta_read_file(const char* path, char** seqs, long* seqs_bytes, char** hdrs, long* hdrs_bytes) {
  FILE* fp = fopen(path, "r");
  if (!fp) return -1;
  std::string s, h, all_s, all_h;
  int n = 0;
  while (Fasta_Read(fp, s, h)) {
    all_s += s;
    all_s.push_back('\0');
    all_h += h;
    all_h.push_back('\0');
    n++;
  }
  fclose(fp);
  *seqs = (char*)malloc(all_s.size() + 1);
  memcpy(*seqs, all_s.data(), all_s.size());
  *seqs_bytes = (long)all_s.size();
  *hdrs = (char*)malloc(all_h.size() + 1);
  memcpy(*hdrs, all_h.data(), all_h.size());
  *hdrs_bytes = (long)all_h.size();
  return n;
}

// Fasta_Qual_Vec_Read (Common/fasta.cc:115-170) over a whole file: *vals = all values (malloc'd), *counts = values per
// record (malloc'd, n entries), headers as above.  Returns the number of records.
int ref_fasta_qual_read_file(const char* path, int** vals, long* n_vals, long** counts, char** hdrs, long* hdrs_bytes) {
  FILE* fp = fopen(path, "r");
  if (!fp) return -1;
  std::vector<int> q, all;
  std::vector<long> cnt;
  std::string h, all_h;
  while (Fasta_Qual_Vec_Read(fp, q, h)) {
    all.insert(all.end(), q.begin(), q.end());
    cnt.push_back((long)q.size());
    all_h += h;
    all_h.push_back('\0');
  }
  fclose(fp);
  *vals = (int*)malloc((all.size() + 1) * sizeof(int));
  memcpy(*vals, all.data(), all.size() * sizeof(int));
  *n_vals = (long)all.size();
  *counts = (long*)malloc((cnt.size() + 1) * sizeof(long));
  memcpy(*counts, cnt.data(), cnt.size() * sizeof(long));
  *hdrs = (char*)malloc(all_h.size() + 1);
  memcpy(*hdrs, all_h.data(), all_h.size());
  *hdrs_bytes = (long)all_h.size();
  return (int)cnt.size();
}

}  // extern "C"
