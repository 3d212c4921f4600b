// Host-side pieces of the boundary: error state, device checks, the copy-number probability table
// (src/probabilitytable.cpp, src/copynumber.cpp — evaluated in x87 long double like the reference and
// stored as natural logs in fp64), histogram peak logic (src/histogram.cpp:41-63,
// src/sequenceutils.cpp:42-84) and the VCF-ordered result layout (src/genotypingresult.cpp:48-67).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "common.cuh"

namespace pg {

static thread_local std::string g_error;
static thread_local int g_code = PG_OK;
thread_local uint64_t g_launches = 0;

int fail(int code, const std::string& msg) {
  g_error = msg;
  g_code = code;
  return code;
}
void clear_error() {
  g_error.clear();
  g_code = PG_OK;
}
int last_code() { return g_code == PG_OK ? PG_ERR_ARG : g_code; }

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PG_ERR_CUDA, "no CUDA device available (libpangenie_b200 has no CPU fallback)");
  }
  if (device < 0 || device >= n) return fail(PG_ERR_ARG, "invalid device ordinal " + std::to_string(device));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(PG_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major < 10) return fail(PG_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100-class; this library is built for sm_100a only");
  return PG_OK;
}

// ---- probability model, long double (x87) as in the reference -----------------------------------
typedef long double ld;

static double get_error_param(double kmer_coverage) {  // src/probabilitytable.cpp:7-19
  if (kmer_coverage < 10.0) return 0.99;
  if (kmer_coverage < 20) return 0.95;
  if (kmer_coverage < 40) return 0.9;
  return 0.8;
}

static ld poisson(ld mean, unsigned value) {  // src/probabilitytable.cpp:75-81
  ld sum = 0.0L;
  const int v = (int)value;
  for (size_t i = 1; i <= value; ++i) sum += std::log((double)i);
  return expl(-mean + v * logl(mean) - sum);
}

static ld geometric(ld p, unsigned value) { return powl(1.0L - p, value) * p; }  // :83-85

static void compute_probability(unsigned cov, unsigned count, ld reg, ld out[3]) {  // :55-65 + copynumber.cpp:14-41
  const ld c0 = geometric(get_error_param(cov), count);
  const ld c1 = poisson(cov / 2.0, count);
  const ld c2 = poisson(cov, count);
  if (reg > 0) {
    const ld sum = c0 + c1 + c2 + 3.0L * reg;
    out[0] = (c0 + reg) / sum;
    out[1] = (c1 + reg) / sum;
    out[2] = 1.0L - out[0] - out[1];
  } else {
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
  }
}

static double to_log(ld p) { return p > 0 ? (double)logl(p) : -std::numeric_limits<double>::infinity(); }

}  // namespace pg

using namespace pg;

extern "C" const char* pg_last_error(void) { return g_error.c_str(); }
extern "C" const char* pg_version(void) { return "pangenie_b200 0.2 (sm_100a)"; }
extern "C" uint64_t pg_kernel_launches(void) { return g_launches; }
extern "C" int pg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int pg_probtable_init(pg_probtable* t, uint16_t cov_min, uint16_t cov_max, uint16_t count_max,
                                 double regularization) {
  clear_error();
  if (!t) return fail(PG_ERR_ARG, "null table");
  if (cov_max < cov_min) return fail(PG_ERR_ARG, "cov_max < cov_min");
  t->cov_min = cov_min;
  t->cov_max = cov_max;
  t->count_max = count_max;
  t->regularization = regularization;
  const size_t ncov = (size_t)(cov_max - cov_min), n = ncov * count_max * 3;
  t->log_p = n ? (double*)malloc(n * sizeof(double)) : nullptr;
  if (n && !t->log_p) return fail(PG_ERR_ARG, "out of host memory");
  for (unsigned count = 0; count < count_max; ++count)
    for (unsigned j = 0; j < ncov; ++j) {
      ld p[3];
      compute_probability(cov_min + j, count, (ld)regularization, p);
      double* e = t->log_p + ((size_t)count * ncov + j) * 3;
      for (int cn = 0; cn < 3; ++cn) e[cn] = to_log(p[cn]);
    }
  return PG_OK;
}

extern "C" int pg_probtable_modify(pg_probtable* t, uint16_t cov, uint16_t count, double p0, double p1, double p2) {
  clear_error();
  if (!t || !t->log_p || cov < t->cov_min || cov >= t->cov_max || count >= t->count_max)
    return fail(PG_ERR_ARG, "ProbabilityTable::modify_probability: no precomputed values for these parameters.");
  double* e = t->log_p + ((size_t)count * (t->cov_max - t->cov_min) + (cov - t->cov_min)) * 3;
  e[0] = to_log(p0);
  e[1] = to_log(p1);
  e[2] = to_log(p2);
  return PG_OK;
}

extern "C" double pg_probtable_get(const pg_probtable* t, uint16_t cov, uint16_t count, int cn) {
  if (!t || cn < 0 || cn > 2) return std::numeric_limits<double>::quiet_NaN();
  if (t->log_p && cov >= t->cov_min && cov < t->cov_max && count < t->count_max) {
    const double l = t->log_p[((size_t)count * (t->cov_max - t->cov_min) + (cov - t->cov_min)) * 3 + cn];
    return std::isinf(l) ? 0.0 : (double)expl((ld)l);
  }
  ld p[3];
  compute_probability(cov, count, (ld)t->regularization, p);
  return (double)p[cn];
}

extern "C" void pg_probtable_free(pg_probtable* t) {
  if (t && t->log_p) free(t->log_p);
  if (t) t->log_p = nullptr;
}

extern "C" int pg_histogram_peak(uint64_t* h, uint64_t n, int largest_peak, uint64_t* peak) {
  clear_error();
  if (!h || !peak || n < 2) return fail(PG_ERR_ARG, "invalid histogram");
  // Histogram::smooth_histogram: in place, so bin i-1 is already smoothed when bin i is computed
  for (uint64_t i = 1; i + 1 < n; ++i) h[i] = (h[i - 1] + h[i] + h[i + 1]) / 3;
  // Histogram::find_peaks
  std::vector<uint64_t> ids, vals;
  bool falling = false;
  uint64_t prev = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t v = h[i];
    if (prev < v) {
      falling = false;
    } else if (prev > v) {
      if (!falling) {
        ids.push_back(i - 1);
        vals.push_back(prev);
      }
      falling = true;
    }
    prev = v;
  }
  // compute_kmer_coverage
  if (ids.empty()) return fail(PG_ERR_ARG, "sequenceutils::computeHistogram: no peak found in kmer-count histogram.");
  if (ids.size() < 2) {
    *peak = ids[0];
    return PG_OK;
  }
  uint64_t best, best_id, second, second_id;
  if (vals[0] < vals[1]) {
    best = vals[1]; best_id = ids[1]; second = vals[0]; second_id = ids[0];
  } else {
    best = vals[0]; best_id = ids[0]; second = vals[1]; second_id = ids[1];
  }
  for (size_t i = 0; i < vals.size(); ++i) {
    if (vals[i] > best) {
      second = best; second_id = best_id; best = vals[i]; best_id = ids[i];
    } else if (vals[i] > second && vals[i] != best) {
      second = vals[i]; second_id = ids[i];
    }
  }
  *peak = largest_peak ? best_id : second_id;
  return PG_OK;
}

extern "C" int pg_result_layout(const pg_panel* p, uint64_t* offsets) {
  clear_error();
  if (!p || !offsets) return fail(PG_ERR_ARG, "null argument");
  uint64_t off = 0;
  for (uint32_t v = 0; v < p->n_variants; ++v) {
    offsets[v] = off;
    uint64_t maxa = 0;
    for (uint32_t a = p->allele_offsets[v]; a < p->allele_offsets[v + 1]; ++a) maxa = std::max<uint64_t>(maxa, p->allele_ids[a]);
    const uint64_t nr = maxa + 1;
    off += nr * (nr + 1) / 2;
  }
  offsets[p->n_variants] = off;
  return PG_OK;
}
