// Engine: panel upload, count fill, emission, forward-backward orchestration, result download.
// Replaces run_genotyping / HMM::HMM (src/commands.cpp:155-185, src/hmm.cpp:25-74), ColumnIndexer
// (src/columnindexer.cpp:8-33), EmissionProbabilityComputer (src/emissionprobabilitycomputer.cpp:9-53),
// fill_read_kmercounts (src/commands.cpp:76-152) and the GT/GQ post-processing of Graph::write_genotypes
// (src/graph.cpp:206-240).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "hmm_kernels.cuh"
#include "hmm_scan.cuh"

namespace pg {

// PG_TRACE=1: host wall-clock of the phases of one call on stderr (where does the time between kernels go?)
struct HostTrace {
  bool on;
  const char* name;
  std::chrono::steady_clock::time_point t0;
  explicit HostTrace(const char* n) : name(n) {
    static const bool v = [] { const char* e = getenv("PG_TRACE"); return e && e[0] == '1'; }();
    on = v;
    t0 = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[pg trace] %s: %s %.3f ms\n", name, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// -------------------------------------------------------------------------------------------------
// device views
// -------------------------------------------------------------------------------------------------
struct PanelDev {
  uint32_t V, P, K, A, F;
  const uint64_t* positions;
  const uint16_t* path_to_allele;
  uint16_t* coverage;
  const uint32_t* kmer_off;
  uint16_t* kmer_counts;
  const uint32_t* allele_off;
  const uint16_t* allele_ids;
  const uint8_t* allele_undef;
  const uint16_t* allele_koff;
  const uint32_t* allele_kmask;
  const uint64_t* kmer_codes;
  const uint32_t* flank_off;
  const uint64_t* flank_codes;
};

struct TableDev {
  uint32_t cov_min, cov_max, count_max;
  double reg;
  const double* log_p;
};

// -------------------------------------------------------------------------------------------------
// probability model on the device (out-of-table entries; src/probabilitytable.cpp:47-65,75-85)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double poisson_log(double mean, double v) { return -mean + v * log(mean) - lgamma(v + 1.0); }

__device__ void logp_triple(const TableDev& T, uint32_t cov, uint32_t count, double lp[3]) {
  if (T.log_p && cov >= T.cov_min && cov < T.cov_max && count < T.count_max) {
    const double* e = T.log_p + ((size_t)count * (T.cov_max - T.cov_min) + (cov - T.cov_min)) * 3;
    lp[0] = e[0];
    lp[1] = e[1];
    lp[2] = e[2];
    return;
  }
  const double c = (double)cov, v = (double)count;
  const double p = c < 10.0 ? 0.99 : c < 20.0 ? 0.95 : c < 40.0 ? 0.9 : 0.8;  // get_error_param
  const double l0 = v * log1p(-p) + log(p);                                   // geometric
  const double l1 = poisson_log(c / 2.0, v);
  const double l2 = poisson_log(c, v);
  if (T.reg > 0.0) {  // CopyNumber(c0,c1,c2,reg) (src/copynumber.cpp:22-28)
    const double c0 = exp(l0), c1 = exp(l1), c2 = exp(l2);
    const double s = c0 + c1 + c2 + 3.0 * T.reg;
    lp[0] = log((c0 + T.reg) / s);
    lp[1] = log((c1 + T.reg) / s);
    lp[2] = log((c2 + T.reg) / s);
  } else {
    lp[0] = l0;
    lp[1] = l1;
    lp[2] = l2;
  }
}

__device__ __forceinline__ double logaddexp(double a, double b) {
  const double m = fmax(a, b), n = fmin(a, b);
  if (isinf(m) && m < 0) return m;
  return m + log1p(exp(n - m));
}

__device__ __forceinline__ uint32_t kmer_on_allele(uint32_t k, uint32_t off, uint32_t mask) {  // src/kmerpath.cpp:33-48
  return (k >= off && k < off + 32u) ? ((mask >> (k - off)) & 1u) : 0u;
}

// -------------------------------------------------------------------------------------------------
// emission: one warp per variant.  em[em_off[v] + i1*A + i2] (allele-index space) = e(i1,i2)/max,
// log_scale[v] = ln max; all-zero tables report 1.0 everywhere (emissionprobabilitycomputer.cpp:24,31-34).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) emission_kernel(PanelDev pd, TableDev T, const uint64_t* __restrict__ em_off,
                                                        double* __restrict__ em, double* __restrict__ log_scale,
                                                        uint32_t v_begin, uint32_t v_end) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = v_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); v < v_end; v += warps) {
    const uint32_t ab = pd.allele_off[v], A = pd.allele_off[v + 1] - ab;
    const uint32_t kb = pd.kmer_off[v], K = pd.kmer_off[v + 1] - kb;
    const uint32_t cov = pd.coverage[v];
    double* out = em + em_off[v];
    double mx = -INFINITY;
    for (uint32_t i1 = 0; i1 < A; ++i1) {
      const uint32_t off1 = pd.allele_koff[ab + i1], m1 = pd.allele_kmask[ab + i1];
      const bool u1 = pd.allele_undef[ab + i1] != 0;
      for (uint32_t i2 = i1; i2 < A; ++i2) {
        const uint32_t off2 = pd.allele_koff[ab + i2], m2 = pd.allele_kmask[ab + i2];
        const bool u2 = pd.allele_undef[ab + i2] != 0;
        double sum = 0.0;
        for (uint32_t k = lane; k < K; k += 32) {
          double lp[3];
          logp_triple(T, cov, pd.kmer_counts[kb + k], lp);
          const uint32_t c = kmer_on_allele(k, off1, m1) + kmer_on_allele(k, off2, m2);
          double term;
          if (u1 && u2) {
            term = logaddexp(logaddexp(lp[0], lp[1]), lp[2]) - 1.0986122886681098;  // ln 3
          } else if (u1 || u2) {
            const uint32_t c0 = c < 2 ? c : 2, c1 = c + 1 < 2 ? c + 1 : 2;  // reference asserts c < 2 (:44)
            term = logaddexp(lp[c0], lp[c1]) - 0.6931471805599453;          // ln 2
          } else {
            term = lp[c];
          }
          sum += term;
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
          out[(size_t)i1 * A + i2] = sum;
          out[(size_t)i2 * A + i1] = sum;
        }
        mx = fmax(mx, sum);
      }
    }
    __syncwarp();
    const bool all_zero = isinf(mx) && mx < 0;
    for (uint32_t q = lane; q < A * A; q += 32) out[q] = all_zero ? 1.0 : exp(out[q] - mx);
    if (lane == 0) log_scale[v] = all_zero ? 0.0 : mx;
  }
}

// -------------------------------------------------------------------------------------------------
// ColumnIndexer (src/columnindexer.cpp:24-31): a variant is an HMM column iff some selected path carries an
// allele that is neither 0 nor undefined.  One warp per variant.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) column_flag_kernel(PanelDev pd, const uint16_t* __restrict__ sel, uint32_t n_sel,
                                                           uint8_t* __restrict__ is_column) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < pd.V; v += warps) {
    const uint32_t ab = pd.allele_off[v], A = pd.allele_off[v + 1] - ab;
    bool any = false;
    for (uint32_t q = lane; q < n_sel; q += 32) {
      const uint16_t a = pd.path_to_allele[(size_t)v * pd.P + sel[q]];
      if (a != 0) {
        bool undef = false;
        for (uint32_t i = 0; i < A; ++i)
          if (pd.allele_ids[ab + i] == a) undef = pd.allele_undef[ab + i] != 0;
        if (!undef) any = true;
      }
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) is_column[v] = any ? 1 : 0;
  }
}

// -------------------------------------------------------------------------------------------------
// per-column descriptor records (layout in hmm_kernels.cuh).  One warp per column.
// -------------------------------------------------------------------------------------------------
struct TransParams {
  double recomb, effN;
  int uniform;
};

__device__ __forceinline__ void transition_coefs(uint64_t from, uint64_t to, uint32_t P, const TransParams& tp, double out[4]) {
  if (tp.uniform) {  // compute_transition_prob returns 1 for every switch count (:34-38)
    out[0] = 0.0; out[1] = 0.0; out[2] = 1.0; out[3] = (double)P * (double)P;
    return;
  }
  // distance = (to-from) * 0.000004 * recomb * N_e ; r = (1-e^{-d/P})/P ; n = e^{-d/P} + r   (:14-18)
  const double x = (double)(to - from) * 0.000004 * tp.recomb * tp.effN / (double)P;
  const double em = exp(-x);
  const double r = -expm1(-x) / (double)P;
  out[0] = em * em;  // t0 - 2 t1 + t2 = (n-r)^2
  out[1] = r * em;   // t1 - t2      = r (n-r)
  out[2] = r * r;    // t2
  const double n = em + r + ((double)P - 1.0) * r;  // (n - r + P r): total mass multiplier
  out[3] = n * n;
}

__global__ void __launch_bounds__(256) desc_build_kernel(PanelDev pd, const uint16_t* __restrict__ sel, uint32_t n_sel,
                                                          const uint32_t* __restrict__ col_variant, const uint32_t* __restrict__ col_chrom_end,
                                                          const uint32_t* __restrict__ col_chrom_begin, uint32_t n_cols,
                                                          const uint64_t* __restrict__ em_off, const double* __restrict__ em,
                                                          TransParams tp, const uint64_t* __restrict__ gl_off, uint8_t* __restrict__ desc,
                                                          uint32_t stride) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_cols; t += warps) {
    const uint32_t v = col_variant[t];
    const uint32_t ab = pd.allele_off[v], A = pd.allele_off[v + 1] - ab;
    double* d = reinterpret_cast<double*>(desc + (size_t)t * stride);
    if (lane == 0) {
      double c[4] = {0, 0, 0, 0};
      if (t > col_chrom_begin[t]) transition_coefs(pd.positions[col_variant[t - 1]], pd.positions[v], n_sel, tp, c);
      d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; d[3] = c[3];
      double e[4] = {0, 0, 0, 0};
      if (t + 1 < col_chrom_end[t]) transition_coefs(pd.positions[v], pd.positions[col_variant[t + 1]], n_sel, tp, e);
      d[4] = e[0]; d[5] = e[1]; d[6] = e[2]; d[7] = e[3];
      uint32_t* h = reinterpret_cast<uint32_t*>(d + 8);
      h[0] = A;
      h[1] = v;
      // [9]: columns with more than HMM_FAST_A alleles: pointer to the emission table; others: the allele ids of the (up to
      // four) allele indices, 16 bits each.  [31]: offset of the variant's posterior row.  With both in the record the posterior
      // writer of the block kernel needs no dependent global loads (allele_off -> allele_ids, gl_off) on its column's path.
      unsigned long long w9 = (unsigned long long)(em + em_off[v]);
      if (A <= HMM_FAST_A) {
        w9 = 0;
        for (uint32_t i = 0; i < A; ++i) w9 |= (unsigned long long)pd.allele_ids[ab + i] << (16 * i);
      }
      *reinterpret_cast<unsigned long long*>(d + 9) = w9;
      *reinterpret_cast<unsigned long long*>(d + 31) = gl_off[v];
    }
    if (lane < 16) {
      const uint32_t i1 = lane >> 2, i2 = lane & 3;
      d[10 + lane] = (A <= HMM_FAST_A && i1 < A && i2 < A) ? em[em_off[v] + (size_t)i1 * A + i2] : 0.0;
    }
    uint16_t* aidx = reinterpret_cast<uint16_t*>(d + DESC_HEAD_DOUBLES);
    const uint32_t padded = (n_sel + 7u) & ~7u;
    unsigned long long* bits = reinterpret_cast<unsigned long long*>(d + DESC_BITS_AT);
    for (uint32_t q0 = 0; q0 < 320; q0 += 32) {  // 5 x 64 bits, zero beyond the selected paths
      const uint32_t q = q0 + lane;
      uint16_t idx = 0;
      if (q < n_sel) {
        const uint16_t a = pd.path_to_allele[(size_t)v * pd.P + sel[q]];
        for (uint32_t i = 0; i < A; ++i)
          if (pd.allele_ids[ab + i] == a) idx = (uint16_t)i;
      }
      if (q < padded) aidx[q] = idx;
      const unsigned b = __ballot_sync(0xffffffffu, (idx & 1u) != 0);
      if (lane == 0) reinterpret_cast<unsigned*>(bits)[q0 >> 5] = b;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// fill (src/commands.cpp:113-137, src/kmerparser.cpp:30-49)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_kmer_counts_kernel(const uint64_t* __restrict__ codes, uint64_t n, uint32_t k,
                                                                const KmerBucket* __restrict__ slots,
                                                                uint64_t cap, uint32_t q, uint32_t sh, uint16_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (uint16_t)table_lookup(codes[i], k, slots, cap, q, sh);  // size_t -> unsigned short (commands.cpp:118,130)
}

__global__ void __launch_bounds__(256) fill_coverage_kernel(const uint32_t* __restrict__ flank_off, const uint64_t* __restrict__ flank_codes,
                                                             uint32_t V, uint32_t k, const KmerBucket* __restrict__ slots,
                                                             uint64_t cap, uint32_t q, uint32_t sh,
                                                             uint64_t peak, uint16_t* __restrict__ coverage) {
  const uint64_t min_cov = peak / 4, max_cov = peak * 4;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    uint64_t total_cov = 0, total_kmers = 0;
    for (uint32_t f = flank_off[v]; f < flank_off[v + 1]; ++f) {
      const uint64_t c = table_lookup(flank_codes[f], k, slots, cap, q, sh);
      if (c < min_cov || c > max_cov) continue;
      total_cov += c;
      total_kmers += 1;
    }
    coverage[v] = (uint16_t)((total_kmers > 0 && total_cov > 0) ? total_cov / total_kmers : peak);
  }
}

// -------------------------------------------------------------------------------------------------
// finalize: per variant normalisation, reference scale, likeliest genotype and quality.
// -------------------------------------------------------------------------------------------------
struct FinalizeArgs {
  uint32_t V;
  const uint8_t* is_column;
  const uint32_t* variant_col;  // column index of each variant (valid if is_column)
  const uint32_t* col_variant;
  const uint32_t* col_chrom_begin;
  const uint32_t* col_chrom_end;
  const uint8_t* desc;
  uint32_t desc_stride;
  const double* tot_fwd;
  const double* tot_bwd;
  const double* log_scale;
  const uint64_t* gl_off;
  const uint32_t* allele_off;
  const uint16_t* allele_ids;
  const uint8_t* allele_undef;
  const uint32_t* kmer_off;
  const uint16_t* coverage;
  double* post;  // in: raw posterior, out: likelihoods
  int16_t* genotype;
  uint32_t* quality;
  uint16_t* unique_kmers;
  uint16_t* coverage_out;
  int normalize;
  uint32_t P;
  // subset combination (run_genotyping, src/commands.cpp:166-176): mode 1 adds the reference-scale likelihoods of this
  // subset to `acc` and ORs the column flags into `col_any` instead of writing results
  int mode;
  double* acc;
  uint8_t* col_any;
};

__device__ __forceinline__ double log_pow2_scale(double T) { return log(pow2_scale_of(T)); }

__global__ void __launch_bounds__(128) finalize_kernel(FinalizeArgs a) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < a.V; v += gridDim.x * blockDim.x) {
    a.unique_kmers[v] = (uint16_t)(a.kmer_off[v + 1] - a.kmer_off[v]);
    a.coverage_out[v] = a.coverage[v];
    double* L = a.post + a.gl_off[v];
    const uint32_t n = (uint32_t)(a.gl_off[v + 1] - a.gl_off[v]);
    uint32_t nr = 0;
    while (nr * (nr + 1) / 2 < n) ++nr;  // nr_alleles = max allele id + 1
    const uint32_t ab = a.allele_off[v], A = a.allele_off[v + 1] - ab;
    double sum = 0.0;
    for (uint32_t g = 0; g < n; ++g) sum += L[g];

    // ---- genotype / quality from the NORMALISED likelihoods restricted to defined alleles ----
    // (graph.cpp:206-240 -> genotypingresult.cpp:70-96,149-180,118-137)
    auto undefined = [&](uint32_t id) {
      for (uint32_t i = 0; i < A; ++i)
        if (a.allele_ids[ab + i] == id) return a.allele_undef[ab + i] != 0;
      return false;
    };
    int16_t g1 = -1, g2 = -1;
    uint32_t gq = 0;
    if (!a.is_column[v] || n == 0) {
      // no likelihoods: 0/0 with likelihood 1 (graph.cpp:225-227); prob_wrong == 0 -> 10000
      g1 = 0; g2 = 0; gq = 10000;
    } else if (sum > 0.0) {
      double ssum = 0.0;
      bool missing = false;
      for (uint32_t a2 = 0; a2 < nr; ++a2) {
        const bool u2 = a2 > 0 && undefined(a2);
        missing |= u2;
        for (uint32_t a1 = 0; a1 <= a2; ++a1) {
          const bool u1 = a1 > 0 && undefined(a1);
          if (!u1 && !u2) ssum += L[a2 * (a2 + 1) / 2 + a1];
        }
      }
      const double denom = missing ? (ssum > 0.0 ? ssum : sum) : sum;  // renormalise only if sum > 0 (:94)
      double best = 0.0, others = 0.0;
      int b1 = 0, b2 = 0, d1 = 0, d2 = 0;  // d*: index within the defined-allele list
      bool have = false;
      d2 = 0;
      for (uint32_t a2 = 0; a2 < nr; ++a2) {
        if (a2 > 0 && undefined(a2)) continue;
        d1 = 0;
        for (uint32_t a1 = 0; a1 <= a2; ++a1) {
          if (a1 > 0 && undefined(a1)) continue;
          const double l = L[a2 * (a2 + 1) / 2 + a1] / denom;
          if (!have || l > best) { best = l; b1 = d1; b2 = d2; have = true; }
          ++d1;
        }
        ++d2;
      }
      bool unique = true;
      d2 = 0;
      for (uint32_t a2 = 0; a2 < nr; ++a2) {
        if (a2 > 0 && undefined(a2)) continue;
        d1 = 0;
        for (uint32_t a1 = 0; a1 <= a2; ++a1) {
          if (a1 > 0 && undefined(a1)) continue;
          const double l = L[a2 * (a2 + 1) / 2 + a1] / denom;
          if (!(d1 == b1 && d2 == b2)) {
            others += l;
            if (fabs(l - best) < 0.0000000001) unique = false;
          }
          ++d1;
        }
        ++d2;
      }
      if (unique && best > 0.0) {
        g1 = (int16_t)b1;
        g2 = (int16_t)b2;
        // prob_wrong = 1.0L - best: best (>= 1/2) sits on the x87 grid of 2^-64, so prob_wrong is a multiple of it
        const double grid = best >= 0.5 ? 5.421010862427522e-20 : 2.710505431213761e-20;
        const double wrong = rint(others / grid) * grid;
        gq = wrong > 0.0 ? (uint32_t)(-10.0 * log10(wrong)) : 10000u;
      }
    }
    a.genotype[2 * v] = g1;
    a.genotype[2 * v + 1] = g2;
    a.quality[v] = gq;

    // ---- likelihood output ----
    if (!a.is_column[v]) continue;
    if (a.normalize) {
      if (sum > 0.0)
        for (uint32_t g = 0; g < n; ++g) L[g] /= sum;
      continue;
    }
    // reference scale alpha_hat * b * forward_norm (hmm.cpp:368); derivation in DESIGN.md "scales"
    const uint32_t t = a.variant_col[v];
    const uint32_t cbeg = a.col_chrom_begin[t], cend = a.col_chrom_end[t];
    const double S = (double)a.P * (double)a.P;
    double logfac = 0.0;
    const double TFt = a.tot_fwd[t];
    if (TFt > 0.0) {
      logfac += a.log_scale[v];  // m_t
      if (t > cbeg && a.tot_fwd[t - 1] > 0.0) logfac -= log_pow2_scale(a.tot_fwd[t - 1]) + log(a.tot_fwd[t - 1]);  // phi_t
    }
    if (t + 1 < cend) {  // lambda_t
      const double TY1 = a.tot_bwd[t + 1];
      if (TY1 > 0.0) {
        double loglam = log_pow2_scale(TY1) - a.log_scale[a.col_variant[t + 1]];
        // NP_{t+1} = sum of pre_{t+1}
        if (t + 2 >= cend) {
          loglam += log(S);
        } else {
          const double TY2 = a.tot_bwd[t + 2];
          if (TY2 > 0.0) {
            const double kappa = reinterpret_cast<const double*>(a.desc + (size_t)(t + 1) * a.desc_stride)[7];
            loglam += log(kappa) + log(TY2) + log_pow2_scale(TY2);
          }  // else: beta_hat_{t+1} uniform -> lambda_t = sb_t / m_{t+1}
        }
        logfac -= loglam;
      }
    }
    const double fac = exp(logfac);
    if (a.mode == 1) {
      double* acc = a.acc + a.gl_off[v];
      for (uint32_t g = 0; g < n; ++g) acc[g] += L[g] * fac;  // GenotypingResult::combine (genotypingresult.cpp:193-198)
      a.col_any[v] = 1;
    } else {
      for (uint32_t g = 0; g < n; ++g) L[g] *= fac;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// launch helpers: tile configuration by number of selected paths
// -------------------------------------------------------------------------------------------------
#ifndef PG_SKELETON_CLUSTER_DEFAULT
#define PG_SKELETON_CLUSTER_DEFAULT 0
#endif

struct TileCfg {
  int id, L, CPL, RPW, nwarps, nthreads;
};

static bool pick_cfg(uint32_t P, TileCfg& c) {
  // {lanes per row, columns per lane, rows per thread}; the first two run a whole chain in ONE warp
  static const int cfgs[][3] = {{1, 9, 1}, {1, 16, 1}, {4, 5, 1}, {4, 9, 1}, {4, 17, 1}, {8, 17, 2}, {32, 9, 8}};
  for (int i = 0; i < 7; ++i) {
    const int L = cfgs[i][0], CPL = cfgs[i][1], RPW = cfgs[i][2];
    const int rows_per_warp = (32 / L) * RPW;
    const int nw = ((int)P + rows_per_warp - 1) / rows_per_warp;
    const int max_warps = i < 2 ? 1 : 32;
    static const int nts[] = {32, 32, 96, 160, 288, 544, 1024};
    if (L * CPL >= (int)P && nw <= max_warps && nw * 32 <= nts[i]) {
      c = {i, L, CPL, RPW, nw, nts[i]};
      return true;
    }
  }
  return false;
}

// Resident CTAs per SM the block kernel is compiled for (register cap = 64K / (MINB * NT)).  PG_BLOCK_MINB overrides
// the default (tuning knob); every variant computes the same thing.
static int block_minb(int NT) {
  static const int env = [] { const char* e = getenv("PG_BLOCK_MINB"); return e ? atoi(e) : 0; }();
  if (env >= 1 && env <= 4) return env;
  return NT <= 160 ? 3 : NT <= 288 ? 2 : 1;  // measured: H=32 blocks 0.91 -> 0.74 ms at 3 CTAs/SM; H=64 spills at 3
}

template <int L, int CPL, int RPW, int NT, int MINB>
static cudaError_t launch_block(const ChainParams& p, int grid_blocks, cudaStream_t s) {
  block_kernel<L, CPL, RPW, NT, MINB><<<grid_blocks, NT, sizeof(ChainSmem), s>>>(p);
  return cudaGetLastError();
}

template <int L, int CPL, int RPW, int NT>
static cudaError_t launch_pair(const ChainParams& p, uint32_t n_chrom, int grid_blocks, cudaStream_t s, bool skeleton, bool blocks) {
  const size_t smem = sizeof(ChainSmem);
  if (skeleton) skeleton_kernel<L, CPL, RPW, NT><<<dim3(n_chrom, 2), NT, smem, s>>>(p);
  if (blocks) {
    if (NT > 288) return launch_block<L, CPL, RPW, NT, 1>(p, grid_blocks, s);
    switch (block_minb(NT)) {
      case 1: return launch_block<L, CPL, RPW, NT, 1>(p, grid_blocks, s);
      case 3: return launch_block<L, CPL, RPW, NT, (NT <= 288 ? 3 : 1)>(p, grid_blocks, s);
      case 4: return launch_block<L, CPL, RPW, NT, (NT <= 160 ? 4 : NT <= 288 ? 3 : 1)>(p, grid_blocks, s);
      default: return launch_block<L, CPL, RPW, NT, (NT <= 288 ? 2 : 1)>(p, grid_blocks, s);
    }
  }
  return cudaGetLastError();
}

// checkpoint walk split over a cluster of C CTAs (hmm_kernels.cuh): NT threads serve ceil(P / C) rows
template <int L, int CPL, int NT>
static cudaError_t launch_cluster(const ChainParams& p, uint32_t n_chrom, int C, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_chrom * (uint32_t)C, 2, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = sizeof(ChainSmem);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, skeleton_cluster_kernel<L, CPL, 1, NT>, p);
}

// Cluster size of the checkpoint walk: PG_SKELETON_CLUSTER = 0 (single CTA), 2 or 4.  Returns true if the walk was
// launched on a cluster (configurations with one row per thread, 20 < P <= 68).
static bool try_cluster_skeleton(const ChainParams& p, uint32_t n_chrom, int cfg_id, cudaStream_t s, cudaError_t& err) {
  const char* e = getenv("PG_SKELETON_CLUSTER");
  const int C = e ? atoi(e) : PG_SKELETON_CLUSTER_DEFAULT;
  if (C != 2 && C != 4) return false;
  const int rows = ((int)p.P + C - 1) / C, nw = (rows + 7) / 8;   // L = 4: 8 rows per warp
  if (cfg_id == 3) {                                              // {4, 9, 1}: P <= 36
    if (nw <= 2) err = launch_cluster<4, 9, 64>(p, n_chrom, C, s);
    else if (nw == 3) err = launch_cluster<4, 9, 96>(p, n_chrom, C, s);
    else return false;
    return true;
  }
  if (cfg_id == 4) {                                              // {4, 17, 1}: P <= 68
    if (nw <= 3) err = launch_cluster<4, 17, 96>(p, n_chrom, C, s);
    else if (nw <= 5) err = launch_cluster<4, 17, 160>(p, n_chrom, C, s);
    else return false;
    return true;
  }
  return false;
}

// Checkpoint walk of 16 < P <= 68 paths: the LEAN walk of hmm_kernels.cuh (TMA descriptor ring, one row per thread, 17 columns
// per lane).  PG_SKELETON_TILE = 0 selects the generic walk (the block kernel's tile configuration) instead.
// Measured per 400 k columns (scripts/bench_hmm.py, profiles/r2_skeleton.md): H = 32 24.0 -> 15.0 ms, H = 64 37.6 -> 28.8 ms.
// Not kept (measured, no gain): the generic walk instantiated with one / two lanes per row (23.3 / 27.2 ms at H = 32, and the
// different summation order of whole rows in one thread cost 7e-6 relative on one oracle case), 9 columns per lane for the
// lean walk (16.9 / 37.0 ms), the walk split over a thread-block cluster (38 / 42-47 ms; kept as PG_SKELETON_CLUSTER).
#ifndef PG_SKELETON_TILE_DEFAULT
#define PG_SKELETON_TILE_DEFAULT 3
#endif
static bool try_lean_skeleton(const ChainParams& p, uint32_t n_chrom, cudaStream_t s, cudaError_t& err) {
  const char* e = getenv("PG_SKELETON_TILE");
  const int mode = e ? atoi(e) : PG_SKELETON_TILE_DEFAULT;
  const int P = (int)p.P;
  if (mode != 3) return false;
  const size_t smem = ((sizeof(ChainSmem) + 15) & ~size_t(15)) + HMM_NSLOT * 8;
  const dim3 grid(n_chrom, 2);
  if (P > 16 && P <= 34) {
    if (34 - P <= 4) skeleton_lean_kernel<2, 17, 96, 4><<<grid, 96, smem, s>>>(p);
    else skeleton_lean_kernel<2, 17, 96, 17><<<grid, 96, smem, s>>>(p);
  } else if (P > 34 && P <= 68) {
    if (68 - P <= 4) skeleton_lean_kernel<4, 17, 288, 4><<<grid, 288, smem, s>>>(p);
    else skeleton_lean_kernel<4, 17, 288, 17><<<grid, 288, smem, s>>>(p);
  } else {
    return false;
  }
  err = cudaGetLastError();
  return true;
}

template <int L, int CPL, int RPW, int NT>
static int occupancy_of() {
  int n = 0;
  if (NT > 288) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, block_kernel<L, CPL, RPW, NT, 1>, NT, sizeof(ChainSmem));
    return n;
  }
  switch (block_minb(NT)) {
    case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, block_kernel<L, CPL, RPW, NT, 1>, NT, sizeof(ChainSmem)); break;
    case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, block_kernel<L, CPL, RPW, NT, (NT <= 288 ? 3 : 1)>, NT, sizeof(ChainSmem)); break;
    case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, block_kernel<L, CPL, RPW, NT, (NT <= 160 ? 4 : NT <= 288 ? 3 : 1)>, NT, sizeof(ChainSmem)); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, block_kernel<L, CPL, RPW, NT, (NT <= 288 ? 2 : 1)>, NT, sizeof(ChainSmem));
  }
  return n;
}

}  // namespace pg

using namespace pg;

// =================================================================================================
// engine
// =================================================================================================
struct pg_engine {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10] = {};
  pg_timings tm = {};
  // panel (all chromosomes concatenated)
  uint32_t n_chrom = 0, V = 0, P = 0;
  uint64_t K = 0, A = 0, F = 0;
  std::vector<uint32_t> chrom_v0;  // variant offset of each chromosome (n_chrom + 1)
  // multi-sample batching (pg_hmm_run_samples): per-sample ProbabilityTables for the emission step of the next engine_hmm
  const pg_probtable* const* sample_tables = nullptr;
  uint32_t n_sample_tables = 0, chroms_per_sample = 0;
  std::vector<uint64_t> chrom_k0, chrom_gl0, chrom_f0;
  DevBuf<uint64_t> positions;
  DevBuf<uint16_t> path_to_allele, coverage, kmer_counts, allele_ids, allele_koff;
  DevBuf<uint32_t> kmer_off, allele_off, allele_kmask, flank_off;
  DevBuf<uint8_t> allele_undef;
  DevBuf<uint64_t> kmer_codes, flank_codes, gl_off, em_off;
  bool has_codes = false;
  uint64_t GL = 0, EM = 0;
  // scratch
  DevBuf<double> log_p, em, log_scale, ckpt_fwd, ckpt_bwd, tot_fwd, tot_bwd, block_buf, post, post_acc;
  DevBuf<uint8_t> is_column, desc, col_any;
  DevBuf<uint16_t> sel, unique_kmers, coverage_out;
  DevBuf<uint32_t> col_variant, col_cbeg, col_cend, variant_col, work_counter, quality;
  DevBuf<int16_t> genotype;
  DevBuf<ChromCols> chroms;
  DevBuf<uint2> jobs;
  // parallel-in-time checkpoint path (hmm_scan.cuh)
  DevBuf<TJob> tjobs;
  DevBuf<ScanChrom> scan_chroms;
  DevBuf<double> scan_mats;
  DevBuf<uint32_t> seq_flags;
  bool scan_attr_set = false;
  // column structure of the loaded panel (depends on the panel and the selected paths only, not on the counts): kept
  // across calls so a resident engine does not rebuild / re-upload it for every sample
  // ProbabilityTable(peak/4, peak*4, 2*peak, regularization) of the last sample: samples of the same k-mer coverage
  // peak share it (about 0.15 ms of x87 long double arithmetic for a peak of 7, more for deeper samples)
  pg_probtable table = {};
  bool table_valid = false;
  uint64_t panel_epoch = 0;
  struct ColCache {
    bool valid = false;
    uint64_t epoch = 0;
    std::vector<uint16_t> sel;
    uint32_t B = 0;
    bool scan = false;
    std::vector<ChromCols> chroms;
    uint32_t C = 0, nblk = 0, n_jobs = 0, n_tj = 0;
  } cols;
  std::vector<uint8_t> h_is_column;
  pg_counter* cached_counter = nullptr;  // reused across pg_engine_run_resident calls
};

static PanelDev panel_view(const pg_engine* e) {
  PanelDev pd;
  pd.V = e->V; pd.P = e->P; pd.K = (uint32_t)e->K; pd.A = (uint32_t)e->A; pd.F = (uint32_t)e->F;
  pd.positions = e->positions.p; pd.path_to_allele = e->path_to_allele.p; pd.coverage = e->coverage.p;
  pd.kmer_off = e->kmer_off.p; pd.kmer_counts = e->kmer_counts.p; pd.allele_off = e->allele_off.p;
  pd.allele_ids = e->allele_ids.p; pd.allele_undef = e->allele_undef.p; pd.allele_koff = e->allele_koff.p;
  pd.allele_kmask = e->allele_kmask.p; pd.kmer_codes = e->kmer_codes.p; pd.flank_off = e->flank_off.p;
  pd.flank_codes = e->flank_codes.p;
  return pd;
}

extern "C" pg_engine* pg_engine_create(int device) {
  clear_error();
  if (check_device(device) != PG_OK) return nullptr;
  DeviceGuard g(device);
  pg_engine* e = new pg_engine();
  e->device = device;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  e->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    fail(PG_ERR_CUDA, "cudaStreamCreate failed");
    delete e;
    return nullptr;
  }
  for (auto& ev : e->ev) cudaEventCreate(&ev);
  return e;
}

extern "C" void pg_engine_destroy(pg_engine* e) {
  if (!e) return;
  {
    DeviceGuard g(e->device);
    if (e->cached_counter) pg_count_destroy(e->cached_counter);
    if (e->table_valid) pg_probtable_free(&e->table);
    if (e->stream) {
      cudaStreamSynchronize(e->stream);
      cudaStreamDestroy(e->stream);
    }
    for (auto& ev : e->ev)
      if (ev) cudaEventDestroy(ev);
    // DevBufs must be released while the right device is current
    e->positions.release(); e->path_to_allele.release(); e->coverage.release(); e->kmer_counts.release();
    e->allele_ids.release(); e->allele_koff.release(); e->kmer_off.release(); e->allele_off.release();
    e->allele_kmask.release(); e->flank_off.release(); e->allele_undef.release(); e->kmer_codes.release();
    e->flank_codes.release(); e->gl_off.release(); e->em_off.release(); e->log_p.release(); e->em.release();
    e->log_scale.release(); e->ckpt_fwd.release(); e->ckpt_bwd.release(); e->tot_fwd.release(); e->tot_bwd.release();
    e->block_buf.release(); e->post.release(); e->is_column.release(); e->desc.release(); e->sel.release();
    e->unique_kmers.release(); e->coverage_out.release(); e->col_variant.release(); e->col_cbeg.release();
    e->col_cend.release(); e->variant_col.release(); e->work_counter.release(); e->quality.release();
    e->genotype.release(); e->chroms.release(); e->jobs.release();
    e->post_acc.release(); e->col_any.release();
    e->tjobs.release(); e->scan_chroms.release(); e->scan_mats.release(); e->seq_flags.release();
  }
  delete e;
}

extern "C" int pg_engine_timings(const pg_engine* e, pg_timings* out) {
  if (!e || !out) return fail(PG_ERR_ARG, "null argument");
  *out = e->tm;
  return PG_OK;
}

// ---- panel upload: chromosomes are concatenated, CSR offsets rebased --------------------------------
template <class T>
static int upload_concat(DevBuf<T>& dst, const std::vector<const T*>& srcs, const std::vector<uint64_t>& counts, cudaStream_t s) {
  uint64_t total = 0;
  for (auto c : counts) total += c;
  PG_TRY(dst.reserve(std::max<uint64_t>(total, 1)));
  uint64_t off = 0;
  for (size_t i = 0; i < srcs.size(); ++i) {
    if (counts[i]) PG_CUDA(cudaMemcpyAsync(dst.p + off, srcs[i], counts[i] * sizeof(T), cudaMemcpyHostToDevice, s));
    off += counts[i];
  }
  return PG_OK;
}

static int upload_offsets(DevBuf<uint32_t>& dst, uint32_t n_chrom, const pg_panel* panels, const uint32_t* pg_panel::*member,
                          std::vector<uint64_t>& chrom_base, cudaStream_t s, std::vector<uint32_t>& host_tmp) {
  uint64_t V = 0;
  for (uint32_t c = 0; c < n_chrom; ++c) V += panels[c].n_variants;
  host_tmp.resize(V + 1);
  chrom_base.assign(n_chrom + 1, 0);
  uint64_t base = 0, vi = 0;
  for (uint32_t c = 0; c < n_chrom; ++c) {
    const uint32_t* off = panels[c].*member;
    chrom_base[c] = base;
    for (uint32_t v = 0; v < panels[c].n_variants; ++v) host_tmp[vi++] = (uint32_t)(base + off[v]);
    base += panels[c].n_variants ? off[panels[c].n_variants] : 0;
    if (base > 0xffffffffull) return fail(PG_ERR_ARG, "panel too large: more than 2^32 entries in one CSR array");
  }
  host_tmp[vi] = (uint32_t)base;
  chrom_base[n_chrom] = base;
  PG_TRY(dst.reserve(V + 1));
  PG_CUDA(cudaMemcpyAsync(dst.p, host_tmp.data(), (V + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  PG_CUDA(cudaStreamSynchronize(s));  // host_tmp is reused by the caller
  return PG_OK;
}

static int engine_load_panels(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_hmm_result* layouts,
                              bool need_counts, bool need_codes) {
  pg::NvtxRange nvtx_("pg: load panels");
  if (!e || !panels || n_chrom == 0) return fail(PG_ERR_ARG, "invalid panel arguments");
  DeviceGuard g(e->device);
  const uint32_t P = panels[0].n_paths;
  uint64_t V = 0;
  for (uint32_t c = 0; c < n_chrom; ++c) {
    if (panels[c].n_paths != P) return fail(PG_ERR_ARG, "all chromosomes must be covered by the same number of paths");
    V += panels[c].n_variants;
    if (panels[c].n_variants && (!panels[c].positions || !panels[c].path_to_allele || !panels[c].kmer_offsets || !panels[c].allele_offsets))
      return fail(PG_ERR_ARG, "panel has null arrays");
    if (need_codes && panels[c].n_variants && (!panels[c].kmer_codes || !panels[c].flank_offsets))
      return fail(PG_ERR_ARG, "panel lacks kmer_codes / flank arrays needed for the count fill");
  }
  if (P == 0) return fail(PG_ERR_ARG, "PanGenie-index: no haplotype paths given.");
  if (V > 0xfffffff0ull) return fail(PG_ERR_ARG, "too many variants");
  e->n_chrom = n_chrom;
  e->V = (uint32_t)V;
  e->P = P;
  ++e->panel_epoch;  // invalidates the cached column structure
  e->chrom_v0.assign(n_chrom + 1, 0);
  for (uint32_t c = 0; c < n_chrom; ++c) e->chrom_v0[c + 1] = e->chrom_v0[c] + panels[c].n_variants;
  cudaStream_t s = e->stream;
  std::vector<uint32_t> tmp;
  std::vector<uint64_t> abase;
  PG_TRY(upload_offsets(e->kmer_off, n_chrom, panels, &pg_panel::kmer_offsets, e->chrom_k0, s, tmp));
  PG_TRY(upload_offsets(e->allele_off, n_chrom, panels, &pg_panel::allele_offsets, abase, s, tmp));
  e->K = e->chrom_k0[n_chrom];
  e->A = abase[n_chrom];
  std::vector<uint64_t> nv(n_chrom), nvp(n_chrom), nk(n_chrom), na(n_chrom), nf(n_chrom, 0);
  for (uint32_t c = 0; c < n_chrom; ++c) {
    nv[c] = panels[c].n_variants;
    nvp[c] = (uint64_t)panels[c].n_variants * P;
    nk[c] = e->chrom_k0[c + 1] - e->chrom_k0[c];
    na[c] = abase[c + 1] - abase[c];
  }
#define PG_GATHER(T, member)                                   \
  std::vector<const T*> member##_src(n_chrom);                 \
  for (uint32_t c = 0; c < n_chrom; ++c) member##_src[c] = panels[c].member;
  PG_GATHER(uint64_t, positions)
  PG_GATHER(uint16_t, path_to_allele)
  PG_GATHER(uint16_t, allele_ids)
  PG_GATHER(uint8_t, allele_undefined)
  PG_GATHER(uint16_t, allele_kmer_offset)
  PG_GATHER(uint32_t, allele_kmer_mask)
  PG_TRY(upload_concat(e->positions, positions_src, nv, s));
  PG_TRY(upload_concat(e->path_to_allele, path_to_allele_src, nvp, s));
  PG_TRY(upload_concat(e->allele_ids, allele_ids_src, na, s));
  PG_TRY(upload_concat(e->allele_undef, allele_undefined_src, na, s));
  PG_TRY(upload_concat(e->allele_koff, allele_kmer_offset_src, na, s));
  PG_TRY(upload_concat(e->allele_kmask, allele_kmer_mask_src, na, s));
  PG_TRY(e->coverage.reserve(std::max<uint64_t>(V, 1)));
  PG_TRY(e->kmer_counts.reserve(std::max<uint64_t>(e->K, 1)));
  if (need_counts) {
    std::vector<const uint16_t*> cov_src(n_chrom), cnt_src(n_chrom);
    for (uint32_t c = 0; c < n_chrom; ++c) {
      const bool has_kmers = panels[c].n_variants && panels[c].kmer_offsets[panels[c].n_variants] > 0;
      if (panels[c].n_variants && (!panels[c].coverage || (has_kmers && !panels[c].kmer_counts))) return fail(PG_ERR_ARG, "panel lacks kmer_counts / coverage");
      cov_src[c] = panels[c].coverage;
      cnt_src[c] = panels[c].kmer_counts;
    }
    PG_TRY(upload_concat(e->coverage, cov_src, nv, s));
    PG_TRY(upload_concat(e->kmer_counts, cnt_src, nk, s));
  }
  e->has_codes = false;
  e->F = 0;
  if (need_codes) {
    PG_TRY(upload_offsets(e->flank_off, n_chrom, panels, &pg_panel::flank_offsets, e->chrom_f0, s, tmp));
    e->F = e->chrom_f0[n_chrom];
    for (uint32_t c = 0; c < n_chrom; ++c) nf[c] = e->chrom_f0[c + 1] - e->chrom_f0[c];
    PG_GATHER(uint64_t, kmer_codes)
    PG_GATHER(uint64_t, flank_codes)
    PG_TRY(upload_concat(e->kmer_codes, kmer_codes_src, nk, s));
    PG_TRY(upload_concat(e->flank_codes, flank_codes_src, nf, s));
    e->has_codes = true;
  }
#undef PG_GATHER
  // result layout (VCF-ordered likelihood rows) and compact emission layout (A x A per variant)
  if (layouts) {
    std::vector<uint64_t> gl(V + 1), emo(V + 1);
    e->chrom_gl0.assign(n_chrom + 1, 0);
    uint64_t gbase = 0, ebase = 0, vi = 0;
    for (uint32_t c = 0; c < n_chrom; ++c) {
      e->chrom_gl0[c] = gbase;
      if (panels[c].n_variants && !layouts[c].gl_offsets) return fail(PG_ERR_ARG, "result lacks gl_offsets (use pg_result_layout)");
      for (uint32_t v = 0; v < panels[c].n_variants; ++v) {
        gl[vi] = gbase + layouts[c].gl_offsets[v];
        emo[vi] = ebase;
        const uint64_t A = panels[c].allele_offsets[v + 1] - panels[c].allele_offsets[v];
        ebase += A * A;
        ++vi;
      }
      gbase += panels[c].n_variants ? layouts[c].gl_offsets[panels[c].n_variants] : 0;
    }
    gl[vi] = gbase;
    emo[vi] = ebase;
    e->chrom_gl0[n_chrom] = gbase;
    e->GL = gbase;
    e->EM = ebase;
    PG_TRY(e->gl_off.reserve(V + 1));
    PG_TRY(e->em_off.reserve(V + 1));
    PG_CUDA(cudaMemcpyAsync(e->gl_off.p, gl.data(), (V + 1) * 8, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(e->em_off.p, emo.data(), (V + 1) * 8, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaStreamSynchronize(s));
  }
  PG_CUDA(cudaStreamSynchronize(s));
  return PG_OK;
}

static size_t table_entries(const pg_probtable* t) { return t->log_p ? (size_t)(t->cov_max - t->cov_min) * t->count_max * 3 : 0; }

// `at`: offset (entries) inside e->log_p, which the caller has reserved
static int upload_table(pg_engine* e, const pg_probtable* t, TableDev& td, size_t at = 0, bool reserve = true) {
  td.cov_min = t->cov_min; td.cov_max = t->cov_max; td.count_max = t->count_max; td.reg = t->regularization;
  td.log_p = nullptr;
  const size_t n = table_entries(t);
  if (n) {
    if (reserve) PG_TRY(e->log_p.reserve(at + n));
    PG_CUDA(cudaMemcpyAsync(e->log_p.p + at, t->log_p, n * 8, cudaMemcpyHostToDevice, e->stream));
    td.log_p = e->log_p.p + at;
  }
  return PG_OK;
}

static int run_emission(pg_engine* e, const TableDev& td, uint32_t v_begin = 0, uint32_t v_end = ~0u) {
  PG_TRY(e->em.reserve(std::max<uint64_t>(e->EM, 1)));
  PG_TRY(e->log_scale.reserve(std::max<uint32_t>(e->V, 1)));
  v_end = std::min<uint32_t>(v_end, e->V);
  if (v_begin >= v_end) return PG_OK;
  const int grid = (int)std::min<uint64_t>(((uint64_t)(v_end - v_begin) * 32 + 255) / 256, (uint64_t)e->sm_count * 16);
  emission_kernel<<<grid, 256, 0, e->stream>>>(panel_view(e), td, e->em_off.p, e->em.p, e->log_scale.p, v_begin, v_end);
  count_launch();
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

// emission tables of the loaded panels: one ProbabilityTable for everything, or (multi-sample batching, pg_hmm_run_samples) one
// per sample = per group of `chroms_per_table` consecutive chromosomes
static int run_emissions(pg_engine* e, const pg_probtable* table, const pg_probtable* const* tables, uint32_t n_tables,
                         uint32_t chroms_per_table) {
  if (!tables) {
    TableDev td;
    PG_TRY(upload_table(e, table, td));
    return run_emission(e, td);
  }
  size_t total = 0;
  for (uint32_t i = 0; i < n_tables; ++i) total += table_entries(tables[i]);
  PG_TRY(e->log_p.reserve(std::max<size_t>(total, 1)));
  size_t at = 0;
  for (uint32_t i = 0; i < n_tables; ++i) {
    TableDev td;
    PG_TRY(upload_table(e, tables[i], td, at, false));
    at += table_entries(tables[i]);
    const uint32_t c0 = std::min<uint32_t>(i * chroms_per_table, e->n_chrom), c1 = std::min<uint32_t>((i + 1) * chroms_per_table, e->n_chrom);
    PG_TRY(run_emission(e, td, e->chrom_v0[c0], e->chrom_v0[c1]));
  }
  return PG_OK;
}

// ---- forward-backward over the loaded panel (counts + coverage resident) -------------------------------
static int engine_hmm(pg_engine* e, const pg_probtable* table, const pg_hmm_params* prm, int fin_mode = 0) {
  pg::NvtxRange nvtx_("pg: emission + forward-backward");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  const uint32_t V = e->V, Pfull = e->P;
  // selected paths (get_path_ids with only_include, biallelicuniquekmers.cpp:100-117)
  std::vector<uint16_t> sel;
  if (prm->only_paths) {
    for (uint32_t i = 0; i < prm->n_only_paths; ++i)
      if (prm->only_paths[i] < Pfull) sel.push_back(prm->only_paths[i]);
  } else {
    for (uint32_t i = 0; i < Pfull; ++i) sel.push_back((uint16_t)i);
  }
  const uint32_t P = (uint32_t)sel.size();
  if (P == 0) return fail(PG_ERR_ARG, "HMM::index_columns: column 0 is not covered by any paths.");
  if (P > HMM_PMAX) return fail(PG_ERR_ARG, "more than 256 selected paths: use path sub-sampling (-a) or haplotype sampling (-x)");
  TileCfg cfg;
  if (!pick_cfg(P, cfg)) return fail(PG_ERR_ARG, "no kernel configuration for this number of paths");
  // checkpoint schedule: P <= 9 propagates a basis through every block in parallel (hmm_scan.cuh), larger path
  // sets walk the chains sequentially (skeleton_kernel).  PG_SKELETON=seq|scan and PG_HMM_B are tuning/test knobs.
  const char* skel_env = getenv("PG_SKELETON");
  const bool scan_candidate = P <= (uint32_t)SCAN_CPL && !(skel_env && skel_env[0] == 's' && skel_env[1] == 'e');
  uint32_t B = 64;
  if (scan_candidate) B = 64; else if (P <= 12) B = 256; else if (P <= 36) B = 128;
  if (const char* be = getenv("PG_HMM_B")) {
    const long v = atol(be);
    if (v >= 2 && v <= HMM_BMAX) B = (uint32_t)v;
  }
  pg_engine::ColCache& cc = e->cols;
  const bool cached = cc.valid && cc.epoch == e->panel_epoch && cc.sel == sel && cc.B == B && cc.scan == scan_candidate;
  if (!cached) {
    cc.valid = false;
    PG_TRY(e->sel.reserve(P));
    PG_CUDA(cudaMemcpyAsync(e->sel.p, sel.data(), P * 2, cudaMemcpyHostToDevice, s));
  }

  cudaEventRecord(e->ev[0], s);
  PG_TRY(run_emissions(e, table, e->sample_tables, e->n_sample_tables, e->chroms_per_sample));
  cudaEventRecord(e->ev[1], s);

  // columns
  PG_TRY(e->is_column.reserve(std::max<uint32_t>(V, 1)));
  PG_TRY(e->post.reserve(std::max<uint64_t>(e->GL, 1)));
  PG_CUDA(cudaMemsetAsync(e->post.p, 0, std::max<uint64_t>(e->GL, 1) * 8, s));
  std::vector<uint2> jobs;
  std::vector<uint32_t> col_variant, col_cbeg, col_cend, variant_col;
  if (!cached) {
    e->h_is_column.assign(V, 0);
    if (V) {
      const int grid = (int)std::min<uint64_t>(((uint64_t)V * 32 + 255) / 256, (uint64_t)e->sm_count * 16);
      column_flag_kernel<<<grid, 256, 0, s>>>(panel_view(e), e->sel.p, P, e->is_column.p);
      count_launch();
      PG_CUDA(cudaMemcpyAsync(e->h_is_column.data(), e->is_column.p, V, cudaMemcpyDeviceToHost, s));
    }
    PG_CUDA(cudaStreamSynchronize(s));
    // host: column list, chromosome ranges, blocks, jobs
    variant_col.assign(std::max<uint32_t>(V, 1), 0);
    cc.chroms.assign(e->n_chrom, ChromCols());
    uint32_t nb_run = 0;
    for (uint32_t c = 0; c < e->n_chrom; ++c) {
      const uint32_t cb = (uint32_t)col_variant.size();
      for (uint32_t v = e->chrom_v0[c]; v < e->chrom_v0[c + 1]; ++v)
        if (e->h_is_column[v]) {
          variant_col[v] = (uint32_t)col_variant.size();
          col_variant.push_back(v);
        }
      const uint32_t ce = (uint32_t)col_variant.size();
      cc.chroms[c].col_begin = cb;
      cc.chroms[c].col_end = ce;
      cc.chroms[c].blk_begin = nb_run;
      cc.chroms[c].n_blocks = (ce - cb + B - 1) / B;
      for (uint32_t k = 0; k < cc.chroms[c].n_blocks; ++k) jobs.push_back(make_uint2(c, k));
      nb_run += cc.chroms[c].n_blocks;
      for (uint32_t t = cb; t < ce; ++t) {
        col_cbeg.push_back(cb);
        col_cend.push_back(ce);
      }
    }
    cc.C = (uint32_t)col_variant.size();
    cc.nblk = nb_run;
    cc.n_jobs = (uint32_t)jobs.size();
  }
  const std::vector<ChromCols>& chroms = cc.chroms;
  const uint32_t nblk = cc.nblk;
  const uint32_t C = cc.C;
  e->tm.hmm_columns = C;
  const uint32_t stride = (uint32_t)desc_bytes(P);
  const size_t PP = (size_t)cfg.CPL * cfg.RPW * cfg.nthreads;  // doubles per stored state (thread-major layout)
  PG_TRY(e->col_variant.reserve(std::max<uint32_t>(C, 1)));
  PG_TRY(e->col_cbeg.reserve(std::max<uint32_t>(C, 1)));
  PG_TRY(e->col_cend.reserve(std::max<uint32_t>(C, 1)));
  PG_TRY(e->variant_col.reserve(std::max<uint32_t>(V, 1)));
  PG_TRY(e->chroms.reserve(e->n_chrom));
  PG_TRY(e->jobs.reserve(std::max<size_t>(cc.n_jobs, 1)));
  PG_TRY(e->desc.reserve(std::max<size_t>((size_t)C * stride, 16)));
  PG_TRY(e->tot_fwd.reserve(std::max<uint32_t>(C, 1)));
  PG_TRY(e->tot_bwd.reserve(std::max<uint32_t>(C, 1)));
  const size_t CS = (size_t)P * P;  // doubles per checkpoint (dense)
  PG_TRY(e->ckpt_fwd.reserve(std::max<size_t>((size_t)nblk * CS, 1)));
  PG_TRY(e->ckpt_bwd.reserve(std::max<size_t>((size_t)nblk * CS, 1)));
  PG_TRY(e->work_counter.reserve(4));
  PG_TRY(e->genotype.reserve(std::max<uint32_t>(2 * V, 2)));
  PG_TRY(e->quality.reserve(std::max<uint32_t>(V, 1)));
  PG_TRY(e->unique_kmers.reserve(std::max<uint32_t>(V, 1)));
  PG_TRY(e->coverage_out.reserve(std::max<uint32_t>(V, 1)));
  if (!cached) {
    if (C) {
      PG_CUDA(cudaMemcpyAsync(e->col_variant.p, col_variant.data(), C * 4, cudaMemcpyHostToDevice, s));
      PG_CUDA(cudaMemcpyAsync(e->col_cbeg.p, col_cbeg.data(), C * 4, cudaMemcpyHostToDevice, s));
      PG_CUDA(cudaMemcpyAsync(e->col_cend.p, col_cend.data(), C * 4, cudaMemcpyHostToDevice, s));
      PG_CUDA(cudaMemcpyAsync(e->jobs.p, jobs.data(), jobs.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
    }
    if (V) PG_CUDA(cudaMemcpyAsync(e->variant_col.p, variant_col.data(), V * 4, cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaMemcpyAsync(e->chroms.p, chroms.data(), e->n_chrom * sizeof(ChromCols), cudaMemcpyHostToDevice, s));
    PG_CUDA(cudaStreamSynchronize(s));  // the host vectors above are locals
  }
  PG_CUDA(cudaMemsetAsync(e->work_counter.p, 0, 16, s));

  cudaEventRecord(e->ev[2], s);
  uint64_t block_launches = 0, scan_used = 0;
  if (C) {
    TransParams tp{prm->recombrate, prm->effective_N, prm->uniform};
    const int grid = (int)std::min<uint64_t>(((uint64_t)C * 32 + 255) / 256, (uint64_t)e->sm_count * 16);
    desc_build_kernel<<<grid, 256, 0, s>>>(panel_view(e), e->sel.p, P, e->col_variant.p, e->col_cend.p, e->col_cbeg.p, C,
                                            e->em_off.p, e->em.p, tp, e->gl_off.p, e->desc.p, stride);
    count_launch();
    PG_CUDA(cudaGetLastError());
    cudaEventRecord(e->ev[3], s);

    ChainParams cp;
    cp.state_stride = (uint32_t)PP;
    cp.ckpt_stride = (uint32_t)CS;
    cp.P = P; cp.B = B; cp.desc = e->desc.p; cp.desc_stride = stride; cp.chroms = e->chroms.p;
    cp.ckpt_fwd = e->ckpt_fwd.p; cp.ckpt_bwd = e->ckpt_bwd.p; cp.tot_fwd = e->tot_fwd.p; cp.tot_bwd = e->tot_bwd.p;
    cp.post = e->post.p; cp.gl_off = e->gl_off.p; cp.allele_off = e->allele_off.p; cp.allele_ids = e->allele_ids.p;
    cp.work_counter = e->work_counter.p; cp.jobs = e->jobs.p; cp.n_jobs = cc.n_jobs;
    cp.seq_flags = nullptr;
    bool need_skel = false;
    for (auto& ch : chroms) need_skel |= ch.n_blocks > 1;
    // ---- transfer jobs of the scan path
    const bool use_scan = scan_candidate && need_skel;
    ScanParams sp;
    memset(&sp, 0, sizeof(sp));
    if (use_scan) {
      const uint32_t NB = P * (P + 1) / 2;
      std::vector<TJob> tj;
      std::vector<ScanChrom> sch(e->n_chrom);
      for (uint32_t c = 0; c < e->n_chrom && !cached; ++c) {
        const ChromCols& ch = chroms[c];
        sch[c].n_tj = ch.n_blocks > 1 ? ch.n_blocks - 1 : 0;
        sch[c].pad = 0;
        sch[c].out_first[0] = ch.blk_begin + 1;                              // forward job k writes slot blk_begin + k + 1
        sch[c].out_first[1] = ch.blk_begin + (ch.n_blocks > 1 ? ch.n_blocks - 2 : 0);  // backward: k = n_blocks-1 first -> slot k - 1
        sch[c].tj_begin[0] = (uint32_t)tj.size();
        for (uint32_t k = 0; k + 1 < ch.n_blocks; ++k) {  // forward: block k -> F at its last column
          TJob j;
          memset(&j, 0, sizeof(j));
          const int cb = (int)(ch.col_begin + k * B), ce = cb + (int)B;
          j.chrom = c; j.dir = 0;
          j.t_first = k == 0 ? cb + 1 : cb;
          j.n_steps = (uint32_t)(ce - j.t_first);
          j.out_blk = ch.blk_begin + k + 1;
          tj.push_back(j);
        }
        sch[c].tj_begin[1] = (uint32_t)tj.size();
        for (uint32_t k = ch.n_blocks; k-- > 1;) {  // backward: block k -> Y at its first column
          TJob j;
          memset(&j, 0, sizeof(j));
          const int cb = (int)(ch.col_begin + k * B);
          const int ce = std::min<int>(cb + (int)B, (int)ch.col_end);
          j.chrom = c; j.dir = 1;
          j.t_first = k + 1 == ch.n_blocks ? (int)ch.col_end - 2 : ce - 1;
          j.n_steps = (uint32_t)(j.t_first - cb + 1);
          j.out_blk = ch.blk_begin + k - 1;
          tj.push_back(j);
        }
      }
      constexpr uint32_t G = 32 / SCAN_CPL;
      if (!cached) cc.n_tj = (uint32_t)tj.size();
      sp.n_tj = cc.n_tj;
      sp.n_groups = (NB + G - 1) / G;
      sp.n_items = sp.n_tj * sp.n_groups;
      sp.mat_stride = ((NB + 1u) * NB + 1u) & ~1u;
      PG_TRY(e->tjobs.reserve(std::max<size_t>(cc.n_tj, 1)));
      PG_TRY(e->scan_chroms.reserve(e->n_chrom));
      PG_TRY(e->scan_mats.reserve((size_t)sp.n_tj * sp.mat_stride));
      PG_TRY(e->seq_flags.reserve(e->n_chrom));
      if (!cached) {
        PG_CUDA(cudaMemcpyAsync(e->tjobs.p, tj.data(), tj.size() * sizeof(TJob), cudaMemcpyHostToDevice, s));
        PG_CUDA(cudaMemcpyAsync(e->scan_chroms.p, sch.data(), sch.size() * sizeof(ScanChrom), cudaMemcpyHostToDevice, s));
        PG_CUDA(cudaStreamSynchronize(s));  // tj / sch are stack-owned host vectors
      }
      PG_CUDA(cudaMemsetAsync(e->seq_flags.p, 0, e->n_chrom * sizeof(uint32_t), s));
      sp.tjobs = e->tjobs.p; sp.mats = e->scan_mats.p; sp.chroms = e->scan_chroms.p;
      sp.seq_flags = e->seq_flags.p;
      cp.seq_flags = e->seq_flags.p;
    }
    int occ = 1;
#define PG_OCC(L, CPL, RPW, NT) occ = occupancy_of<L, CPL, RPW, NT>();
    switch (cfg.id) {
      case 0: PG_OCC(1, 9, 1, 32) break;
      case 1: PG_OCC(1, 16, 1, 32) break;
      case 2: PG_OCC(4, 5, 1, 96) break;
      case 3: PG_OCC(4, 9, 1, 160) break;
      case 4: PG_OCC(4, 17, 1, 288) break;
      case 5: PG_OCC(8, 17, 2, 544) break;
      default: PG_OCC(32, 9, 8, 1024) break;
    }
#undef PG_OCC
    if (occ < 1) occ = 1;
    const int grid_blocks = (int)std::min<uint64_t>(cc.n_jobs, (uint64_t)e->sm_count * occ);
    PG_TRY(e->block_buf.reserve((size_t)grid_blocks * B * PP));
    cp.block_buf = e->block_buf.p;
    cudaError_t le;
    // skeleton and block kernels are launched separately so each phase is timed by its own events
    cudaEventRecord(e->ev[4], s);
#define PG_LAUNCH(L, CPL, RPW, NT, SK, BL) le = launch_pair<L, CPL, RPW, NT>(cp, e->n_chrom, grid_blocks, s, SK, BL);
#define PG_DISPATCH(SK, BL)                                    \
    switch (cfg.id) {                                          \
      case 0: PG_LAUNCH(1, 9, 1, 32, SK, BL) break;            \
      case 1: PG_LAUNCH(1, 16, 1, 32, SK, BL) break;           \
      case 2: PG_LAUNCH(4, 5, 1, 96, SK, BL) break;            \
      case 3: PG_LAUNCH(4, 9, 1, 160, SK, BL) break;           \
      case 4: PG_LAUNCH(4, 17, 1, 288, SK, BL) break;          \
      case 5: PG_LAUNCH(8, 17, 2, 544, SK, BL) break;          \
      default: PG_LAUNCH(32, 9, 8, 1024, SK, BL) break;        \
    }
    if (use_scan) {
      basis_kernel<SCAN_CPL><<<(sp.n_items + BASIS_WARPS - 1) / BASIS_WARPS, BASIS_WARPS * 32, 0, s>>>(cp, sp);
      if (!e->scan_attr_set) {  // per device: the engine owns one device
        PG_CUDA(cudaFuncSetAttribute(scan_kernel<SCAN_CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem<SCAN_CPL>)));
        e->scan_attr_set = true;
      }
      scan_kernel<SCAN_CPL><<<dim3(e->n_chrom, 2), SCAN_THREADS, sizeof(ScanSmem<SCAN_CPL>), s>>>(cp, sp);
      count_launch(2);
      PG_CUDA(cudaGetLastError());
      scan_used = 1;
    }
    if (need_skel) {
      le = cudaSuccess;
      if (!try_cluster_skeleton(cp, e->n_chrom, cfg.id, s, le) && !try_lean_skeleton(cp, e->n_chrom, s, le)) { PG_DISPATCH(true, false) }
      if (le != cudaSuccess) return fail(PG_ERR_CUDA, std::string("skeleton_kernel launch: ") + cudaGetErrorString(le));
      count_launch();
    }
    cudaEventRecord(e->ev[5], s);
    PG_DISPATCH(false, true)
    if (le != cudaSuccess) return fail(PG_ERR_CUDA, std::string("block_kernel launch: ") + cudaGetErrorString(le));
    count_launch();
    block_launches = 1;
#undef PG_DISPATCH
#undef PG_LAUNCH
    cudaEventRecord(e->ev[6], s);
  } else {
    for (int i = 3; i <= 6; ++i) cudaEventRecord(e->ev[i], s);
  }
  if (V) {
    FinalizeArgs fa;
    fa.V = V; fa.is_column = e->is_column.p; fa.variant_col = e->variant_col.p; fa.col_variant = e->col_variant.p;
    fa.col_chrom_begin = e->col_cbeg.p; fa.col_chrom_end = e->col_cend.p; fa.desc = e->desc.p; fa.desc_stride = stride;
    fa.tot_fwd = e->tot_fwd.p; fa.tot_bwd = e->tot_bwd.p; fa.log_scale = e->log_scale.p; fa.gl_off = e->gl_off.p;
    fa.allele_off = e->allele_off.p; fa.allele_ids = e->allele_ids.p; fa.allele_undef = e->allele_undef.p;
    fa.kmer_off = e->kmer_off.p; fa.coverage = e->coverage.p; fa.post = e->post.p; fa.genotype = e->genotype.p;
    fa.quality = e->quality.p; fa.unique_kmers = e->unique_kmers.p; fa.coverage_out = e->coverage_out.p;
    fa.normalize = prm->normalize; fa.P = P;
    fa.mode = fin_mode; fa.acc = e->post_acc.p; fa.col_any = e->col_any.p;
    const int grid = (int)std::min<uint64_t>(((uint64_t)V + 127) / 128, (uint64_t)e->sm_count * 16);
    finalize_kernel<<<grid, 128, 0, s>>>(fa);
    count_launch();
    PG_CUDA(cudaGetLastError());
  }
  cudaEventRecord(e->ev[7], s);
  PG_CUDA(cudaStreamSynchronize(s));
  float ms;
  cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->tm.emission_ms = ms;
  cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->tm.emission_ms += ms;  // descriptor build belongs to the emission stage
  cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]); e->tm.hmm_skeleton_ms = ms;
  cudaEventElapsedTime(&ms, e->ev[5], e->ev[6]); e->tm.hmm_blocks_ms = ms;
  cudaEventElapsedTime(&ms, e->ev[6], e->ev[7]); e->tm.finalize_ms = ms;
  e->tm.hmm_block_launches = block_launches;
  e->tm.hmm_scan_used = scan_used;
  if (!cached) {  // everything uploaded and consumed without error: later calls on the same panel may reuse it
    cc.epoch = e->panel_epoch;
    cc.sel = sel;
    cc.B = B;
    cc.scan = scan_candidate;
    cc.valid = true;
  }
  return PG_OK;
}

static int engine_fetch_results(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, pg_hmm_result* results) {
  pg::NvtxRange nvtx_("pg: fetch results");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  for (uint32_t c = 0; c < n_chrom; ++c) {
    const uint32_t v0 = e->chrom_v0[c], nv = panels[c].n_variants;
    if (!nv) continue;
    pg_hmm_result& r = results[c];
    const uint64_t g0 = e->chrom_gl0[c], gn = e->chrom_gl0[c + 1] - g0;
    if (r.likelihoods && gn) PG_CUDA(cudaMemcpyAsync(r.likelihoods, e->post.p + g0, gn * 8, cudaMemcpyDeviceToHost, s));
    if (r.is_column) PG_CUDA(cudaMemcpyAsync(r.is_column, e->is_column.p + v0, nv, cudaMemcpyDeviceToHost, s));
    if (r.genotype) PG_CUDA(cudaMemcpyAsync(r.genotype, e->genotype.p + 2 * (size_t)v0, nv * 4, cudaMemcpyDeviceToHost, s));
    if (r.quality) PG_CUDA(cudaMemcpyAsync(r.quality, e->quality.p + v0, nv * 4, cudaMemcpyDeviceToHost, s));
    if (r.unique_kmers) PG_CUDA(cudaMemcpyAsync(r.unique_kmers, e->unique_kmers.p + v0, nv * 2, cudaMemcpyDeviceToHost, s));
    if (r.coverage) PG_CUDA(cudaMemcpyAsync(r.coverage, e->coverage_out.p + v0, nv * 2, cudaMemcpyDeviceToHost, s));
  }
  PG_CUDA(cudaStreamSynchronize(s));
  return PG_OK;
}

static int engine_fill(pg_engine* e, const pg_counter* c, uint64_t peak) {
  pg::NvtxRange nvtx_("pg: fill counts");
  if (!e->has_codes) return fail(PG_ERR_ARG, "panel was loaded without k-mer codes");
  if (c->device != e->device) return fail(PG_ERR_ARG, "counter and engine live on different devices");
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  cudaEventRecord(e->ev[8], s);
  if (e->K) {
    const int grid = (int)std::min<uint64_t>((e->K + 255) / 256, (uint64_t)e->sm_count * 16);
    fill_kmer_counts_kernel<<<grid, 256, 0, s>>>(e->kmer_codes.p, e->K, c->k, c->slots, c->capacity, c->cap_q, c->cap_sh, e->kmer_counts.p);
    count_launch();
  }
  if (e->V) {
    const int grid = (int)std::min<uint64_t>(((uint64_t)e->V + 255) / 256, (uint64_t)e->sm_count * 16);
    fill_coverage_kernel<<<grid, 256, 0, s>>>(e->flank_off.p, e->flank_codes.p, e->V, c->k, c->slots, c->capacity, c->cap_q, c->cap_sh, peak, e->coverage.p);
    count_launch();
  }
  PG_CUDA(cudaGetLastError());
  cudaEventRecord(e->ev[9], s);
  PG_CUDA(cudaStreamSynchronize(s));
  float ms;
  cudaEventElapsedTime(&ms, e->ev[8], e->ev[9]);
  e->tm.fill_ms = ms;
  return PG_OK;
}

static int engine_fetch_counts(pg_engine* e, uint32_t n_chrom, pg_panel* panels) {
  DeviceGuard g(e->device);
  for (uint32_t c = 0; c < n_chrom; ++c) {
    const uint32_t nv = panels[c].n_variants;
    if (!nv) continue;
    const uint64_t k0 = e->chrom_k0[c], kn = e->chrom_k0[c + 1] - k0;
    if (panels[c].kmer_counts && kn) PG_CUDA(cudaMemcpyAsync(panels[c].kmer_counts, e->kmer_counts.p + k0, kn * 2, cudaMemcpyDeviceToHost, e->stream));
    if (panels[c].coverage) PG_CUDA(cudaMemcpyAsync(panels[c].coverage, e->coverage.p + e->chrom_v0[c], nv * 2, cudaMemcpyDeviceToHost, e->stream));
  }
  PG_CUDA(cudaStreamSynchronize(e->stream));
  return PG_OK;
}

// the sample's ProbabilityTable (src/commands.cpp:846), rebuilt only when the peak or the regularisation changes
static int engine_table(pg_engine* e, uint64_t peak, double regularization, const pg_probtable** out) {
  const uint16_t cmin = (uint16_t)(peak / 4), cmax = (uint16_t)(peak * 4), nmax = (uint16_t)(2 * peak);
  if (!(e->table_valid && e->table.cov_min == cmin && e->table.cov_max == cmax && e->table.count_max == nmax &&
        e->table.regularization == regularization)) {
    if (e->table_valid) pg_probtable_free(&e->table);
    e->table_valid = false;
    PG_TRY(pg_probtable_init(&e->table, cmin, cmax, nmax, regularization));
    e->table_valid = true;
  }
  *out = &e->table;
  return PG_OK;
}

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" int pg_hmm_run(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                          const pg_hmm_params* params, pg_hmm_result* results) {
  clear_error();
  if (!e || !panels || !table || !params || !results) return fail(PG_ERR_ARG, "null argument");
  const uint64_t l0 = g_launches;
  memset(&e->tm, 0, sizeof(e->tm));
  PG_TRY(engine_load_panels(e, n_chrom, panels, results, true, false));
  PG_TRY(engine_hmm(e, table, params));
  PG_TRY(engine_fetch_results(e, n_chrom, panels, results));
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

// Genotyping over several path subsets (`-a`, src/commands.cpp:916-993): every subset is run un-normalised, the
// likelihoods are added up (run_genotyping, :166-176) and normalised at the end (:982-988); GT / GQ follow from the sum.
static int engine_hmm_subsets(pg_engine* e, const pg_probtable* table, const pg_hmm_params* prm, uint32_t n_subsets,
                              const uint32_t* subset_offsets, const uint16_t* subset_paths) {
  DeviceGuard g(e->device);
  cudaStream_t s = e->stream;
  const uint32_t V = e->V;
  PG_TRY(e->post_acc.reserve(std::max<uint64_t>(e->GL, 1)));
  PG_TRY(e->col_any.reserve(std::max<uint32_t>(V, 1)));
  PG_CUDA(cudaMemsetAsync(e->post_acc.p, 0, std::max<uint64_t>(e->GL, 1) * 8, s));
  PG_CUDA(cudaMemsetAsync(e->col_any.p, 0, std::max<uint32_t>(V, 1), s));
  pg_timings acc_tm;
  memset(&acc_tm, 0, sizeof(acc_tm));
  for (uint32_t k = 0; k < n_subsets; ++k) {
    pg_hmm_params ps = *prm;
    ps.only_paths = subset_paths + subset_offsets[k];
    ps.n_only_paths = subset_offsets[k + 1] - subset_offsets[k];
    ps.normalize = 0;
    PG_TRY(engine_hmm(e, table, &ps, 1));
    acc_tm.emission_ms += e->tm.emission_ms; acc_tm.hmm_skeleton_ms += e->tm.hmm_skeleton_ms;
    acc_tm.hmm_blocks_ms += e->tm.hmm_blocks_ms; acc_tm.finalize_ms += e->tm.finalize_ms;
    acc_tm.hmm_columns += e->tm.hmm_columns; acc_tm.hmm_block_launches += e->tm.hmm_block_launches;
  }
  if (V) {
    // the sum becomes the result: normalise, genotype and quality as for a single run
    PG_CUDA(cudaMemcpyAsync(e->post.p, e->post_acc.p, std::max<uint64_t>(e->GL, 1) * 8, cudaMemcpyDeviceToDevice, s));
    PG_CUDA(cudaMemcpyAsync(e->is_column.p, e->col_any.p, V, cudaMemcpyDeviceToDevice, s));
    FinalizeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.V = V; fa.is_column = e->is_column.p; fa.gl_off = e->gl_off.p; fa.allele_off = e->allele_off.p;
    fa.allele_ids = e->allele_ids.p; fa.allele_undef = e->allele_undef.p; fa.kmer_off = e->kmer_off.p;
    fa.coverage = e->coverage.p; fa.post = e->post.p; fa.genotype = e->genotype.p; fa.quality = e->quality.p;
    fa.unique_kmers = e->unique_kmers.p; fa.coverage_out = e->coverage_out.p;
    fa.normalize = 1;  // the normalised branch reads none of the per-column arrays
    const int grid = (int)std::min<uint64_t>(((uint64_t)V + 127) / 128, (uint64_t)e->sm_count * 16);
    finalize_kernel<<<grid, 128, 0, s>>>(fa);
    count_launch();
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaStreamSynchronize(s));
  }
  e->cols.valid = false;  // is_column now holds the union over the subsets, not the last subset's columns
  e->tm.emission_ms = acc_tm.emission_ms; e->tm.hmm_skeleton_ms = acc_tm.hmm_skeleton_ms;
  e->tm.hmm_blocks_ms = acc_tm.hmm_blocks_ms; e->tm.finalize_ms = acc_tm.finalize_ms;
  e->tm.hmm_columns = acc_tm.hmm_columns; e->tm.hmm_block_launches = acc_tm.hmm_block_launches;
  return PG_OK;
}

extern "C" int pg_hmm_run_subsets(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                                  const pg_hmm_params* params, uint32_t n_subsets, const uint32_t* subset_offsets,
                                  const uint16_t* subset_paths, pg_hmm_result* results) {
  clear_error();
  if (!e || !panels || !table || !params || !results || !subset_offsets || !subset_paths || n_subsets == 0)
    return fail(PG_ERR_ARG, "null argument");
  const uint64_t l0 = g_launches;
  memset(&e->tm, 0, sizeof(e->tm));
  PG_TRY(engine_load_panels(e, n_chrom, panels, results, true, false));
  PG_TRY(engine_hmm_subsets(e, table, params, n_subsets, subset_offsets, subset_paths));
  PG_TRY(engine_fetch_results(e, n_chrom, panels, results));
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

extern "C" int pg_hmm_run_samples(pg_engine* e, uint32_t n_samples, uint32_t n_chrom, const pg_panel* panels,
                                  const uint16_t* const* kmer_counts, const uint16_t* const* coverage,
                                  const pg_probtable* const* tables, const pg_hmm_params* params, pg_hmm_result* results) {
  clear_error();
  if (!e || !panels || !kmer_counts || !coverage || !tables || !params || !results) return fail(PG_ERR_ARG, "null argument");
  if (n_samples == 0 || n_chrom == 0) return fail(PG_ERR_ARG, "no samples or no chromosomes");
  for (uint32_t i = 0; i < n_samples; ++i)
    if (!tables[i]) return fail(PG_ERR_ARG, "null ProbabilityTable");
  // S x C panel views over ONE copy of the structure on the host; every sample brings its own counts and coverages
  std::vector<pg_panel> views((size_t)n_samples * n_chrom);
  for (uint32_t smp = 0; smp < n_samples; ++smp)
    for (uint32_t c = 0; c < n_chrom; ++c) {
      const size_t i = (size_t)smp * n_chrom + c;
      if (!kmer_counts[i] || !coverage[i]) return fail(PG_ERR_ARG, "null counts / coverage of a sample");
      views[i] = panels[c];
      views[i].kmer_counts = const_cast<uint16_t*>(kmer_counts[i]);
      views[i].coverage = const_cast<uint16_t*>(coverage[i]);
    }
  const uint64_t l0 = g_launches;
  memset(&e->tm, 0, sizeof(e->tm));
  PG_TRY(engine_load_panels(e, n_samples * n_chrom, views.data(), results, true, false));
  e->sample_tables = tables; e->n_sample_tables = n_samples; e->chroms_per_sample = n_chrom;
  const int st = engine_hmm(e, tables[0], params);
  e->sample_tables = nullptr; e->n_sample_tables = 0; e->chroms_per_sample = 0;
  PG_TRY(st);
  PG_TRY(engine_fetch_results(e, n_samples * n_chrom, views.data(), results));
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

extern "C" int pg_emission_run(pg_engine* e, const pg_panel* panel, const pg_probtable* table, const uint64_t* em_offsets,
                               double* emissions, double* log_scale) {
  clear_error();
  if (!e || !panel || !table || !em_offsets || !emissions || !log_scale) return fail(PG_ERR_ARG, "null argument");
  const uint32_t V = panel->n_variants;
  // engine_load_panels needs a layout to size the compact emission array; likelihood layout is irrelevant here
  std::vector<uint64_t> gl(V + 1, 0);
  pg_hmm_result lay;
  memset(&lay, 0, sizeof(lay));
  lay.gl_offsets = gl.data();
  PG_TRY(engine_load_panels(e, 1, panel, &lay, true, false));
  DeviceGuard g(e->device);
  TableDev td;
  PG_TRY(upload_table(e, table, td));
  PG_TRY(run_emission(e, td));
  std::vector<double> em(std::max<uint64_t>(e->EM, 1)), ls(std::max<uint32_t>(V, 1));
  if (e->EM) PG_CUDA(cudaMemcpyAsync(em.data(), e->em.p, e->EM * 8, cudaMemcpyDeviceToHost, e->stream));
  if (V) PG_CUDA(cudaMemcpyAsync(ls.data(), e->log_scale.p, V * 8, cudaMemcpyDeviceToHost, e->stream));
  PG_CUDA(cudaStreamSynchronize(e->stream));
  uint64_t eo = 0;
  for (uint32_t v = 0; v < V; ++v) {  // scatter allele-index space -> dense allele-id space
    const uint32_t ab = panel->allele_offsets[v], A = panel->allele_offsets[v + 1] - ab;
    uint32_t maxa = 0;
    for (uint32_t i = 0; i < A; ++i) maxa = std::max<uint32_t>(maxa, panel->allele_ids[ab + i]);
    const size_t dim = (size_t)maxa + 1;
    double* out = emissions + em_offsets[v];
    std::fill(out, out + dim * dim, 0.0);
    for (uint32_t i1 = 0; i1 < A; ++i1)
      for (uint32_t i2 = 0; i2 < A; ++i2) out[panel->allele_ids[ab + i1] * dim + panel->allele_ids[ab + i2]] = em[eo + (size_t)i1 * A + i2];
    log_scale[v] = ls[v];
    eo += (uint64_t)A * A;
  }
  return PG_OK;
}

extern "C" int pg_fill_counts(pg_engine* e, const pg_counter* c, uint64_t kmer_abundance_peak, uint32_t n_chrom, pg_panel* panels) {
  clear_error();
  if (!e || !c || !panels) return fail(PG_ERR_ARG, "null argument");
  const uint64_t l0 = g_launches;
  PG_TRY(engine_load_panels(e, n_chrom, panels, nullptr, false, true));
  PG_TRY(engine_fill(e, c, kmer_abundance_peak));
  PG_TRY(engine_fetch_counts(e, n_chrom, panels));
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

extern "C" int pg_genotype_run(pg_engine* e, const pg_genotype_input* in, uint32_t n_chrom, pg_panel* panels,
                               const pg_hmm_params* params, pg_hmm_result* results, uint64_t* kmer_abundance_peak) {
  pg::NvtxRange nvtx_("pg_genotype_run");
  clear_error();
  if (!e || !in || !panels || !params || !results) return fail(PG_ERR_ARG, "null argument");
  const uint64_t l0 = g_launches;
  memset(&e->tm, 0, sizeof(e->tm));
  HostTrace tr("genotype_run");
  // 1) count (src/commands.cpp:829-833); the table and its staging buffers are kept across calls
  const uint64_t max_distinct = in->segments ? std::max<uint64_t>(in->segments_len, 1024) : std::max<uint64_t>(in->hash_size, 1024);
  if (e->cached_counter && (e->cached_counter->k != in->k || e->cached_counter->max_distinct < max_distinct)) {
    pg_count_destroy(e->cached_counter);
    e->cached_counter = nullptr;
  }
  if (!e->cached_counter) {
    e->cached_counter = pg_count_new(in->k, max_distinct, e->device);
    if (!e->cached_counter) return last_code();
  } else {
    PG_TRY(pg_count_clear(e->cached_counter));
  }
  tr.mark("clear");
  pg_counter* c = e->cached_counter;
  bool panels_loaded = false;
  if (in->segments) {
    // the panel upload (host-side flattening + 12 small copies) runs while the first read chunks cross PCIe
    const std::function<int()> upload = [&]() -> int {
      PG_TRY(engine_load_panels(e, n_chrom, panels, results, false, true));
      panels_loaded = true;
      return PG_OK;
    };
    PG_TRY(count_prime_update(c, in->segments, in->segments_len, in->reads, in->reads_len, &upload));
    e->tm.prime_ms = c->last_prime_ms;
  } else {
    PG_TRY(pg_count_feed(c, in->reads, in->reads_len, PG_OP_COUNT));
  }
  e->tm.count_ms = c->last_feed_ms;
  e->tm.count_probe_ms = c->last_probe_ms;
  e->tm.count_probe_passes = (uint64_t)c->n_probe;
  e->tm.kmers_counted = c->kmers_seen;
  e->tm.text_bytes = in->reads_len;
  tr.mark("count");
  // 2) histogram peak (:840); largest_peak == count_only_graph
  uint64_t peak = 0;
  PG_TRY(pg_count_compute_histogram(c, 10000, in->segments != nullptr, in->histogram_path, &peak));
  if (kmer_abundance_peak) *kmer_abundance_peak = peak;
  // 3) ProbabilityTable(peak/4, peak*4, 2*peak, regularization) (:846)
  const pg_probtable* table = nullptr;
  PG_TRY(engine_table(e, peak, in->regularization, &table));
  tr.mark("histogram+table");
  // 4) fill + 5) HMM, panel uploaded once
  if (!panels_loaded) PG_TRY(engine_load_panels(e, n_chrom, panels, results, false, true));
  tr.mark("load_panels");
  PG_TRY(engine_fill(e, c, peak));
  tr.mark("fill");
  PG_TRY(engine_hmm(e, table, params));
  tr.mark("hmm");
  PG_TRY(engine_fetch_counts(e, n_chrom, panels));
  PG_TRY(engine_fetch_results(e, n_chrom, panels, results));
  tr.mark("fetch");
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

// histogram peak -> ProbabilityTable -> fill -> emission + forward-backward on the loaded panels
static int engine_after_count(pg_engine* e, const pg_counter* c, bool largest_peak, double regularization,
                              const pg_hmm_params* params, uint64_t* kmer_abundance_peak) {
  uint64_t peak = 0;
  HostTrace tr("after_count");
  cudaEvent_t h0 = c->ev_t0, h1 = c->ev_t1;
  cudaEventRecord(h0, c->stream);
  int st = pg_count_compute_histogram(c, 10000, largest_peak, nullptr, &peak);
  cudaEventRecord(h1, c->stream);
  cudaEventSynchronize(h1);
  float hms = 0;
  cudaEventElapsedTime(&hms, h0, h1);
  if (st != PG_OK) return st;
  if (kmer_abundance_peak) *kmer_abundance_peak = peak;
  const pg_probtable* table = nullptr;
  PG_TRY(engine_table(e, peak, regularization, &table));
  tr.mark("histogram+table");
  PG_TRY(engine_fill(e, c, peak));
  tr.mark("fill");
  const double fill_ms = e->tm.fill_ms;
  PG_TRY(engine_hmm(e, table, params));
  tr.mark("hmm");
  e->tm.fill_ms = fill_ms;
  e->tm.histogram_ms = hms;
  return PG_OK;
}

extern "C" int pg_engine_load(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_hmm_result* layouts) {
  clear_error();
  if (!e || !panels || !layouts) return fail(PG_ERR_ARG, "null argument");
  return engine_load_panels(e, n_chrom, panels, layouts, false, true);
}

extern "C" int pg_engine_run_resident(pg_engine* e, const char* d_reads, uint64_t reads_len, const char* d_segments,
                                      uint64_t segments_len, uint32_t k, uint64_t hash_size, double regularization,
                                      const pg_hmm_params* params, uint64_t* kmer_abundance_peak) {
  pg::NvtxRange nvtx_("pg_engine_run_resident");
  clear_error();
  if (!e || !d_reads || !params) return fail(PG_ERR_ARG, "null argument");
  if (!e->has_codes) return fail(PG_ERR_ARG, "call pg_engine_load first");
  const uint64_t l0 = g_launches;
  memset(&e->tm, 0, sizeof(e->tm));
  HostTrace tr("run_resident");
  const uint64_t max_distinct = d_segments ? std::max<uint64_t>(segments_len, 1024) : std::max<uint64_t>(hash_size, 1024);
  if (e->cached_counter && (e->cached_counter->k != k || e->cached_counter->max_distinct < max_distinct)) {
    pg_count_destroy(e->cached_counter);
    e->cached_counter = nullptr;
  }
  if (!e->cached_counter) {
    e->cached_counter = pg_count_new(k, max_distinct, e->device);
    if (!e->cached_counter) return last_code();
  } else {
    PG_TRY(pg_count_clear(e->cached_counter));
  }
  tr.mark("clear");
  pg_counter* c = e->cached_counter;
  if (d_segments) {
    PG_TRY(count_prime_update(c, d_segments, segments_len, d_reads, reads_len));
    e->tm.prime_ms = c->last_prime_ms;
  } else {
    PG_TRY(pg_count_feed_device(c, d_reads, reads_len, PG_OP_COUNT));
  }
  tr.mark("count");
  e->tm.count_ms = c->last_feed_ms;
  e->tm.count_probe_ms = c->last_probe_ms;
  e->tm.count_probe_passes = (uint64_t)c->n_probe;
  e->tm.kmers_counted = c->kmers_seen;
  e->tm.text_bytes = reads_len;
  PG_TRY(engine_after_count(e, c, d_segments != nullptr, regularization, params, kmer_abundance_peak));
  tr.mark("after_count");
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

extern "C" int pg_engine_run_counted(pg_engine* e, const pg_counter* c, int largest_peak, double regularization,
                                     const pg_hmm_params* params, uint64_t* kmer_abundance_peak) {
  clear_error();
  if (!e || !c || !params) return fail(PG_ERR_ARG, "null argument");
  if (!e->has_codes) return fail(PG_ERR_ARG, "call pg_engine_load first");
  const uint64_t l0 = g_launches;
  const pg_timings keep = e->tm;
  PG_TRY(engine_after_count(e, c, largest_peak != 0, regularization, params, kmer_abundance_peak));
  e->tm.count_ms = keep.count_ms;
  e->tm.prime_ms = keep.prime_ms;
  e->tm.kernel_launches = g_launches - l0;
  return PG_OK;
}

extern "C" int pg_engine_fetch(pg_engine* e, uint32_t n_chrom, pg_panel* panels, pg_hmm_result* results) {
  clear_error();
  if (!e || !panels || !results) return fail(PG_ERR_ARG, "null argument");
  PG_TRY(engine_fetch_counts(e, n_chrom, panels));
  return engine_fetch_results(e, n_chrom, panels, results);
}
