/*
 * hmm_binding.cpp — the reference-side binding: a drop-in replacement for the reference's src/hmm.cpp
 * that keeps src/hmm.hpp UNCHANGED (same class, same constructor, same accessors) and forwards the
 * forward-backward pass to the C-ABI of include/pangenie_b200.h.  A PanGenie maintainer adds this file
 * to PanGenieLib instead of hmm.cpp and links libpangenie_b200.so.
 *
 * It is compiled against the reference headers where they lie (-I<reference>/src); nothing is copied.
 * Backend selection at build time:
 *     -DPG_BINDING_BACKEND_GPU     pg_hmm_run   (libpangenie_b200.so, the product)
 *     -DPG_BINDING_BACKEND_ORACLE  pgo_hmm_run  (oracle/libpg_oracle.so — test infrastructure only, used to
 *                                  pin the CPU restatement against the reference's own test-suite)
 * Viterbi phasing (run_phasing, experimental in the reference and out of scope here) is not forwarded.
 */
#include <cmath>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "hmm.hpp"  // the reference's header, unmodified
#include "pangenie_b200.h"

#if defined(PG_BINDING_BACKEND_ORACLE)
#include "../oracle/pg_oracle.h"
#endif

using namespace std;

namespace {

struct FlatPanel {
  vector<uint64_t> positions;
  vector<uint16_t> path_to_allele, coverage, kmer_counts, allele_ids, allele_koff;
  vector<uint32_t> kmer_offsets, allele_offsets, allele_kmask;
  vector<uint8_t> allele_undefined;
  pg_panel view;
};

/* vector<shared_ptr<UniqueKmers>>  ->  pg_panel, through the public virtual interface only
 * (src/uniquekmers.hpp:22-69). */
void flatten(vector<shared_ptr<UniqueKmers>>* uk, FlatPanel& f) {
  const size_t V = uk->size();
  const unsigned short P = V ? uk->at(0)->get_nr_paths() : 0;
  f.kmer_offsets.push_back(0);
  f.allele_offsets.push_back(0);
  for (size_t v = 0; v < V; ++v) {
    UniqueKmers& u = *uk->at(v);
    if (u.get_nr_paths() != P) throw runtime_error("hmm_binding: variants covered by different numbers of paths");
    f.positions.push_back(u.get_variant_position());
    f.coverage.push_back(u.get_coverage());
    for (unsigned short p = 0; p < P; ++p) f.path_to_allele.push_back(u.get_allele(p));
    const size_t K = u.size();
    for (size_t k = 0; k < K; ++k) f.kmer_counts.push_back(u.get_readcount_of(k));
    f.kmer_offsets.push_back((uint32_t)f.kmer_counts.size());
    vector<unsigned short> ids;
    u.get_allele_ids(ids);  // keys of the alleles map, ascending
    for (unsigned short a : ids) {
      // KmerPath window: offset = first k-mer on the allele, 32-bit mask from there (src/kmerpath.cpp:13-48)
      uint32_t off = 0, mask = 0;
      bool first = true;
      for (size_t k = 0; k < K; ++k) {
        if (!u.kmer_on_allele(k, a)) continue;
        if (first) {
          off = (uint32_t)k;
          first = false;
        }
        mask |= 1u << (k - off);
      }
      f.allele_ids.push_back(a);
      f.allele_undefined.push_back(u.is_undefined_allele(a) ? 1 : 0);
      f.allele_koff.push_back((uint16_t)off);
      f.allele_kmask.push_back(mask);
    }
    f.allele_offsets.push_back((uint32_t)f.allele_ids.size());
  }
  pg_panel& p = f.view;
  p.n_variants = (uint32_t)V;
  p.n_paths = P;
  p.positions = f.positions.data();
  p.path_to_allele = f.path_to_allele.data();
  p.coverage = f.coverage.data();
  p.kmer_offsets = f.kmer_offsets.data();
  p.kmer_counts = f.kmer_counts.data();
  p.allele_offsets = f.allele_offsets.data();
  p.allele_ids = f.allele_ids.data();
  p.allele_undefined = f.allele_undefined.data();
  p.allele_kmer_offset = f.allele_koff.data();
  p.allele_kmer_mask = f.allele_kmask.data();
  p.kmer_codes = nullptr;
  p.flank_offsets = nullptr;
  p.flank_codes = nullptr;
}

/* ProbabilityTable keeps its range private, so the binding tabulates the (coverage, count) bounding box the
 * panel actually uses by calling get_probability — this also carries modify_probability() edits across.
 * (A production integration constructs pg_probtable directly from the histogram peak instead.) */
void tabulate(ProbabilityTable* probs, const FlatPanel& f, pg_probtable& t, vector<double>& storage) {
  unsigned cov_min = 65535, cov_max = 0, count_max = 0;
  for (auto c : f.coverage) {
    cov_min = min<unsigned>(cov_min, c);
    cov_max = max<unsigned>(cov_max, c);
  }
  for (auto c : f.kmer_counts) count_max = max<unsigned>(count_max, c);
  if (f.coverage.empty()) cov_min = cov_max = 0;
  if ((uint64_t)(cov_max - cov_min + 1) * (count_max + 1) > (1u << 24))
    throw runtime_error("hmm_binding: (coverage, count) range too large to tabulate");
  t.cov_min = (uint16_t)cov_min;
  t.cov_max = (uint16_t)(cov_max + 1);
  t.count_max = (uint16_t)min<unsigned>(count_max + 1, 65535);
  t.regularization = 0.0;  // every entry the panel can touch is inside the table
  const size_t ncov = t.cov_max - t.cov_min;
  storage.assign(ncov * t.count_max * 3, 0.0);
  for (unsigned count = 0; count < t.count_max; ++count)
    for (unsigned cov = t.cov_min; cov < t.cov_max; ++cov) {
      CopyNumber cn = probs->get_probability(cov, count);
      for (int i = 0; i < 3; ++i) {
        long double p = cn.get_probability_of(i);
        storage[(count * ncov + (cov - t.cov_min)) * 3 + i] = p > 0 ? (double)logl(p) : -INFINITY;
      }
    }
  t.log_p = storage.data();
}

#if defined(PG_BINDING_BACKEND_GPU)
pg_engine* engine() {
  static pg_engine* e = pg_engine_create(0);
  if (!e) throw runtime_error(string("pg_engine_create: ") + pg_last_error());
  return e;
}
#endif

}  // namespace

HMM::HMM(vector<shared_ptr<UniqueKmers>>* unique_kmers, ProbabilityTable* probabilities, bool run_genotyping, bool run_phasing,
         double recombrate, bool uniform, long double effective_N, vector<unsigned short>* only_paths, bool normalize)
    : unique_kmers(unique_kmers),
      probabilities(probabilities),
      genotyping_result(unique_kmers->size()),
      recombrate(recombrate),
      uniform(uniform),
      effective_N(effective_N) {
  this->column_indexer = nullptr;
  this->previous_backward_column = nullptr;
  (void)run_phasing;  // Viterbi is not on the accelerated path
  if (!run_genotyping) return;
  // the reference constructs a ColumnIndexer first, which throws for variants not covered by any path
  {
    ColumnIndexer check(unique_kmers, only_paths);
  }
  const size_t V = unique_kmers->size();
  if (V == 0) return;

  FlatPanel f;
  flatten(unique_kmers, f);
  pg_probtable table;
  vector<double> table_storage;
  tabulate(probabilities, f, table, table_storage);

  pg_hmm_params prm;
  prm.recombrate = recombrate;
  prm.effective_N = (double)effective_N;
  prm.uniform = uniform ? 1 : 0;
  prm.normalize = normalize ? 1 : 0;
  prm.only_paths = only_paths ? only_paths->data() : nullptr;
  prm.n_only_paths = only_paths ? (uint32_t)only_paths->size() : 0;

  vector<uint64_t> gl_off(V + 1);
  pg_result_layout(&f.view, gl_off.data());
  vector<double> lik(gl_off[V]);
  vector<uint8_t> is_col(V);
  vector<int16_t> gt(2 * V);
  vector<uint32_t> gq(V);
  vector<uint16_t> uks(V), cov(V);
  pg_hmm_result res;
  res.gl_offsets = gl_off.data();
  res.likelihoods = lik.data();
  res.is_column = is_col.data();
  res.genotype = gt.data();
  res.quality = gq.data();
  res.unique_kmers = uks.data();
  res.coverage = cov.data();

#if defined(PG_BINDING_BACKEND_GPU)
  if (pg_hmm_run(engine(), 1, &f.view, &table, &prm, &res) != PG_OK) throw runtime_error(string("pg_hmm_run: ") + pg_last_error());
#elif defined(PG_BINDING_BACKEND_ORACLE)
  if (pgo_hmm_run(1, &f.view, &table, &prm, &res) != PG_OK) throw runtime_error(string("pgo_hmm_run: ") + pgo_last_error());
#else
#error "define PG_BINDING_BACKEND_GPU or PG_BINDING_BACKEND_ORACLE"
#endif

  // back into the reference's result type.  The reference stores an entry for every allele pair realised by
  // a pair of selected paths (hmm.cpp:368), including zeros; skipped variants stay empty (columnindexer.cpp:24-31).
  vector<unsigned short> paths, alleles;
  for (size_t v = 0; v < V; ++v) {
    GenotypingResult& g = this->genotyping_result[v];
    g.set_unique_kmers(uks[v]);
    g.set_coverage(cov[v]);
    if (!is_col[v]) continue;
    paths.clear();
    alleles.clear();
    unique_kmers->at(v)->get_path_ids(paths, alleles, only_paths);
    set<unsigned short> present(alleles.begin(), alleles.end());
    for (unsigned short a1 : present)
      for (unsigned short a2 : present) {
        if (a1 > a2) continue;
        const uint64_t idx = gl_off[v] + ((uint64_t)a2 * (a2 + 1)) / 2 + a1;
        g.add_to_likelihood(a1, a2, (long double)lik[idx]);
      }
  }
}

HMM::~HMM() {}

vector<GenotypingResult> HMM::get_genotyping_result() const { return this->genotyping_result; }

vector<GenotypingResult> HMM::move_genotyping_result() { return move(this->genotyping_result); }

void HMM::combine_likelihoods(HMM& other) {  // src/hmm.cpp:513-519
  for (size_t i = 0; i < this->genotyping_result.size(); ++i) this->genotyping_result.at(i).combine(other.genotyping_result.at(i));
}

void HMM::normalize() {  // src/hmm.cpp:521-525
  for (size_t i = 0; i < this->genotyping_result.size(); ++i) this->genotyping_result[i].normalize();
}
