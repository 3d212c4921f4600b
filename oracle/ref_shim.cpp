/*
 * ref_shim.cpp — C entry points around the reference's own UNMODIFIED hot-path sources.
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the sources where they lie
 * under /root/reference/src (nothing is copied) into oracle/_ref/libpg_ref.so, against the
 * header-only cereal stand-in in oracle/cereal_standin/.  Same signatures as include/pangenie_b200.h
 * with prefix pgr_.  Used to pin oracle/pg_oracle.cpp and as the `cpu_baseline.kind = "reference"` arm.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <unistd.h>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include <chrono>

#include "../include/pangenie_b200.h"
#include "biallelicuniquekmers.hpp"
#include "emissionprobabilitycomputer.hpp"
#include "histogram.hpp"
#include "haplotypesampler.hpp"
#include "hmm.hpp"
#include "multiallelicuniquekmers.hpp"
#include "probabilitytable.hpp"
#include "sequenceutils.hpp"
#include "transitionprobabilitycomputer.hpp"

namespace {
thread_local std::string g_err;

typedef std::vector<std::shared_ptr<UniqueKmers>> UKVec;

/* Rebuilds the reference's objects from the flat panel through their public API only. */
void build_unique_kmers(const pg_panel* p, UKVec& out) {
  out.clear();
  out.reserve(p->n_variants);
  for (uint32_t v = 0; v < p->n_variants; ++v) {
    std::vector<unsigned short> alleles(p->path_to_allele + (size_t)v * p->n_paths,
                                        p->path_to_allele + (size_t)(v + 1) * p->n_paths);
    bool biallelic = true;
    for (auto a : alleles)
      if (a > 1) biallelic = false;
    uint32_t ab = p->allele_offsets[v], ae = p->allele_offsets[v + 1];
    std::set<unsigned short> on_paths(alleles.begin(), alleles.end());
    if (on_paths.size() != ae - ab) throw std::runtime_error("panel allele list differs from alleles on paths");
    std::shared_ptr<UniqueKmers> u;
    if (biallelic) u.reset(new BiallelicUniqueKmers(p->positions[v], alleles));
    else u.reset(new MultiallelicUniqueKmers(p->positions[v], alleles));
    uint32_t kb = p->kmer_offsets[v], ke = p->kmer_offsets[v + 1];
    for (uint32_t k = kb; k < ke; ++k) {
      std::vector<unsigned short> on;
      for (uint32_t a = ab; a < ae; ++a) {
        uint32_t idx = k - kb, off = p->allele_kmer_offset[a];
        if (idx >= off && idx < off + 32 && ((p->allele_kmer_mask[a] >> (idx - off)) & 1u)) on.push_back(p->allele_ids[a]);
      }
      u->insert_kmer(p->kmer_counts[k], on);
    }
    for (uint32_t a = ab; a < ae; ++a)
      if (p->allele_undefined[a]) u->set_undefined_allele(p->allele_ids[a]);
    u->set_coverage(p->coverage[v]);
    out.push_back(u);
  }
}

/* A reference ProbabilityTable holding the caller's dense entries. Entries that already equal the
 * reference's own value to 1e-13 relative (in log) are left untouched, so a standard table stays
 * bit-identical to the reference's long double one. */
ProbabilityTable build_table(const pg_probtable* t) {
  ProbabilityTable probs(t->cov_min, t->cov_max, t->count_max, (long double)t->regularization);
  if (!t->log_p) return probs;
  size_t ncov = t->cov_max - t->cov_min;
  for (unsigned count = 0; count < t->count_max; ++count)
    for (unsigned cov = t->cov_min; cov < t->cov_max; ++cov) {
      const double* e = t->log_p + ((size_t)count * ncov + (cov - t->cov_min)) * 3;
      CopyNumber cur = probs.get_probability(cov, count);
      bool same = true;
      long double p[3];
      for (int cn = 0; cn < 3; ++cn) {
        p[cn] = (std::isinf(e[cn]) && e[cn] < 0) ? 0.0L : expl((long double)e[cn]);
        long double r = cur.get_probability_of(cn);
        if (fabsl(p[cn] - r) > 1e-13L * fabsl(r) + 1e-300L) same = false;
      }
      if (!same) probs.modify_probability(cov, count, CopyNumber(p[0], p[1], p[2]));
    }
  return probs;
}

int run_one(const pg_panel* panel, ProbabilityTable* probs, const pg_hmm_params* prm, pg_hmm_result* res, double* hmm_seconds = nullptr) {
  UKVec uk;
  build_unique_kmers(panel, uk);   // (the reference deserialises these objects; not part of run_genotyping)
  std::vector<unsigned short> only;
  if (prm->only_paths) only.assign(prm->only_paths, prm->only_paths + prm->n_only_paths);
  const auto t0 = std::chrono::steady_clock::now();
  HMM hmm(&uk, probs, true, false, prm->recombrate, prm->uniform != 0, (long double)prm->effective_N,
          prm->only_paths ? &only : nullptr, prm->normalize != 0);
  std::vector<GenotypingResult> gr = hmm.move_genotyping_result();
  if (hmm_seconds) *hmm_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  ColumnIndexer indexer(&uk, prm->only_paths ? &only : nullptr);
  std::memset(res->is_column, 0, panel->n_variants);
  for (size_t c = 0; c < indexer.size(); ++c) res->is_column[indexer.get_variant_id(c)] = 1;
  for (uint32_t v = 0; v < panel->n_variants; ++v) {
    uint64_t off = res->gl_offsets[v], n = res->gl_offsets[v + 1] - off;
    size_t nr_alleles = 0;
    while (nr_alleles * (nr_alleles + 1) / 2 < n) ++nr_alleles;
    std::vector<long double> l = gr[v].get_all_likelihoods(nr_alleles);
    for (size_t i = 0; i < n; ++i) res->likelihoods[off + i] = (double)l[i];
    res->unique_kmers[v] = gr[v].nr_unique_kmers();
    res->coverage[v] = gr[v].coverage();
    // post-processing of Graph::write_genotypes (graph.cpp:206-240) for a single-record bubble
    GenotypingResult tmp = gr[v];
    tmp.normalize();
    if (tmp.contains_no_likelihoods()) tmp.add_to_likelihood(0, 0, 1.0);
    std::vector<unsigned short> defined{0};
    for (unsigned short a = 1; a < nr_alleles; ++a)
      if (!uk[v]->is_undefined_allele(a)) defined.push_back(a);
    GenotypingResult fin = defined.size() < nr_alleles ? tmp.get_specific_likelihoods(defined) : tmp;
    std::pair<int, int> g = fin.get_likeliest_genotype();
    res->genotype[2 * v] = (int16_t)g.first;
    res->genotype[2 * v + 1] = (int16_t)g.second;
    res->quality[v] = g.first != -1 ? (uint32_t)fin.get_genotype_quality(g.first, g.second) : 0;
  }
  return PG_OK;
}
}  // namespace

extern "C" const char* pgr_last_error(void) { return g_err.c_str(); }

/** As pgr_hmm_run_mt; additionally chrom_seconds[c] (may be NULL) receives the wall seconds chromosome c spent inside the
 *  reference's `HMM` constructor + move_genotyping_result (what run_genotyping executes per job, commands.cpp:155-171),
 *  excluding the re-creation of the UniqueKmers objects from the flat panel. */
extern "C" int pgr_hmm_run_timed(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                                 const pg_hmm_params* params, pg_hmm_result* results, int threads, double* chrom_seconds) {
  try {
    ProbabilityTable probs = build_table(table);
    if (threads <= 1 || n_chrom == 1) {
      for (uint32_t c = 0; c < n_chrom; ++c) run_one(&panels[c], &probs, params, &results[c], chrom_seconds ? chrom_seconds + c : nullptr);
      return PG_OK;
    }
    // one job per chromosome on a fixed pool (commands.cpp:949-978)
    std::mutex mu;
    uint32_t next = 0;
    std::string err;
    std::vector<std::thread> pool;
    for (int t = 0; t < std::min<int>(threads, (int)n_chrom); ++t)
      pool.emplace_back([&]() {
        while (true) {
          uint32_t c;
          {
            std::lock_guard<std::mutex> lk(mu);
            if (next >= n_chrom) return;
            c = next++;
          }
          try {
            run_one(&panels[c], &probs, params, &results[c], chrom_seconds ? chrom_seconds + c : nullptr);
          } catch (std::exception& e) {
            std::lock_guard<std::mutex> lk(mu);
            err = e.what();
          }
        }
      });
    for (auto& th : pool) th.join();
    if (!err.empty()) {
      g_err = err;
      return PG_ERR_ARG;
    }
    return PG_OK;
  } catch (std::exception& e) {
    g_err = e.what();
    return PG_ERR_ARG;
  }
}

extern "C" int pgr_hmm_run_mt(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                              const pg_hmm_params* params, pg_hmm_result* results, int threads) {
  return pgr_hmm_run_timed(n_chrom, panels, table, params, results, threads, nullptr);
}

extern "C" int pgr_hmm_run(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                           const pg_hmm_params* params, pg_hmm_result* results) {
  return pgr_hmm_run_mt(n_chrom, panels, table, params, results, 1);
}

extern "C" int pgr_emission_run(const pg_panel* panel, const pg_probtable* table, const uint64_t* em_offsets,
                                double* emissions, double* log_scale) {
  try {
    ProbabilityTable probs = build_table(table);
    UKVec uk;
    build_unique_kmers(panel, uk);
    for (uint32_t v = 0; v < panel->n_variants; ++v) {
      EmissionProbabilityComputer em(uk[v], &probs);
      std::vector<unsigned short> ids;
      uk[v]->get_allele_ids(ids);
      unsigned short maxa = 0;
      for (auto a : ids) maxa = std::max(maxa, a);
      size_t dim = (size_t)maxa + 1;
      double* out = emissions + em_offsets[v];
      std::fill(out, out + dim * dim, 0.0);
      long double mx = 0.0L;
      for (auto a1 : ids)
        for (auto a2 : ids) mx = std::max(mx, em.get_emission_probability(a1, a2));
      log_scale[v] = mx > 0 ? (double)logl(mx) : 0.0;
      for (auto a1 : ids)
        for (auto a2 : ids) out[a1 * dim + a2] = (double)(em.get_emission_probability(a1, a2) / (mx > 0 ? mx : 1.0L));
    }
    return PG_OK;
  } catch (std::exception& e) {
    g_err = e.what();
    return PG_ERR_ARG;
  }
}

/** ln of ProbabilityTable(cov_min,cov_max,count_max,reg).get_probability(cov,count).get_probability_of(cn). */
extern "C" double pgr_log_probability(uint16_t cov_min, uint16_t cov_max, uint16_t count_max, double reg,
                                      uint16_t cov, uint16_t count, int cn) {
  ProbabilityTable probs(cov_min, cov_max, count_max, (long double)reg);
  return (double)logl(probs.get_probability(cov, count).get_probability_of(cn));
}

/** TransitionProbabilityComputer(from,to,recomb,nr_paths,uniform,N).compute_transition_prob(s), s=0..2 */
extern "C" void pgr_transitions(uint64_t from, uint64_t to, double recomb, uint16_t nr_paths, int uniform,
                                double effective_N, double out[3]) {
  TransitionProbabilityComputer t(from, to, recomb, nr_paths, uniform != 0, (long double)effective_N);
  for (unsigned short s = 0; s < 3; ++s) out[s] = (double)t.compute_transition_prob(s);
}

/** Histogram smoothing + peaks + choice (histogram.cpp:41-63, sequenceutils.cpp:42-84). */
extern "C" int pgr_histogram_peak(const uint64_t* bins, uint64_t n, int largest_peak, uint64_t* peak) {
  try {
    // Histogram(filename, max_value) is the reference's only bulk-load path (histogram.cpp:12-23)
    char tmpl[] = "/tmp/pgr_histoXXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) throw std::runtime_error("mkstemp failed");
    FILE* f = fdopen(fd, "w");
    for (uint64_t v = 0; v < n; ++v) fprintf(f, "%llu\t%llu\n", (unsigned long long)v, (unsigned long long)bins[v]);
    fclose(f);
    Histogram h(std::string(tmpl), n - 1);
    remove(tmpl);
    h.smooth_histogram();
    std::vector<size_t> ids, vals;
    h.find_peaks(ids, vals);
    *peak = compute_kmer_coverage(ids, vals, largest_peak != 0);
    return PG_OK;
  } catch (std::exception& e) {
    g_err = e.what();
    return PG_ERR_ARG;
  }
}

/** HaplotypeSampler(&unique_kmers, size, recombrate, effective_N, &best_scores, add_reference, "", "None", allele_penalty)
 *  (src/haplotypesampler.cpp:20-78), the reference's own class on objects rebuilt from the flat panel.
 *  n_out = size + (add_reference ? 1 : 0).  Outputs:
 *    sampled_paths       [n_out][V]  path id chosen by Viterbi pass i at every variant (get_sampled_paths())
 *    best_scores         [size]      DP score of every pass
 *    new_path_to_allele  [V][n_out]  get_allele(j) of the UPDATED UniqueKmers (update_unique_kmers(), :289-303)
 *    new_kmer_count      [V]         size() of the updated objects, new_counts: their get_readcount_of(i), concatenated */
extern "C" int pgr_haplotype_sample(const pg_panel* panel, uint32_t size, double recombrate, double effective_N,
                                    int add_reference, uint16_t allele_penalty, uint64_t* sampled_paths,
                                    uint32_t* best_scores, uint16_t* new_path_to_allele, uint32_t* new_kmer_count,
                                    uint16_t* new_counts) {
  try {
    UKVec uk;
    build_unique_kmers(panel, uk);
    std::vector<unsigned int> scores;
    HaplotypeSampler sampler(&uk, size, recombrate, (long double)effective_N, &scores, add_reference != 0, "", "None", allele_penalty);
    SampledPaths sp = sampler.get_sampled_paths();
    const size_t V = panel->n_variants, n_out = sp.sampled_paths.size();
    for (size_t i = 0; i < n_out; ++i)
      for (size_t v = 0; v < V; ++v) sampled_paths[i * V + v] = sp.sampled_paths[i][v];
    for (size_t i = 0; i < scores.size(); ++i) best_scores[i] = scores[i];
    size_t k = 0;
    for (size_t v = 0; v < V; ++v) {
      for (size_t j = 0; j < n_out; ++j) new_path_to_allele[v * n_out + j] = uk[v]->get_allele(j);
      new_kmer_count[v] = (uint32_t)uk[v]->size();
      for (size_t i = 0; i < uk[v]->size(); ++i) new_counts[k++] = uk[v]->get_readcount_of(i);
    }
    return PG_OK;
  } catch (std::exception& e) {
    g_err = e.what();
    return PG_ERR_ARG;
  }
}
