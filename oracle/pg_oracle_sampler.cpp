/*
 * pg_oracle_sampler.cpp — CPU restatement of the reference's HaplotypeSampler (the integer Viterbi
 * that real >100-haplotype runs execute between the count fill and the HMM).  TEST INFRASTRUCTURE ONLY (see pg_oracle.h):
 * groundwork for a device implementation, pinned against the reference's own class compiled unmodified
 * (oracle/_ref/libpg_ref.so, pgr_haplotype_sample) by tests/test_oracle_vs_reference.py.
 *
 * What is restated (all file:line relative to /root/reference/src):
 *   - allele penalties           samplingemissions.cpp:9-44   (-10 log10 of the fraction of an allele's k-mers seen >= 3 times,
 *                                                              float log, truncated to u16; undefined 50, unseen 25; penalize())
 *   - recombination cost         samplingtransitions.cpp:5-22 (phred-scaled Li-Stephens switch probability, double exp/log10)
 *   - one Viterbi pass           haplotypesampler.cpp:111-176, 178-286 (min / second-min of the previous column with the
 *                                reference's tie rules, overflow saturation, paths of earlier passes masked per column)
 *   - the sampled panel          haplotypesampler.cpp:289-303, {bi,multi}allelicuniquekmers.cpp update_paths (k-mers that lie
 *                                on no remaining allele are dropped)
 * The reference keeps ~sqrt(V) columns and recomputes during backtracking (:116-126, 151-158); that changes memory, not
 * results, so this restatement simply keeps every backtrace column.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <set>
#include <vector>

#include "pg_oracle.h"

namespace {

constexpr unsigned int UMAX = std::numeric_limits<unsigned int>::max();

inline bool kmer_on(const pg_panel* p, uint32_t a, uint32_t k) {  // kmerpath.cpp:33-48
  const uint32_t off = p->allele_kmer_offset[a];
  return k >= off && k < off + 32u && ((p->allele_kmer_mask[a] >> (k - off)) & 1u);
}

// samplingtransitions.cpp:5-14.  That file has no `using namespace std`, so its unqualified exp / log10 are the C
// double-precision functions; the arithmetic around them is long double.
unsigned int recombination_cost(uint64_t from, uint64_t to, double recomb_rate, unsigned short nr_paths, long double effective_N) {
  const long double distance = (to - from) * 0.000004L * ((long double)recomb_rate) * effective_N;
  const long double recomb_prob = (1.0L - ::exp((double)(-distance / (long double)nr_paths))) * (1.0L / (long double)nr_paths);
  return (unsigned int)(-10.0 * ::log10((double)recomb_prob));
}

struct Penalties {  // SamplingEmissions
  std::vector<unsigned short> by_allele;
  static constexpr unsigned short DEFAULT = 25;
  void penalize(unsigned short allele, unsigned short penalty) {  // samplingemissions.cpp:38-44
    by_allele[allele] += penalty;
    if (by_allele[allele] > DEFAULT) by_allele[allele] = DEFAULT;
  }
};

Penalties initial_penalties(const pg_panel* p, uint32_t v) {  // samplingemissions.cpp:9-32
  Penalties pen;
  const uint32_t ab = p->allele_offsets[v], ae = p->allele_offsets[v + 1];
  const uint32_t kb = p->kmer_offsets[v], K = p->kmer_offsets[v + 1] - kb;
  unsigned short max_allele = 0;
  for (uint32_t a = ab; a < ae; ++a) max_allele = std::max(max_allele, p->allele_ids[a]);
  pen.by_allele.assign((size_t)max_allele + 1, 0);
  for (uint32_t a = ab; a < ae; ++a) {
    const unsigned short id = p->allele_ids[a];
    if (p->allele_undefined[a]) {
      pen.by_allele[id] = 50;
      continue;
    }
    unsigned short total = 0, present = 0;  // kmers_on_allele / present_kmers_on_allele (multiallelicuniquekmers.cpp:150-162)
    for (uint32_t k = 0; k < K; ++k)
      if (kmer_on(p, a, k)) {
        ++total;
        if (p->kmer_counts[kb + k] >= 3) ++present;
      }
    const float fraction = total > 0 ? present / (float)total : 1.0f;
    if (fraction > 0.0) pen.by_allele[id] = (unsigned short)(-10.0 * std::log10(fraction));  // float log10, as in the reference
    else pen.by_allele[id] = Penalties::DEFAULT;
  }
  return pen;
}

}  // namespace

extern "C" int pgo_haplotype_sample(const pg_panel* panel, uint32_t size, double recombrate, double effective_N,
                                    int add_reference, uint16_t allele_penalty, uint64_t* sampled_paths,
                                    uint32_t* best_scores, uint16_t* new_path_to_allele, uint32_t* new_kmer_count,
                                    uint16_t* new_counts) {
  const size_t V = panel->n_variants, P = panel->n_paths;
  if (size < 1 || V == 0) return PG_OK;
  std::vector<Penalties> pen(V);
  for (size_t v = 0; v < V; ++v) pen[v] = initial_penalties(panel, (uint32_t)v);
  std::vector<unsigned int> switch_cost(V, 0);
  for (size_t v = 1; v < V; ++v)
    switch_cost[v] = recombination_cost(panel->positions[v - 1], panel->positions[v], recombrate, (unsigned short)P, (long double)effective_N);
  std::vector<uint8_t> used(V * P, 0);  // (column, path) taken by an earlier pass
  std::vector<unsigned int> prev(P), cur(P);
  std::vector<uint32_t> back(V * P);
  auto sat_add = [](unsigned int a, unsigned int b) {  // haplotypesampler.cpp:253-254, 262, 273-274
    const unsigned int s = a + b;
    return s < a ? UMAX : s;
  };
  for (uint32_t pass = 0; pass < size; ++pass) {
    for (size_t v = 0; v < V; ++v) {
      const uint16_t* alleles = panel->path_to_allele + v * P;
      size_t first_id = UMAX, second_id = UMAX;
      unsigned int first_val = UMAX, second_val = UMAX;
      if (v > 0) {  // get_column_minima over the paths still free in the previous column (:82-108)
        for (size_t i = 0; i < P; ++i) {
          if (used[(v - 1) * P + i]) continue;
          if (prev[i] < first_val) {
            second_val = first_val;
            second_id = first_id;
            first_val = prev[i];
            first_id = i;
          } else if (prev[i] < second_val && i != first_id) {
            second_val = prev[i];
            second_id = i;
          }
        }
      }
      for (size_t i = 0; i < P; ++i) {
        if (used[v * P + i]) {
          cur[i] = UMAX;
          back[v * P + i] = UMAX;
          continue;
        }
        unsigned int cell = 0;
        uint32_t from = UMAX;
        if (v > 0) {
          const bool is_first = i == first_id;
          cell = sat_add(is_first ? second_val : first_val, switch_cost[v]);
          from = (uint32_t)(is_first ? second_id : first_id);
          if (!used[(v - 1) * P + i]) {
            const unsigned int same = prev[i];  // no recombination: cost 0
            if (same < cell) {
              cell = same;
              from = (uint32_t)i;
            }
          }
        }
        cur[i] = sat_add(cell, pen[v].by_allele[alleles[i]]);
        back[v * P + i] = from;
      }
      prev.swap(cur);
    }
    // best end state: first minimal entry of the last column (:131-141)
    size_t best = 0;
    for (size_t i = 1; i < P; ++i)
      if (prev[i] < prev[best]) best = i;
    best_scores[pass] = prev[best];
    // backtrace; every visited allele is penalised for the following passes (:147-166)
    for (size_t v = V; v-- > 0;) {
      sampled_paths[(size_t)pass * V + v] = best;
      pen[v].penalize(panel->path_to_allele[v * P + best], allele_penalty);
      used[v * P + best] = 1;
      if (v > 0) best = back[v * P + best];
    }
  }
  const size_t n_out = size + (add_reference ? 1 : 0);
  if (add_reference)
    for (size_t v = 0; v < V; ++v) sampled_paths[(size_t)size * V + v] = 0;  // :48
  // update_unique_kmers (:289-303): the panel restricted to the sampled paths
  size_t k_out = 0;
  for (size_t v = 0; v < V; ++v) {
    std::set<unsigned short> kept;
    for (size_t j = 0; j < n_out; ++j) {
      const unsigned short a = panel->path_to_allele[v * P + sampled_paths[j * V + v]];
      new_path_to_allele[v * n_out + j] = a;
      kept.insert(a);
    }
    const uint32_t ab = panel->allele_offsets[v], ae = panel->allele_offsets[v + 1];
    const uint32_t kb = panel->kmer_offsets[v], K = panel->kmer_offsets[v + 1] - kb;
    uint32_t n = 0;
    for (uint32_t k = 0; k < K; ++k) {  // a k-mer survives iff it lies on a remaining allele (update_paths)
      bool on = false;
      for (uint32_t a = ab; a < ae && !on; ++a) on = kept.count(panel->allele_ids[a]) && kmer_on(panel, a, k);
      if (on) {
        new_counts[k_out++] = panel->kmer_counts[kb + k];
        ++n;
      }
    }
    new_kmer_count[v] = n;
  }
  return PG_OK;
}
