/*
 * pg_oracle_index.cpp — CPU restatement of the reference's unique-k-mer selection (index stage).
 * TEST INFRASTRUCTURE ONLY (see pg_oracle.h).
 *
 * Follows src/stepwiseuniquekmercomputer.cpp statement by statement with std::map in place of
 * std::map<jellyfish::mer_dna, ...> (mer_dna orders by the 2-bit code with the first base most significant, the order
 * of the u64 codes used here).  jellyfish is absent from /root/reference; the three behaviours of `mer_dna` this file
 * depends on are restated from jellyfish 2.x: (1) `mer_dna("")` is the all-A k-mer, (2) `shift_left(char)` with a
 * character outside ACGTacgt leaves the k-mer unchanged, (3) operator< is numeric on the packed code.
 * PINNED by the reference's own index fixture: tests/data/index_chr1_Graph.cereal (input) ->
 * tests/data/index_chr1_kmers.tsv.gz + tests/data/index_UniqueKmersMap.cereal (output of the real PanGenie-index);
 * see tests/test_index_build.py.  (1) and (2) are only reachable with undefined bases inside flanks or overhangs shorter
 * than k; for those inputs parity is unpinned.
 */
#include <cstring>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "pg_oracle.h"

namespace {

struct Window {  // jellyfish::mer_dna of size k: shift_left(char)
  uint64_t code = 0, mask;
  explicit Window(uint32_t k) : mask(k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1)) {}
  void shift_left(char c) {
    int x;
    switch (c) {
      case 'A': case 'a': x = 0; break;
      case 'C': case 'c': x = 1; break;
      case 'G': case 'g': x = 2; break;
      case 'T': case 't': x = 3; break;
      default: return;
    }
    code = ((code << 2) | (uint64_t)x) & mask;
  }
};

/* stepwise_unique_kmers (src/stepwiseuniquekmercomputer.cpp:11-34) */
void stepwise_unique_kmers(const char* allele, uint64_t len, uint16_t index, uint32_t k,
                           std::map<uint64_t, std::vector<uint16_t>>& occurences) {
  std::map<uint64_t, size_t> counts;
  size_t extra_shifts = k;
  Window current(k);
  for (uint64_t i = 0; i < len; ++i) {
    const char base = allele[i];
    if (extra_shifts == 0) counts[current.code] += 1;
    if (base != 'A' && base != 'C' && base != 'G' && base != 'T') extra_shifts = k + 1;
    current.shift_left(base);
    if (extra_shifts > 0) extra_shifts -= 1;
  }
  counts[current.code] += 1;
  for (auto const& e : counts)
    if (e.second == 1) occurences[e.first].push_back(index);
}

}  // namespace

struct pgo_unique_kmers {
  uint32_t V = 0, P = 0;
  std::vector<uint64_t> positions;
  std::vector<uint16_t> p2a, coverage, kmer_counts, allele_ids, allele_koff;
  std::vector<uint32_t> kmer_offsets, allele_offsets, allele_kmask, flank_offsets;
  std::vector<uint8_t> allele_undef;
  std::vector<uint64_t> kmer_codes, flank_codes;
};

extern "C" pgo_unique_kmers* pgo_unique_kmers_compute(const pgo_counter* graph, const pg_variants* in) {
  if (!graph || !in) return nullptr;
  auto* u = new pgo_unique_kmers;
  const uint32_t V = in->n_variants, P = in->n_paths, k = in->k;
  u->V = V; u->P = P;
  u->positions.assign(in->positions, in->positions + V);
  u->p2a.assign(in->path_to_allele, in->path_to_allele + (size_t)V * P);
  u->coverage.assign(V, 0);
  u->kmer_offsets.push_back(0); u->allele_offsets.push_back(0); u->flank_offsets.push_back(0);
  auto abundance = [&](uint64_t code) { uint64_t c; pgo_count_lookup(graph, &code, 1, &c); return c; };
  for (uint32_t v = 0; v < V; ++v) {
    const uint16_t* paths = in->path_to_allele + (size_t)v * P;
    const uint32_t a0 = in->allele_offsets[v], n_alleles = in->allele_offsets[v + 1] - a0;
    /* compute_unique_kmers :118-152 */
    bool is_biallelic = true;
    for (uint32_t p = 0; p < P; ++p)
      if (paths[p] != 0 && paths[p] != 1) is_biallelic = false;
    struct Info { bool undefined = false; uint16_t offset = 0; uint32_t mask = 0; };
    std::map<uint16_t, Info> alleles;  // UniqueKmers::alleles: the alleles carried by a path (constructors of both classes)
    for (uint32_t p = 0; p < P; ++p) alleles[paths[p]];
    std::map<uint64_t, std::vector<uint16_t>> occurences;
    for (uint32_t a = 0; a < n_alleles; ++a) {
      if (in->allele_undefined[a0 + a]) {
        auto it = alleles.find((uint16_t)a);
        if (it == alleles.end()) { delete u; return nullptr; }  // set_undefined_allele throws (multiallelicuniquekmers.cpp:180-186)
        it->second.undefined = true;
        continue;
      }
      const uint64_t s0 = in->seq_offsets[a0 + a], s1 = in->seq_offsets[a0 + a + 1];
      stepwise_unique_kmers(in->seq + s0, s1 - s0, (uint16_t)a, k, occurences);
    }
    /* select_kmers :46-93 */
    std::map<uint16_t, std::queue<uint64_t>> allele_to_kmers;
    for (auto const& kmer : occurences) {
      const size_t genomic_count = abundance(kmer.first), local_count = kmer.second.size();
      if ((genomic_count - local_count) != 0) continue;
      if (local_count > 1) continue;
      bool covered = false;
      for (uint32_t p = 0; p < P; ++p) covered |= paths[p] == kmer.second[0];
      if (!covered) continue;
      allele_to_kmers[kmer.second[0]].push(kmer.first);
    }
    size_t nr_selected = 0;
    std::map<uint16_t, std::vector<uint64_t>> result;
    bool keep_adding = true;
    uint16_t max_alleles = (uint16_t)P;
    if (max_alleles < 301) max_alleles = 301;
    const size_t max_kmers = is_biallelic ? 16 : 32;
    while (nr_selected < max_alleles && keep_adding) {
      bool kmer_added = false;
      for (auto& a : allele_to_kmers) {
        if (a.second.size() > 0 && result[a.first].size() < max_kmers) {
          result[a.first].push_back(a.second.front());
          a.second.pop();
          kmer_added = true;
          nr_selected += 1;
        }
        if (nr_selected >= max_alleles) break;
      }
      keep_adding = kmer_added;
    }
    /* insert_kmer in (allele, selection) order :154-165; KmerPath::set_position (src/kmerpath.cpp:13-31) */
    uint32_t index = 0;
    for (auto& a : result)
      for (uint64_t code : a.second) {
        Info& info = alleles[a.first];
        if (info.mask == 0) info.offset = (uint16_t)index;
        info.mask |= 1u << (index - info.offset);
        u->kmer_codes.push_back(code);
        u->kmer_counts.push_back(0);
        ++index;
      }
    u->kmer_offsets.push_back((uint32_t)u->kmer_codes.size());
    for (auto const& a : alleles) {
      u->allele_ids.push_back(a.first);
      u->allele_undef.push_back(a.second.undefined ? 1 : 0);
      u->allele_koff.push_back(a.second.offset);
      u->allele_kmask.push_back(a.second.mask);
    }
    u->allele_offsets.push_back((uint32_t)u->allele_ids.size());
    /* determine_unique_flanking_kmers :227-264 */
    for (int side = 0; side < 2; ++side) {
      const uint64_t* off = side ? in->right_offsets : in->left_offsets;
      const char* seq = side ? in->right_seq : in->left_seq;
      std::map<uint64_t, std::vector<uint16_t>> occ;
      stepwise_unique_kmers(seq + off[v], off[v + 1] - off[v], (uint16_t)side, k, occ);
      size_t selected = 0;
      for (auto& kmer : occ) {
        if (selected >= 12) break;
        if (abundance(kmer.first) == 1) {
          u->flank_codes.push_back(kmer.first);
          selected += 1;
        }
      }
    }
    u->flank_offsets.push_back((uint32_t)u->flank_codes.size());
  }
  return u;
}

extern "C" int pgo_unique_kmers_panel(pgo_unique_kmers* u, pg_panel* out) {
  if (!u || !out) return PG_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  out->n_variants = u->V; out->n_paths = u->P;
  out->positions = u->positions.data(); out->path_to_allele = u->p2a.data(); out->coverage = u->coverage.data();
  out->kmer_offsets = u->kmer_offsets.data(); out->kmer_counts = u->kmer_counts.data();
  out->allele_offsets = u->allele_offsets.data(); out->allele_ids = u->allele_ids.data();
  out->allele_undefined = u->allele_undef.data(); out->allele_kmer_offset = u->allele_koff.data();
  out->allele_kmer_mask = u->allele_kmask.data(); out->kmer_codes = u->kmer_codes.data();
  out->flank_offsets = u->flank_offsets.data(); out->flank_codes = u->flank_codes.data();
  return PG_OK;
}

extern "C" void pgo_unique_kmers_free(pgo_unique_kmers* u) { delete u; }
