// C++ host for the `PanGenie -f <prefix> -i <reads>` stage over the C-ABI only (include/pangenie_b200.h): what
// run_genotype_command does between "UniqueKmersMap loaded" and "write VCF" (src/commands.cpp:730-1084), without
// jellyfish or cereal.  Prints one line per variant: chromosome, position, GT, GQ, UK, KC, genotype likelihoods in VCF
// order — the values Graph::write_genotypes puts into the GT:GQ:GL:KC fields (src/graph.cpp:206-260; the reference prints
// GL as log10, this tool prints the probabilities).
//
//   make tools        (g++ -std=c++17 -O2 -Iinclude ... -Lpangenie_b200 -lpangenie_b200, rpath to the library)
//   integration/genotype_from_index <prefix> <reads.fa|fq> [device]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "pangenie_b200.h"

static std::vector<char> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) {
    fprintf(stderr, "File %s cannot be opened.\n", path.c_str());
    exit(1);
  }
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <index prefix> <reads.fa|fq> [device]\n", argv[0]);
    return 2;
  }
  const int device = argc > 3 ? atoi(argv[3]) : 0;
  pg_index* ix = pg_index_open(argv[1], 1);
  if (!ix) {
    fprintf(stderr, "%s\n", pg_last_error());
    return 1;
  }
  const uint32_t n_chrom = pg_index_n_chromosomes(ix);
  std::vector<pg_panel> panels(n_chrom);
  std::vector<pg_hmm_result> results(n_chrom);
  std::vector<std::vector<uint64_t>> gl_off(n_chrom);
  std::vector<std::vector<double>> lik(n_chrom);
  std::vector<std::vector<uint8_t>> is_col(n_chrom);
  std::vector<std::vector<int16_t>> gt(n_chrom);
  std::vector<std::vector<uint32_t>> gq(n_chrom);
  std::vector<std::vector<uint16_t>> uk(n_chrom), kc(n_chrom);
  for (uint32_t c = 0; c < n_chrom; ++c) {
    if (pg_index_panel(ix, c, &panels[c])) {
      fprintf(stderr, "%s\n", pg_last_error());
      return 1;
    }
    const uint32_t V = panels[c].n_variants;
    gl_off[c].resize(V + 1);
    pg_result_layout(&panels[c], gl_off[c].data());
    lik[c].assign(gl_off[c][V], 0.0);
    is_col[c].assign(V, 0);
    gt[c].assign(2 * (size_t)V, 0);
    gq[c].assign(V, 0);
    uk[c].assign(V, 0);
    kc[c].assign(V, 0);
    results[c] = pg_hmm_result{gl_off[c].data(), lik[c].data(), is_col[c].data(), gt[c].data(), gq[c].data(), uk[c].data(), kc[c].data()};
  }
  const std::vector<char> reads = slurp(argv[2]);
  const std::vector<char> segments = slurp(pg_index_segments_path(ix));
  pg_genotype_input in{};
  in.reads = reads.data();
  in.reads_len = reads.size();
  in.segments = segments.data();
  in.segments_len = segments.size();
  in.k = pg_index_kmer_size(ix);
  in.hash_size = 3000000000ull;  // -e default (src/pangenie-genotype.cpp)
  in.regularization = 0.01;
  in.histogram_path = nullptr;
  pg_hmm_params prm{};
  prm.recombrate = 1.26;
  prm.effective_N = 0.00001;
  prm.uniform = 0;
  prm.normalize = 1;
  prm.only_paths = nullptr;
  prm.n_only_paths = 0;
  pg_engine* e = pg_engine_create(device);
  if (!e) {
    fprintf(stderr, "%s\n", pg_last_error());
    return 1;
  }
  uint64_t peak = 0;
  if (pg_genotype_run(e, &in, n_chrom, panels.data(), &prm, results.data(), &peak)) {
    fprintf(stderr, "%s\n", pg_last_error());
    return 1;
  }
  fprintf(stderr, "Computed kmer abundance peak: %llu\n", (unsigned long long)peak);
  printf("#chromosome\tposition\tGT\tGQ\tUK\tKC\tlikelihoods\n");
  for (uint32_t c = 0; c < n_chrom; ++c) {
    for (uint32_t v = 0; v < panels[c].n_variants; ++v) {
      const int a = gt[c][2 * v], b = gt[c][2 * v + 1];
      printf("%s\t%llu\t", pg_index_chromosome_name(ix, c), (unsigned long long)panels[c].positions[v]);
      if (a < 0) printf("./.");
      else printf("%d/%d", a, b);
      printf("\t%u\t%u\t%u\t", gq[c][v], uk[c][v], kc[c][v]);
      for (uint64_t g = gl_off[c][v]; g < gl_off[c][v + 1]; ++g) printf("%s%.6g", g == gl_off[c][v] ? "" : ",", lik[c][g]);
      printf("\n");
    }
  }
  pg_engine_destroy(e);
  pg_index_close(ix);
  return 0;
}
