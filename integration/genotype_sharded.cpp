// C++ host for ONE sample sharded over the GPUs of a box, over the C-ABI (include/pangenie_b200.h) and NCCL
// only: what `PanGenie -f <prefix> -i <reads>` does between "UniqueKmersMap loaded" and "write VCF"
// (src/commands.cpp:730-1084), with the chromosomes LPT-assigned to the GPUs (one pool job per chromosome in the reference,
// :955-978) and the reads cut into record-aligned ranges.
//
//   every GPU:  PRIME its own k-mer table with the whole segment file, bring it into the canonical layout
//               (pg_count_canonicalize: identical array on every GPU, no broadcast), UPDATE with its read range
//   exchange:   ONE collective - ncclAllReduce(sum, uint32) of the count array, moved through the contiguous exchange
//               buffer in pieces (pg_count_export_range / pg_count_import_range)
//   every GPU:  histogram peak, ProbabilityTable, fill, emission + forward-backward of its chromosomes
//               (pg_engine_run_counted), no further exchange; results are printed in chromosome order
//
// One process, one host thread and one NCCL rank per GPU (ncclCommInitAll).  Output format = genotype_from_index.cpp.
//
//   make tools
//   integration/genotype_sharded <index prefix> <reads.fa|fq> <n_gpus>
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <numeric>
#include <string>
#include <thread>
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

// record-aligned cuts of a FASTA ('>') or 4-line FASTQ ('@') buffer (same rule as pangenie_b200/distributed.py)
static std::vector<size_t> record_cuts(const std::vector<char>& t, int n_shards) {
  const size_t n = t.size();
  std::vector<size_t> cuts{0};
  const bool fastq = n && t[0] == '@';
  for (int s = 1; s < n_shards; ++s) {
    size_t p = n * (size_t)s / (size_t)n_shards;
    while (p < n && p > 0 && t[p - 1] != '\n') ++p;
    while (p < n) {
      if (fastq) {
        size_t q = p;
        int nl = 0;
        while (q < n && nl < 2) nl += t[q++] == '\n';
        if (t[p] == '@' && q < n && t[q] == '+') break;
      } else if (t[p] == '>') {
        break;
      }
      while (p < n && t[p] != '\n') ++p;
      ++p;
    }
    cuts.push_back(std::max(std::min(p, n), cuts.back()));
  }
  cuts.push_back(n);
  return cuts;
}

struct Out {
  std::vector<uint64_t> gl_off;
  std::vector<double> lik;
  std::vector<uint8_t> is_col;
  std::vector<int16_t> gt;
  std::vector<uint32_t> gq;
  std::vector<uint16_t> uk, kc;
};

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <index prefix> <reads.fa|fq> <n_gpus>\n", argv[0]);
    return 2;
  }
  const int n_gpus = std::max(1, atoi(argv[3]));
  if (pg_device_count() < n_gpus) {
    fprintf(stderr, "%d GPUs requested, %d usable\n", n_gpus, pg_device_count());
    return 1;
  }
  pg_index* ix = pg_index_open(argv[1], 1);
  if (!ix) {
    fprintf(stderr, "%s\n", pg_last_error());
    return 1;
  }
  const uint32_t n_chrom = pg_index_n_chromosomes(ix);
  std::vector<pg_panel> panels(n_chrom);
  std::vector<Out> out(n_chrom);
  std::vector<pg_hmm_result> results(n_chrom);
  for (uint32_t c = 0; c < n_chrom; ++c) {
    if (pg_index_panel(ix, c, &panels[c])) {
      fprintf(stderr, "%s\n", pg_last_error());
      return 1;
    }
    const uint32_t V = panels[c].n_variants;
    Out& o = out[c];
    o.gl_off.resize(V + 1);
    pg_result_layout(&panels[c], o.gl_off.data());
    o.lik.assign(o.gl_off[V], 0.0);
    o.is_col.assign(V, 0);
    o.gt.assign(2 * (size_t)V, 0);
    o.gq.assign(V, 0);
    o.uk.assign(V, 0);
    o.kc.assign(V, 0);
    results[c] = pg_hmm_result{o.gl_off.data(), o.lik.data(), o.is_col.data(), o.gt.data(), o.gq.data(), o.uk.data(), o.kc.data()};
  }
  const std::vector<char> reads = slurp(argv[2]);
  const std::vector<char> segments = slurp(pg_index_segments_path(ix));
  const uint32_t k = pg_index_kmer_size(ix);
  // chromosomes: longest processing time first (descending variant count, to the least loaded GPU)
  std::vector<uint32_t> order(n_chrom);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return panels[a].n_variants > panels[b].n_variants; });
  std::vector<std::vector<uint32_t>> mine(n_gpus);
  std::vector<uint64_t> load(n_gpus, 0);
  for (uint32_t c : order) {
    const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    mine[g].push_back(c);
    load[g] += panels[c].n_variants;
  }
  for (auto& m : mine) std::sort(m.begin(), m.end());
  const std::vector<size_t> cuts = record_cuts(reads, n_gpus);

  std::vector<int> devs(n_gpus);
  std::iota(devs.begin(), devs.end(), 0);
  std::vector<ncclComm_t> comms(n_gpus);
  if (ncclCommInitAll(comms.data(), n_gpus, devs.data()) != ncclSuccess) {
    fprintf(stderr, "ncclCommInitAll failed\n");
    return 1;
  }
  std::atomic<int> failed{0};
  std::vector<uint64_t> peaks(n_gpus, 0);
  auto worker = [&](int g) {
    auto die = [&](const char* what) {
      fprintf(stderr, "GPU %d: %s: %s\n", g, what, pg_last_error());
      failed = 1;
    };
    cudaSetDevice(g);
    cudaStream_t s;
    cudaStreamCreate(&s);
    pg_engine* e = pg_engine_create(g);
    // every rank sizes its table from the same number, so the capacities (and the canonical layouts) are identical
    pg_counter* c = e ? pg_count_new(k, std::max<uint64_t>(segments.size(), 1024), g) : nullptr;
    if (!e || !c) return die("create");
    if (pg_count_feed(c, segments.data(), segments.size(), PG_OP_PRIME)) return die("PRIME");
    if (pg_count_canonicalize(c)) return die("canonicalize");
    if (cuts[g + 1] > cuts[g] && pg_count_feed(c, reads.data() + cuts[g], cuts[g + 1] - cuts[g], PG_OP_UPDATE)) return die("UPDATE");
    // THE collective: the count arrays of all GPUs are added position by position
    const uint64_t cap = pg_count_capacity(c), chunk = std::min<uint64_t>(cap, 1ull << 28);
    uint64_t addr = 0;
    if (pg_count_exchange_buffer(c, chunk, &addr)) return die("exchange buffer");
    for (uint64_t first = 0; first < cap && !failed; first += chunk) {
      const uint64_t n = std::min<uint64_t>(chunk, cap - first);
      if (pg_count_export_range(c, first, n)) return die("export");
      if (ncclAllReduce((const void*)addr, (void*)addr, n, ncclUint32, ncclSum, comms[g], s) != ncclSuccess) return die("ncclAllReduce");
      cudaStreamSynchronize(s);
      if (pg_count_import_range(c, first, n)) return die("import");
    }
    if (!mine[g].empty()) {
      std::vector<pg_panel> ps;
      std::vector<pg_hmm_result> rs;
      for (uint32_t ch : mine[g]) {
        ps.push_back(panels[ch]);
        rs.push_back(results[ch]);
      }
      pg_hmm_params prm{};
      prm.recombrate = 1.26;
      prm.effective_N = 0.00001;
      prm.normalize = 1;
      if (pg_engine_load(e, (uint32_t)ps.size(), ps.data(), rs.data())) return die("load");
      if (pg_engine_run_counted(e, c, 1, 0.01, &prm, &peaks[g])) return die("run");
      if (pg_engine_fetch(e, (uint32_t)ps.size(), ps.data(), rs.data())) return die("fetch");
    } else {
      if (pg_count_compute_histogram(c, 10000, 1, nullptr, &peaks[g])) return die("histogram");
    }
    pg_count_destroy(c);
    pg_engine_destroy(e);
    cudaStreamDestroy(s);
  };
  std::vector<std::thread> pool;
  for (int g = 0; g < n_gpus; ++g) pool.emplace_back(worker, g);
  for (auto& t : pool) t.join();
  for (auto& cm : comms) ncclCommDestroy(cm);
  if (failed) return 1;
  for (int g = 1; g < n_gpus; ++g)
    if (peaks[g] != peaks[0]) {
      fprintf(stderr, "k-mer abundance peaks differ between GPUs (%llu vs %llu)\n", (unsigned long long)peaks[g], (unsigned long long)peaks[0]);
      return 1;
    }
  fprintf(stderr, "Computed kmer abundance peak: %llu\n", (unsigned long long)peaks[0]);
  printf("#chromosome\tposition\tGT\tGQ\tUK\tKC\tlikelihoods\n");
  for (uint32_t c = 0; c < n_chrom; ++c) {
    const Out& o = out[c];
    for (uint32_t v = 0; v < panels[c].n_variants; ++v) {
      const int a = o.gt[2 * v], b = o.gt[2 * v + 1];
      printf("%s\t%llu\t", pg_index_chromosome_name(ix, c), (unsigned long long)panels[c].positions[v]);
      if (a < 0) printf("./.");
      else printf("%d/%d", a, b);
      printf("\t%u\t%u\t%u\t", o.gq[v], o.uk[v], o.kc[v]);
      for (uint64_t gi = o.gl_off[v]; gi < o.gl_off[v + 1]; ++gi) printf("%s%.6g", gi == o.gl_off[v] ? "" : ",", o.lik[gi]);
      printf("\n");
    }
  }
  pg_index_close(ix);
  return 0;
}
