/*
 * pg_oracle.cpp — CPU restatement of PanGenie's genotyping hot path.  TEST INFRASTRUCTURE ONLY
 * (see pg_oracle.h).  x86-64 `long double` (x87 80-bit) everywhere the reference uses it, no
 * -ffast-math.  Each function cites the reference file:line (relative to the reference tree) it restates.
 */
#include "pg_oracle.h"

#include <algorithm>
#include <atomic>
#include <memory>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
typedef long double ld;
}  // namespace

extern "C" const char* pgo_last_error(void) { return g_err.c_str(); }

/* =================================================================================================
 * K-mer counting.  Reference call sites: src/jellyfishcounter.hpp:46-68 (COUNT/PRIME/UPDATE),
 * src/jellyfishcounter.cpp:26-104.  The arithmetic lives in jellyfish 2.x (not in /root/reference):
 *   - mer_dna code A/a=0 C/c=1 G/g=2 T/t=3, first base most significant; any other character resets
 *     the rolling window (mer_iterator);
 *   - canonical = numerically smaller of the k-mer and its reverse complement;
 *   - mer_overlap_sequence_parser: file type from the first character ('>' FASTA, '@' FASTQ);
 *     sequence lines of a record are concatenated (newlines dropped); records are separated by an
 *     'N' so k-mers never span records; FASTQ quality is skipped by LENGTH (so '@'/'+' inside the
 *     quality string are harmless) and multi-line FASTQ is accepted.
 * ================================================================================================= */
/* Lock-free open-addressing table (linear probing, power-of-two capacity) so the CPU baseline counts at a
 * rate comparable to jellyfish's own lock-free hash; resized (single-threaded) before a feed when the text
 * could overflow it. */
struct pgo_counter {
  uint32_t k = 0;
  uint64_t cap = 0, mask = 0;
  std::unique_ptr<std::atomic<uint64_t>[]> keys;
  std::unique_ptr<std::atomic<uint64_t>[]> vals;
  std::atomic<uint64_t> distinct{0};
  static constexpr uint64_t EMPTY = ~0ULL;

  static uint64_t mix(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
  }
  void alloc(uint64_t c) {
    cap = 1024;
    while (cap < c) cap <<= 1;
    mask = cap - 1;
    keys.reset(new std::atomic<uint64_t>[cap]);
    vals.reset(new std::atomic<uint64_t>[cap]);
    // first touch in parallel: the initialisation of a table for a whole genome is otherwise the slowest part of PRIME
    const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), (unsigned)(cap >> 20) + 1u));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
      pool.emplace_back([this, t, nt]() {
        for (uint64_t i = cap * t / nt; i < cap * (t + 1) / nt; ++i) {
          keys[i].store(EMPTY, std::memory_order_relaxed);
          vals[i].store(0, std::memory_order_relaxed);
        }
      });
    for (auto& th : pool) th.join();
  }
  void reserve(uint64_t want_distinct) {
    const uint64_t need = (uint64_t)(want_distinct / 0.5) + 1024;
    if (need <= cap) return;
    pgo_counter old;
    old.cap = cap; old.keys.swap(keys); old.vals.swap(vals);
    alloc(need);
    for (uint64_t i = 0; i < old.cap; ++i) {
      const uint64_t key = old.keys[i].load(std::memory_order_relaxed);
      if (key == EMPTY) continue;
      uint64_t s = mix(key) & mask;
      while (keys[s].load(std::memory_order_relaxed) != EMPTY) s = (s + 1) & mask;
      keys[s].store(key, std::memory_order_relaxed);
      vals[s].store(old.vals[i].load(std::memory_order_relaxed), std::memory_order_relaxed);
    }
  }
  // returns the slot of `key`, inserting it if `insert`; -1 if absent
  int64_t find(uint64_t key, bool insert) {
    uint64_t s = mix(key) & mask;
    while (true) {
      uint64_t cur = keys[s].load(std::memory_order_acquire);
      if (cur == key) return (int64_t)s;
      if (cur == EMPTY) {
        if (!insert) return -1;
        if (keys[s].compare_exchange_strong(cur, key, std::memory_order_acq_rel)) {
          distinct.fetch_add(1, std::memory_order_relaxed);
          return (int64_t)s;
        }
        if (cur == key) return (int64_t)s;
      }
      s = (s + 1) & mask;
    }
  }
  int64_t find_const(uint64_t key) const {
    if (!cap) return -1;
    uint64_t s = mix(key) & mask;
    while (true) {
      const uint64_t cur = keys[s].load(std::memory_order_relaxed);
      if (cur == key) return (int64_t)s;
      if (cur == EMPTY) return -1;
      s = (s + 1) & mask;
    }
  }
  void apply(uint64_t key, int op) {
    if (op == PG_OP_UPDATE) {
      const int64_t s = find(key, false);
      if (s >= 0) vals[s].fetch_add(1, std::memory_order_relaxed);
    } else {
      const int64_t s = find(key, true);
      if (op == PG_OP_COUNT) vals[s].fetch_add(1, std::memory_order_relaxed);
    }
  }
  uint64_t get(uint64_t key) const {
    const int64_t s = find_const(key);
    return s < 0 ? 0 : vals[s].load(std::memory_order_relaxed);
  }
};

namespace {

inline int base_code(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

/* Streams the concatenated sequence characters of a FASTA/FASTQ buffer to `emit(char)`, emitting 'N'
 * between records, exactly as jellyfish's parser lays them out for the mer iterator. */
template <class Emit>
int parse_records(const char* text, uint64_t len, Emit emit) {
  if (len == 0) return PG_OK;
  uint64_t i = 0;
  auto skip_line = [&]() {
    while (i < len && text[i] != '\n') ++i;
    if (i < len) ++i;
  };
  auto skip_newlines = [&]() {
    while (i < len && text[i] == '\n') ++i;
  };
  if (text[0] == '>') {
    skip_line();  // first header
    while (i < len) {
      skip_newlines();
      if (i >= len) break;
      if (text[i] == '>') {
        emit('N');
        skip_line();
        continue;
      }
      while (i < len && text[i] != '\n') emit(text[i++]);
    }
    return PG_OK;
  }
  if (text[0] == '@') {
    skip_line();  // first header
    while (i < len) {
      uint64_t seq_len = 0;
      // sequence lines until a line starting with '+'
      while (true) {
        skip_newlines();
        if (i >= len || text[i] == '+') break;
        while (i < len && text[i] != '\n') {
          emit(text[i++]);
          ++seq_len;
        }
      }
      if (i >= len) break;
      skip_line();  // '+' line
      uint64_t quals = 0;
      skip_newlines();
      while (i < len && quals < seq_len) {
        while (i < len && text[i] != '\n' && quals < seq_len) {
          ++i;
          ++quals;
        }
        skip_newlines();
      }
      if (quals != seq_len) return fail(PG_ERR_FORMAT, "invalid fastq: quality/sequence length mismatch");
      if (i < len) {
        emit('N');
        skip_line();  // next header
      }
    }
    return PG_OK;
  }
  return fail(PG_ERR_FORMAT, "unsupported sequence format (expected '>' or '@')");
}

/* Rolling canonical k-mer extraction (jellyfish mer_iterator). */
template <class Sink>
struct KmerRoller {
  uint32_t k;
  uint64_t mask, fwd = 0, rev = 0;
  uint32_t filled = 0;
  Sink sink;
  KmerRoller(uint32_t k_, Sink s) : k(k_), mask(k_ == 32 ? ~0ULL : ((1ULL << (2 * k_)) - 1)), sink(s) {}
  void operator()(char ch) {
    int c = base_code((unsigned char)ch);
    if (c < 0) {
      filled = 0;
      return;
    }
    fwd = ((fwd << 2) | (uint64_t)c) & mask;
    rev = (rev >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
    if (++filled >= k) sink(fwd < rev ? fwd : rev);
  }
};

uint64_t canonical_of(uint64_t code, uint32_t k) {
  uint64_t rev = 0, f = code;
  for (uint32_t i = 0; i < k; ++i) {
    rev = (rev << 2) | (3 - (f & 3));
    f >>= 2;
  }
  return code < rev ? code : rev;
}

bool encode_ascii(const char* s, uint32_t k, uint64_t* out) {
  uint64_t v = 0;
  for (uint32_t i = 0; i < k; ++i) {
    int c = base_code((unsigned char)s[i]);
    if (c < 0) return false;
    v = (v << 2) | (uint64_t)c;
  }
  *out = v;
  return true;
}

}  // namespace

extern "C" pgo_counter* pgo_count_new(uint32_t k) {
  if (k < 1 || k > 32) {
    fail(PG_ERR_ARG, "k must be in [1,32]");
    return nullptr;
  }
  pgo_counter* c = new pgo_counter();
  c->k = k;
  c->alloc(1024);
  return c;
}

namespace {
/* record-aligned cut points for multi-threaded parsing */
std::vector<uint64_t> record_cuts(const char* text, uint64_t len, int threads) {
  std::vector<uint64_t> cuts{0};
  if (len == 0) return {0, 0};
  const bool fastq = text[0] == '@';
  for (int t = 1; t < threads; ++t) {
    uint64_t p = len / threads * t;
    while (p < len && p > 0 && text[p - 1] != '\n') ++p;
    if (fastq) {
      // a record starts at a line beginning with '@' whose line+2 begins with '+' (4-line records)
      while (p < len) {
        uint64_t q = p;
        int nl = 0;
        while (q < len && nl < 2) {
          if (text[q] == '\n') ++nl;
          ++q;
        }
        if (text[p] == '@' && q < len && text[q] == '+') break;
        while (p < len && text[p] != '\n') ++p;
        ++p;
      }
    } else {
      while (p < len && text[p] != '>') {
        while (p < len && text[p] != '\n') ++p;
        ++p;
      }
    }
    if (p < len && p > cuts.back()) cuts.push_back(p);
  }
  cuts.push_back(len);
  return cuts;
}
}  // namespace

extern "C" int pgo_count_feed_mt(pgo_counter* c, const char* text, uint64_t len, int op, int threads) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  if (len == 0) return PG_OK;
  if (text[0] != '>' && text[0] != '@') return fail(PG_ERR_FORMAT, "unsupported sequence format (expected '>' or '@')");
  if (op != PG_OP_UPDATE) c->reserve(c->distinct.load() + len);  // every new key comes from one text byte
  if (threads < 1) threads = 1;
  if (len < (1u << 16)) threads = 1;
  // chunks that do not start at the beginning of the file lack the leading header line the parser expects:
  // each shard is parsed as its own file, which is valid because it starts at a record boundary.
  std::vector<uint64_t> cuts = record_cuts(text, len, threads);
  const size_t n = cuts.size() - 1;
  std::vector<int> status(n, PG_OK);
  std::vector<std::string> errs(n);
  auto work = [&](size_t t) {
    auto sink = [&](uint64_t key) { c->apply(key, op); };
    KmerRoller<decltype(sink)> roller(c->k, sink);
    status[t] = parse_records(text + cuts[t], cuts[t + 1] - cuts[t], [&](char ch) { roller(ch); });
    if (status[t] != PG_OK) errs[t] = g_err;
  };
  if (n == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (size_t t = 0; t < n; ++t)
    if (status[t] != PG_OK) return fail(status[t], errs[t]);
  return PG_OK;
}

extern "C" int pgo_count_feed(pgo_counter* c, const char* text, uint64_t len, int op) { return pgo_count_feed_mt(c, text, len, op, 1); }

/* src/jellyfishcounter.cpp:26-49 (segments == NULL) and :51-85. */
extern "C" pgo_counter* pgo_count_create_from_buffers(const char* reads, uint64_t reads_len, const char* segments,
                                                      uint64_t segments_len, uint32_t k) {
  pgo_counter* c = pgo_count_new(k);
  if (!c) return nullptr;
  int st = PG_OK;
  if (segments) {
    st = pgo_count_feed(c, segments, segments_len, PG_OP_PRIME);
    if (st == PG_OK) st = pgo_count_feed(c, reads, reads_len, PG_OP_UPDATE);
  } else {
    st = pgo_count_feed(c, reads, reads_len, PG_OP_COUNT);
  }
  if (st != PG_OK) {
    delete c;
    return nullptr;
  }
  return c;
}

/* src/jellyfishcounter.cpp:87-95. */
extern "C" int pgo_count_lookup_ascii(const pgo_counter* c, const char* kmers, uint64_t n, uint64_t* out) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t code;
    out[i] = 0;
    if (!encode_ascii(kmers + i * c->k, c->k, &code)) continue;
    out[i] = c->get(canonical_of(code, c->k));
  }
  return PG_OK;
}

/* src/jellyfishcounter.cpp:97-104. */
extern "C" int pgo_count_lookup(const pgo_counter* c, const uint64_t* kmers, uint64_t n, uint64_t* out) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  for (uint64_t i = 0; i < n; ++i) out[i] = c->get(canonical_of(kmers[i], c->k));
  return PG_OK;
}

/* src/jellyfishcounter.cpp:106-117. */
extern "C" int pgo_count_kmer_coverage(const pgo_counter* c, uint64_t genome_kmers, uint64_t* out) {
  ld result = 0.0L, genome = 1.0L * genome_kmers;
  for (uint64_t i = 0; i < c->cap; ++i)
    if (c->keys[i].load(std::memory_order_relaxed) != pgo_counter::EMPTY) result += (1.0L * c->vals[i].load(std::memory_order_relaxed)) / genome;
  *out = (uint64_t)ceill(result);
  return PG_OK;
}

/* src/jellyfishcounter.cpp:119-126 + src/histogram.cpp:26-30. */
extern "C" int pgo_count_histogram(const pgo_counter* c, uint64_t max_count, uint64_t* bins) {
  std::fill(bins, bins + max_count + 1, 0);
  for (uint64_t i = 0; i < c->cap; ++i) {
    if (c->keys[i].load(std::memory_order_relaxed) == pgo_counter::EMPTY) continue;
    const uint64_t v = c->vals[i].load(std::memory_order_relaxed);
    if (v > 0 && v <= max_count) ++bins[v];
  }
  return PG_OK;
}

/* src/histogram.cpp:41-63 + src/sequenceutils.cpp:42-84. */
extern "C" int pgo_histogram_peak(uint64_t* h, uint64_t n, int largest_peak, uint64_t* peak) {
  if (n < 2) return fail(PG_ERR_ARG, "histogram too small");
  for (uint64_t i = 1; i + 1 < n; ++i) h[i] = (h[i - 1] + h[i] + h[i + 1]) / 3;  // in place (:41-45)
  std::vector<uint64_t> ids, vals;
  bool direction = false;
  uint64_t prev = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = h[i];
    if (prev < v) {
      direction = false;
    } else if (prev > v) {
      if (!direction) {
        ids.push_back(i - 1);
        vals.push_back(prev);
      }
      direction = true;
    }
    prev = v;
  }
  if (ids.empty()) return fail(PG_ERR_ARG, "no peak found in kmer-count histogram");
  if (ids.size() < 2) {
    *peak = ids[0];
    return PG_OK;
  }
  uint64_t largest, second, largest_id, second_id;
  if (vals[0] < vals[1]) {
    largest = vals[1]; largest_id = ids[1]; second = vals[0]; second_id = ids[0];
  } else {
    largest = vals[0]; largest_id = ids[0]; second = vals[1]; second_id = ids[1];
  }
  for (size_t i = 0; i < vals.size(); ++i) {
    if (vals[i] > largest) {
      second = largest; second_id = largest_id; largest = vals[i]; largest_id = ids[i];
    } else if (vals[i] > second && vals[i] != largest) {
      second = vals[i]; second_id = ids[i];
    }
  }
  *peak = largest_peak ? largest_id : second_id;
  return PG_OK;
}

/* src/jellyfishcounter.cpp:119-153. */
extern "C" int pgo_count_compute_histogram(const pgo_counter* c, uint64_t max_count, int largest_peak,
                                           const char* filename, uint64_t* peak) {
  std::vector<uint64_t> bins(max_count + 1);
  pgo_count_histogram(c, max_count, bins.data());
  if (filename && *filename) {
    std::ofstream f(filename);
    if (!f.good()) return fail(PG_ERR_IO, "cannot create histogram file");
    for (uint64_t i = 0; i <= max_count; ++i) f << i << '\t' << bins[i] << std::endl;
  }
  int st = pgo_histogram_peak(bins.data(), bins.size(), largest_peak, peak);
  if (st != PG_OK) return st;
  if (filename && *filename) {
    std::ofstream f(filename, std::ios::app);
    f << "parameters\t" << *peak / 2.0 << '\t' << *peak << std::endl;
  }
  return PG_OK;
}

extern "C" uint64_t pgo_count_distinct(const pgo_counter* c) { return c ? c->distinct.load() : 0; }
extern "C" void pgo_count_destroy(pgo_counter* c) { delete c; }

/* =================================================================================================
 * Probability model: src/probabilitytable.cpp:7-19,47-65,75-85 and src/copynumber.cpp:14-41.
 * ================================================================================================= */
namespace {

double get_error_param(double cov) {  // probabilitytable.cpp:7-19
  if (cov < 10.0) return 0.99;
  if (cov < 20) return 0.95;
  if (cov < 40) return 0.9;
  return 0.8;
}

ld poisson(ld mean, unsigned value) {  // probabilitytable.cpp:75-81
  ld sum = 0.0L;
  int v = (int)value;
  for (size_t i = 1; i <= value; ++i) sum += std::log((double)i);
  ld log_val = -mean + v * logl(mean) - sum;
  return expl(log_val);
}

ld geometric(ld p, unsigned value) { return powl(1.0L - p, value) * p; }  // :83-85

struct CN {
  ld p[3];
};

CN compute_probability(uint16_t cov, uint16_t count, ld reg) {  // probabilitytable.cpp:55-65
  ld c0 = geometric(get_error_param(cov), count);
  ld c1 = poisson(cov / 2.0, count);
  ld c2 = poisson(cov, count);
  CN r;
  if (reg > 0) {  // copynumber.cpp:22-28,36-37
    ld sum = c0 + c1 + c2 + 3.0L * reg;
    r.p[0] = (c0 + reg) / sum;
    r.p[1] = (c1 + reg) / sum;
    r.p[2] = 1.0L - r.p[0] - r.p[1];
  } else {
    r.p[0] = c0; r.p[1] = c1; r.p[2] = c2;
  }
  return r;
}

CN table_get(const pg_probtable* t, uint16_t cov, uint16_t count) {  // probabilitytable.cpp:47-53
  if (t->log_p && cov >= t->cov_min && cov < t->cov_max && count < t->count_max) {
    const double* e = t->log_p + ((size_t)count * (t->cov_max - t->cov_min) + (cov - t->cov_min)) * 3;
    CN r;
    for (int i = 0; i < 3; ++i) r.p[i] = std::isinf(e[i]) && e[i] < 0 ? 0.0L : expl((ld)e[i]);
    return r;
  }
  return compute_probability(cov, count, (ld)t->regularization);
}

}  // namespace

extern "C" double pgo_log_probability(uint16_t cov, uint16_t count, double regularization, int cn) {
  CN c = compute_probability(cov, count, (ld)regularization);
  return (double)logl(c.p[cn]);
}

/* =================================================================================================
 * Panel accessors (src/kmerpath.cpp:33-48, src/biallelicuniquekmers.cpp / multiallelicuniquekmers.cpp).
 * ================================================================================================= */
namespace {

struct VariantView {
  const pg_panel* p;
  uint32_t v;
  uint32_t a_begin, a_end, k_begin, k_end;
  VariantView(const pg_panel* panel, uint32_t vi) : p(panel), v(vi) {
    a_begin = p->allele_offsets[v]; a_end = p->allele_offsets[v + 1];
    k_begin = p->kmer_offsets[v]; k_end = p->kmer_offsets[v + 1];
  }
  uint32_t n_alleles() const { return a_end - a_begin; }
  uint32_t n_kmers() const { return k_end - k_begin; }
  uint16_t allele_id(uint32_t ai) const { return p->allele_ids[a_begin + ai]; }
  int find_allele(uint16_t id) const {
    for (uint32_t ai = 0; ai < n_alleles(); ++ai)
      if (allele_id(ai) == id) return (int)ai;
    return -1;
  }
  bool undefined_idx(uint32_t ai) const { return p->allele_undefined[a_begin + ai] != 0; }
  bool is_undefined(uint16_t id) const {  // biallelicuniquekmers.cpp: is_undefined_allele
    int ai = find_allele(id);
    return ai >= 0 && undefined_idx((uint32_t)ai);
  }
  unsigned on_allele_idx(uint32_t kmer, uint32_t ai) const {  // kmerpath.cpp:33-48
    uint32_t off = p->allele_kmer_offset[a_begin + ai];
    if (kmer < off || kmer >= off + 32) return 0;
    return (p->allele_kmer_mask[a_begin + ai] >> (kmer - off)) & 1u;
  }
  uint16_t count(uint32_t kmer) const { return p->kmer_counts[k_begin + kmer]; }
  uint16_t coverage() const { return p->coverage[v]; }
  uint16_t path_allele(uint32_t path) const { return p->path_to_allele[(size_t)v * p->n_paths + path]; }
  uint16_t max_allele() const {
    uint16_t m = 0;
    for (uint32_t ai = 0; ai < n_alleles(); ++ai) m = std::max(m, allele_id(ai));
    return m;
  }
};

/* src/emissionprobabilitycomputer.cpp:9-53.  e is indexed by allele INDEX (position in the variant's
 * allele list). */
struct Emission {
  uint32_t n;
  std::vector<ld> e;
  bool all_zeros;
  ld get(uint32_t i1, uint32_t i2) const { return all_zeros ? 1.0L : e[(size_t)i1 * n + i2]; }
};

Emission compute_emission(const VariantView& vv, const pg_probtable* t) {
  Emission em;
  em.n = vv.n_alleles();
  em.e.assign((size_t)em.n * em.n, 0.0L);
  em.all_zeros = true;
  uint32_t K = vv.n_kmers();
  std::vector<CN> cn(K);
  for (uint32_t k = 0; k < K; ++k) cn[k] = table_get(t, vv.coverage(), vv.count(k));
  for (uint32_t i1 = 0; i1 < em.n; ++i1) {
    for (uint32_t i2 = 0; i2 < em.n; ++i2) {
      bool u1 = vv.undefined_idx(i1), u2 = vv.undefined_idx(i2);
      ld result = 1.0L;
      for (uint32_t k = 0; k < K; ++k) {
        unsigned c = vv.on_allele_idx(k, i1) + vv.on_allele_idx(k, i2);
        if (u1 && u2) {
          result *= (1.0L / 3.0L) * (cn[k].p[0] + cn[k].p[1] + cn[k].p[2]);
        } else if (u1 || u2) {
          unsigned c1 = std::min(c + 1, 2u);  // reference asserts c < 2 here (:44)
          result *= 0.5L * (cn[k].p[std::min(c, 2u)] + cn[k].p[c1]);
        } else {
          result *= cn[k].p[c];
        }
      }
      em.e[(size_t)i1 * em.n + i2] = result;
      if (result > 0) em.all_zeros = false;
    }
  }
  return em;
}

}  // namespace

extern "C" int pgo_emission_run(const pg_panel* panel, const pg_probtable* table, const uint64_t* em_offsets,
                                double* emissions, double* log_scale) {
  for (uint32_t v = 0; v < panel->n_variants; ++v) {
    VariantView vv(panel, v);
    Emission em = compute_emission(vv, table);
    uint32_t dim = (uint32_t)vv.max_allele() + 1;
    double* out = emissions + em_offsets[v];
    std::fill(out, out + (size_t)dim * dim, 0.0);
    ld mx = 0.0L;
    for (ld x : em.e) mx = std::max(mx, x);
    if (em.all_zeros) {
      log_scale[v] = 0.0;
      for (uint32_t i1 = 0; i1 < em.n; ++i1)
        for (uint32_t i2 = 0; i2 < em.n; ++i2) out[(size_t)vv.allele_id(i1) * dim + vv.allele_id(i2)] = 1.0;
      continue;
    }
    log_scale[v] = (double)logl(mx);
    for (uint32_t i1 = 0; i1 < em.n; ++i1)
      for (uint32_t i2 = 0; i2 < em.n; ++i2)
        out[(size_t)vv.allele_id(i1) * dim + vv.allele_id(i2)] = (double)(em.e[(size_t)i1 * em.n + i2] / mx);
  }
  return PG_OK;
}

/* =================================================================================================
 * Fill: src/commands.cpp:76-152 (count lookup, u16 truncation at update_readcount) and
 * src/kmerparser.cpp:30-49 (local coverage).
 * ================================================================================================= */
extern "C" int pgo_fill_counts(const pgo_counter* c, uint64_t peak, uint32_t n_chrom, pg_panel* panels) {
  for (uint32_t ch = 0; ch < n_chrom; ++ch) {
    pg_panel& p = panels[ch];
    if (!p.kmer_codes || !p.flank_offsets) return fail(PG_ERR_ARG, "panel lacks kmer_codes/flank arrays");
    for (uint32_t v = 0; v < p.n_variants; ++v) {
      for (uint32_t k = p.kmer_offsets[v]; k < p.kmer_offsets[v + 1]; ++k) {
        uint64_t cnt;
        pgo_count_lookup(c, &p.kmer_codes[k], 1, &cnt);
        p.kmer_counts[k] = (uint16_t)cnt;  // size_t -> unsigned short (commands.cpp:118,130)
      }
      uint64_t total_cov = 0, total_kmers = 0, min_cov = peak / 4, max_cov = peak * 4;
      for (uint32_t f = p.flank_offsets[v]; f < p.flank_offsets[v + 1]; ++f) {
        uint64_t cnt;
        pgo_count_lookup(c, &p.flank_codes[f], 1, &cnt);
        if (cnt < min_cov || cnt > max_cov) continue;
        total_cov += cnt;
        total_kmers += 1;
      }
      p.coverage[v] = (uint16_t)((total_kmers > 0 && total_cov > 0) ? total_cov / total_kmers : peak);
    }
  }
  return PG_OK;
}

/* =================================================================================================
 * HMM: src/columnindexer.cpp:8-33, src/transitionprobabilitycomputer.cpp:8-39, src/hmm.cpp:76-110,
 * 175-405, src/genotypingresult.cpp, output post-processing of src/graph.cpp:206-240.
 * ================================================================================================= */
namespace {

typedef std::map<std::pair<uint16_t, uint16_t>, ld> GLMap;  // genotypingresult.hpp: genotype_to_likelihood

void add_to_likelihood(GLMap& m, uint16_t a1, uint16_t a2, ld value) {  // genotypingresult.cpp:16-28
  if (a1 < a2) m[{a1, a2}] += value;
  else m[{a2, a1}] += value;
}

void normalize_map(GLMap& m) {  // genotypingresult.cpp:200-210
  ld s = 0.0L;
  for (auto& kv : m) s += kv.second;
  if (s > 0)
    for (auto& kv : m) kv.second /= s;
}

std::pair<int, int> likeliest(const GLMap& m) {  // genotypingresult.cpp:149-180
  if (m.empty()) return {-1, -1};
  ld best = 0.0L;
  std::pair<uint16_t, uint16_t> bg(0, 0);
  for (auto& kv : m)
    if (kv.second >= best) { best = kv.second; bg = kv.first; }
  for (auto& kv : m)
    if (kv.first != bg && fabsl(kv.second - best) < 0.0000000001) return {-1, -1};
  if (best > 0.0L) return {bg.first, bg.second};
  return {-1, -1};
}

int run_chromosome(const pg_panel* panel, const pg_probtable* table, const pg_hmm_params* prm, pg_hmm_result* res) {
  const uint32_t V = panel->n_variants, P = panel->n_paths;
  if (V == 0) return PG_OK;
  // paths in use (get_path_ids with only_include, biallelicuniquekmers.cpp:100-117)
  std::vector<uint16_t> paths;
  if (prm->only_paths) {
    for (uint32_t i = 0; i < prm->n_only_paths; ++i)
      if (prm->only_paths[i] < P) paths.push_back(prm->only_paths[i]);
  } else {
    for (uint32_t i = 0; i < P; ++i) paths.push_back((uint16_t)i);
  }
  const size_t np = paths.size();
  if (np == 0) return fail(PG_ERR_ARG, "HMM::index_columns: column 0 is not covered by any paths.");
  const size_t S = np * np;

  // ColumnIndexer (columnindexer.cpp:8-33)
  std::vector<uint32_t> cols;
  for (uint32_t v = 0; v < V; ++v) {
    VariantView vv(panel, v);
    bool all_absent = true;
    for (size_t i = 0; i < np; ++i) {
      uint16_t a = vv.path_allele(paths[i]);
      if (a != 0 && !vv.is_undefined(a)) all_absent = false;
    }
    res->is_column[v] = all_absent ? 0 : 1;
    if (!all_absent) cols.push_back(v);
  }
  const size_t C = cols.size();
  std::vector<GLMap> gl(V);

  // per-column allele index of each path (index into the variant's allele list)
  auto path_allele_idx = [&](const VariantView& vv, std::vector<uint32_t>& out) {
    out.resize(np);
    for (size_t i = 0; i < np; ++i) {
      int ai = vv.find_allele(vv.path_allele(paths[i]));
      out[i] = (uint32_t)ai;
    }
  };
  auto transitions = [&](uint64_t from, uint64_t to, ld t[3]) {  // transitionprobabilitycomputer.cpp:8-39
    if (prm->uniform) {
      t[0] = t[1] = t[2] = 1.0L;
      return;
    }
    ld distance = (to - from) * 0.000004L * ((ld)prm->recombrate) * (ld)prm->effective_N;
    ld recomb = (1.0L - expl(-distance / (ld)np)) * (1.0L / (ld)np);
    ld no_recomb = expl(-distance / (ld)np) + recomb;
    t[0] = no_recomb * no_recomb; t[1] = no_recomb * recomb; t[2] = recomb * recomb;
  };

  // forward pass (hmm.cpp:175-273), all columns kept (the reference's sqrt(N) checkpointing only
  // trades memory for recomputation; values are identical)
  std::vector<std::vector<ld>> fwd(C);
  std::vector<ld> fnorm(C);
  std::vector<uint32_t> aidx;
  for (size_t c = 0; c < C; ++c) {
    VariantView vv(panel, cols[c]);
    Emission em = compute_emission(vv, table);
    path_allele_idx(vv, aidx);
    std::vector<ld>& cur = fwd[c];
    cur.resize(S);
    std::vector<ld> hi(np, 0.0L), hj(np, 0.0L);
    ld hij = 0.0L, t[3] = {0, 0, 0};
    if (c > 0) {
      const std::vector<ld>& prev = fwd[c - 1];
      size_t i = 0;
      for (size_t p1 = 0; p1 < np; ++p1)
        for (size_t p2 = 0; p2 < np; ++p2) {
          hi[p1] += prev[i]; hj[p2] += prev[i]; hij += prev[i]; ++i;
        }
      transitions(panel->positions[cols[c - 1]], panel->positions[cols[c]], t);
    }
    ld norm = 0.0L;
    size_t i = 0;
    for (size_t p1 = 0; p1 < np; ++p1)
      for (size_t p2 = 0; p2 < np; ++p2) {
        ld prev_cell = 1.0L;
        if (c > 0) {
          ld pc = fwd[c - 1][i];
          prev_cell = t[0] * pc + t[1] * (hi[p1] + hj[p2] - 2 * pc) + t[2] * (hij - hi[p1] - hj[p2] + pc);
        }
        ld cell = prev_cell * em.get(aidx[p1], aidx[p2]);
        cur[i++] = cell;
        norm += cell;
      }
    if (norm > 0.0L) {
      for (auto& x : cur) x /= norm;
      fnorm[c] = norm;
    } else {
      ld u = 1.0L / (ld)S;
      for (auto& x : cur) x = u;
      fnorm[c] = 1.0L;
    }
  }

  // backward pass with posterior accumulation (hmm.cpp:275-405)
  std::vector<ld> prev_b;
  std::vector<uint32_t> aidx_next;
  for (size_t cc = C; cc-- > 0;) {
    VariantView vv(panel, cols[cc]);
    path_allele_idx(vv, aidx);
    std::vector<ld> hi(np, 0.0L), hj(np, 0.0L), helper(S, 0.0L);
    ld hij = 0.0L, t[3] = {0, 0, 0};
    const bool last = cc + 1 == C;
    if (!last) {
      VariantView vn(panel, cols[cc + 1]);
      Emission emn = compute_emission(vn, table);
      path_allele_idx(vn, aidx_next);
      transitions(panel->positions[cols[cc]], panel->positions[cols[cc + 1]], t);
      size_t i = 0;
      for (size_t p1 = 0; p1 < np; ++p1)
        for (size_t p2 = 0; p2 < np; ++p2) {
          ld h = prev_b[i] * emn.get(aidx_next[p1], aidx_next[p2]);
          helper[i] = h; hi[p1] += h; hj[p2] += h; hij += h; ++i;
        }
    }
    std::vector<ld> cur(S);
    ld norm = 0.0L;
    size_t i = 0;
    for (size_t p1 = 0; p1 < np; ++p1)
      for (size_t p2 = 0; p2 < np; ++p2) {
        ld cell = 1.0L;
        if (!last) {
          ld h = helper[i];
          cell = t[0] * h + t[1] * (hi[p1] + hj[p2] - 2 * h) + t[2] * (hij - hi[p1] - hj[p2] + h);
        }
        cur[i] = cell;
        norm += cell;
        ld fb = fwd[cc][i] * cell;
        add_to_likelihood(gl[cols[cc]], vv.allele_id(aidx[p1]), vv.allele_id(aidx[p2]), fb * fnorm[cc]);
        ++i;
      }
    if (norm > 0.0L) {
      for (auto& x : cur) x /= norm;
    } else {
      ld u = 1.0L / (ld)S;
      for (auto& x : cur) x = u;
    }
    prev_b.swap(cur);
    fwd[cc].clear();
    fwd[cc].shrink_to_fit();
  }

  // outputs
  for (uint32_t v = 0; v < V; ++v) {
    VariantView vv(panel, v);
    res->unique_kmers[v] = (uint16_t)vv.n_kmers();
    res->coverage[v] = vv.coverage();
    GLMap m = gl[v];
    if (prm->normalize) normalize_map(m);  // hmm.cpp:41-45
    const uint64_t off = res->gl_offsets[v];
    const uint64_t n = res->gl_offsets[v + 1] - off;
    std::fill(res->likelihoods + off, res->likelihoods + off + n, 0.0);
    for (auto& kv : m) {  // genotypingresult.cpp:48-67
      uint64_t idx = ((uint64_t)kv.first.second * (kv.first.second + 1)) / 2 + kv.first.first;
      if (idx >= n) return fail(PG_ERR_ARG, "genotype does not match number of alleles");
      res->likelihoods[off + idx] = (double)kv.second;
    }
    // what Graph::write_genotypes prints (graph.cpp:206-240): always from normalised likelihoods
    GLMap nm = gl[v];
    normalize_map(nm);  // commands.cpp:981-987
    if (nm.empty()) add_to_likelihood(nm, 0, 0, 1.0L);
    std::vector<uint16_t> defined{0};
    uint16_t maxa = vv.max_allele();
    for (uint16_t a = 1; a <= maxa; ++a)
      if (!vv.is_undefined(a)) defined.push_back(a);
    GLMap spec;
    if (defined.size() < (size_t)maxa + 1) {  // get_specific_likelihoods (genotypingresult.cpp:70-96)
      std::map<uint16_t, uint16_t> index;
      for (uint16_t i = 0; i < defined.size(); ++i) index[defined[i]] = i;
      ld sum = 0.0L;
      for (auto& kv : nm) {
        if (!index.count(kv.first.first) || !index.count(kv.first.second)) continue;
        add_to_likelihood(spec, index[kv.first.first], index[kv.first.second], kv.second);
        sum += kv.second;
      }
      if (sum > 0)
        for (auto& kv : spec) kv.second /= sum;
    } else {
      spec = nm;
    }
    std::pair<int, int> g = likeliest(spec);
    res->genotype[2 * v] = (int16_t)g.first;
    res->genotype[2 * v + 1] = (int16_t)g.second;
    res->quality[v] = 0;
    if (g.first != -1) {  // get_genotype_quality (genotypingresult.cpp:118-137)
      ld like = 0.0L;
      auto it = spec.find({(uint16_t)g.first, (uint16_t)g.second});
      if (it != spec.end()) like = it->second;
      ld wrong = 1.0L - like;
      res->quality[v] = wrong > 0.0 ? (uint32_t)(size_t)(-10 * log10l(wrong)) : 10000;
    }
  }
  return PG_OK;
}

}  // namespace

extern "C" int pgo_hmm_run(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                           const pg_hmm_params* params, pg_hmm_result* results) {
  for (uint32_t c = 0; c < n_chrom; ++c) {
    int st = run_chromosome(&panels[c], table, params, &results[c]);
    if (st != PG_OK) return st;
  }
  return PG_OK;
}

extern "C" int pgo_hmm_run_mt(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                              const pg_hmm_params* params, pg_hmm_result* results, int threads) {
  if (threads <= 1) return pgo_hmm_run(n_chrom, panels, table, params, results);
  std::vector<int> status(n_chrom, PG_OK);
  std::vector<std::string> errs(n_chrom);
  std::mutex mu;
  uint32_t next = 0;
  std::vector<std::thread> pool;
  for (int t = 0; t < std::min<int>(threads, (int)n_chrom); ++t) {
    pool.emplace_back([&]() {
      while (true) {
        uint32_t c;
        {
          std::lock_guard<std::mutex> lk(mu);
          if (next >= n_chrom) return;
          c = next++;
        }
        status[c] = run_chromosome(&panels[c], table, params, &results[c]);
        if (status[c] != PG_OK) errs[c] = g_err;
      }
    });
  }
  for (auto& th : pool) th.join();
  for (uint32_t c = 0; c < n_chrom; ++c)
    if (status[c] != PG_OK) return fail(status[c], errs[c]);
  return PG_OK;
}
