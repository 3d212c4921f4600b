/*
 * pg_oracle.h — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (pangenie_b200/) never does.  Every function mirrors the signature of the
 * product entry point of the same name in include/pangenie_b200.h (prefix pgo_ instead of pg_) so the
 * parity tests swap one for the other.
 *
 * Parity status: emission + HMM restatement is pinned against the reference's own sources compiled
 * unmodified (oracle/_ref, built by oracle/Makefile) and against the reference's test vectors
 * (tests/golden/).  The k-mer counting restatement follows the published behaviour of jellyfish 2.x
 * (pinned 2.2.10 in the reference's environment.yml:24; source absent from /root/reference) and is
 * pinned by the reference's golden fixtures at that boundary (tests/golden/counting/*: KmerCounterTest
 * vectors and the filled counts of tests/data/region_UniqueKmersList.cereal).
 */
#ifndef PG_ORACLE_H
#define PG_ORACLE_H
#include "../include/pangenie_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* pgo_last_error(void);

/* ---- counting (src/jellyfishcounter.{hpp,cpp} + jellyfish 2.x semantics) ---- */
typedef struct pgo_counter pgo_counter;
pgo_counter* pgo_count_new(uint32_t k);
int pgo_count_feed(pgo_counter* c, const char* text, uint64_t len, int op);
/** multi-threaded feed over `threads` host threads (used by the CPU baseline). */
int pgo_count_feed_mt(pgo_counter* c, const char* text, uint64_t len, int op, int threads);
pgo_counter* pgo_count_create_from_buffers(const char* reads, uint64_t reads_len, const char* segments,
                                           uint64_t segments_len, uint32_t k);
int pgo_count_lookup_ascii(const pgo_counter* c, const char* kmers, uint64_t n, uint64_t* out);
int pgo_count_lookup(const pgo_counter* c, const uint64_t* kmers, uint64_t n, uint64_t* out);
int pgo_count_kmer_coverage(const pgo_counter* c, uint64_t genome_kmers, uint64_t* out);
int pgo_count_histogram(const pgo_counter* c, uint64_t max_count, uint64_t* bins);
int pgo_count_compute_histogram(const pgo_counter* c, uint64_t max_count, int largest_peak,
                                const char* filename, uint64_t* peak);
uint64_t pgo_count_distinct(const pgo_counter* c);
void pgo_count_destroy(pgo_counter* c);
int pgo_histogram_peak(uint64_t* bins, uint64_t n, int largest_peak, uint64_t* peak);

/* ---- probability model (src/probabilitytable.cpp, src/copynumber.cpp) ---- */
/** P(CN=cn ; cov,count) straight from the formulas (no table), long double, returned as ln. */
double pgo_log_probability(uint16_t cov, uint16_t count, double regularization, int cn);

/* ---- fill (src/commands.cpp:76-152, src/kmerparser.cpp:30-49) ---- */
int pgo_fill_counts(const pgo_counter* c, uint64_t kmer_abundance_peak, uint32_t n_chrom, pg_panel* panels);

/* ---- emission + HMM (src/emissionprobabilitycomputer.cpp, src/hmm.cpp, ...) ---- */
int pgo_emission_run(const pg_panel* panel, const pg_probtable* table, const uint64_t* em_offsets,
                     double* emissions, double* log_scale);
int pgo_hmm_run(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                const pg_hmm_params* params, pg_hmm_result* results);
/** As pgo_hmm_run, one host thread per chromosome up to `threads` (the reference's run_genotyping
 *  dispatch, src/commands.cpp:949-978). */
int pgo_hmm_run_mt(uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                   const pg_hmm_params* params, pg_hmm_result* results, int threads);

/* ---- HaplotypeSampler (src/haplotypesampler.cpp, samplingemissions.cpp, samplingtransitions.cpp): groundwork for
 * SURVEY.md 8f row 3; no product counterpart yet.  n_out = size + (add_reference ? 1 : 0);
 * sampled_paths [n_out][V], best_scores [size], new_path_to_allele [V][n_out], new_kmer_count [V], new_counts [<= K]. */
int pgo_haplotype_sample(const pg_panel* panel, uint32_t size, double recombrate, double effective_N, int add_reference,
                         uint16_t allele_penalty, uint64_t* sampled_paths, uint32_t* best_scores,
                         uint16_t* new_path_to_allele, uint32_t* new_kmer_count, uint16_t* new_counts);

/* ---- index stage: unique-k-mer selection (src/stepwiseuniquekmercomputer.cpp:11-93, 95-197, 227-264); see
 * pg_oracle_index.cpp.  `graph` = COUNT of the path-segment file. ---- */
typedef struct pgo_unique_kmers pgo_unique_kmers;
pgo_unique_kmers* pgo_unique_kmers_compute(const pgo_counter* graph, const pg_variants* in);
int pgo_unique_kmers_panel(pgo_unique_kmers* u, pg_panel* out);
void pgo_unique_kmers_free(pgo_unique_kmers* u);

#ifdef __cplusplus
}
#endif
#endif
