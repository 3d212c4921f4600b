/*
 * pangenie_b200.h — C-ABI of the B200-native PanGenie genotyping hot path.
 *
 * The reference (eblerjana/pangenie, C++20, CPU only) has no FFI; its hot path is reached through two
 * C++ seams.  Every entry point below names the reference interface it replaces (file:line relative
 * to the reference tree):
 *
 *   Seam 1  KmerCounter                 src/kmercounter.hpp:9-24, src/jellyfishcounter.{hpp,cpp}
 *   Seam 2  HMM over UniqueKmers        src/hmm.hpp:38-46 (ctor + move_genotyping_result),
 *                                       src/emissionprobabilitycomputer.cpp:9-53,
 *                                       src/columnindexer.cpp:8-33, src/transitionprobabilitycomputer.cpp:8-39
 *   Glue    fill_read_kmercounts        src/commands.cpp:76-152, src/kmerparser.cpp:30-49
 *           run_genotype_command        src/commands.cpp:730-1084 (the `PanGenie -f` stage)
 *
 * Conventions: plain C structs, caller-owned HOST buffers unless a name ends in `_device`, int status
 * (0 = PG_OK) + pg_last_error() (thread-local string), no exceptions / STL / torch types across the
 * boundary, explicit device ordinal per handle, no global state.  Read-only calls on a finished
 * counter (lookups, histogram, coverage) are safe from several host threads: lookups run on private streams and buffers,
 * the histogram / coverage calls serialise on a per-counter lock around their shared scratch.  There is NO CPU fallback: every compute entry point
 * fails with PG_ERR_CUDA when no sm_100-class device can be opened.
 */
#ifndef PANGENIE_B200_H
#define PANGENIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_ARG 1      /* invalid argument / malformed panel */
#define PG_ERR_CUDA 2     /* CUDA runtime failure or no usable device */
#define PG_ERR_FORMAT 3   /* unsupported sequence file layout (e.g. multi-line FASTQ) */
#define PG_ERR_FULL 4     /* k-mer table capacity exhausted (count-all mode; raise hash_size) */
#define PG_ERR_IO 5

/** Last error message of the calling thread ("" if none). */
const char* pg_last_error(void);
/** Library version string. */
const char* pg_version(void);
/** Number of usable CUDA devices (0 if none; never throws). */
int pg_device_count(void);
/** Kernels this library has launched from the calling host thread so far (measurement hook: bench.py reports the
 *  difference over its timed region as `gpu_launches`). */
uint64_t pg_kernel_launches(void);

/* ------------------------------------------------------------------------------------------------
 * Seam 1: KmerCounter  (src/kmercounter.hpp:9-24)
 * ---------------------------------------------------------------------------------------------- */

typedef struct pg_counter pg_counter;

/* Counting operations, as in `enum OPERATION { COUNT, PRIME, UPDATE }` (src/jellyfishcounter.hpp:21,
 * 49-65): COUNT = hash.add(mer,1); PRIME = hash.set(mer) (insert key with value 0);
 * UPDATE = hash.update_add(mer,1) (+1 only if the key is already present). */
#define PG_OP_COUNT 0
#define PG_OP_PRIME 1
#define PG_OP_UPDATE 2

/**
 * Replaces `JellyfishCounter(readfile, kmer_size, nr_threads, hash)` (src/jellyfishcounter.cpp:26-49,
 * count everything) when `segments_path == NULL`, and `JellyfishCounter(readfile, {kmerfiles}, ...)`
 * (src/jellyfishcounter.cpp:51-85, PRIME with the graph k-mers then UPDATE with the reads — the
 * default of `PanGenie`, src/pangenie-genotype.cpp:37,110) otherwise.  Canonical k-mers, k <= 32.
 * `hash_size` is the `-e` value: in PRIME/UPDATE mode the table is sized from the segment file and
 * `hash_size` is only a lower bound; in count-all mode it is the number of distinct k-mers the table
 * must hold (jellyfish grows its table, this one fails with PG_ERR_FULL).  `nr_threads` of the
 * reference has no meaning here.  Both files are streamed: file -> ring of pinned staging buffers -> device, never held in
 * host memory as a whole (uncompressed FASTA / 4-line FASTQ; anything else fails with PG_ERR_FORMAT).
 * Returns NULL on error.
 */
pg_counter* pg_count_create(const char* reads_path, const char* segments_path, uint32_t k,
                            uint64_t hash_size, int device);

/** Same, with the file CONTENTS given as host buffers (segments may be NULL). */
pg_counter* pg_count_create_from_buffers(const char* reads, uint64_t reads_len, const char* segments,
                                         uint64_t segments_len, uint32_t k, uint64_t hash_size,
                                         int device);

/** Incremental form used for sharded / streamed counting: an empty table able to hold
 *  `max_distinct` keys, then any sequence of pg_count_feed calls.  Each fed buffer must start at a
 *  record boundary and hold whole records ('>' FASTA or '@' 4-line FASTQ). */
pg_counter* pg_count_new(uint32_t k, uint64_t max_distinct, int device);
int pg_count_feed(pg_counter* c, const char* text, uint64_t len, int op);
/** As pg_count_feed but `d_text` is DEVICE memory on the counter's device (no copies). */
int pg_count_feed_device(pg_counter* c, const char* d_text, uint64_t len, int op);

/** `getKmerAbundance(std::string kmer)` (src/jellyfishcounter.cpp:87-95) for n k-mers of length k laid
 *  out back to back in `kmers` (n*k chars, no separators): canonicalise, look up, 0 if absent or if
 *  the k-mer contains a non-ACGT character. */
int pg_count_lookup_ascii(const pg_counter* c, const char* kmers, uint64_t n, uint64_t* out);
/** Same for 2-bit packed k-mers (A=0,C=1,G=2,T=3, first base most significant, value < 4^k);
 *  canonicalised on the device. Replaces `getKmerAbundance(jellyfish::mer_dna)` (:97-104). */
int pg_count_lookup(const pg_counter* c, const uint64_t* kmers, uint64_t n, uint64_t* out);

/** `computeKmerCoverage(genome_kmers)` (src/jellyfishcounter.cpp:106-117): ceil(sum(counts)/genome_kmers). */
int pg_count_kmer_coverage(const pg_counter* c, uint64_t genome_kmers, uint64_t* out);

/** Abundance histogram of all keys with count > 0 (src/jellyfishcounter.cpp:119-126): bins[v] = number
 *  of keys with count v for v in [0, max_count] (bins[0] stays 0; counts > max_count are dropped). */
int pg_count_histogram(const pg_counter* c, uint64_t max_count, uint64_t* bins);

/** `computeHistogram(max_count, largest_peak, filename)` (src/jellyfishcounter.cpp:119-153): histogram,
 *  optional `<filename>` dump in the reference's format, in-place smoothing (src/histogram.cpp:41-45),
 *  peak search (:47-63) and peak choice (src/sequenceutils.cpp:42-84). `filename` may be NULL. */
int pg_count_compute_histogram(const pg_counter* c, uint64_t max_count, int largest_peak,
                               const char* filename, uint64_t* kmer_abundance_peak);

/** Device views for the cross-GPU exchange.  `slots_addr`: the table itself, capacity/4 buckets of 64 bytes
 *  (4 u64 keys, 4 u32 counts, 16 B padding; 16 bytes per slot) — broadcast it (as bytes) after PRIME so every GPU holds
 *  a layout-identical table.
 *  `counts_addr`: a contiguous u32[capacity] staging array; pg_count_export_counts copies the counts into it,
 *  the host framework all-reduces it (torch.distributed / NCCL), pg_count_import_counts writes the sums back. */
int pg_count_device_arrays(const pg_counter* c, uint64_t* slots_addr, uint64_t* counts_addr, uint64_t* capacity);
int pg_count_export_counts(pg_counter* c);
int pg_count_import_counts(pg_counter* c);
/** Rewrites the keys of a PRIMEd table (all counts still zero) into the canonical layout: every run of consecutive full
 *  buckets sorted by (home bucket, k-mer).  The layout then depends on the SET of primed k-mers only, not on the order the
 *  insertions happened to win, so every GPU that PRIMEs the same segment file into a table of the same capacity holds the
 *  identical array and the per-GPU count arrays can be added position by position: the sample needs ONE all-reduce and no
 *  broadcast of the table.  Lookups are unaffected. */
int pg_count_canonicalize(pg_counter* c);
/** The same exchange in pieces, so the contiguous staging array need not be as large as the table: the buffer holds
 *  `n_slots` u32 counts (multiple of 4); export/import move the counts of slots [first_slot, first_slot + n_slots). */
int pg_count_exchange_buffer(pg_counter* c, uint64_t n_slots, uint64_t* addr);
int pg_count_export_range(pg_counter* c, uint64_t first_slot, uint64_t n_slots);
int pg_count_import_range(pg_counter* c, uint64_t first_slot, uint64_t n_slots);
/** k-mers processed / device milliseconds of the last pg_count_feed* call (measurement hooks); pg_count_last_probe_ms:
 *  the part of it spent in the probe passes over the partition buffers (0 when the pass probed directly), and their number. */
uint64_t pg_count_kmers_seen(const pg_counter* c);
double pg_count_last_ms(const pg_counter* c);
double pg_count_last_probe_ms(const pg_counter* c, uint32_t* n_passes);

/** Number of distinct keys / slots (diagnostics). */
uint64_t pg_count_distinct(const pg_counter* c);
uint64_t pg_count_capacity(const pg_counter* c);
void pg_count_destroy(pg_counter* c);

/** Pure host helper (no device): smoothing + peak search + peak choice on a given histogram
 *  (src/histogram.cpp:41-63, src/sequenceutils.cpp:42-84). `bins` has n entries and is modified.
 *  Returns PG_ERR_ARG if no peak exists (the reference throws). */
int pg_histogram_peak(uint64_t* bins, uint64_t n, int largest_peak, uint64_t* peak);

/* ------------------------------------------------------------------------------------------------
 * Panel: flat SoA form of `std::vector<std::shared_ptr<UniqueKmers>>` for ONE chromosome
 * (src/uniquekmers.hpp:22-69, src/biallelicuniquekmers.hpp:106-114, src/multiallelicuniquekmers.hpp:
 * 105-113, src/kmerpath.hpp:26-33).  All arrays are host memory owned by the caller.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t n_variants;              /* V */
  uint32_t n_paths;                 /* P: get_nr_paths(); identical for every variant          */
  const uint64_t* positions;        /* [V]   get_variant_position()                            */
  const uint16_t* path_to_allele;   /* [V*P] get_allele(path) ; row = variant                  */
  uint16_t* coverage;               /* [V]   get_coverage()  (written by pg_fill_*)            */
  const uint32_t* kmer_offsets;     /* [V+1] CSR over the unique k-mers of each variant        */
  uint16_t* kmer_counts;            /* [K]   get_readcount_of(i) (written by pg_fill_*)        */
  const uint32_t* allele_offsets;   /* [V+1] CSR over the alleles map of each variant          */
  const uint16_t* allele_ids;       /* [A]   keys of `alleles`, ascending within a variant     */
  const uint8_t* allele_undefined;  /* [A]   AlleleInfo::is_undefined                          */
  const uint16_t* allele_kmer_offset; /* [A] KmerPath::offset                                  */
  const uint32_t* allele_kmer_mask; /* [A]   KmerPath::kmers: bit b set <=> k-mer offset+b is on the
                                             allele; k-mers outside [offset, offset+32) are absent
                                             (src/kmerpath.cpp:33-48)                           */
  /* Only needed by pg_fill_* / pg_genotype_run (the content of <prefix>_<chrom>_kmers.tsv.gz,
   * src/stepwiseuniquekmercomputer.cpp:105, 2-bit packed): */
  const uint64_t* kmer_codes;       /* [K]   unique k-mers, same CSR as kmer_counts; may be NULL */
  const uint32_t* flank_offsets;    /* [V+1] CSR over flanking k-mers; may be NULL              */
  const uint64_t* flank_codes;      /* [F]   flanking k-mers                                    */
} pg_panel;

/* ------------------------------------------------------------------------------------------------
 * Index artefacts of `PanGenie-index` — what `PanGenie -f <prefix>` loads before the hot path
 * starts: `<prefix>_UniqueKmersMap.cereal` (cereal binary archive of UniqueKmersMap, src/commands.hpp:11-28, read at
 * src/commands.cpp:772-778), `<prefix>_<chrom>_kmers.tsv.gz` (src/kmerparser.cpp:16-28, read at src/commands.cpp:98-137)
 * and the path of `<prefix>_path_segments.fasta` (:764).  Host code, no device needed.  Chromosomes come in the archive's
 * std::map order, the order the `-f` stage processes them.  The panel arrays are owned by the index (valid until
 * pg_index_close); `coverage` / `kmer_counts` hold what the archive holds and are overwritten by pg_fill_counts /
 * pg_genotype_run.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_index pg_index;
/** with_kmers != 0: also read the k-mer table of every chromosome (needed by the fill step). NULL on error. */
pg_index* pg_index_open(const char* prefix, int with_kmers);
/** One archive file only (e.g. a serialised result of the `-f` stage). NULL on error. */
pg_index* pg_index_open_archive(const char* archive_path);
void pg_index_close(pg_index* ix);
uint32_t pg_index_kmer_size(const pg_index* ix);
uint32_t pg_index_n_chromosomes(const pg_index* ix);
const char* pg_index_chromosome_name(const pg_index* ix, uint32_t i);
int pg_index_add_reference(const pg_index* ix);
const char* pg_index_segments_path(const pg_index* ix);
int pg_index_panel(pg_index* ix, uint32_t i, pg_panel* out);

/* ------------------------------------------------------------------------------------------------
 * ProbabilityTable (src/probabilitytable.hpp:12-30).  Dense natural-log probabilities for
 * cov in [cov_min, cov_max) x count in [0, count_max); entries outside are computed on the fly with
 * the reference's formulas (src/probabilitytable.cpp:47-65,75-85, src/copynumber.cpp:14-41).
 * log_p[(count*(cov_max-cov_min) + (cov-cov_min))*3 + cn] = ln P(CN = cn ; cov, count), -inf for 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint16_t cov_min, cov_max, count_max;
  double regularization;
  double* log_p;
} pg_probtable;

/** `ProbabilityTable(cov_min, cov_max, count_max, regularization)` (src/probabilitytable.cpp:28-45);
 *  evaluated in long double on the host, stored as ln in fp64. Allocates t->log_p. */
int pg_probtable_init(pg_probtable* t, uint16_t cov_min, uint16_t cov_max, uint16_t count_max,
                      double regularization);
/** `modify_probability(cov, count, CopyNumber(p0,p1,p2))` (src/probabilitytable.cpp:67-73). */
int pg_probtable_modify(pg_probtable* t, uint16_t cov, uint16_t count, double p0, double p1, double p2);
/** `get_probability(cov,count).get_probability_of(cn)` as a linear probability (host, long double
 *  internally) — for tests. */
double pg_probtable_get(const pg_probtable* t, uint16_t cov, uint16_t count, int cn);
void pg_probtable_free(pg_probtable* t);

/* ------------------------------------------------------------------------------------------------
 * Seam 2: HMM (src/hmm.hpp:38) — forward-backward genotyping of whole chromosomes.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double recombrate;          /* default 1.26   (src/pangenie-genotype.cpp:42)                   */
  double effective_N;         /* default 1e-5   (src/pangenie-genotype.cpp:33); hmm.hpp default 25000 */
  int uniform;                /* uniform transition probabilities (src/transitionprobabilitycomputer.cpp:34-38) */
  int normalize;              /* 1: per-variant GenotypingResult::normalize() (hmm.cpp:41-45);
                                 0: the reference's un-normalised alpha*beta*forward_norm scale    */
  const uint16_t* only_paths; /* NULL = all paths; else path ids to use (hmm.hpp:36)             */
  uint32_t n_only_paths;
} pg_hmm_params;

/** Output for one chromosome; all arrays caller-allocated.  `gl_offsets` is an INPUT describing the
 *  layout (fill it with pg_result_layout).  likelihoods[gl_offsets[v] + a2*(a2+1)/2 + a1], a1 <= a2, is
 *  `get_genotype_likelihood(a1,a2)` (src/genotypingresult.cpp:39-46) of variant v in VCF order
 *  (:48-67); variants that are not HMM columns (src/columnindexer.cpp:24-31) keep all-zero rows. */
typedef struct {
  const uint64_t* gl_offsets; /* [V+1] */
  double* likelihoods;        /* [gl_offsets[V]] */
  uint8_t* is_column;         /* [V] 1 if the variant was an HMM column                           */
  int16_t* genotype;          /* [2V] likeliest genotype of the NORMALISED likelihoods restricted to
                                 defined alleles (src/genotypingresult.cpp:70-96,149-180; what
                                 Graph::write_genotypes prints, src/graph.cpp:206-240); -1,-1 = "./." */
  uint32_t* quality;          /* [V] get_genotype_quality of that genotype (:118-137); 0 if "./."  */
  uint16_t* unique_kmers;     /* [V] nr_unique_kmers() echo (hmm.cpp:106-108)                     */
  uint16_t* coverage;         /* [V] coverage() echo (hmm.cpp:109)                                */
} pg_hmm_result;

/** Fills offsets[0..V] with the VCF-ordered likelihood layout: nr_alleles(v) = max allele id + 1,
 *  row length nr_alleles*(nr_alleles+1)/2. Pure host helper. */
int pg_result_layout(const pg_panel* panel, uint64_t* offsets);

typedef struct pg_engine pg_engine;
/** Per-device engine: owns streams, scratch HBM and compiled kernel configuration. */
pg_engine* pg_engine_create(int device);
void pg_engine_destroy(pg_engine* e);

/**
 * Replaces `HMM(unique_kmers, probabilities, true, false, recombrate, uniform, effective_N, only_paths,
 * normalize)` + `move_genotyping_result()` (src/hmm.hpp:38,46; called from run_genotyping,
 * src/commands.cpp:155-185) for `n_chrom` chromosomes at once: one forward and one backward chain per
 * chromosome run concurrently on the device.  Panels must carry kmer_counts and coverage.
 */
int pg_hmm_run(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
               const pg_hmm_params* params, pg_hmm_result* results);

/**
 * Genotyping over several subsets of paths — the `-a` mode of the reference (src/commands.cpp:916-993): every subset is
 * run like `HMM(..., normalize = false, only_paths = subset)` (run_genotyping, :155-160), the likelihoods are added per
 * variant (`GenotypingResult::combine`, :166-176) and normalised once at the end (:982-988).  subset k uses the path ids
 * subset_paths[subset_offsets[k] .. subset_offsets[k+1]).  The partition of the paths itself (PathSampler) stays with the
 * caller.  Results as for pg_hmm_run with normalize = 1; `is_column` is the union over the subsets.
 */
int pg_hmm_run_subsets(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_probtable* table,
                       const pg_hmm_params* params, uint32_t n_subsets, const uint32_t* subset_offsets,
                       const uint16_t* subset_paths, pg_hmm_result* results);

/** Multi-sample batching (reference README.md:128, 217: one index, many samples).  `n_samples` samples that share
 *  the panel STRUCTURE of `panels` (n_chrom chromosomes; one host copy) are genotyped by one forward-backward launch set: the
 *  forward and backward chains of all n_samples x n_chrom (sample, chromosome) pairs walk concurrently and the block kernel pulls
 *  the jobs of all of them.  Sample s brings kmer_counts[s*n_chrom + c] (same CSR as panels[c].kmer_counts) and
 *  coverage[s*n_chrom + c] ([V_c]) as pg_fill_counts wrote them, and its own ProbabilityTable tables[s] (its own k-mer coverage
 *  peak, src/commands.cpp:846).  results[s*n_chrom + c] as for pg_hmm_run.  Identical to n_samples separate pg_hmm_run calls. */
int pg_hmm_run_samples(pg_engine* e, uint32_t n_samples, uint32_t n_chrom, const pg_panel* panels,
                       const uint16_t* const* kmer_counts, const uint16_t* const* coverage,
                       const pg_probtable* const* tables, const pg_hmm_params* params, pg_hmm_result* results);

/** Emission tables only — `EmissionProbabilityComputer` (src/emissionprobabilitycomputer.cpp:9-34).
 *  emissions: for variant v a dense (maxA+1)x(maxA+1) row-major matrix at em_offsets[v] (maxA = largest
 *  allele id), entries for allele pairs not in the variant's allele map are 0; values are scaled by
 *  exp(-log_scale[v]) (log_scale = ln of the largest entry; all-zero variants report 1.0 everywhere
 *  with log_scale 0, as get_emission_probability does, :31-34). */
int pg_emission_run(pg_engine* e, const pg_panel* panel, const pg_probtable* table,
                    const uint64_t* em_offsets, double* emissions, double* log_scale);

/**
 * `fill_read_kmercounts` (src/commands.cpp:76-152) without the HaplotypeSampler step: looks up every
 * unique and flanking k-mer of the panel in the counter, writes kmer_counts (truncated to u16 like
 * update_readcount(i, count), :130) and coverage (compute_local_coverage, src/kmerparser.cpp:30-49).
 */
int pg_fill_counts(pg_engine* e, const pg_counter* c, uint64_t kmer_abundance_peak, uint32_t n_chrom,
                   pg_panel* panels);

/**
 * The `PanGenie -f` stage end to end (src/commands.cpp:730-1084 minus archive I/O and VCF text):
 * count (PRIME segments, UPDATE reads) -> histogram peak -> ProbabilityTable(peak/4, peak*4, 2*peak,
 * regularization) -> fill -> emission + forward-backward.  Host buffers in, host buffers out.
 */
typedef struct {
  const char* reads; uint64_t reads_len;          /* FASTA/FASTQ file content                     */
  const char* segments; uint64_t segments_len;    /* <prefix>_path_segments.fasta content or NULL  */
  uint32_t k;                                     /* k-mer size (UniqueKmersMap::kmersize)         */
  uint64_t hash_size;                             /* -e                                            */
  double regularization;                          /* default 0.01 (src/pangenie-genotype.cpp:36)   */
  const char* histogram_path;                     /* <out>_histogram.histo or NULL                 */
} pg_genotype_input;

int pg_genotype_run(pg_engine* e, const pg_genotype_input* in, uint32_t n_chrom, pg_panel* panels,
                    const pg_hmm_params* params, pg_hmm_result* results, uint64_t* kmer_abundance_peak);

/* ---- resident form: inputs already in HBM (what bench.py times as `value`) ----------------------------
 * pg_engine_load uploads the panels (with k-mer codes) once; pg_engine_run_resident executes the whole
 * `PanGenie -f` stage from DEVICE-resident read / segment text without any host<->device bulk copy;
 * pg_engine_fetch downloads counts, coverage and results into the caller's buffers. */
int pg_engine_load(pg_engine* e, uint32_t n_chrom, const pg_panel* panels, const pg_hmm_result* layouts);
int pg_engine_run_resident(pg_engine* e, const char* d_reads, uint64_t reads_len, const char* d_segments,
                           uint64_t segments_len, uint32_t k, uint64_t hash_size, double regularization,
                           const pg_hmm_params* params, uint64_t* kmer_abundance_peak);
int pg_engine_fetch(pg_engine* e, uint32_t n_chrom, pg_panel* panels, pg_hmm_result* results);
/** Everything after counting (histogram peak -> ProbabilityTable -> fill -> emission + forward-backward) on the
 *  loaded panels, against a counter the caller filled — e.g. one whose count array was all-reduced over the GPUs
 *  that each counted a shard of the reads.  `largest_peak` as in computeHistogram. */
int pg_engine_run_counted(pg_engine* e, const pg_counter* c, int largest_peak, double regularization,
                          const pg_hmm_params* params, uint64_t* kmer_abundance_peak);

/* ------------------------------------------------------------------------------------------------
 * HaplotypeSampler - `HaplotypeSampler(&unique_kmers, size, recombrate, effective_N, &best_scores,
 * add_reference, "", chromosome, allele_penalty)` (src/haplotypesampler.cpp:20-78), what fill_read_kmercounts runs on every
 * chromosome when the panel has more than 100 paths (src/commands.cpp:139-146, 800-803): `size` integer Viterbi passes over the
 * paths (costs: SamplingEmissions, src/samplingemissions.cpp:9-44; SamplingTransitions, src/samplingtransitions.cpp:5-22),
 * each masking the cells and penalising the alleles of the passes before it, then the panel restricted to the sampled paths
 * (update_unique_kmers, :289-303).  The panel must carry the filled kmer_counts.  n_out = size + (add_reference ? 1 : 0).
 *   sampled_paths       [n_out][V]  path id chosen by pass i at every variant (get_sampled_paths(); the last row is the
 *                                   reference path 0 if add_reference)
 *   best_scores         [size]      Viterbi score of every pass
 *   new_path_to_allele  [V][n_out]  alleles of the sampled paths = path_to_allele of the sampled panel
 *   new_kmer_count      [V]         unique k-mers that still lie on a remaining allele; new_counts: their read counts,
 *                                   concatenated in variant order
 * Integer arithmetic throughout: identical to the reference, ties included.  Up to 1024 paths.
 * ---------------------------------------------------------------------------------------------- */
int pg_haplotype_sample(int device, const pg_panel* panel, uint32_t size, double recombrate, double effective_N,
                        int add_reference, uint16_t allele_penalty, uint64_t* sampled_paths, uint32_t* best_scores,
                        uint16_t* new_path_to_allele, uint32_t* new_kmer_count, uint16_t* new_counts);

/** Empties a counter (all keys removed, counts zero) so its HBM allocation can be reused. */
int pg_count_clear(pg_counter* c);

/* ------------------------------------------------------------------------------------------------
 * Index stage: unique-k-mer selection, `StepwiseUniqueKmerComputer::compute_unique_kmers`
 * (src/stepwiseuniquekmercomputer.cpp:95-197) with `select_kmers` (:46-93), `stepwise_unique_kmers` (:11-34) and
 * `determine_unique_flanking_kmers` (:227-264), run by `PanGenie-index` once per chromosome (src/commands.cpp:647-700)
 * against the k-mer counts of `<prefix>_path_segments.fasta` (COUNT mode: `pg_count_create(segments, NULL, ...)`).
 *
 * Input = what the computer reads from the reference's `Graph` for ONE chromosome, flattened (all arrays host memory,
 * caller-owned): per variant bubble v its alleles 0..n_v-1 (`allele_offsets`), per allele `Variant::get_allele_sequence(a)`
 * (flanks of k-1 bases included, ASCII, any character outside ACGT is an undefined base) and
 * `Variant::is_undefined_allele(a)`, per path `Variant::get_allele_on_path(p)`, and `Graph::get_left_overhang(v, 2k)` /
 * `get_right_overhang(v, 2k)` (src/graph.cpp:554-592).
 *
 * Result = the chromosome's `std::vector<std::shared_ptr<UniqueKmers>>` as a `pg_panel` (coverage and counts zero, as the
 * index stage leaves them; `kmer_codes` / `flank_codes` = the k-mers of `<prefix>_<chrom>_kmers.tsv.gz` as they occur on the
 * allele, i.e. NOT canonicalised, 2-bit packed with the first base most significant), owned by the handle.
 * Per variant: every k-mer of a defined allele that occurs exactly once in that allele, on exactly one allele of the bubble
 * and nowhere else in the graph, and whose allele is carried by a path, in ascending k-mer order per allele; alleles take
 * turns (ascending allele id) until every allele has 16 (all paths on alleles 0/1) or 32 k-mers or max(301, P) k-mers are
 * selected; k-mer i of the variant is the i-th in (allele id, k-mer) order.  Flanks: up to 12 k-mers per side, ascending, that
 * occur once in the overhang and once in the graph.  One CTA per variant: enumeration, a bitonic sort in shared memory (global
 * scratch for bubbles with more than 2048 k-mers), probes of the graph table, selection.  Integer work: identical to the
 * reference.  Limits: k <= 32, at most 65535 alleles per bubble.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t n_variants;              /* V */
  uint32_t n_paths;                 /* P */
  uint32_t k;
  const uint64_t* positions;        /* [V]   Variant::get_start_position()                       */
  const uint64_t* end_positions;    /* [V]   Variant::get_end_position(); only written to the tsv; may be NULL */
  const uint16_t* path_to_allele;   /* [V*P] Variant::get_allele_on_path(p)                      */
  const uint32_t* allele_offsets;   /* [V+1] CSR over the alleles of each bubble                 */
  const uint8_t* allele_undefined;  /* [A]   Variant::is_undefined_allele(a)                     */
  const uint64_t* seq_offsets;      /* [A+1] CSR over `seq`                                      */
  const char* seq;                  /*       Variant::get_allele_sequence(a), concatenated      */
  const uint64_t* left_offsets;     /* [V+1] CSR over `left_seq`                                 */
  const char* left_seq;             /*       Graph::get_left_overhang(v, 2k)                    */
  const uint64_t* right_offsets;    /* [V+1] */
  const char* right_seq;            /*       Graph::get_right_overhang(v, 2k)                   */
} pg_variants;

typedef struct pg_unique_kmers pg_unique_kmers;
/** NULL on error (pg_last_error()).  `graph_counts` must hold the COUNT of the path-segment file. */
pg_unique_kmers* pg_unique_kmers_compute(int device, const pg_counter* graph_counts, const pg_variants* in);
/** The panel of the chromosome; arrays stay valid until pg_unique_kmers_free. */
int pg_unique_kmers_panel(pg_unique_kmers* u, pg_panel* out);
/** Device time of the selection kernel (ms) and the k-mers it enumerated. */
int pg_unique_kmers_stats(const pg_unique_kmers* u, double* kernel_ms, uint64_t* kmers_enumerated);
/** Writes `<prefix>_<chrom>_kmers.tsv.gz` (src/stepwiseuniquekmercomputer.cpp:104-105, 151-182; read back by
 *  src/kmerparser.cpp:16-28): header + one line per variant `chrom start end k-mers flank-k-mers`, "nan" for none. */
int pg_unique_kmers_write_tsv(const pg_unique_kmers* u, const char* chromosome, const uint64_t* end_positions,
                              const char* path);
void pg_unique_kmers_free(pg_unique_kmers* u);

/* ---- measurement hooks (bench.py): per-stage device times of the last call on this engine ---- */
typedef struct {
  double count_ms, histogram_ms, fill_ms, emission_ms, hmm_skeleton_ms, hmm_blocks_ms, finalize_ms;
  uint64_t hmm_columns;        /* HMM columns processed                                            */
  uint64_t hmm_block_launches; /* kernel launches of the block forward-backward kernel             */
  uint64_t kernel_launches;    /* all kernels launched by the last call                            */
  uint64_t kmers_counted;      /* k-mers streamed through the UPDATE/COUNT pass                     */
  uint64_t text_bytes;         /* read text bytes streamed                                          */
  double prime_ms;             /* PRIME pass over the segment file                                  */
  uint64_t hmm_scan_used;      /* 1 if the checkpoints came from the parallel-in-time scan path (P <= 9) */
  double count_probe_ms;       /* part of count_ms spent in probe_parts_kernel (partitioned counting)  */
  uint64_t count_probe_passes; /* launches of probe_parts_kernel in the UPDATE/COUNT pass              */
} pg_timings;
int pg_engine_timings(const pg_engine* e, pg_timings* out);

#ifdef __cplusplus
}
#endif
#endif /* PANGENIE_B200_H */
