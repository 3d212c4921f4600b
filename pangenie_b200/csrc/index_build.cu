// Index stage on the device: the unique-k-mer selection of `StepwiseUniqueKmerComputer`
// (reference src/stepwiseuniquekmercomputer.cpp:11-34 stepwise_unique_kmers, :46-93 select_kmers, :95-197
// compute_unique_kmers, :227-264 determine_unique_flanking_kmers), one CTA per variant bubble.
//
// The reference keeps the k-mers of a bubble in a std::map<mer_dna, vector<allele>> (ordered by the 2-bit code, first base
// most significant) and walks it.  Here a CTA
//   1. enumerates the k-mers of every defined allele straight from the ASCII sequences (one thread per start position),
//   2. sorts (k-mer, allele) with a bitonic network in shared memory (bubbles with more than SEL_CAP k-mers: a second launch
//      with a small grid sorting in a global scratch),
//   3. flags in parallel what the map walk keeps: exactly one occurrence inside its allele, no second allele with such an
//      occurrence, graph count == local count (probe of the graph k-mer table, csrc/common.cuh table_lookup), allele carried
//      by a path,
//   4. brings the kept k-mers into (allele, k-mer) order (second bitonic sort) and lets one thread replay the reference's
//      round-robin (`select_kmers`: alleles take turns until 16 / 32 k-mers per allele or max(301, P) in total) - that walk is
//      a few dozen steps for a SNP and writes the k-mers in their final order together with the KmerPath offset / mask of every
//      allele (src/kmerpath.cpp:13-31).
// Flanks: the same enumeration + sort per side, kept = one occurrence in the overhang and graph count 1, first 12 per side.
// Results are written at per-variant capacity offsets; the host compacts them into the pg_panel arrays.
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace pg {

constexpr int SEL_THREADS = 128;
constexpr int SEL_CAP = 2048;        // k-mers of one sort that fit the shared-memory arrays
constexpr uint32_t SEL_INVALID = 0xFFFFu;
constexpr int SEL_COVER_BITS = 8192; // alleles tracked by the shared-memory "carried by a path" bitmap
constexpr int SEL_MAX_FLANK = 12;    // stepwiseuniquekmercomputer.cpp:230

struct SelArgs {
  uint32_t V, P, k;
  const uint32_t* order;            // variants handled by this launch
  uint32_t n_order;
  const uint16_t* p2a;              // [V*P]
  const uint32_t* allele_off;       // [V+1]
  const uint8_t* allele_undef;      // [A]
  const uint64_t* seq_off;          // [A+1]
  const char* seq;
  const uint64_t* slot_off;         // [A+1] CSR of k-mer slots per allele (0 for undefined alleles)
  const uint64_t* left_off;         // [V+1]
  const char* left_seq;
  const uint64_t* right_off;
  const char* right_seq;
  const KmerBucket* tab;            // graph k-mer table
  uint64_t cap;
  uint32_t cap_q, cap_sh;
  // outputs
  const uint64_t* out_off;          // [V+1] capacity CSR of the selected k-mers
  uint64_t* out_codes;
  uint32_t* out_count;              // [V]
  uint64_t* flank_codes;            // [V*24]
  uint32_t* flank_count;            // [V]
  uint8_t* covered;                 // [A] allele carried by a path
  uint16_t* a_koff;                 // [A]
  uint32_t* a_kmask;                // [A]
  uint32_t* err;                    // bit 0: undefined allele that no path carries (the reference throws)
  // global scratch of the BIG launch: per CTA n2_max entries
  unsigned long long* g_key;
  uint32_t* g_al;
  uint64_t n2_max;
};

__device__ __forceinline__ int base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}
// the reference tests the raw character (`!= 'A' && != 'C' ...`, :21) for the window reset, while mer_dna::shift_left also
// accepts lower case; allele and overhang sequences come from DnaSequence::to_string and are upper case
__device__ __forceinline__ bool is_upper_base(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// k-mer counted for slot `e - (k-1)` of a sequence of length len (the one ending at base e), or invalid.
// Regular slots (e < len-1): counted iff the window lies inside the sequence and holds no undefined base (:20-24).
// Last slot (e == len-1, or the only slot of a sequence shorter than k): always counted (:27); when its window is broken the
// reference's rolling k-mer holds the last k bases that were shifted in (undefined characters are not shifted), on top of the
// all-A start value.
__device__ bool kmer_at(const char* s, uint64_t len, uint32_t k, uint64_t slot, uint64_t n_slots, uint64_t& code) {
  const uint64_t mask = kmer_mask(k);
  const bool last = slot + 1 == n_slots;
  if (len >= k) {
    const uint64_t b = slot;  // window [b, b+k)
    uint64_t c = 0;
    bool ok = true;
    for (uint32_t i = 0; i < k; ++i) {
      const char ch = s[b + i];
      ok &= is_upper_base(ch);
      c = (c << 2) | (uint64_t)(base_code(ch) & 3);
    }
    if (ok) { code = c & mask; return true; }
    if (!last) return false;
  }
  // broken last window: replay the shifts from the start (rare: undefined base near the end, or len < k)
  uint64_t c = 0;
  for (uint64_t i = 0; i < len; ++i) {
    const int x = base_code(s[i]);
    if (x >= 0) c = ((c << 2) | (uint64_t)x) & mask;
  }
  code = c;
  return true;
}

struct SortView {
  unsigned long long* key;
  uint32_t* al;
};

__device__ __forceinline__ bool entry_less(unsigned long long ka, uint32_t aa, unsigned long long kb, uint32_t ab, bool by_allele) {
  const bool ia = aa == SEL_INVALID, ib = ab == SEL_INVALID;
  if (ia != ib) return ib;          // invalid entries last
  if (by_allele) return aa != ab ? aa < ab : ka < kb;
  return ka != kb ? ka < kb : aa < ab;
}

__device__ void bitonic_sort(SortView v, uint32_t n2, bool by_allele) {
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < (n2 >> 1); i += blockDim.x) {
        const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const unsigned long long kl = v.key[lo], kh = v.key[hi];
        const uint32_t al = v.al[lo], ah = v.al[hi];
        const bool swap = asc ? entry_less(kh, ah, kl, al, by_allele) : entry_less(kl, al, kh, ah, by_allele);
        if (swap) {
          v.key[lo] = kh; v.key[hi] = kl;
          v.al[lo] = ah; v.al[hi] = al;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ uint32_t pow2_at_least(uint32_t n) {
  uint32_t p = 2;
  while (p < n) p <<= 1;
  return p;
}

// Flank k-mers of one side: enumerate, sort, keep (one occurrence in the overhang, graph count 1), first 12 in k-mer order.
__device__ void select_flank(const SelArgs& a, SortView sv, const char* s, uint64_t len, uint64_t* out, uint32_t& n_out,
                             uint32_t* s_flag) {
  const uint64_t n_slots = len >= a.k ? len - a.k + 1 : 1;
  const uint32_t n2 = pow2_at_least((uint32_t)n_slots);
  for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
    uint64_t code = ~0ULL;
    uint32_t al = SEL_INVALID;
    if (i < n_slots && kmer_at(s, len, a.k, i, n_slots, code)) al = 0;
    sv.key[i] = code;
    sv.al[i] = al;
  }
  __syncthreads();
  bitonic_sort(sv, n2, false);
  // keep flags in place of the allele field (1 = kept)
  for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
    uint32_t keep = 0;
    if (sv.al[i] != SEL_INVALID) {
      const unsigned long long key = sv.key[i];
      const bool dup = (i > 0 && sv.al[i - 1] != SEL_INVALID && sv.key[i - 1] == key) ||
                       (i + 1 < n2 && sv.al[i + 1] != SEL_INVALID && sv.key[i + 1] == key);
      if (!dup && table_lookup(key, a.k, a.tab, a.cap, a.cap_q, a.cap_sh) == 1) keep = 1;
    }
    s_flag[i] = keep;  // (s_flag aliases nothing sv uses; sized like the sort arrays)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < n2 && n < SEL_MAX_FLANK; ++i)
      if (s_flag[i]) out[n++] = sv.key[i];
    n_out = n;
  }
  __syncthreads();
}

template <bool BIG>
__global__ void __launch_bounds__(SEL_THREADS) select_kmers_kernel(const SelArgs a) {
  __shared__ unsigned long long s_key[BIG ? 1 : SEL_CAP];
  __shared__ uint32_t s_al[BIG ? 1 : SEL_CAP];
  __shared__ uint32_t s_flag[BIG ? 1 : SEL_CAP];
  __shared__ uint32_t s_cover[SEL_COVER_BITS / 32];
  __shared__ uint32_t s_n, s_nl, s_nr;
  __shared__ uint32_t s_round[33];
  SortView sv;
  uint32_t* flag;
  if (BIG) {
    sv.key = a.g_key + (size_t)blockIdx.x * a.n2_max;
    sv.al = a.g_al + (size_t)blockIdx.x * 2 * a.n2_max;
    flag = sv.al + a.n2_max;
  } else {
    sv.key = s_key;
    sv.al = s_al;
    flag = s_flag;
  }
  for (uint32_t oi = blockIdx.x; oi < a.n_order; oi += gridDim.x) {
    const uint32_t v = a.order[oi];
    const uint32_t a0 = a.allele_off[v], a1 = a.allele_off[v + 1], nA = a1 - a0;
    const uint16_t* paths = a.p2a + (size_t)v * a.P;
    // ---- alleles carried by a path (Variant::get_paths_of_allele non-empty, :67-69); biallelic test (:118-124) ----
    for (uint32_t i = threadIdx.x; i < SEL_COVER_BITS / 32; i += blockDim.x) s_cover[i] = 0;
    if (threadIdx.x == 0) s_n = 1;  // is_biallelic
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < a.P; p += blockDim.x) {
      const uint32_t al = paths[p];
      if (al < nA) a.covered[a0 + al] = 1;
      if (al < SEL_COVER_BITS) atomicOr(&s_cover[al >> 5], 1u << (al & 31));
      if (al > 1) s_n = 0;
    }
    __syncthreads();
    const bool biallelic = s_n != 0;
    __syncthreads();
    // ---- enumerate the k-mers of the defined alleles (stepwise_unique_kmers per allele, :140-149) ----
    const uint64_t sl0 = a.slot_off[a0];
    const uint32_t N = (uint32_t)(a.slot_off[a1] - sl0);
    const uint32_t n2 = pow2_at_least(N);
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
      uint64_t code = ~0ULL;
      uint32_t al = SEL_INVALID;
      if (i < N) {
        uint32_t lo = 0, hi = nA;  // allele whose slot range holds i
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (a.slot_off[a0 + mid] - sl0 <= i) lo = mid;
          else hi = mid;
        }
        const uint64_t so = a.seq_off[a0 + lo], len = a.seq_off[a0 + lo + 1] - so;
        const uint64_t first = a.slot_off[a0 + lo] - sl0, n_slots = a.slot_off[a0 + lo + 1] - sl0 - first;
        if (kmer_at(a.seq + so, len, a.k, i - first, n_slots, code)) al = lo;
      }
      sv.key[i] = code;
      sv.al[i] = al;
    }
    for (uint32_t al = threadIdx.x; al < nA; al += blockDim.x) {
      if (a.allele_undef[a0 + al]) {  // set_undefined_allele throws for an allele without an entry (multiallelicuniquekmers.cpp:180-186)
        bool carried = false;
        for (uint32_t p = 0; p < a.P && !carried; ++p) carried = paths[p] == al;
        if (!carried) atomicOr(a.err, 1u);
      }
      a.a_koff[a0 + al] = 0;
      a.a_kmask[a0 + al] = 0;
    }
    __syncthreads();
    bitonic_sort(sv, n2, false);
    // ---- what the walk over `occurences` keeps (:52-72) ----
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
      uint32_t keep = 0;
      const uint32_t al = sv.al[i];
      if (al != SEL_INVALID) {
        const unsigned long long key = sv.key[i];
        auto same_pair = [&](uint32_t j) { return sv.al[j] == al && sv.key[j] == key; };
        const bool once_in_allele = !(i > 0 && same_pair(i - 1)) && !(i + 1 < n2 && same_pair(i + 1));
        if (once_in_allele) {
          // local_count = alleles holding the k-mer exactly once: scan the run of equal k-mers
          uint32_t local = 1;
          auto unique_at = [&](uint32_t j) {
            const uint32_t aj = sv.al[j];
            const bool l = j > 0 && sv.al[j - 1] == aj && sv.key[j - 1] == key;
            const bool r = j + 1 < n2 && sv.al[j + 1] == aj && sv.key[j + 1] == key;
            return !l && !r;
          };
          for (uint32_t j = i; j-- > 0 && sv.al[j] != SEL_INVALID && sv.key[j] == key && local < 2;) local += unique_at(j) ? 1 : 0;
          for (uint32_t j = i + 1; j < n2 && sv.al[j] != SEL_INVALID && sv.key[j] == key && local < 2; ++j) local += unique_at(j) ? 1 : 0;
          if (local == 1) {
            bool carried;
            if (al < SEL_COVER_BITS) carried = (s_cover[al >> 5] >> (al & 31)) & 1u;
            else {
              carried = false;
              for (uint32_t p = 0; p < a.P && !carried; ++p) carried = paths[p] == al;
            }
            if (carried && table_lookup(key, a.k, a.tab, a.cap, a.cap_q, a.cap_sh) == 1) keep = 1;  // genomic_count - local_count == 0
          }
        }
      }
      flag[i] = keep;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x)
      if (!flag[i]) sv.al[i] = SEL_INVALID;
    __syncthreads();
    bitonic_sort(sv, n2, true);  // kept k-mers first, by (allele, k-mer) = the per-allele queues in allele order
    // ---- select_kmers round-robin (:74-92) + insert_kmer order (:154-165), one thread ----
    if (threadIdx.x == 0) {
      uint32_t max_total = a.P < 301 ? 301 : a.P;  // (unsigned short in the reference; P < 65535)
      const uint32_t max_kmers = biallelic ? 16 : 32;
      for (int r = 0; r <= 32; ++r) s_round[r] = 0;
      uint32_t M = 0;
      {  // candidates per round: the j-th k-mer of an allele's queue is taken in round j
        uint32_t cur = SEL_INVALID, j = 0;
        for (; M < n2 && sv.al[M] != SEL_INVALID; ++M) {
          if (sv.al[M] != cur) { cur = sv.al[M]; j = 0; }
          if (j < max_kmers) s_round[j] += 1;
          ++j;
        }
      }
      uint32_t full_rounds = 0, taken = 0;  // rounds taken completely; then `extra` k-mers of round full_rounds in allele order
      while (full_rounds < max_kmers && s_round[full_rounds] > 0 && taken + s_round[full_rounds] <= max_total) taken += s_round[full_rounds++];
      uint32_t extra = (full_rounds < max_kmers && taken < max_total) ? max_total - taken : 0;
      if (full_rounds < max_kmers && s_round[full_rounds] < extra) extra = s_round[full_rounds];
      uint64_t* out = a.out_codes + a.out_off[v];
      uint32_t n = 0, cur = SEL_INVALID, j = 0;
      for (uint32_t i = 0; i < M; ++i) {
        const uint32_t al = sv.al[i];
        if (al != cur) { cur = al; j = 0; }
        bool take = j < full_rounds;
        if (!take && j == full_rounds && extra > 0) { take = true; --extra; }
        if (take) {
          if (a.a_kmask[a0 + al] == 0) a.a_koff[a0 + al] = (uint16_t)n;   // KmerPath::set_position (kmerpath.cpp:13-31)
          a.a_kmask[a0 + al] |= 1u << (n - a.a_koff[a0 + al]);
          out[n++] = sv.key[i];
        }
        ++j;
      }
      a.out_count[v] = n;
    }
    __syncthreads();
    // ---- flanking k-mers (determine_unique_flanking_kmers, :227-264) ----
    select_flank(a, sv, a.left_seq + a.left_off[v], a.left_off[v + 1] - a.left_off[v], a.flank_codes + (size_t)v * 24, s_nl, flag);
    select_flank(a, sv, a.right_seq + a.right_off[v], a.right_off[v + 1] - a.right_off[v], a.flank_codes + (size_t)v * 24 + s_nl, s_nr, flag);
    if (threadIdx.x == 0) a.flank_count[v] = s_nl + s_nr;
    __syncthreads();
  }
}

}  // namespace pg

using namespace pg;

struct pg_unique_kmers {
  uint32_t V = 0, P = 0, k = 0;
  std::vector<uint64_t> positions, kmer_codes, flank_codes;
  std::vector<uint16_t> p2a, coverage, kmer_counts, allele_ids, allele_koff;
  std::vector<uint32_t> kmer_offsets, allele_offsets, allele_kmask, flank_offsets;
  std::vector<uint8_t> allele_undef;
  double kernel_ms = 0.0;
  uint64_t kmers_enumerated = 0;
};

namespace {

template <class T>
int upload(DevBuf<T>& d, const T* h, size_t n, cudaStream_t s) {
  PG_TRY(d.reserve(n ? n : 1));
  if (n) PG_CUDA(cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  return PG_OK;
}

int compute_impl(int device, const pg_counter* gc, const pg_variants* in, pg_unique_kmers* u) {
  NvtxRange nvtx_("pg_unique_kmers_compute");
  if (!gc || !in) return fail(PG_ERR_ARG, "null argument");
  if (in->k < 1 || in->k > 32 || in->k != gc->k) return fail(PG_ERR_ARG, "k must be in [1,32] and equal to the counter's k");
  if (in->n_paths >= 65535) return fail(PG_ERR_ARG, "number of paths exceeds 65534 (src/stepwiseuniquekmercomputer.cpp:120)");
  if (gc->device != device) return fail(PG_ERR_ARG, "graph k-mer counter lives on another device");
  PG_TRY(check_device(device));
  DeviceGuard guard(device);
  const uint32_t V = in->n_variants, P = in->n_paths, k = in->k;
  const uint64_t A = V ? in->allele_offsets[V] : 0;
  u->V = V; u->P = P; u->k = k;
  u->positions.assign(in->positions, in->positions + V);
  u->p2a.assign(in->path_to_allele, in->path_to_allele + (size_t)V * P);
  u->coverage.assign(V, 0);
  u->kmer_offsets.assign(V + 1, 0); u->allele_offsets.assign(V + 1, 0); u->flank_offsets.assign(V + 1, 0);
  if (V == 0) return PG_OK;
  // host: k-mer slots per allele, capacity of the selection per variant, split into shared-memory / global-scratch variants
  std::vector<uint64_t> slot_off(A + 1, 0), out_off(V + 1, 0);
  std::vector<uint32_t> small, big;
  uint64_t n2_max = 2;
  for (uint32_t v = 0; v < V; ++v) {
    const uint32_t a0 = in->allele_offsets[v], a1 = in->allele_offsets[v + 1];
    if (a1 < a0 || a1 - a0 > 65535) return fail(PG_ERR_ARG, "a bubble has more than 65535 alleles or the allele CSR is not ascending");
    for (uint32_t a = a0; a < a1; ++a) {
      const uint64_t len = in->seq_offsets[a + 1] - in->seq_offsets[a];
      slot_off[a + 1] = slot_off[a] + (in->allele_undefined[a] ? 0 : (len >= k ? len - k + 1 : 1));
    }
    uint64_t N = slot_off[a1] - slot_off[a0];
    for (int side = 0; side < 2; ++side) {
      const uint64_t* off = side ? in->right_offsets : in->left_offsets;
      const uint64_t len = off[v + 1] - off[v];
      N = std::max<uint64_t>(N, len >= k ? len - k + 1 : 1);
    }
    if (N > (1ull << 26)) return fail(PG_ERR_ARG, "a bubble has more than 2^26 k-mers");
    uint64_t n2 = 2;
    while (n2 < N) n2 <<= 1;
    if (n2 <= SEL_CAP) small.push_back(v);
    else { big.push_back(v); n2_max = std::max(n2_max, n2); }
    const uint64_t cap_total = std::max<uint32_t>(P, 301);
    out_off[v + 1] = out_off[v] + std::min<uint64_t>(std::min<uint64_t>(slot_off[a1] - slot_off[a0], cap_total), 32ull * (a1 - a0));
  }
  u->kmers_enumerated = slot_off[A];
  cudaStream_t s = gc->stream;
  DevBuf<uint32_t> d_order, d_aoff, d_out_count, d_flank_count, d_kmask, d_err, d_gal;
  DevBuf<uint16_t> d_p2a, d_koff;
  DevBuf<uint8_t> d_undef, d_cov;
  DevBuf<uint64_t> d_seq_off, d_slot_off, d_left_off, d_right_off, d_out_off, d_out_codes, d_flank_codes;
  DevBuf<unsigned long long> d_gkey;
  DevBuf<char> d_seq, d_left, d_right;
  std::vector<uint32_t> order(small);
  order.insert(order.end(), big.begin(), big.end());
  PG_TRY(upload(d_order, order.data(), order.size(), s));
  PG_TRY(upload(d_p2a, in->path_to_allele, (size_t)V * P, s));
  PG_TRY(upload(d_aoff, in->allele_offsets, (size_t)V + 1, s));
  PG_TRY(upload(d_undef, in->allele_undefined, A, s));
  PG_TRY(upload(d_seq_off, in->seq_offsets, A + 1, s));
  PG_TRY(upload(d_seq, in->seq, in->seq_offsets[A], s));
  PG_TRY(upload(d_slot_off, slot_off.data(), A + 1, s));
  PG_TRY(upload(d_left_off, in->left_offsets, (size_t)V + 1, s));
  PG_TRY(upload(d_left, in->left_seq, in->left_offsets[V], s));
  PG_TRY(upload(d_right_off, in->right_offsets, (size_t)V + 1, s));
  PG_TRY(upload(d_right, in->right_seq, in->right_offsets[V], s));
  PG_TRY(upload(d_out_off, out_off.data(), (size_t)V + 1, s));
  PG_TRY(d_out_codes.reserve(out_off[V] ? out_off[V] : 1));
  PG_TRY(d_out_count.reserve(V)); PG_TRY(d_flank_count.reserve(V)); PG_TRY(d_flank_codes.reserve((size_t)V * 24));
  PG_TRY(d_cov.reserve(A ? A : 1)); PG_TRY(d_koff.reserve(A ? A : 1)); PG_TRY(d_kmask.reserve(A ? A : 1)); PG_TRY(d_err.reserve(1));
  PG_CUDA(cudaMemsetAsync(d_cov.p, 0, A ? A : 1, s));
  PG_CUDA(cudaMemsetAsync(d_err.p, 0, 4, s));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  SelArgs a = {};
  a.V = V; a.P = P; a.k = k;
  a.p2a = d_p2a.p; a.allele_off = d_aoff.p; a.allele_undef = d_undef.p; a.seq_off = d_seq_off.p; a.seq = d_seq.p;
  a.slot_off = d_slot_off.p; a.left_off = d_left_off.p; a.left_seq = d_left.p; a.right_off = d_right_off.p; a.right_seq = d_right.p;
  a.tab = gc->slots; a.cap = gc->capacity; a.cap_q = gc->cap_q; a.cap_sh = gc->cap_sh;
  a.out_off = d_out_off.p; a.out_codes = d_out_codes.p; a.out_count = d_out_count.p; a.flank_codes = d_flank_codes.p;
  a.flank_count = d_flank_count.p; a.covered = d_cov.p; a.a_koff = d_koff.p; a.a_kmask = d_kmask.p; a.err = d_err.p;
  cudaEvent_t e0, e1;
  PG_CUDA(cudaEventCreate(&e0)); PG_CUDA(cudaEventCreate(&e1));
  cudaEventRecord(e0, s);
  if (!small.empty()) {
    a.order = d_order.p; a.n_order = (uint32_t)small.size();
    const int grid = (int)std::min<uint64_t>(small.size(), (uint64_t)sms * 8);
    select_kmers_kernel<false><<<grid, SEL_THREADS, 0, s>>>(a);
    count_launch();
  }
  if (!big.empty()) {
    const int grid = (int)std::min<uint64_t>(big.size(), (uint64_t)sms);
    PG_TRY(d_gkey.reserve((size_t)grid * n2_max));
    PG_TRY(d_gal.reserve((size_t)grid * 2 * n2_max));
    a.order = d_order.p + small.size(); a.n_order = (uint32_t)big.size();
    a.g_key = d_gkey.p; a.g_al = d_gal.p; a.n2_max = n2_max;
    select_kmers_kernel<true><<<grid, SEL_THREADS, 0, s>>>(a);
    count_launch();
  }
  cudaEventRecord(e1, s);
  PG_CUDA(cudaGetLastError());
  // results -> host, compacted into the panel arrays
  std::vector<uint64_t> h_codes(out_off[V]), h_flank((size_t)V * 24);
  std::vector<uint32_t> h_count(V), h_fcount(V), h_kmask(A);
  std::vector<uint16_t> h_koff(A);
  std::vector<uint8_t> h_cov(A);
  uint32_t h_err = 0;
  if (out_off[V]) PG_CUDA(cudaMemcpyAsync(h_codes.data(), d_out_codes.p, out_off[V] * 8, cudaMemcpyDeviceToHost, s));
  PG_CUDA(cudaMemcpyAsync(h_flank.data(), d_flank_codes.p, (size_t)V * 24 * 8, cudaMemcpyDeviceToHost, s));
  PG_CUDA(cudaMemcpyAsync(h_count.data(), d_out_count.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
  PG_CUDA(cudaMemcpyAsync(h_fcount.data(), d_flank_count.p, (size_t)V * 4, cudaMemcpyDeviceToHost, s));
  if (A) {
    PG_CUDA(cudaMemcpyAsync(h_kmask.data(), d_kmask.p, A * 4, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaMemcpyAsync(h_koff.data(), d_koff.p, A * 2, cudaMemcpyDeviceToHost, s));
    PG_CUDA(cudaMemcpyAsync(h_cov.data(), d_cov.p, A, cudaMemcpyDeviceToHost, s));
  }
  PG_CUDA(cudaMemcpyAsync(&h_err, d_err.p, 4, cudaMemcpyDeviceToHost, s));
  PG_CUDA(cudaStreamSynchronize(s));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  u->kernel_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (h_err & 1u) return fail(PG_ERR_ARG, "an undefined allele is not carried by any path (the reference's set_undefined_allele throws: allele_id does not exist)");
  uint64_t K = 0, F = 0, NA = 0;
  for (uint32_t v = 0; v < V; ++v) {
    K += h_count[v]; F += h_fcount[v];
    for (uint32_t al = in->allele_offsets[v]; al < in->allele_offsets[v + 1]; ++al) NA += h_cov[al];
  }
  u->kmer_codes.reserve(K); u->flank_codes.reserve(F);
  u->allele_ids.reserve(NA); u->allele_undef.reserve(NA); u->allele_koff.reserve(NA); u->allele_kmask.reserve(NA);
  for (uint32_t v = 0; v < V; ++v) {
    u->kmer_codes.insert(u->kmer_codes.end(), h_codes.begin() + out_off[v], h_codes.begin() + out_off[v] + h_count[v]);
    u->flank_codes.insert(u->flank_codes.end(), h_flank.begin() + (size_t)v * 24, h_flank.begin() + (size_t)v * 24 + h_fcount[v]);
    const uint32_t a0 = in->allele_offsets[v];
    for (uint32_t al = a0; al < in->allele_offsets[v + 1]; ++al) {
      if (!h_cov[al]) continue;  // UniqueKmers::alleles holds the alleles some path carries
      u->allele_ids.push_back((uint16_t)(al - a0));
      u->allele_undef.push_back(in->allele_undefined[al] ? 1 : 0);
      u->allele_koff.push_back(h_koff[al]);
      u->allele_kmask.push_back(h_kmask[al]);
    }
    u->kmer_offsets[v + 1] = (uint32_t)u->kmer_codes.size();
    u->flank_offsets[v + 1] = (uint32_t)u->flank_codes.size();
    u->allele_offsets[v + 1] = (uint32_t)u->allele_ids.size();
  }
  u->kmer_counts.assign(u->kmer_codes.size(), 0);
  return PG_OK;
}

}  // namespace

extern "C" pg_unique_kmers* pg_unique_kmers_compute(int device, const pg_counter* graph_counts, const pg_variants* in) {
  clear_error();
  pg_unique_kmers* u = new pg_unique_kmers;
  if (compute_impl(device, graph_counts, in, u) != PG_OK) {
    delete u;
    return nullptr;
  }
  return u;
}

extern "C" int pg_unique_kmers_panel(pg_unique_kmers* u, pg_panel* out) {
  clear_error();
  if (!u || !out) return fail(PG_ERR_ARG, "null argument");
  memset(out, 0, sizeof(*out));
  out->n_variants = u->V; out->n_paths = u->P;
  out->positions = u->positions.data(); out->path_to_allele = u->p2a.data(); out->coverage = u->coverage.data();
  out->kmer_offsets = u->kmer_offsets.data(); out->kmer_counts = u->kmer_counts.data();
  out->allele_offsets = u->allele_offsets.data(); out->allele_ids = u->allele_ids.data();
  out->allele_undefined = u->allele_undef.data(); out->allele_kmer_offset = u->allele_koff.data();
  out->allele_kmer_mask = u->allele_kmask.data(); out->kmer_codes = u->kmer_codes.data();
  out->flank_offsets = u->flank_offsets.data(); out->flank_codes = u->flank_codes.data();
  return PG_OK;
}

extern "C" int pg_unique_kmers_stats(const pg_unique_kmers* u, double* kernel_ms, uint64_t* kmers_enumerated) {
  if (!u) return fail(PG_ERR_ARG, "null argument");
  if (kernel_ms) *kernel_ms = u->kernel_ms;
  if (kmers_enumerated) *kmers_enumerated = u->kmers_enumerated;
  return PG_OK;
}

extern "C" int pg_unique_kmers_write_tsv(const pg_unique_kmers* u, const char* chromosome, const uint64_t* end_positions,
                                         const char* path) {
  clear_error();
  if (!u || !chromosome || !end_positions || !path) return fail(PG_ERR_ARG, "null argument");
  gzFile f = gzopen(path, "wb");
  if (!f) return fail(PG_ERR_IO, std::string("File ") + path + " cannot be created. Note that the filename must not contain non-existing directories.");
  auto put = [&](const std::string& s) { return gzwrite(f, s.data(), (unsigned)s.size()) == (int)s.size(); };
  auto kmer_str = [&](uint64_t code) {
    std::string s(u->k, 'A');
    for (uint32_t i = 0; i < u->k; ++i) s[u->k - 1 - i] = "ACGT"[(code >> (2 * i)) & 3];
    return s;
  };
  bool ok = put("#chromosome\tstart\tend\tunique_kmers\tunique_kmers_overhang\n");
  for (uint32_t v = 0; v < u->V && ok; ++v) {
    std::string line = std::string(chromosome) + "\t" + std::to_string(u->positions[v]) + "\t" + std::to_string(end_positions[v]) + "\t";
    for (uint32_t i = u->kmer_offsets[v]; i < u->kmer_offsets[v + 1]; ++i) line += (i > u->kmer_offsets[v] ? "," : "") + kmer_str(u->kmer_codes[i]);
    if (u->kmer_offsets[v] == u->kmer_offsets[v + 1]) line += "nan";
    line += "\t";
    for (uint32_t i = u->flank_offsets[v]; i < u->flank_offsets[v + 1]; ++i) line += (i > u->flank_offsets[v] ? "," : "") + kmer_str(u->flank_codes[i]);
    if (u->flank_offsets[v] == u->flank_offsets[v + 1]) line += "nan";
    line += "\n";
    ok = put(line);
  }
  if (gzclose(f) != Z_OK || !ok) return fail(PG_ERR_IO, std::string("write to ") + path + " failed");
  return PG_OK;
}

extern "C" void pg_unique_kmers_free(pg_unique_kmers* u) { delete u; }
