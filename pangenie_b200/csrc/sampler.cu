// HaplotypeSampler on the device: the integer Viterbi that every run with more than 100 haplotype
// paths executes between the count fill and the HMM (reference src/commands.cpp:800-803, src/haplotypesampler.cpp:20-78,
// 110-303, src/samplingemissions.cpp:9-44, src/samplingtransitions.cpp:5-22).
//
// `size` Viterbi passes over the P paths of one chromosome; every pass masks the (column, path) cells taken by earlier
// passes and penalises the alleles it visited.  A pass is a chain over the columns whose state is P unsigned costs:
//   cost_v[i] = pen_v[allele(i)] + min(cost_{v-1}[i], min_{j != i} cost_{v-1}[j] + switch_v)      (saturating adds)
// so a column needs only the smallest and second smallest free entry of the previous column.  One WARP walks a
// chromosome with the costs in registers (path i = lane + 32 j): no block barriers, the column minima by shuffles, the
// per-column inputs (penalty of every path, mask bits) streamed as coalesced rows two columns ahead.  The backtrace is
// the same walk backwards over the stored back-pointer rows.  Everything is integer: results are identical to the
// reference's, ties included (first / second minimum by (value, path id), first minimal end state).
//
// Host parts: the recombination costs (x87 long double + double exp/log10, samplingtransitions.cpp:5-14) and the 33 x 33
// table of allele penalties -10 log10(present / total) in the reference's float arithmetic, so no device libm is involved.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"

namespace pg {

constexpr uint32_t UMAXV = 0xffffffffu;
constexpr uint16_t PEN_DEFAULT = 25;  // SamplingEmissions default_penalty
constexpr int SAMPLER_MAX_PL = 32;    // paths per lane -> P <= 1024

struct SamplerArgs {
  uint32_t V, P, Pw;            // Pw = words per mask row = ceil(P / 32)
  const uint16_t* p2a;          // [V*P] allele id of every path
  const uint32_t* aoff;         // [V+1]
  const uint16_t* aids;         // [A]
  const uint8_t* aundef;        // [A]
  const uint16_t* akoff;        // [A]
  const uint32_t* akmask;       // [A]
  const uint32_t* koff;         // [V+1]
  const uint16_t* kcounts;      // [K]
  uint16_t* pidx;               // [V*P] index of the path's allele in the variant's allele list
  uint16_t* pen;                // [A]   current penalty of every (variant, allele)
  uint16_t* cpen;               // [V*P] penalty of every (column, path) for the current pass
  const uint32_t* sw;           // [V]   recombination cost into column v (sw[0] unused)
  uint32_t* used;               // [V*Pw] bit i of row v: path i taken by an earlier pass
  uint16_t* back;               // [V*P] back pointers of the current pass (0xffff = none)
  unsigned long long* out_paths;  // [n_out*V]
  uint32_t* best_scores;        // [size]
  uint32_t* end_state;          // [1] best end state of the current pass
  const uint16_t* pen_lut;      // [33*33] penalty by (total, present) k-mers
  uint16_t allele_penalty;
};

// ---- setup: list index of every path's allele, initial penalties ----------------------------------------------
__global__ void __launch_bounds__(256) sampler_pidx_kernel(SamplerArgs a) {
  const uint64_t n = (uint64_t)a.V * a.P;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = (uint32_t)(x / a.P);
    const uint16_t id = a.p2a[x];
    const uint32_t ab = a.aoff[v], ae = a.aoff[v + 1];
    uint32_t k = 0;
    for (uint32_t q = ab; q < ae; ++q)
      if (a.aids[q] == id) k = q - ab;
    a.pidx[x] = (uint16_t)k;
  }
}

// SamplingEmissions ctor (samplingemissions.cpp:9-32): undefined 50; else -10 log10(fraction of the allele's k-mers seen at
// least 3 times) (LUT), 25 if none was seen, 0 if the allele has no k-mers.
__global__ void __launch_bounds__(256) sampler_penalty_kernel(SamplerArgs a, uint32_t A_total, const uint32_t* __restrict__ allele_variant) {
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < A_total; q += gridDim.x * blockDim.x) {
    if (a.aundef[q]) {
      a.pen[q] = 50;
      continue;
    }
    const uint32_t v = allele_variant[q];
    const uint32_t kb = a.koff[v], K = a.koff[v + 1] - kb;
    const uint32_t off = a.akoff[q];
    uint32_t mask = a.akmask[q], total = 0, present = 0;
    while (mask) {
      const uint32_t b = (uint32_t)__ffs(mask) - 1u;
      mask &= mask - 1u;
      const uint32_t k = off + b;
      if (k < K) {
        ++total;
        present += a.kcounts[kb + k] >= 3 ? 1u : 0u;
      }
    }
    a.pen[q] = a.pen_lut[total * 33 + present];
  }
}

__global__ void __launch_bounds__(256) sampler_cpen_kernel(SamplerArgs a) {
  const uint64_t n = (uint64_t)a.V * a.P;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = (uint32_t)(x / a.P);
    a.cpen[x] = a.pen[a.aoff[v] + a.pidx[x]];
  }
}

__device__ __forceinline__ uint32_t sat_add(uint32_t x, uint32_t y) {  // haplotypesampler.cpp:253-254, 262, 273-274
  const uint32_t s = x + y;
  return s < x ? UMAXV : s;
}

// (value, id) pairs ordered lexicographically; a pair with value UMAX never qualifies (get_column_minima compares with <)
struct Best2 {
  uint32_t v1, i1, v2, i2;
  __device__ __forceinline__ void init() { v1 = v2 = UMAXV; i1 = i2 = UMAXV; }
  __device__ __forceinline__ void add(uint32_t v, uint32_t i) {
    if (v == UMAXV) return;
    if (v < v1 || (v == v1 && i < i1)) {
      v2 = v1; i2 = i1; v1 = v; i1 = i;
    } else if (v < v2 || (v == v2 && i < i2)) {
      v2 = v; i2 = i;
    }
  }
};

// ---- one forward pass: one warp per launch ----------------------------------------------------------------------
template <int PL>
__global__ void __launch_bounds__(32) sampler_forward_kernel(SamplerArgs a, uint32_t pass) {
  const uint32_t lane = threadIdx.x, V = a.V, P = a.P, Pw = a.Pw;
  uint32_t val[PL];
  uint32_t um_prev[PL], um_cur[PL];       // mask word j of the previous / current column (bit `lane` is mine)
  uint16_t pn_cur[PL], pn_nxt[PL];
  uint32_t um_nxt[PL];
  auto load_row = [&](uint32_t v, uint16_t (&pn)[PL], uint32_t (&um)[PL]) {
#pragma unroll
    for (int j = 0; j < PL; ++j) {
      const uint32_t i = lane + 32u * j;
      pn[j] = (v < V && i < P) ? a.cpen[(size_t)v * P + i] : (uint16_t)0;
      um[j] = (v < V && (uint32_t)j < Pw) ? a.used[(size_t)v * Pw + j] : 0u;
    }
  };
  load_row(0, pn_cur, um_cur);
  load_row(1, pn_nxt, um_nxt);
#pragma unroll
  for (int j = 0; j < PL; ++j) {
    val[j] = 0;
    um_prev[j] = 0;
  }
  for (uint32_t v = 0; v < V; ++v) {
    // minima of the previous column over the paths that were free there
    Best2 b;
    b.init();
    if (v > 0) {
#pragma unroll
      for (int j = 0; j < PL; ++j) {
        const uint32_t i = lane + 32u * j;
        if (i < P && !((um_prev[j] >> lane) & 1u)) b.add(val[j], i);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ov1 = __shfl_xor_sync(0xffffffffu, b.v1, o), oi1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
        const uint32_t ov2 = __shfl_xor_sync(0xffffffffu, b.v2, o), oi2 = __shfl_xor_sync(0xffffffffu, b.i2, o);
        b.add(ov1, oi1);
        b.add(ov2, oi2);
      }
    }
    const uint32_t sw = v > 0 ? a.sw[v] : 0u;
#pragma unroll
    for (int j = 0; j < PL; ++j) {
      const uint32_t i = lane + 32u * j;
      if (i >= P) continue;
      uint32_t cell = 0, from = UMAXV, out;
      if ((um_cur[j] >> lane) & 1u) {
        out = UMAXV;                       // taken by an earlier pass (haplotypesampler.cpp:205-212)
      } else {
        if (v > 0) {
          const bool is_first = i == b.i1;
          cell = sat_add(is_first ? b.v2 : b.v1, sw);
          from = is_first ? b.i2 : b.i1;
          if (!((um_prev[j] >> lane) & 1u) && val[j] < cell) {
            cell = val[j];
            from = i;
          }
        }
        out = sat_add(cell, (uint32_t)pn_cur[j]);
      }
      val[j] = out;
      a.back[(size_t)v * P + i] = (uint16_t)(from == UMAXV ? 0xffffu : from);
    }
    // rotate the streamed rows; request column v + 2
#pragma unroll
    for (int j = 0; j < PL; ++j) {
      um_prev[j] = um_cur[j];
      um_cur[j] = um_nxt[j];
      pn_cur[j] = pn_nxt[j];
    }
    load_row(v + 2, pn_nxt, um_nxt);
  }
  // best end state: first minimal entry of the last column (haplotypesampler.cpp:131-141)
  uint32_t bv = UMAXV, bi = UMAXV;
#pragma unroll
  for (int j = 0; j < PL; ++j) {
    const uint32_t i = lane + 32u * j;
    if (i < P && (val[j] < bv || (val[j] == bv && i < bi))) {
      bv = val[j];
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t ov = __shfl_xor_sync(0xffffffffu, bv, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    a.best_scores[pass] = bv;
    a.end_state[0] = bi;
  }
}

// ---- backtrace of one pass: visited cells are masked, visited alleles penalised (haplotypesampler.cpp:147-166) ----
__global__ void __launch_bounds__(32) sampler_backtrace_kernel(SamplerArgs a, uint32_t pass) {
  __shared__ uint16_t s_row[2][SAMPLER_MAX_PL * 32];
  const uint32_t lane = threadIdx.x, V = a.V, P = a.P, Pw = a.Pw;
  uint32_t best = a.end_state[0];
  if (best >= P) best = 0;
  auto stage = [&](uint32_t v, int buf) {
    for (uint32_t i = lane; i < P; i += 32) s_row[buf][i] = a.back[(size_t)v * P + i];
  };
  if (V) stage(V - 1, (V - 1) & 1);
  for (uint32_t v = V; v-- > 0;) {
    if (v > 0) stage(v - 1, (v - 1) & 1);  // next row on its way while this column is handled
    __syncwarp();
    if (lane == 0) {
      a.out_paths[(size_t)pass * V + v] = best;
      const uint32_t q = a.aoff[v] + a.pidx[(size_t)v * P + best];
      uint32_t p = (uint32_t)a.pen[q] + a.allele_penalty;   // SamplingEmissions::penalize (samplingemissions.cpp:38-44)
      if ((uint16_t)p > PEN_DEFAULT) p = PEN_DEFAULT;
      a.pen[q] = (uint16_t)p;
      a.used[(size_t)v * Pw + (best >> 5)] |= 1u << (best & 31u);
    }
    const uint32_t nb = s_row[v & 1][best];
    best = nb == 0xffffu ? 0u : nb;
    __syncwarp();
  }
}

// the sampled panel (update_unique_kmers, haplotypesampler.cpp:289-303; update_paths of the UniqueKmers classes): allele of
// every sampled path, and for every k-mer whether it still lies on a remaining allele
__global__ void __launch_bounds__(128) sampler_update_kernel(SamplerArgs a, uint32_t n_out, uint16_t* __restrict__ new_p2a,
                                                             uint8_t* __restrict__ kmer_keep) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < a.V; v += gridDim.x * blockDim.x) {
    const uint32_t ab = a.aoff[v], ae = a.aoff[v + 1];
    const uint32_t kb = a.koff[v], K = a.koff[v + 1] - kb;
    for (uint32_t k = 0; k < K; ++k) kmer_keep[kb + k] = 0;
    for (uint32_t j = 0; j < n_out; ++j) {
      const uint32_t path = (uint32_t)a.out_paths[(size_t)j * a.V + v];
      new_p2a[(size_t)v * n_out + j] = a.p2a[(size_t)v * a.P + path];
      const uint32_t q = ab + a.pidx[(size_t)v * a.P + path];
      if (q >= ae) continue;
      const uint32_t off = a.akoff[q];
      uint32_t mask = a.akmask[q];
      while (mask) {
        const uint32_t b = (uint32_t)__ffs(mask) - 1u;
        mask &= mask - 1u;
        if (off + b < K) kmer_keep[kb + off + b] = 1;
      }
    }
  }
}

// samplingtransitions.cpp:5-14.  That file has no `using namespace std`, so its unqualified exp / log10 are the C
// double-precision functions; the arithmetic around them is long double.
static uint32_t recombination_cost(uint64_t from, uint64_t to, double recomb_rate, unsigned short nr_paths, long double effective_N) {
  const long double distance = (to - from) * 0.000004L * ((long double)recomb_rate) * effective_N;
  const long double recomb_prob = (1.0L - ::exp((double)(-distance / (long double)nr_paths))) * (1.0L / (long double)nr_paths);
  return (unsigned int)(-10.0 * ::log10((double)recomb_prob));
}

template <int PL>
static void launch_forward(const SamplerArgs& a, uint32_t pass, cudaStream_t s) {
  sampler_forward_kernel<PL><<<1, 32, 0, s>>>(a, pass);
}

}  // namespace pg

using namespace pg;

extern "C" int pg_haplotype_sample(int device, const pg_panel* panel, uint32_t size, double recombrate, double effective_N,
                                   int add_reference, uint16_t allele_penalty, uint64_t* sampled_paths, uint32_t* best_scores,
                                   uint16_t* new_path_to_allele, uint32_t* new_kmer_count, uint16_t* new_counts) {
  pg::NvtxRange nvtx_("pg_haplotype_sample");
  clear_error();
  if (!panel || !sampled_paths || !best_scores || !new_path_to_allele || !new_kmer_count || !new_counts) return fail(PG_ERR_ARG, "null argument");
  const uint32_t V = panel->n_variants, P = panel->n_paths;
  if (size < 1 || V == 0) return PG_OK;
  if (P < 1 || P > 32u * SAMPLER_MAX_PL) return fail(PG_ERR_ARG, "haplotype sampling supports up to 1024 paths");
  PG_TRY(check_device(device));
  DeviceGuard g(device);
  const uint64_t A = panel->allele_offsets[V], K = panel->kmer_offsets[V];
  const uint32_t Pw = (P + 31) / 32;
  const size_t n_out = size + (add_reference ? 1 : 0);
  // host parts: recombination costs, penalty table (reference float arithmetic), variant of every allele entry
  std::vector<uint32_t> sw(V, 0), avar(std::max<uint64_t>(A, 1));
  for (uint32_t v = 1; v < V; ++v) sw[v] = recombination_cost(panel->positions[v - 1], panel->positions[v], recombrate, (unsigned short)P, (long double)effective_N);
  for (uint32_t v = 0; v < V; ++v)
    for (uint32_t q = panel->allele_offsets[v]; q < panel->allele_offsets[v + 1]; ++q) avar[q] = v;
  std::vector<uint16_t> lut(33 * 33, 0);
  for (unsigned total = 0; total <= 32; ++total)
    for (unsigned present = 0; present <= total; ++present) {
      const float fraction = total > 0 ? present / (float)total : 1.0f;   // samplingemissions.cpp:21-27
      lut[total * 33 + present] = fraction > 0.0 ? (unsigned short)(-10.0 * std::log10(fraction)) : PEN_DEFAULT;
    }
  DevBuf<uint16_t> d_p2a, d_aids, d_akoff, d_kcounts, d_pidx, d_pen, d_cpen, d_back, d_lut, d_newp2a;
  DevBuf<uint32_t> d_aoff, d_akmask, d_koff, d_sw, d_used, d_scores, d_end, d_avar;
  DevBuf<uint8_t> d_aundef, d_keep;
  DevBuf<unsigned long long> d_paths;
  const size_t VP = (size_t)V * P;
  PG_TRY(d_p2a.reserve(VP)); PG_TRY(d_pidx.reserve(VP)); PG_TRY(d_cpen.reserve(VP)); PG_TRY(d_back.reserve(VP));
  PG_TRY(d_aoff.reserve(V + 1)); PG_TRY(d_koff.reserve(V + 1)); PG_TRY(d_sw.reserve(V));
  PG_TRY(d_aids.reserve(std::max<uint64_t>(A, 1))); PG_TRY(d_aundef.reserve(std::max<uint64_t>(A, 1)));
  PG_TRY(d_akoff.reserve(std::max<uint64_t>(A, 1))); PG_TRY(d_akmask.reserve(std::max<uint64_t>(A, 1)));
  PG_TRY(d_pen.reserve(std::max<uint64_t>(A, 1))); PG_TRY(d_avar.reserve(std::max<uint64_t>(A, 1)));
  PG_TRY(d_kcounts.reserve(std::max<uint64_t>(K, 1))); PG_TRY(d_keep.reserve(std::max<uint64_t>(K, 1)));
  PG_TRY(d_used.reserve((size_t)V * Pw)); PG_TRY(d_scores.reserve(size)); PG_TRY(d_end.reserve(4)); PG_TRY(d_lut.reserve(33 * 33));
  PG_TRY(d_paths.reserve(n_out * V)); PG_TRY(d_newp2a.reserve(n_out * V));
  cudaStream_t s;
  PG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  auto up = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess; };
  cudaError_t ce = cudaSuccess;
  auto chk = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
  chk(up(d_p2a.p, panel->path_to_allele, VP * 2));
  chk(up(d_aoff.p, panel->allele_offsets, (V + 1) * 4));
  chk(up(d_koff.p, panel->kmer_offsets, (V + 1) * 4));
  chk(up(d_aids.p, panel->allele_ids, A * 2));
  chk(up(d_aundef.p, panel->allele_undefined, A));
  chk(up(d_akoff.p, panel->allele_kmer_offset, A * 2));
  chk(up(d_akmask.p, panel->allele_kmer_mask, A * 4));
  chk(up(d_kcounts.p, panel->kmer_counts, K * 2));
  chk(up(d_sw.p, sw.data(), V * 4));
  chk(up(d_avar.p, avar.data(), A * 4));
  chk(up(d_lut.p, lut.data(), lut.size() * 2));
  chk(cudaMemsetAsync(d_used.p, 0, (size_t)V * Pw * 4, s));
  SamplerArgs a;
  memset(&a, 0, sizeof(a));
  a.V = V; a.P = P; a.Pw = Pw; a.p2a = d_p2a.p; a.aoff = d_aoff.p; a.aids = d_aids.p; a.aundef = d_aundef.p; a.akoff = d_akoff.p;
  a.akmask = d_akmask.p; a.koff = d_koff.p; a.kcounts = d_kcounts.p; a.pidx = d_pidx.p; a.pen = d_pen.p; a.cpen = d_cpen.p;
  a.sw = d_sw.p; a.used = d_used.p; a.back = d_back.p; a.out_paths = d_paths.p; a.best_scores = d_scores.p; a.end_state = d_end.p;
  a.pen_lut = d_lut.p; a.allele_penalty = allele_penalty;
  const int grid = (int)std::min<uint64_t>((VP + 255) / 256, 148 * 16);
  sampler_pidx_kernel<<<grid, 256, 0, s>>>(a);
  if (A) sampler_penalty_kernel<<<(int)std::min<uint64_t>((A + 255) / 256, 148 * 16), 256, 0, s>>>(a, (uint32_t)A, d_avar.p);
  count_launch(2);
  const int PL = (int)((P + 31) / 32);
  for (uint32_t pass = 0; pass < size; ++pass) {
    sampler_cpen_kernel<<<grid, 256, 0, s>>>(a);
    if (PL <= 1) launch_forward<1>(a, pass, s);
    else if (PL <= 2) launch_forward<2>(a, pass, s);
    else if (PL <= 4) launch_forward<4>(a, pass, s);
    else if (PL <= 8) launch_forward<8>(a, pass, s);
    else if (PL <= 16) launch_forward<16>(a, pass, s);
    else launch_forward<32>(a, pass, s);
    sampler_backtrace_kernel<<<1, 32, 0, s>>>(a, pass);
    count_launch(3);
  }
  if (add_reference) chk(cudaMemsetAsync(d_paths.p + (size_t)size * V, 0, (size_t)V * 8, s));  // the reference path 0 (:48)
  sampler_update_kernel<<<(int)std::min<uint32_t>((V + 127) / 128, 148 * 8), 128, 0, s>>>(a, (uint32_t)n_out, d_newp2a.p, d_keep.p);
  count_launch();
  chk(cudaGetLastError());
  std::vector<uint8_t> keep(std::max<uint64_t>(K, 1));
  chk(cudaMemcpyAsync(sampled_paths, d_paths.p, n_out * V * 8, cudaMemcpyDeviceToHost, s));
  chk(cudaMemcpyAsync(best_scores, d_scores.p, size * 4, cudaMemcpyDeviceToHost, s));
  chk(cudaMemcpyAsync(new_path_to_allele, d_newp2a.p, n_out * V * 2, cudaMemcpyDeviceToHost, s));
  if (K) chk(cudaMemcpyAsync(keep.data(), d_keep.p, K, cudaMemcpyDeviceToHost, s));
  chk(cudaStreamSynchronize(s));
  cudaStreamDestroy(s);
  if (ce != cudaSuccess) return fail(PG_ERR_CUDA, std::string("haplotype sampling: ") + cudaGetErrorString(ce));
  // compaction of the surviving k-mer counts (order preserved)
  size_t k_out = 0;
  for (uint32_t v = 0; v < V; ++v) {
    uint32_t n = 0;
    for (uint32_t k = panel->kmer_offsets[v]; k < panel->kmer_offsets[v + 1]; ++k)
      if (keep[k]) {
        new_counts[k_out++] = panel->kmer_counts[k];
        ++n;
      }
    new_kmer_count[v] = n;
  }
  return PG_OK;
}
