// Forward-backward kernels over diploid haplotype-pair states (replaces src/hmm.cpp:175-405,
// src/transitionprobabilitycomputer.cpp:8-39 and the per-cell ColumnIndexer lookups of the reference).
//
// Recurrence (per HMM column t, P selected paths, state matrix X is P x P):
//   forward : F_t = e'_t o trans_t(F_{t-1})          F_0 = e'_0
//   backward: Y_t = e'_t o pre_t,  pre_t = trans_{t+1}(Y_{t+1}),  pre_{C-1} = 1
//   trans(X)_{ij} = a X_ij + b (R_i + R_j) + c T,   R = row sums of X, T = total,
//   a = e^{-2x}, b = r e^{-x}, c = r^2 with x = d/P, r = (1-e^{-x})/P  (== t0-2t1+t2, t1-t2, t2 of
//   src/transitionprobabilitycomputer.cpp:14-18; the reference's column sums equal the row sums because
//   X stays symmetric)
//   posterior_t(g) = sum_{cells in genotype g} F_t o pre_t                       (hmm.cpp:362-368)
// Every column is rescaled by an exact power of two taken from the exponent of the previous total (the
// reference divides by the total, hmm.cpp:253-267; any per-column scale cancels in the per-variant
// normalisation, and the un-normalised reference scale is reconstructed in finalize_kernel).
// Underflow fallbacks follow hmm.cpp:258-260 / :377-379: a zero total replaces the column by uniform.
//
// Decomposition (DESIGN.md "HMM schedule"):
//   phase 1 "skeleton": one CTA per (chromosome, direction) walks the whole chain, state in registers,
//            storing a dense P x P checkpoint every B columns.  Latency bound (one barrier + two reductions per
//            column).  For P <= 9 the checkpoints come from the parallel-in-time kernels of hmm_scan.cuh instead and
//            this kernel only runs for chromosomes those kernels flagged (a column total underflowed to zero).
//   phase 2 "blocks":   every block of B columns is independent given its two checkpoints; a CTA recomputes
//            forward (storing F_t: 8 P^2 bytes per column), then backward fusing the posterior (reading F_t
//            back).  All SMs busy: this is the HBM-bound kernel the roofline is reported for.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace pg {

constexpr int HMM_PMAX = 256;       // largest number of selected paths supported
constexpr int HMM_RS_PAD = 288;     // row-sum array length in shared memory
constexpr int HMM_NSLOT = 8;        // descriptor ring slots (power of two: slot = column & 7)
constexpr int HMM_PREFETCH = 4;     // descriptor prefetch distance (columns); 8 columns + L2 hints measured no faster
constexpr int HMM_FAST_A = 4;       // columns with <= 4 alleles: staged emission table + class sums
constexpr int HMM_BMAX = 256;       // largest block length

// Per-column descriptor record, contiguous in HBM so one cp.async burst fetches it.
// doubles: [0..3] a,b,c,kappa of the transition (t-1 -> t); [4..7] the same for (t -> t+1);
//          [8] header: u32 A (alleles), u32 variant index; [9] emission pointer (u64, A > HMM_FAST_A);
//          [10..25] 4x4 emission table (row-major, allele-index space) when A <= HMM_FAST_A;
//          (A <= HMM_FAST_A: [9] holds the allele ids of the allele indices instead, 16 bits each);
//          [26..30] 5 x u64: bit p = allele index (0/1) of selected path p, for columns with A <= 2;
//          [31] u64 offset of the variant's posterior row (gl_off[variant]);
// then u16 aidx[P] (allele index of each selected path), padded to 16 bytes.
constexpr int DESC_HEAD_DOUBLES = 32;
constexpr int DESC_BITS_AT = 26;
__host__ __device__ inline size_t desc_bytes(uint32_t P) { return (size_t)DESC_HEAD_DOUBLES * 8 + (((size_t)P * 2 + 15) & ~(size_t)15); }
constexpr int DESC_SLOT_WORDS = (DESC_HEAD_DOUBLES * 8 + 2 * HMM_PMAX + 16) / 8;

struct ChromCols {
  uint32_t col_begin;  // first global column index of the chromosome
  uint32_t col_end;    // one past the last
  uint32_t blk_begin;  // first global block index of the chromosome
  uint32_t n_blocks;
};

struct ChainParams {
  uint32_t P;                  // selected paths
  uint32_t B;                  // block length (columns)
  const uint8_t* desc;         // [n_cols] descriptor records
  uint32_t desc_stride;
  const ChromCols* chroms;     // [n_chrom]
  double* ckpt_fwd;            // [n_blocks_total][state_stride]  F of the column before block k (k >= 1 in a chrom)
  double* ckpt_bwd;            // [n_blocks_total][state_stride]  Y of the column after block k (k < n_blocks-1)
  double* tot_fwd;             // [n_cols] TF_t = sum F_t
  double* tot_bwd;             // [n_cols] TY_t = sum Y_t
  double* block_buf;           // [grid][B][state_stride] forward columns of the block being processed
  uint32_t state_stride;       // doubles per stored block_buf column = cells per thread x threads (thread-major private layout)
  uint32_t ckpt_stride;        // doubles per checkpoint = P*P: dense row-major, independent of the tile configuration, so the
                               // skeleton may use a wider CTA than the block kernel (and the scan path writes the same format)
  double* post;                // VCF-ordered raw posteriors (zero-initialised)
  const uint64_t* gl_off;      // [n_variants+1]
  const uint32_t* allele_off;  // [n_variants+1]
  const uint16_t* allele_ids;  // [A_total]
  uint32_t* work_counter;      // phase-2 job queue head
  const uint2* jobs;           // [n_jobs] (chromosome, block)
  uint32_t n_jobs;
  const uint32_t* seq_flags;   // skeleton_kernel: if non-null, only chromosomes with a non-zero flag are walked (hmm_scan.cuh)
};

__device__ __forceinline__ double pow2_scale_of(double T) {
  // exact 2^-(unbiased exponent of T) for T > 0
  const int e = (__double2hiint(T) >> 20) & 0x7ff;
  return __hiloint2double((2046 - e) << 20, 0);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t pair_index(uint32_t a, uint32_t b) {  // VCF order (genotypingresult.cpp:61)
  const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
  return hi * (hi + 1) / 2 + lo;
}

struct ChainSmem {
  double rs[2][HMM_RS_PAD];              // row sums of the current / next state
  double wr[2][HMM_FAST_A][HMM_RS_PAD];  // per-row, per-column-class sums of F o pre, double-buffered
  uint8_t wr_ai[2][HMM_RS_PAD];          // allele index of each row for the column held in wr[.]
  double tf[HMM_BMAX];                   // forward totals of the block's columns (phase 2)
  unsigned long long desc[HMM_NSLOT][DESC_SLOT_WORDS];
};

// =================================================================================================
// Register-resident chain.  L lanes per row, each lane owns CPL CONTIGUOUS columns [lc*CPL, lc*CPL+CPL),
// RPW rows per thread, G = 32/L rows per warp pass; warp w owns rows [w*RPW*G, (w+1)*RPW*G).
// States are stored thread-major ([cell][thread]) so every warp access is a coalesced 256 B segment.
// =================================================================================================
template <int L, int CPL, int RPW, int NT>
struct Chain {
  static constexpr int G = 32 / L;
  static constexpr bool ONEWARP = NT == 32;
  static constexpr int CELLS = RPW * CPL;
  double x[RPW][CPL];
  double rrow[RPW];
  int w, lane, lr, lc, col0;
  int P;
  int row0, row_lim;           // rows [row0, row_lim) of the state are this CTA's (the whole matrix unless the chain is split over a cluster)
  int n_peers;                 // other CTAs of the cluster that need my row sums
  double* rs_peer[3];          // their rs[][] arrays (distributed shared memory)
  uint32_t vmask;  // bit s set <=> column col0+s < P
  ChainSmem* sm;
  const ChainParams* prm;
  double* post_col;
  const uint16_t* ids_col;

  __device__ __forceinline__ int row(int r) const { return row0 + (w * RPW + r) * G + lr; }
  // a row sum goes into my rs[buf] and, when the chain is split over a cluster, into every peer's (DSMEM stores)
  __device__ __forceinline__ void put_rowsum(int buf, int i, double v) const {
    sm->rs[buf][i] = v;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q < n_peers) rs_peer[q][buf * HMM_RS_PAD + i] = v;
  }

  __device__ __forceinline__ void init(ChainSmem* s, const ChainParams* p) {
    sm = s;
    prm = p;
    P = (int)p->P;
    w = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    lr = lane / L;
    lc = lane % L;
    col0 = lc * CPL;
    int nv = P - col0;
    nv = nv < 0 ? 0 : (nv > CPL ? CPL : nv);
    vmask = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
    post_col = nullptr;
    ids_col = nullptr;
    row0 = 0;
    row_lim = P;
    n_peers = 0;
  }

  __device__ __forceinline__ void sync() const {
    if (ONEWARP) __syncwarp();
    else __syncthreads();
  }

  __device__ __forceinline__ double row_reduce(double v) const {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }

  // total of the row sums in rs[buf]; every warp evaluates the same expression -> bitwise identical T
  __device__ __forceinline__ double total(int buf) const {
    if (L == 1) {  // lane-per-row layout: all CPL row sums are read by every lane anyway; same order in all lanes
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int q = 0; q + 1 < CPL; q += 2) {
        t0 += sm->rs[buf][q];
        t1 += sm->rs[buf][q + 1];
      }
      if (CPL & 1) t0 += sm->rs[buf][CPL - 1];
      return t0 + t1;
    }
    double t = 0.0;
    constexpr int NR = (L * CPL + 31) / 32;  // rs[] is zero beyond P, so the trip count can be static
#pragma unroll
    for (int q = 0; q < NR; ++q) t += sm->rs[buf][lane + 32 * q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  }

  // ---- descriptor ring: exactly one cp.async group per call, issued in column order ---------------
  // `slot` is the ring slot of column t: callers advance it incrementally (no division in the column loop)
  __device__ __forceinline__ void prefetch_desc(int t, int lo, int hi, int slot) const {
    if (w == 0) {  // one warp moves the record (<= 752 B = 47 x 16 B); the column barrier publishes it
      if (t >= lo && t < hi) {
        const uint8_t* src = prm->desc + (size_t)(uint32_t)t * prm->desc_stride;
        uint8_t* dst = reinterpret_cast<uint8_t*>(sm->desc[slot]);
        for (uint32_t o = lane * 16; o < prm->desc_stride; o += 32 * 16) cp_async16(dst + o, src + o);
      }
      cp_async_commit();
    }
  }
  __device__ __forceinline__ const double* desc_d(int slot) const { return reinterpret_cast<const double*>(sm->desc[slot]); }
  static __device__ __forceinline__ int slot_next(int s) { return (s + 1) & (HMM_NSLOT - 1); }
  static __device__ __forceinline__ int slot_prev(int s) { return (s - 1) & (HMM_NSLOT - 1); }
  static __device__ __forceinline__ int slot_of(int t) { return t & (HMM_NSLOT - 1); }

  // ---- state I/O: thread-major private layout, only valid cells touch memory ----------------------
  __device__ __forceinline__ void store_state(double* dst) const {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row(r) < row_lim) {
#pragma unroll
        for (int s = 0; s < CPL; ++s)
          if ((vmask >> s) & 1u) dst[(size_t)(r * CPL + s) * NT + threadIdx.x] = x[r][s];
      }
    }
  }
  __device__ __forceinline__ void load_state(const double* src) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const bool rok = row(r) < row_lim;
#pragma unroll
      for (int s = 0; s < CPL; ++s) x[r][s] = (rok && ((vmask >> s) & 1u)) ? src[(size_t)(r * CPL + s) * NT + threadIdx.x] : 0.0;
    }
  }
  // checkpoints: dense row-major P x P (layout shared by every tile configuration and by hmm_scan.cuh)
  __device__ __forceinline__ void store_dense(double* dst) const {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = row(r);
      if (i < row_lim) {
#pragma unroll
        for (int s = 0; s < CPL; ++s)
          if ((vmask >> s) & 1u) dst[(size_t)i * P + col0 + s] = x[r][s];
      }
    }
  }
  __device__ __forceinline__ void load_dense(const double* src) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = row(r);
      const bool rok = i < row_lim;
#pragma unroll
      for (int s = 0; s < CPL; ++s) x[r][s] = (rok && ((vmask >> s) & 1u)) ? src[(size_t)i * P + col0 + s] : 0.0;
    }
  }
  // stored forward column -> registers (issued early so the latency overlaps the reductions of the step)
  __device__ __forceinline__ void load_cells(const double* src, bool dead_uniform, double (&u)[RPW][CPL]) const {
    const double uni = 1.0 / ((double)P * (double)P);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const bool rok = row(r) < row_lim;
#pragma unroll
      for (int s = 0; s < CPL; ++s) {
        const bool ok = rok && ((vmask >> s) & 1u);
        u[r][s] = ok ? (dead_uniform ? uni : src[(size_t)(r * CPL + s) * NT + threadIdx.x]) : 0.0;
      }
    }
  }
  // publish the row sums of x into rs[buf]; caller must sync() before anyone reads them
  __device__ __forceinline__ void publish_rowsums(int buf) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < CPL; ++s) acc += x[r][s];
      acc = row_reduce(acc);
      rrow[r] = acc;
      if (lc == 0 && row(r) < row_lim) put_rowsum(buf, row(r), acc);
    }
  }

  __device__ __forceinline__ void bind_posterior(const double* d) {
    const uint32_t v = reinterpret_cast<const uint32_t*>(d + 8)[1];
    post_col = prm->post + prm->gl_off[v];
    ids_col = prm->allele_ids + prm->allele_off[v];
  }

  // CPL bits of the 256-bit path mask starting at bit `start`
  static __device__ __forceinline__ uint32_t bits_at(const unsigned long long* wds, int start) {
    const int idx = start >> 6, sh = start & 63;
    unsigned long long lo = wds[idx] >> sh;
    if (sh) lo |= wds[idx + 1] << (64 - sh);  // 5 words are stored, so idx+1 is always readable
    return (uint32_t)lo;
  }

  // ---- one column.  FIRST: no transition (pre = 1).  WITH_POST: accumulate F o pre (u = stored F_t). ----
  // Reads rs[cbuf] (row sums of x), writes rs[nbuf]; coefficients (t-1 -> t) forward, (t -> t+1) backward.
  template <bool BACKWARD, bool FIRST, bool WITH_POST>
  __device__ __forceinline__ void step(int slot, int cbuf, int nbuf, double Tprev, const double (&u)[RPW][CPL], int wbuf) {
    const double* d = desc_d(slot);
    const uint32_t A = reinterpret_cast<const uint32_t*>(d + 8)[0];
    const double S = (double)P * (double)P;
    const bool dead = !FIRST && !(Tprev > 0.0);  // previous column underflowed -> uniform replacement
    double ca = 0.0, cb = 0.0, cc = 1.0;
    if (!FIRST) {
      const int o = BACKWARD ? 4 : 0;
      if (!dead) {
        const double sc = pow2_scale_of(Tprev);
        ca = d[o] * sc;
        cb = d[o + 1] * sc;
        cc = d[o + 2] * Tprev * sc;
      } else {
        cc = BACKWARD ? 1.0 / S : d[o + 3] / S;
      }
    }
    // the reference's backward cells of a dead column are all zero (hmm.cpp:348-352 with zero helpers)
    const double wscale = dead ? 0.0 : 1.0;
    const double* rj = &sm->rs[cbuf][col0];  // row sums of my columns (rs is zero-padded beyond P)

    if (A <= 2) {
      // ------------- biallelic column: allele indices come as bitmasks, emission by select -------------
      const unsigned long long* bw = reinterpret_cast<const unsigned long long*>(d + DESC_BITS_AT);
      const uint32_t jb = bits_at(bw, col0);
      const double e00 = d[10], e01 = d[11], e10 = d[14], e11 = d[15];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int i = row(r);
        const bool rok = i < row_lim;
        const uint32_t ib = rok ? (uint32_t)((bw[i >> 6] >> (i & 63)) & 1ull) : 0u;
        const double er0 = rok ? (ib ? e10 : e00) : 0.0, er1 = rok ? (ib ? e11 : e01) : 0.0;
        const double rho = FIRST ? 1.0 : cb * rrow[r] + cc;
        double acc = 0.0, acc2 = 0.0, w0 = 0.0, w1 = 0.0;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
          const bool ok = (vmask >> s) & 1u;  // lane-constant; selects only, no branches in the cell body
          const bool c1 = (jb >> s) & 1u;
          const double pre = FIRST ? 1.0 : fma(ca, x[r][s], fma(cb, rj[s], rho));
          if (WITH_POST) {
            const double wv = u[r][s] * pre * wscale;
            w0 += c1 ? 0.0 : wv;  // accumulated separately: total - w1 would cancel
            w1 += c1 ? wv : 0.0;
          }
          const double pe = pre * (c1 ? er1 : er0);
          const double v = ok ? pe : 0.0;
          x[r][s] = v;
          if (s & 1) acc2 += v;
          else acc += v;
        }
        acc += acc2;
        acc = row_reduce(acc);
        rrow[r] = acc;
        if (lc == 0 && rok) put_rowsum(nbuf, i, acc);
        if (WITH_POST) {
          w0 = row_reduce(w0);
          w1 = row_reduce(w1);
          if (lc == 0 && rok) {
            sm->wr[wbuf][0][i] = w0;
            sm->wr[wbuf][1][i] = w1;
            sm->wr[wbuf][2][i] = 0.0;
            sm->wr[wbuf][3][i] = 0.0;
            sm->wr_ai[wbuf][i] = (uint8_t)ib;
          }
        }
      }
      return;
    }

    // ------------- general column (A > 2): emission by table lookup -------------
    const uint16_t* aidx = reinterpret_cast<const uint16_t*>(d + DESC_HEAD_DOUBLES);
    const bool fastA = A <= HMM_FAST_A;
    const double* eg = reinterpret_cast<const double*>(*reinterpret_cast<const unsigned long long*>(d + 9));
    if (WITH_POST && !fastA) bind_posterior(d);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = row(r);
      const bool rok = i < row_lim;
      const uint32_t ai = rok ? aidx[i] : 0;
      const double rho = FIRST ? 1.0 : cb * rrow[r] + cc;
      double acc = 0.0;
      double wc[HMM_FAST_A];
#pragma unroll
      for (int q = 0; q < HMM_FAST_A; ++q) wc[q] = 0.0;
#pragma unroll
      for (int s = 0; s < CPL; ++s) {
        const bool ok = rok && ((vmask >> s) & 1u);
        const uint32_t aj = ok ? aidx[col0 + s] : 0;
        const double pre = FIRST ? 1.0 : fma(ca, x[r][s], fma(cb, rj[s], rho));
        if (WITH_POST) {
          const double wv = u[r][s] * pre * wscale;
          if (fastA) {
#pragma unroll
            for (int q = 0; q < HMM_FAST_A; ++q) wc[q] += (aj == (uint32_t)q) ? wv : 0.0;
          } else if (wv != 0.0) {
            atomicAdd(post_col + pair_index(ids_col[ai], ids_col[aj]), wv);  // rare general path
          }
        }
        double em = 0.0;
        if (ok) em = fastA ? d[10 + ai * HMM_FAST_A + aj] : __ldg(eg + (size_t)ai * A + aj);
        const double v = pre * em;
        x[r][s] = v;
        acc += v;
      }
      acc = row_reduce(acc);
      rrow[r] = acc;
      if (lc == 0 && rok) put_rowsum(nbuf, i, acc);
      if (WITH_POST && fastA) {
#pragma unroll
        for (int q = 0; q < HMM_FAST_A; ++q) {
          const double c = row_reduce(wc[q]);
          if (lc == 0 && rok) sm->wr[wbuf][q][i] = c;
        }
        if (lc == 0 && rok) sm->wr_ai[wbuf][i] = (uint8_t)ai;
      }
    }
  }

  // ---- posterior writer for a fast-A column whose class sums sit in wr[wbuf] (call after the barrier) ----
  // One warp folds the rows: M[alpha][beta] = sum over rows i with allele index alpha of wr[beta][i].  The writer sits on
  // the critical path of its column (everyone meets it at the next barrier), so all 32 lanes share the row loop:
  // biallelic columns use 8 lanes per (alpha, beta), columns with 3-4 alleles 2 lanes per (alpha, beta).
  __device__ __forceinline__ void write_posterior(int slot, int wbuf) {
    const double* d = desc_d(slot);
    const uint32_t A = reinterpret_cast<const uint32_t*>(d + 8)[0];
    if (A > HMM_FAST_A) return;
    if (A <= 2) {
      const uint32_t combo = lane >> 3, sub = lane & 7, alpha = combo >> 1, beta = combo & 1;
      double m = 0.0;
      for (int i = (int)sub; i < P; i += 8) m += (sm->wr_ai[wbuf][i] == alpha) ? sm->wr[wbuf][beta][i] : 0.0;
      m += __shfl_xor_sync(0xffffffffu, m, 1);
      m += __shfl_xor_sync(0xffffffffu, m, 2);
      m += __shfl_xor_sync(0xffffffffu, m, 4);
      const double mt = __shfl_sync(0xffffffffu, m, (int)(((beta << 1) | alpha) << 3));  // M[beta][alpha]
      if (sub == 0 && alpha <= beta && beta < A) {  // destination from the record itself: no dependent global loads
        const unsigned long long idp = reinterpret_cast<const unsigned long long*>(d)[9];
        const unsigned long long base = reinterpret_cast<const unsigned long long*>(d)[31];
        prm->post[base + pair_index((uint32_t)(idp >> (16 * alpha)) & 0xffffu, (uint32_t)(idp >> (16 * beta)) & 0xffffu)] = alpha == beta ? m : m + mt;
      }
      return;
    }
    const uint32_t half = lane >> 4, combo = lane & 15, alpha = combo >> 2, beta = combo & 3;
    double m = 0.0;
    if (alpha < A && beta < A)
      for (int i = (int)half; i < P; i += 2) m += (sm->wr_ai[wbuf][i] == alpha) ? sm->wr[wbuf][beta][i] : 0.0;
    m += __shfl_xor_sync(0xffffffffu, m, 16);
    const double mt = __shfl_sync(0xffffffffu, m, (int)((beta << 2) | alpha));  // M[beta][alpha]
    if (half == 0 && alpha <= beta && beta < A) {
      const unsigned long long idp = reinterpret_cast<const unsigned long long*>(d)[9];
      const unsigned long long base = reinterpret_cast<const unsigned long long*>(d)[31];
      prm->post[base + pair_index((uint32_t)(idp >> (16 * alpha)) & 0xffffu, (uint32_t)(idp >> (16 * beta)) & 0xffffu)] = alpha == beta ? m : m + mt;
    }
  }
};

// -------------------------------------------------------------------------------------------------
// phase 1: skeleton chains.  grid = (n_chrom, 2): y = 0 forward, y = 1 backward.
// -------------------------------------------------------------------------------------------------
template <int L, int CPL, int RPW, int NT>
__global__ void __launch_bounds__(NT) skeleton_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem* sm = reinterpret_cast<ChainSmem*>(smem_raw);
  Chain<L, CPL, RPW, NT> ch;
  ch.init(sm, &p);
  const ChromCols cc = p.chroms[blockIdx.x];
  if (cc.n_blocks <= 1) return;
  if (p.seq_flags && !p.seq_flags[blockIdx.x]) return;  // checkpoints already produced by the scan path
  for (int i = threadIdx.x; i < 2 * HMM_RS_PAD; i += NT) (&sm->rs[0][0])[i] = 0.0;  // zero padding beyond P
  ch.sync();
  const int c0 = (int)cc.col_begin, c1 = (int)cc.col_end, B = (int)p.B;
  const size_t CS = p.ckpt_stride;
  double nou[RPW][CPL];  // unused (no posterior in the skeleton); the compiler drops it
  constexpr int D = HMM_PREFETCH;
  int cur = 0;
  if (blockIdx.y == 0) {
    // forward: needs F at columns c0 + k*B - 1 for k = 1 .. n_blocks-1 (stored as ckpt_fwd[k])
    const int last = c0 + (int)(cc.n_blocks - 1) * B - 1;
    int slot = ch.slot_of(c0), pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(c0 + d, c0, c1, pslot);
      pslot = ch.slot_next(pslot);
    }
    cp_async_wait<D - 1>();
    ch.sync();
    ch.template step<false, true, false>(slot, 0, 0, 0.0, nou, 0);
    ch.prefetch_desc(c0 + D, c0, c1, pslot);
    pslot = ch.slot_next(pslot);
    cp_async_wait<D - 1>();
    ch.sync();
    slot = ch.slot_next(slot);
    int until_ckpt = B - 1;  // columns to go before the state entering the next block is complete
    uint32_t blk = cc.blk_begin + 1;
    for (int t = c0 + 1; t <= last; ++t) {
      if (until_ckpt == 0) {
        ch.store_dense(p.ckpt_fwd + (size_t)blk * CS);
        ++blk;
        until_ckpt = B;
      }
      --until_ckpt;
      const double T = ch.total(cur);
      ch.template step<false, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      ch.prefetch_desc(t + D, c0, c1, pslot);
      pslot = ch.slot_next(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      slot = ch.slot_next(slot);
      cur ^= 1;
    }
    ch.store_dense(p.ckpt_fwd + (size_t)(cc.blk_begin + cc.n_blocks - 1) * CS);
  } else {
    // backward: needs Y at columns c0 + k*B for k = 1 .. n_blocks-1 (stored as ckpt_bwd[k-1]).
    // The ring is walked downwards: slot(t-1) = slot(t) - 1 (mod NSLOT).
    auto slot_prev = [](int s) { return (s - 1) & (HMM_NSLOT - 1); };
    const int first = c0 + B;
    int slot = ch.slot_of(c1 - 1), pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(c1 - 1 - d, c0, c1, pslot);
      pslot = slot_prev(pslot);
    }
    cp_async_wait<D - 1>();
    ch.sync();
    ch.template step<true, true, false>(slot, 0, 0, 0.0, nou, 0);
    int rel = (c1 - 1 - c0) % B;        // position of column t inside its block (one division per chain)
    uint32_t blk = cc.blk_begin + (uint32_t)((c1 - 1 - c0) / B);
    if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
    ch.prefetch_desc(c1 - 1 - D, c0, c1, pslot);
    pslot = slot_prev(pslot);
    cp_async_wait<D - 1>();
    ch.sync();
    slot = slot_prev(slot);
    for (int t = c1 - 2; t >= first; --t) {
      if (rel == 0) {
        rel = B;
        --blk;
      }
      --rel;
      const double T = ch.total(cur);
      ch.template step<true, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
      ch.prefetch_desc(t - D, c0, c1, pslot);
      pslot = slot_prev(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      slot = slot_prev(slot);
      cur ^= 1;
    }
  }
  cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// phase 1, LEAN walk (one row per thread, L = 2 or 4 lanes per row, CPL = 17 columns per lane: P <= 34 / P <= 68).
//
// What bounds the walk (profiles/r2_skeleton.md): one chain = one CTA = one or two warps per scheduler, so the schedulers
// have nothing to switch to and every dependent instruction costs its full latency - measured on B200
// (scripts/microbench/latency.cu): fp64 op 8 cycles, one shuffle level of a double 35, shared store -> barrier -> load
// 50-60.  The generic walk above spends ~1450 cycles per column at P = 33 WHATEVER its tile shape; its source-level profile
// shows where: ~400 instructions per warp and column around 4 fp64 operations per cell (selects, address arithmetic,
// re-materialised shared-window bases), the descriptor prefetch (address arithmetic + cp.async + wait_group in warp 0 while
// the other warps sit at the barrier) and the loads of the next record.  This kernel removes those:
//   * descriptor records arrive by TMA: ONE elected thread issues `cp.async.bulk` (global -> shared, 336 B at P = 33) per
//     column, 6 columns ahead, completion through one mbarrier per ring slot (complete_tx); consumers `try_wait` on the slot's
//     phase - no per-lane address arithmetic, no wait_group, and completion makes the record visible without a CTA barrier;
//   * the record of column t+1 is read at the TOP of column t (software pipelining): its shared-memory latency overlaps the
//     row-sum loads of column t;
//   * the 17 row sums of the thread's columns are loaded once; their sum, completed by log2(L) shuffle levels, IS the total
//     (every lane group adds the same numbers in the same order: bitwise the same T, hence the same power-of-two scale, in
//     every thread);
//   * 4 fp64 operations + one select per cell; only the trailing NMASK cells carry the "column < P" select;
//   * shared memory is addressed through a 32-bit base taken once (no S2R / LEA re-materialisation in the loop);
//   * the dead-column case (total == 0 -> uniform) is branch-free.
// Columns with more than two alleles take the generic step (same registers, same shared-memory row sums).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  unsigned long long a;
  asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
  return (uint32_t)a;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds_v2f64(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void lds_v2u64(uint32_t a, unsigned long long& x, unsigned long long& y) {
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one descriptor record, global -> shared, by the TMA engine; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_record(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct LeanCol {  // chain-independent inputs of one column
  double da, db, dc, dk, er0, er1;
  uint32_t jb, A;
};

template <int L, int CPL, int NT>
struct LeanCtx {
  uint32_t rs_base;    // shared address of sm->rs[0][0]
  uint32_t desc_base;  // shared address of sm->desc[0]
  uint32_t bar_base;   // shared address of the 8 slot mbarriers
  uint32_t phase;      // bit s: parity to wait for on slot s
  int col0, row, lc;
  bool rok;
  double invS;
};

template <bool BACKWARD, int L, int CPL, int NT>
__device__ __forceinline__ void lean_load(LeanCtx<L, CPL, NT>& cx, int slot, LeanCol& c) {
  mbar_wait(cx.bar_base + 8u * (uint32_t)slot, (cx.phase >> slot) & 1u);
  cx.phase ^= 1u << slot;
  const uint32_t d = cx.desc_base + (uint32_t)slot * (uint32_t)(DESC_SLOT_WORDS * 8);
  const uint32_t o = BACKWARD ? 32u : 0u;
  lds_v2f64(d + o, c.da, c.db);
  lds_v2f64(d + o + 16u, c.dc, c.dk);
  c.A = lds_u32(d + 64u);
  double e00, e01, e10, e11;
  lds_v2f64(d + 80u, e00, e01);
  lds_v2f64(d + 112u, e10, e11);
  unsigned long long w0, w1;
  lds_v2u64(d + (uint32_t)(DESC_BITS_AT * 8), w0, w1);  // allele bits of paths 0..127 (P <= 68 here)
  const int c0 = cx.col0;
  unsigned long long lo = c0 < 64 ? (w0 >> c0) : (w1 >> (c0 - 64));
  if (c0 > 0 && c0 < 64) lo |= w1 << (64 - c0);
  c.jb = (uint32_t)lo;
  const uint32_t ib = (uint32_t)(((cx.row < 64 ? w0 : w1) >> (cx.row & 63)) & 1ull);
  c.er0 = cx.rok ? (ib ? e10 : e00) : 0.0;
  c.er1 = cx.rok ? (ib ? e11 : e01) : 0.0;
}

template <bool BACKWARD, int NMASK, int L, int CPL, int NT>
__device__ __forceinline__ void lean_step(Chain<L, CPL, 1, NT>& ch, const LeanCtx<L, CPL, NT>& cx, const LeanCol& c, int cbuf, int nbuf) {
  const uint32_t ra = cx.rs_base + (uint32_t)(cbuf * HMM_RS_PAD + cx.col0) * 8u;
  double rj[CPL];
#pragma unroll
  for (int s = 0; s < CPL; ++s) rj[s] = lds_f64(ra + 8u * s);  // zero beyond P
  double t0 = rj[0], t1 = rj[1], t2 = rj[2], t3 = rj[3];
#pragma unroll
  for (int s = 4; s < CPL; ++s) {
    if ((s & 3) == 0) t0 += rj[s];
    else if ((s & 3) == 1) t1 += rj[s];
    else if ((s & 3) == 2) t2 += rj[s];
    else t3 += rj[s];
  }
  double T = (t0 + t1) + (t2 + t3);
#pragma unroll
  for (int o = 1; o < L; o <<= 1) T += __shfl_xor_sync(0xffffffffu, T, o);
  // previous column underflowed (T == 0) -> uniform replacement (hmm.cpp:258-260 / :377-379), without a branch
  const bool live = T > 0.0;
  const double sc = pow2_scale_of(live ? T : 1.0);
  const double ca = live ? c.da * sc : 0.0;
  const double cb = live ? c.db * sc : 0.0;
  const double cc = live ? c.dc * T * sc : (BACKWARD ? cx.invS : c.dk * cx.invS);
  const double rho = fma(cb, ch.rrow[0], cc);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int s = 0; s < CPL; ++s) {
    const double pre = fma(ca, ch.x[0][s], fma(cb, rj[s], rho));
    double v = pre * (((c.jb >> s) & 1u) ? c.er1 : c.er0);
    if (s >= CPL - NMASK) v = ((ch.vmask >> s) & 1u) ? v : 0.0;  // columns beyond P: trailing cells of the last lane group
    ch.x[0][s] = v;
    if ((s & 3) == 0) a0 += v;
    else if ((s & 3) == 1) a1 += v;
    else if ((s & 3) == 2) a2 += v;
    else a3 += v;
  }
  double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  ch.rrow[0] = acc;
  if (cx.lc == 0 && cx.rok) sts_f64(cx.rs_base + (uint32_t)(nbuf * HMM_RS_PAD + cx.row) * 8u, acc);
}

constexpr int LEAN_AHEAD = 6;  // records in flight ahead of the column being computed (ring of HMM_NSLOT = 8)

template <int L, int CPL, int NT, int NMASK>
__global__ void __launch_bounds__(NT) skeleton_lean_kernel(const ChainParams p) {
  static_assert(CPL <= 32 && L * CPL <= 128, "allele bits of a lane's columns must fit one word; paths must fit two mask words");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem* sm = reinterpret_cast<ChainSmem*>(smem_raw);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + ((sizeof(ChainSmem) + 15) & ~size_t(15)));
  Chain<L, CPL, 1, NT> ch;
  ch.init(sm, &p);
  const ChromCols cc = p.chroms[blockIdx.x];
  if (cc.n_blocks <= 1) return;
  if (p.seq_flags && !p.seq_flags[blockIdx.x]) return;
  LeanCtx<L, CPL, NT> cx;
  cx.rs_base = smem_addr(&sm->rs[0][0]);
  cx.desc_base = smem_addr(&sm->desc[0][0]);
  cx.bar_base = smem_addr(bars);
  cx.phase = 0;
  cx.col0 = ch.col0;
  cx.row = ch.row(0);
  cx.lc = ch.lc;
  cx.rok = cx.row < ch.row_lim;
  cx.invS = 1.0 / ((double)ch.P * (double)ch.P);
  for (int i = threadIdx.x; i < 2 * HMM_RS_PAD; i += NT) (&sm->rs[0][0])[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < HMM_NSLOT; ++i) mbar_init(cx.bar_base + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int c0 = (int)cc.col_begin, c1 = (int)cc.col_end, B = (int)p.B;
  const size_t CS = p.ckpt_stride;
  const uint32_t rec = p.desc_stride;
  double nou[1][CPL];
  LeanCol col;
  int cur = 0;
  auto issue = [&](int t) {  // elected thread: record of column t -> ring slot t & 7
    const uint32_t sl = (uint32_t)(t & (HMM_NSLOT - 1));
    tma_load_record(cx.desc_base + sl * (uint32_t)(DESC_SLOT_WORDS * 8), p.desc + (size_t)(uint32_t)t * rec, rec, cx.bar_base + 8u * sl);
  };
  if (blockIdx.y == 0) {
    const int last = c0 + (int)(cc.n_blocks - 1) * B - 1;   // columns c0 .. last are walked
    if (threadIdx.x == 0)
      for (int t = c0; t <= last && t <= c0 + LEAN_AHEAD; ++t) issue(t);
    int slot = ch.slot_of(c0);
    mbar_wait(cx.bar_base + 8u * (uint32_t)slot, 0);
    cx.phase ^= 1u << slot;
    ch.template step<false, true, false>(slot, 0, 0, 0.0, nou, 0);
    slot = ch.slot_next(slot);
    if (c0 + 1 <= last) lean_load<false>(cx, slot, col);
    __syncthreads();
    int until_ckpt = B - 1;
    uint32_t blk = cc.blk_begin + 1;
    for (int t = c0 + 1; t <= last; ++t) {
      // the slot of record t-2 is free (its last reader passed the barrier of column t-1): refill it with record t+6
      if (threadIdx.x == 0 && t + LEAN_AHEAD <= last) issue(t + LEAN_AHEAD);
      const LeanCol cur_col = col;
      const int nslot = ch.slot_next(slot);
      if (t < last) lean_load<false>(cx, nslot, col);  // record of column t+1, overlapped with the step below
      if (until_ckpt == 0) {
        ch.store_dense(p.ckpt_fwd + (size_t)blk * CS);
        ++blk;
        until_ckpt = B;
      }
      --until_ckpt;
      if (cur_col.A <= 2) {
        lean_step<false, NMASK>(ch, cx, cur_col, cur, cur ^ 1);
      } else {
        const double T = ch.total(cur);
        ch.template step<false, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      }
      slot = nslot;
      __syncthreads();
      cur ^= 1;
    }
    ch.store_dense(p.ckpt_fwd + (size_t)(cc.blk_begin + cc.n_blocks - 1) * CS);
  } else {
    const int first = c0 + B;                               // columns c1-1 .. first are walked
    if (threadIdx.x == 0)
      for (int t = c1 - 1; t >= first && t >= c1 - 1 - LEAN_AHEAD; --t) issue(t);
    int slot = ch.slot_of(c1 - 1);
    mbar_wait(cx.bar_base + 8u * (uint32_t)slot, 0);
    cx.phase ^= 1u << slot;
    ch.template step<true, true, false>(slot, 0, 0, 0.0, nou, 0);
    int rel = (c1 - 1 - c0) % B;
    uint32_t blk = cc.blk_begin + (uint32_t)((c1 - 1 - c0) / B);
    if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
    slot = ch.slot_prev(slot);
    if (c1 - 2 >= first) lean_load<true>(cx, slot, col);
    __syncthreads();
    for (int t = c1 - 2; t >= first; --t) {
      if (threadIdx.x == 0 && t - LEAN_AHEAD >= first) issue(t - LEAN_AHEAD);
      const LeanCol cur_col = col;
      const int nslot = ch.slot_prev(slot);
      if (t > first) lean_load<true>(cx, nslot, col);
      if (rel == 0) {
        rel = B;
        --blk;
      }
      --rel;
      if (cur_col.A <= 2) {
        lean_step<true, NMASK>(ch, cx, cur_col, cur, cur ^ 1);
      } else {
        const double T = ch.total(cur);
        ch.template step<true, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      }
      if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
      slot = nslot;
      __syncthreads();
      cur ^= 1;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// phase 1 over a thread-block CLUSTER.  The walk above is bound by what one SM can issue and move per column (ncu at
// P = 65: 2300 cycles per column with 9 warps, issue + shared-memory pipes; shortening or re-ordering the dependent chain
// did not move it), and the longest chromosome's walk is the critical path of the whole stage once the sample is sharded
// over GPUs.  Here the ROWS of the state are split over the CTAs of a cluster (2 or 4 SMs per chain): every CTA computes
// the cells of its rows for all P columns, writes the row sums it produces into its own shared memory AND into its peers'
// (distributed shared memory), and one cluster barrier per column (arrive.release / wait.acquire) replaces the CTA
// barrier.  Everything else - descriptors, scaling, the uniform replacement of dead columns, the dense checkpoints (rows
// are disjoint) - is the code of the single-CTA walk; the results are bitwise the same because every cell is computed by
// the same expression from the same row sums (the total is summed in the same order by every warp).
// grid = (n_chrom * C, 2), cluster = (C, 1, 1); NT threads serve ceil(P / C) rows.
// -------------------------------------------------------------------------------------------------
template <int L, int CPL, int RPW, int NT>
__global__ void __launch_bounds__(NT) skeleton_cluster_kernel(const ChainParams p) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem* sm = reinterpret_cast<ChainSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  Chain<L, CPL, RPW, NT> ch;
  ch.init(sm, &p);
  const int rows_per = ((int)p.P + C - 1) / C;
  ch.row0 = rank * rows_per;
  ch.row_lim = min((int)p.P, ch.row0 + rows_per);
  ch.n_peers = C - 1;
#pragma unroll
  for (int q = 0; q < 3; ++q) ch.rs_peer[q] = cluster.map_shared_rank(&sm->rs[0][0], (rank + 1 + q) % C);   // (entries >= n_peers unused)
  const ChromCols cc = p.chroms[blockIdx.x / C];
  if (cc.n_blocks <= 1) return;                               // (the whole cluster takes the same branch)
  if (p.seq_flags && !p.seq_flags[blockIdx.x / C]) return;
  for (int i = threadIdx.x; i < 2 * HMM_RS_PAD; i += NT) (&sm->rs[0][0])[i] = 0.0;  // zero padding beyond P
  cluster.sync();                                             // nobody writes into a peer's rs[] before it is zeroed
  const int c0 = (int)cc.col_begin, c1 = (int)cc.col_end, B = (int)p.B;
  const size_t CS = p.ckpt_stride;
  double nou[RPW][CPL];
  constexpr int D = HMM_PREFETCH;
  int cur = 0;
  if (blockIdx.y == 0) {
    const int last = c0 + (int)(cc.n_blocks - 1) * B - 1;
    int slot = ch.slot_of(c0), pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(c0 + d, c0, c1, pslot);
      pslot = ch.slot_next(pslot);
    }
    cp_async_wait<D - 1>();
    __syncthreads();
    ch.template step<false, true, false>(slot, 0, 0, 0.0, nou, 0);
    ch.prefetch_desc(c0 + D, c0, c1, pslot);
    pslot = ch.slot_next(pslot);
    cp_async_wait<D - 1>();
    cluster.sync();
    slot = ch.slot_next(slot);
    int until_ckpt = B - 1;
    uint32_t blk = cc.blk_begin + 1;
    for (int t = c0 + 1; t <= last; ++t) {
      if (until_ckpt == 0) {
        ch.store_dense(p.ckpt_fwd + (size_t)blk * CS);
        ++blk;
        until_ckpt = B;
      }
      --until_ckpt;
      const double T = ch.total(cur);
      ch.template step<false, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      ch.prefetch_desc(t + D, c0, c1, pslot);
      pslot = ch.slot_next(pslot);
      cp_async_wait<D - 1>();
      cluster.sync();
      slot = ch.slot_next(slot);
      cur ^= 1;
    }
    ch.store_dense(p.ckpt_fwd + (size_t)(cc.blk_begin + cc.n_blocks - 1) * CS);
  } else {
    const int first = c0 + B;
    int slot = ch.slot_of(c1 - 1), pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(c1 - 1 - d, c0, c1, pslot);
      pslot = ch.slot_prev(pslot);
    }
    cp_async_wait<D - 1>();
    __syncthreads();
    ch.template step<true, true, false>(slot, 0, 0, 0.0, nou, 0);
    int rel = (c1 - 1 - c0) % B;
    uint32_t blk = cc.blk_begin + (uint32_t)((c1 - 1 - c0) / B);
    if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
    ch.prefetch_desc(c1 - 1 - D, c0, c1, pslot);
    pslot = ch.slot_prev(pslot);
    cp_async_wait<D - 1>();
    cluster.sync();
    slot = ch.slot_prev(slot);
    for (int t = c1 - 2; t >= first; --t) {
      if (rel == 0) {
        rel = B;
        --blk;
      }
      --rel;
      const double T = ch.total(cur);
      ch.template step<true, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      if (rel == 0) ch.store_dense(p.ckpt_bwd + (size_t)(blk - 1) * CS);
      ch.prefetch_desc(t - D, c0, c1, pslot);
      pslot = ch.slot_prev(pslot);
      cp_async_wait<D - 1>();
      cluster.sync();
      slot = ch.slot_prev(slot);
      cur ^= 1;
    }
  }
  cp_async_wait<0>();
  cluster.sync();  // no CTA leaves while a peer may still write into its shared memory
}

// -------------------------------------------------------------------------------------------------
// phase 2: block forward-backward with fused posterior.  Persistent CTAs pull (chromosome, block) jobs.
// -------------------------------------------------------------------------------------------------
template <int L, int CPL, int RPW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) block_kernel(const ChainParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem* sm = reinterpret_cast<ChainSmem*>(smem_raw);
  __shared__ uint32_t s_job;
  Chain<L, CPL, RPW, NT> ch;
  ch.init(sm, &p);
  for (int i = threadIdx.x; i < 2 * HMM_RS_PAD; i += NT) (&sm->rs[0][0])[i] = 0.0;  // zero padding beyond P
  const size_t PP = p.state_stride;
  double* buf = p.block_buf + (size_t)blockIdx.x * p.B * PP;
  constexpr int D = HMM_PREFETCH;
  constexpr int NW = NT / 32;
  double nou[RPW][CPL], ureg[RPW][CPL];
  auto slot_prev = [](int s) { return (s - 1) & (HMM_NSLOT - 1); };
  while (true) {
    ch.sync();
    if (threadIdx.x == 0) s_job = atomicAdd(p.work_counter, 1u);
    ch.sync();
    const uint32_t job = s_job;
    if (job >= p.n_jobs) break;
    const uint2 jb = p.jobs[job];
    const ChromCols cc = p.chroms[jb.x];
    const int c0 = (int)cc.col_begin, c1 = (int)cc.col_end;
    const int cb = c0 + (int)jb.y * (int)p.B;
    const int ce = (cb + (int)p.B < c1) ? cb + (int)p.B : c1;
    const uint32_t gblk = cc.blk_begin + jb.y;
    int cur = 0;

    // ---------------- forward sub-pass: F_t for t in [cb, ce) -> buf ----------------
    int slot = ch.slot_of(cb), pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(cb + d, cb, ce, pslot);
      pslot = ch.slot_next(pslot);
    }
    int t = cb;
    double* bcol = buf;
    if (jb.y == 0) {
      cp_async_wait<D - 1>();
      ch.sync();
      ch.template step<false, true, false>(slot, 0, 0, 0.0, nou, 0);
      ch.store_state(bcol);
      bcol += PP;
      ch.prefetch_desc(cb + D, cb, ce, pslot);
      pslot = ch.slot_next(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      slot = ch.slot_next(slot);
      t = cb + 1;
    } else {
      ch.load_dense(p.ckpt_fwd + (size_t)gblk * p.ckpt_stride);
      ch.publish_rowsums(0);
      cp_async_wait<D - 1>();
      ch.sync();
    }
    for (; t < ce; ++t) {
      const double T = ch.total(cur);
      if (threadIdx.x == 0 && t > cb) sm->tf[t - 1 - cb] = T;
      ch.template step<false, false, false>(slot, cur, cur ^ 1, T, nou, 0);
      ch.store_state(bcol);
      bcol += PP;
      ch.prefetch_desc(t + D, cb, ce, pslot);
      pslot = ch.slot_next(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      slot = ch.slot_next(slot);
      cur ^= 1;
    }
    {
      const double T = ch.total(cur);
      if (threadIdx.x == 0) sm->tf[ce - 1 - cb] = T;
    }
    cp_async_wait<0>();
    ch.sync();
    for (int q = cb + (int)threadIdx.x; q < ce; q += NT) p.tot_fwd[q] = sm->tf[q - cb];

    // ---------------- backward sub-pass with posterior ----------------
    slot = ch.slot_of(ce - 1);
    pslot = slot;
    for (int d = 0; d < D; ++d) {
      ch.prefetch_desc(ce - 1 - d, cb, ce, pslot);
      pslot = slot_prev(pslot);
    }
    cur = 0;
    t = ce - 1;
    bcol = buf + (size_t)(ce - 1 - cb) * PP;
    int pending_slot = -1, pending_w = 0, pending_wbuf = 0;  // column whose class sums await the posterior writer
    int wsel = 0, wbuf = 0;
    if (ce == c1) {
      cp_async_wait<D - 1>();
      ch.sync();
      ch.load_cells(bcol, !(sm->tf[t - cb] > 0.0), ureg);
      ch.template step<true, true, true>(slot, 0, 0, 0.0, ureg, wbuf);
      bcol -= PP;
      ch.prefetch_desc(t - D, cb, ce, pslot);
      pslot = slot_prev(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      pending_slot = slot; pending_w = wsel; pending_wbuf = wbuf;
      wsel = wsel + 1 == NW ? 0 : wsel + 1;
      wbuf ^= 1;
      slot = slot_prev(slot);
      --t;
    } else {
      ch.load_dense(p.ckpt_bwd + (size_t)gblk * p.ckpt_stride);
      ch.publish_rowsums(0);
      cp_async_wait<D - 1>();
      ch.sync();
    }
    for (; t >= cb; --t) {
      ch.load_cells(bcol, !(sm->tf[t - cb] > 0.0), ureg);  // issued first: overlaps the writer and the reductions
      if (pending_slot >= 0 && ch.w == pending_w) ch.write_posterior(pending_slot, pending_wbuf);
      const double T = ch.total(cur);
      // tot_bwd[ce] belongs to the block that computed column ce: a checkpoint may carry a different power-of-two
      // scale than that block's own column (hmm_scan.cuh), and finalize_kernel needs the owner's exponent
      if (threadIdx.x == 0 && t + 1 < ce) p.tot_bwd[t + 1] = T;
      if (t - 2 >= cb) {  // pull the forward column needed two steps from now towards L2
        const char* nxt = reinterpret_cast<const char*>(bcol - 2 * PP);
        for (size_t o = (size_t)threadIdx.x * 128; o < PP * 8; o += (size_t)NT * 128) prefetch_l2(nxt + o);
      }
      ch.template step<true, false, true>(slot, cur, cur ^ 1, T, ureg, wbuf);
      bcol -= PP;
      ch.prefetch_desc(t - D, cb, ce, pslot);
      pslot = slot_prev(pslot);
      cp_async_wait<D - 1>();
      ch.sync();
      pending_slot = slot; pending_w = wsel; pending_wbuf = wbuf;
      wsel = wsel + 1 == NW ? 0 : wsel + 1;
      wbuf ^= 1;
      slot = slot_prev(slot);
      cur ^= 1;
    }
    if (pending_slot >= 0 && ch.w == pending_w) ch.write_posterior(pending_slot, pending_wbuf);
    {
      const double T = ch.total(cur);
      if (threadIdx.x == 0) p.tot_bwd[cb] = T;
    }
    cp_async_wait<0>();
  }
}

}  // namespace pg
