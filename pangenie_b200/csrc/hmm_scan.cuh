// Parallel-in-time checkpoint computation for SMALL path sets (P <= 9): replaces the sequential skeleton walk of
// hmm_kernels.cuh when the haplotype-pair state space is small enough to propagate a whole basis.
//
// The forward column update F_t = e_t o (Q_t F_{t-1} Q_t^T) (src/hmm.cpp:175-273) and the backward update
// (src/hmm.cpp:275-405) are LINEAR maps on symmetric P x P matrices with non-negative coefficients.  For a block
// of B columns the map "state entering the block -> state leaving it" is therefore an NB x NB matrix,
// NB = P(P+1)/2, whose column b is obtained by pushing the b-th symmetric unit matrix through the block.
//   1. basis_kernel : every (block, direction, basis element) is an independent chain of B columns -> all SMs busy
//                     (the sequential skeleton used 2 warps of the whole GPU);
//   2. scan_kernel  : per (chromosome, direction) the checkpoints follow by NB x NB mat-vec products, one per block.
// All terms are non-negative, so the combination is as well conditioned as the sequential recurrence (relative
// error ~ NB * 2^-53).  Every chain carries its own exact power-of-two scale (exponent sum E_b); checkpoints are
// re-scaled by a power of two, which the block kernel and finalize_kernel are invariant to.
// Exactness guard: the reference replaces a column whose total underflowed to zero by a uniform column
// (hmm.cpp:258-260, 377-379) - a non-linear step.  If any basis chain of a chromosome ever sees a zero total, or a
// combined checkpoint has a zero total, the chromosome is flagged and skeleton_kernel recomputes its checkpoints
// sequentially (it returns immediately for unflagged chromosomes).
#pragma once
#include <climits>

#include "hmm_kernels.cuh"

namespace pg {

constexpr int SCAN_CPL = 9;                                  // largest P handled by the scan path
constexpr int SCAN_NB = SCAN_CPL * (SCAN_CPL + 1) / 2;       // 45 basis elements / upper-triangle cells
constexpr int SCAN_THREADS = 192;                            // scan_kernel CTA (>= 4 * SCAN_NB)
constexpr int BASIS_WARPS = 4;                               // warps per basis_kernel CTA

struct TJob {
  uint32_t chrom;
  int32_t t_first;   // first column stepped
  uint32_t n_steps;  // columns stepped (forward: ascending, backward: descending)
  uint32_t dir;      // 0 forward, 1 backward
  uint32_t out_blk;  // checkpoint slot written (index into ckpt_fwd / ckpt_bwd)
  uint32_t pad[3];
};

struct ScanChrom {
  uint32_t tj_begin[2];  // first transfer job of the forward / backward chain (processing order)
  uint32_t n_tj;         // jobs per direction (= n_blocks - 1, 0 if the chromosome has a single block)
  uint32_t out_first[2]; // checkpoint slot written by the first job of each chain; forward counts up, backward down
  uint32_t pad;
};

struct ScanParams {
  const TJob* tjobs;
  uint32_t n_tj;
  uint32_t n_groups;     // basis groups per job = ceil(NB / chains per warp)
  uint32_t n_items;      // n_tj * n_groups
  uint32_t mat_stride;   // doubles per transfer matrix ((NB+1)*NB rounded up to even)
  double* mats;          // [n_tj][mat_stride]: row b < NB = image of basis element b (upper-triangle cell order),
                         // row NB = power-of-two exponent of each row (as doubles, so one cp.async stream brings both)
  const ScanChrom* chroms;
  uint32_t* seq_flags;   // [n_chrom] != 0 -> sequential recomputation required
};

__device__ __forceinline__ int tri_index(int i, int j, int P) { return i * P - (i * (i - 1)) / 2 + (j - i); }  // i <= j

// emission of cell (row r, column s) of the column described by d (layout: hmm_kernels.cuh)
struct ColEm {
  uint32_t A;
  uint32_t bits;        // A <= 2: allele index bit of every path
  double er0, er1;      // A <= 2: emission of (row allele, 0) / (row allele, 1)
  const uint16_t* aidx; // A > 2
  const double* tab;    // A > 2: row of the emission table belonging to this lane's allele
  __device__ __forceinline__ void load(const double* d, int r, bool rok) {
    A = reinterpret_cast<const uint32_t*>(d + 8)[0];
    aidx = reinterpret_cast<const uint16_t*>(d + DESC_HEAD_DOUBLES);
    if (A <= 2) {
      bits = reinterpret_cast<const uint32_t*>(d + DESC_BITS_AT)[0];
      const bool ib = (bits >> r) & 1u;
      er0 = rok ? (ib ? d[14] : d[10]) : 0.0;
      er1 = rok ? (ib ? d[15] : d[11]) : 0.0;
      tab = nullptr;
    } else {
      const uint32_t ai = rok ? aidx[r] : 0;
      bits = 0;
      er0 = er1 = 0.0;
      tab = A <= HMM_FAST_A ? d + 10 + ai * HMM_FAST_A
                            : reinterpret_cast<const double*>(*reinterpret_cast<const unsigned long long*>(d + 9)) + (size_t)ai * A;
    }
  }
  // `ok` guards the table reads: columns s >= P have no allele index (and may lie beyond the record)
  __device__ __forceinline__ double at(int s, bool ok) const {
    if (A <= 2) return ((bits >> s) & 1u) ? er1 : er0;
    if (!ok) return 0.0;
    return tab[aidx[s]];
  }
};

// -------------------------------------------------------------------------------------------------
// 1. basis chains.  One warp runs G = 32/CPL chains of the same transfer job: lane = chain * CPL + row.
// -------------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(BASIS_WARPS * 32, 8) basis_kernel(const ChainParams p, const ScanParams sp) {
  constexpr int G = 32 / CPL;
  __shared__ double rsm[BASIS_WARPS][2][32];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t item = blockIdx.x * BASIS_WARPS + wib;
  if (item >= sp.n_items) return;
  const uint32_t tj = item / sp.n_groups, grp = item - tj * sp.n_groups;
  const TJob job = sp.tjobs[tj];
  const int P = (int)p.P, NB = P * (P + 1) / 2;
  const int g = lane / CPL, r = lane - g * CPL;
  const int b = (int)grp * G + g;
  const bool chain_ok = g < G && b < NB;
  const bool rok = chain_ok && r < P;
  int bi = 0, bj = 0;
  {
    int q = chain_ok ? b : 0;
    while (q >= P - bi) {
      q -= P - bi;
      ++bi;
    }
    bj = bi + q;
  }
  double x[CPL];
  double myR = 0.0;
#pragma unroll
  for (int s = 0; s < CPL; ++s) {
    x[s] = (rok && ((r == bi && s == bj) || (r == bj && s == bi))) ? 1.0 : 0.0;
    myR += x[s];
  }
  int buf = 0;
  rsm[wib][0][lane] = myR;
  __syncwarp();
  const int base = (g < G ? g : 0) * CPL;
  const uint32_t cmask = P >= 32 ? 0xffffffffu : ((1u << P) - 1u);
  int E = 0;
  bool dead = false;
  int t = job.t_first;
  const int dt = job.dir ? -1 : 1;
  const int o = job.dir ? 4 : 0;
  for (uint32_t n = 0; n < job.n_steps; ++n, t += dt) {
    const double* d = reinterpret_cast<const double*>(p.desc + (size_t)(uint32_t)t * p.desc_stride);
    const double ta = d[o], tb = d[o + 1], tc = d[o + 2];
    ColEm em;
    em.load(d, r, rok);
    double R[CPL];
    double T0 = 0.0, T1 = 0.0;
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
      R[s] = rsm[wib][buf][base + s];
      if (s & 1) T1 += R[s];
      else T0 += R[s];
    }
    const double T = T0 + T1;
    const bool alive = T > 0.0;
    dead |= !alive;
    const int ex = (__double2hiint(T) >> 20) & 0x7ff;
    const double sc = alive ? __hiloint2double((2046 - ex) << 20, 0) : 0.0;
    E += alive ? ex - 1023 : 0;
    const double ca = ta * sc, cb = tb * sc, cc = tc * T * sc;
    const double rho = cb * myR + cc;
    double a0 = 0.0, a1 = 0.0;
    if (em.A <= 2) {  // (warp-uniform) biallelic column: the emission is a select on the path's allele bit
      const uint32_t okbits = rok ? cmask : 0u;
#pragma unroll
      for (int s = 0; s < CPL; ++s) {
        const double pre = fma(ca, x[s], fma(cb, R[s], rho));
        const double e = ((em.bits >> s) & 1u) ? em.er1 : em.er0;
        const double v = ((okbits >> s) & 1u) ? pre * e : 0.0;
        x[s] = v;
        if (s & 1) a1 += v;
        else a0 += v;
      }
    } else {
#pragma unroll
      for (int s = 0; s < CPL; ++s) {
        const bool ok = rok && ((cmask >> s) & 1u);
        const double pre = fma(ca, x[s], fma(cb, R[s], rho));
        const double v = ok ? pre * em.at(s, ok) : 0.0;
        x[s] = v;
        if (s & 1) a1 += v;
        else a0 += v;
      }
    }
    myR = a0 + a1;
    buf ^= 1;
    rsm[wib][buf][lane] = myR;
    __syncwarp();
  }
  if (rok) {
    double* out = sp.mats + (size_t)tj * sp.mat_stride + (size_t)b * NB;
#pragma unroll
    for (int s = 0; s < CPL; ++s)
      if (s >= r && s < P) out[tri_index(r, s, P)] = x[s];
    if (r == 0) sp.mats[(size_t)tj * sp.mat_stride + (size_t)NB * NB + b] = (double)E;  // row NB: the exponents
  }
  if (chain_ok && dead) atomicOr(sp.seq_flags + job.chrom, 1u);
}

// -------------------------------------------------------------------------------------------------
// 2. checkpoint scan.  grid = (n_chrom, 2): y = 0 forward, y = 1 backward; one upper-triangle cell per thread.
//    Checkpoints are dense row-major P x P (ChainParams::ckpt_stride), like those of skeleton_kernel.
// -------------------------------------------------------------------------------------------------
constexpr int SCAN_RING = 4;   // transfer matrices in flight (L2 -> shared memory), one cp.async group each
constexpr int SCAN_SPLIT = 4;  // lanes sharing one cell: each sums a quarter of the basis range
template <int CPL>
struct ScanSmem {
  static constexpr int NBMAX = CPL * (CPL + 1) / 2;
  static constexpr int MATMAX = ((NBMAX + 1) * NBMAX + 1) & ~1;
  double mat[SCAN_RING][MATMAX];
  double st[NBMAX];  // current state, upper triangle
  double w[NBMAX];
  int key[SCAN_THREADS / 32];
};

// One step: new_state[c] = sum_b w_b * M[b][c], w_b = state[b] * 2^(E_b - kmax).  The weights only depend on RELATIVE
// exponents and are renormalised every step (largest term in [1,2)), so the state never needs a normalisation of its own:
// its magnitude is that of the basis images, which basis_kernel keeps near 1.  Three barriers per step.
template <int CPL>
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const ChainParams p, const ScanParams sp) {
  extern __shared__ __align__(16) unsigned char scan_smem_raw[];
  ScanSmem<CPL>& sm = *reinterpret_cast<ScanSmem<CPL>*>(scan_smem_raw);
  const uint32_t chrom = blockIdx.x, dir = blockIdx.y;
  const ScanChrom sc = sp.chroms[chrom];
  if (sc.n_tj == 0) return;
  if (sp.seq_flags[chrom]) return;  // a basis chain hit a zero total: skeleton_kernel takes over
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int P = (int)p.P, NB = P * (P + 1) / 2;
  const ChromCols cc = p.chroms[chrom];
  const uint32_t tj0 = sc.tj_begin[dir];
  const uint32_t chunks = sp.mat_stride / 2;  // 16-byte pieces per matrix

  // the matrices do not depend on the chain: they stream in SCAN_RING - 1 steps ahead of their use
  auto prefetch = [&](uint32_t q) {
    if (q < sc.n_tj) {
      const double* src = sp.mats + (size_t)(tj0 + q) * sp.mat_stride;
      double* dst = sm.mat[q % SCAN_RING];
      for (uint32_t c = tid; c < chunks; c += SCAN_THREADS) cp_async16(dst + 2 * c, src + 2 * c);
    }
    cp_async_commit();
  };
  for (uint32_t q = 0; q + 1 < SCAN_RING; ++q) prefetch(q);

  // my cell (ci <= cj) and my share of the basis range
  const int cell = tid / SCAN_SPLIT, part = tid % SCAN_SPLIT;
  const bool cell_ok = cell < NB;
  int ci = 0, cj = 0;
  {
    int q = cell_ok ? cell : 0;
    while (q >= P - ci) {
      q -= P - ci;
      ++ci;
    }
    cj = ci + q;
  }
  // initial state: the chain's first column without transition (pre = 1): F = e'_{c0}, Y = e'_{c1-1}
  if (part == 0) {
    const int t = dir ? (int)cc.col_end - 1 : (int)cc.col_begin;
    const double* d = reinterpret_cast<const double*>(p.desc + (size_t)(uint32_t)t * p.desc_stride);
    ColEm em;
    em.load(d, ci, cell_ok);
    if (cell_ok) sm.st[cell] = em.at(cj, true);
  }

  for (uint32_t q = 0; q <= sc.n_tj; ++q) {
    cp_async_wait<SCAN_RING - 2>();  // all but the newest SCAN_RING - 2 groups are complete: matrix q has landed
    __syncthreads();                 // ... the state written at the end of step q - 1 is visible, and every warp is done
                                     // reading matrix q - 1, whose ring slot the next prefetch overwrites
    prefetch(q + SCAN_RING - 1);
    const double* M = sm.mat[q % SCAN_RING];
    // ---- weights: coefficient of basis b = state cell b, times the row's power-of-two scale, relative to the largest
    double coef = 0.0;
    int key = INT_MIN, Eb = 0;
    if (tid < NB) {
      coef = sm.st[tid];
      Eb = q < sc.n_tj ? (int)M[NB * NB + tid] : 0;
      if (coef > 0.0) key = (((__double2hiint(coef) >> 20) & 0x7ff) - 1023) + Eb;
    }
    int kmax = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    if (lane == 0) sm.key[wid] = kmax;
    __syncthreads();
    kmax = sm.key[0];
#pragma unroll
    for (int i = 1; i < SCAN_THREADS / 32; ++i) kmax = max(kmax, sm.key[i]);
    if (kmax == INT_MIN) {  // dead state: the reference's uniform replacement is not linear -> sequential recomputation
      if (tid == 0) atomicOr(sp.seq_flags + chrom, 1u);
      cp_async_wait<0>();
      return;
    }
    if (q == sc.n_tj) break;  // (the extra round only checks the last checkpoint)
    if (tid < NB) {
      double wv = 0.0;
      if (coef > 0.0) {
        const int sh = Eb - kmax;  // <= 0 up to the coefficient's own exponent; the product is <= 2
        wv = sh < -1000 ? 0.0 : coef * __hiloint2double((1023 + sh) << 20, 0);  // exact 2^sh; terms below 2^-1000 are noise
      }
      sm.w[tid] = wv;
    }
    __syncthreads();
    // ---- new cell value: SCAN_SPLIT lanes per cell, each over b = part, part + SCAN_SPLIT, ...
    double a0 = 0.0, a1 = 0.0;
    if (cell_ok) {
      const double* m = M + cell;
      constexpr int ITER = (ScanSmem<CPL>::NBMAX + SCAN_SPLIT - 1) / SCAN_SPLIT;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {  // static trip count: all loads of the step issue back to back
        const int b = part + i * SCAN_SPLIT;
        const double wv = b < NB ? sm.w[b] : 0.0;
        const double mv = b < NB ? m[b * NB] : 0.0;
        if (i & 1) a1 = fma(wv, mv, a1);
        else a0 = fma(wv, mv, a0);
      }
    }
    double v = a0 + a1;
#pragma unroll
    for (int o = 1; o < SCAN_SPLIT; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (cell_ok && part == 0) {
      sm.st[cell] = v;  // read again only after the next barrier; the readers of this step passed the second one
      const uint32_t out_blk = dir ? sc.out_first[1] - q : sc.out_first[0] + q;
      double* out = (dir ? p.ckpt_bwd : p.ckpt_fwd) + (size_t)out_blk * p.ckpt_stride;  // dense row-major P x P
      out[(size_t)ci * P + cj] = v;
      if (ci != cj) out[(size_t)cj * P + ci] = v;
    }
  }
  cp_async_wait<0>();
}

}  // namespace pg
