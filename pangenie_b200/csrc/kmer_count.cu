// K-mer counting on the device: replaces the jellyfish-backed JellyfishCounter of the reference
// (src/jellyfishcounter.hpp:46-68 COUNT/PRIME/UPDATE, src/jellyfishcounter.cpp:26-153).
//
// Data layout in HBM
//   64-byte buckets (KmerBucket, common.cuh): 4 canonical 2-bit k-mers (first base most significant, ~0 = empty),
//   their 4 u32 counts (jellyfish's counters are effectively unbounded; u32 here) and padding = one DRAM burst.
//   Bucket-linear probing from a 2-multiply mix of the k-mer, load <= 0.6: a lookup is normally ONE latency round.
//
// Text pipeline (per chunk of the read/segment file resident in HBM), all on one stream:
//   tile_lines_kernel   newline count + last-newline position per 3968-byte tile advance   (streams the text once)
//   tile_scan_kernel    exclusive scan over tiles -> line phase / header state at each tile start + carry
//   count_tile_kernel   per 4096-byte tile: 128-bit loads -> per-byte line classification (block scan) -> emitted
//                       symbols packed 2 bits each into a shared-memory stream -> k-mers by funnel-shift extraction,
//                       start positions dealt round-robin -> either probed directly (4 k-mers = 8 key loads in
//                       flight per thread, warp-aggregated atomics) or appended to the buffer of their table partition
//   probe_parts_kernel  (partitioned mode: table >> L2, resident text) works through the partition buffers in order,
//                       so the slice of the table being probed stays L2-resident
// Host text reaches the device through a ring of four 16 MiB staging buffers on a copy stream that runs ahead of the
// kernels; PRIME and UPDATE of one sample are enqueued back to back.
// Algorithmic bytes: text bytes + 16 B per k-mer (8 B key probe + RMW of the count sector).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace pg {

constexpr int CT_THREADS = 256;                       // threads per counting CTA
constexpr int CT_TILE = CT_THREADS * 16;              // bytes loaded per tile (one uint4 per thread)
constexpr int CT_HALO = 128;                          // look-ahead so k-mers may span tiles / chunks
constexpr int CT_ADV = CT_TILE - CT_HALO;             // tile advance (multiple of 16)
constexpr uint64_t CHUNK_BYTES = 64ull << 20;         // chunk of device-resident text processed per launch set (multiple of 16)
constexpr uint64_t STAGE_BYTES = 16ull << 20;         // host text is streamed through a ring of staging buffers of this size

// scalars[] slots
enum { SC_DISTINCT = 0, SC_ERROR = 1, SC_CARRY = 2, SC_KMERS = 3, SC_COUNT_SUM = 4, SC_N = 8 };
// error bits
enum { ERR_PROBE = 1, ERR_HALO = 2, ERR_FASTQ = 4 };
// FASTA line state
enum { LS_LINE_START = 0, LS_HEADER = 1, LS_SEQ = 2 };

__device__ __forceinline__ uint32_t base_code(uint32_t c) {
  // A/a=0 C/c=1 G/g=2 T/t=3, anything else 4 (jellyfish mer_dna::code)
  c &= 0xDFu;  // fold case
  return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// ------------------------------------------------------------------------------------------------
// pass 1: per-tile newline statistics over the OWNED bytes [t*ADV, min((t+1)*ADV, n))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_lines_kernel(const char* __restrict__ text, uint64_t n,
                                                          uint32_t* __restrict__ nl_count,
                                                          int64_t* __restrict__ last_nl) {
  const uint64_t base = (uint64_t)blockIdx.x * CT_ADV;
  const uint64_t end = min(base + (uint64_t)CT_ADV, n);
  uint32_t cnt = 0;
  long long last = -1;
  for (uint64_t p = base + (uint64_t)threadIdx.x * 16; p < end; p += 256 * 16) {
    if (p + 16 <= end) {
      uint4 v = *reinterpret_cast<const uint4*>(text + p);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (((w[i] >> (8 * b)) & 0xff) == '\n') {
            ++cnt;
            last = (long long)(p + 4 * i + b);
          }
      }
    } else {
      for (uint64_t q = p; q < end; ++q)
        if (text[q] == '\n') {
          ++cnt;
          last = (long long)q;
        }
    }
  }
  __shared__ uint32_t s_cnt[8];
  __shared__ long long s_last[8];
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_cnt[threadIdx.x >> 5] = cnt;
    s_last[threadIdx.x >> 5] = last;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    long long l = -1;
    for (int i = 0; i < 8; ++i) {
      c += s_cnt[i];
      l = max(l, s_last[i]);
    }
    nl_count[blockIdx.x] = c;
    last_nl[blockIdx.x] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// pass 2: one CTA scans the tile statistics -> tile_meta[t] = (line phase mod 4) | (fasta state << 2),
// and updates the carry for the next chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tile_scan_kernel(const char* __restrict__ text, uint64_t n, uint32_t n_tiles,
                                                          const uint32_t* __restrict__ nl_count,
                                                          const int64_t* __restrict__ last_nl, int is_fastq,
                                                          uint32_t* __restrict__ tile_meta,
                                                          unsigned long long* __restrict__ scalars) {
  __shared__ uint32_t s_sum[1024];
  __shared__ long long s_max[1024];
  const uint32_t carry = (uint32_t)scalars[SC_CARRY];
  const uint32_t per = (n_tiles + 1023) / 1024;
  const uint32_t t0 = threadIdx.x * per, t1 = min(t0 + per, n_tiles);
  uint32_t sum = 0;
  long long mx = -1;
  for (uint32_t t = t0; t < t1; ++t) {
    sum += nl_count[t];
    mx = max(mx, (long long)last_nl[t]);
  }
  s_sum[threadIdx.x] = sum;
  s_max[threadIdx.x] = mx;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t a = 0;
    long long b = -1;
    if ((int)threadIdx.x >= o) {
      a = s_sum[threadIdx.x - o];
      b = s_max[threadIdx.x - o];
    }
    __syncthreads();
    s_sum[threadIdx.x] += a;
    s_max[threadIdx.x] = max(s_max[threadIdx.x], b);
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? s_sum[threadIdx.x - 1] : 0;
  long long runmax = threadIdx.x ? s_max[threadIdx.x - 1] : -1;
  auto fasta_state_at = [&](long long last_before, uint64_t pos) -> uint32_t {
    // state of the line containing byte `pos`, given the last newline strictly before pos
    if (last_before >= 0) {
      if ((uint64_t)(last_before + 1) == pos) return LS_LINE_START;
      return text[last_before + 1] == '>' ? LS_HEADER : LS_SEQ;
    }
    if (carry == LS_LINE_START) return pos == 0 ? LS_LINE_START : (text[0] == '>' ? LS_HEADER : LS_SEQ);
    return carry;
  };
  for (uint32_t t = t0; t < t1; ++t) {
    const uint64_t base = (uint64_t)t * CT_ADV;
    uint32_t meta = is_fastq ? ((carry + run) & 3u) : (fasta_state_at(runmax, base) << 2);
    tile_meta[t] = meta;
    run += nl_count[t];
    runmax = max(runmax, (long long)last_nl[t]);
  }
  __syncthreads();
  if (threadIdx.x == 1023) {
    const uint32_t total = s_sum[1023];
    const long long lastall = s_max[1023];
    scalars[SC_CARRY] = is_fastq ? ((carry + total) & 3u) : fasta_state_at(lastall, n);
  }
}

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scans (512 threads = 16 warps)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exscan_add(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane < CT_THREADS / 32 ? s_warp[lane] : 0;
    uint32_t xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += y;
    }
    if (lane < CT_THREADS / 32) s_warp[lane] = xi - x;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = xi;                      // grand total
  }
  __syncthreads();
  uint32_t r = inc - v + s_warp[w];
  *total = s_warp[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ int block_exscan_max(int v, int* s_warp) {
  // exclusive prefix maximum (identity -1)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, y);
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < CT_THREADS / 32 ? s_warp[lane] : -1;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi = max(xi, y);
    }
    int ex = __shfl_up_sync(0xffffffffu, xi, 1);
    if (lane == 0) ex = -1;
    if (lane < CT_THREADS / 32) s_warp[lane] = ex;
  }
  __syncthreads();
  int prev_lane = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) prev_lane = -1;
  int r = max(prev_lane, s_warp[w]);
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------------
// table operations
// ------------------------------------------------------------------------------------------------
// Slow path of an INSERTING operation (PRIME / COUNT): `b` is the bucket to (re)examine, position by position with
// FRESH reads (the snapshot taken by the fast path may be stale after a lost CAS).  On success `slot` = 4*bucket + pos.
// Positions < j0 of the FIRST bucket are known to hold other keys (keys never change once set), so the walk starts at j0.
// Every bucket is examined through ONE fresh 2 x 128-bit snapshot (ld.volatile semantics: the fast path's snapshot may be
// stale after a lost CAS), then the first empty position is claimed; a lost CAS re-reads the same bucket behind it.
template <int OP>
__device__ __noinline__ bool resolve_insert(uint64_t kmer, uint64_t b, uint64_t& slot, KmerBucket* tab, uint64_t nb,
                                            unsigned long long* scalars, uint32_t& inserted, int j0 = 0) {
  for (uint32_t probes = 0; probes < (1u << 22); ++probes) {
    const ulonglong2 ka = __ldcv(reinterpret_cast<const ulonglong2*>(&tab[b].key[0]));
    const ulonglong2 kb = __ldcv(reinterpret_cast<const ulonglong2*>(&tab[b].key[2]));
    const unsigned long long k4[4] = {ka.x, ka.y, kb.x, kb.y};
    int pos = -1, empty = 4;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      if (j >= j0 && k4[j] == kmer) pos = j;
      if (j >= j0 && k4[j] == EMPTY_KEY) empty = j;
    }
    if (pos >= 0 && pos < empty) {
      slot = 4 * b + pos;
      return true;
    }
    if (empty < 4) {
      const unsigned long long old = atomicCAS(&tab[b].key[empty], (unsigned long long)EMPTY_KEY, (unsigned long long)kmer);
      if (old == EMPTY_KEY) ++inserted;
      if (old == EMPTY_KEY || old == kmer) {
        slot = 4 * b + empty;
        return true;
      }
      j0 = empty + 1;  // lost the position to another key: look behind it
      if (j0 < 4) continue;
    }
    b = b + 1 == nb ? 0 : b + 1;
    j0 = 0;
  }
  atomicOr(scalars + SC_ERROR, (unsigned long long)ERR_PROBE);
  return false;
}

// ------------------------------------------------------------------------------------------------
// pass 3: classify, pack, extract, probe
// ------------------------------------------------------------------------------------------------
// 16 text bytes -> 16 two-bit codes, MSB first (byte 0 in bits 31..30), and their "not a base" flags (byte 0 in bit 15).
// A/a=0 C/c=1 G/g=2 T/t=3 (jellyfish mer_dna::code); the code of a non-base is irrelevant (its flag kills the window).
__device__ __forceinline__ void pack16(const uint32_t (&w)[4], uint32_t& codes, uint32_t& notbase) {
  codes = 0;
  notbase = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t c = ((w[i] >> 1) ^ (w[i] >> 2)) & 0x03030303u;
    codes = (codes << 8) | ((c * 0x40100401u) >> 24);  // gathers the four 2-bit fields, first byte highest
    const uint32_t u = w[i] & 0xDFDFDFDFu;             // fold case
    const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    notbase = (notbase << 4) | ((((~ok) & 0x01010101u) * 0x08040201u) >> 24 & 0xfu);
  }
}

// ORs the `len` (1..32) right-aligned bits `v` into the MSB-first bit stream `dst` at bit position `bitpos`.
__device__ __forceinline__ void stream_or(uint32_t* dst, uint32_t bitpos, uint32_t v, uint32_t len) {
  const uint64_t x = ((uint64_t)v << (64u - len)) >> (bitpos & 31u);
  const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
  if (hi) atomicOr(dst + (bitpos >> 5), hi);
  if (lo) atomicOr(dst + (bitpos >> 5) + 1, lo);
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[4], uint32_t i) {  // selects, no local-memory indexing
  const uint32_t x = i < 8 ? (i < 4 ? w[0] : w[1]) : (i < 12 ? w[2] : w[3]);
  return (x >> (8 * (i & 3))) & 0xff;
}

struct TableRef {
  KmerBucket* tab;
  uint64_t nbuckets;
  uint32_t cap_q, cap_sh;
  unsigned long long* scalars;
  uint32_t flags;  // bit 0: plain atomics instead of warp-aggregated ones (experiment knob PG_COUNT_NOAGG)
};

// position of `kmer` among the four keys of a bucket snapshot, -1 if absent; branch-free (the short-circuit form compiles to
// a ladder of divergent branches: the lanes of a warp match at different positions)
__device__ __forceinline__ int match_pos(const ulonglong2& ka, const ulonglong2& kb, uint64_t kmer) {
  const uint32_t m = (ka.x == kmer ? 1u : 0u) | (ka.y == kmer ? 2u : 0u) | (kb.x == kmer ? 4u : 0u) | (kb.y == kmer ? 8u : 0u);
  return (int)__ffs(m) - 1;
}

// Lookups whose home bucket is full of other keys (~8% at load 0.6) continue in the next bucket.  Doing that inside the
// batched probe makes every warp execute the walk for a handful of lanes; instead they are parked in a shared-memory
// queue and worked off afterwards with all lanes busy (drain_walks).
constexpr uint32_t WQ_CAP = 1536;
struct WalkQueue {
  uint32_t n;
  uint32_t pad;
  unsigned long long kmer[WQ_CAP];
  unsigned long long bkt[WQ_CAP];
};

// walks from bucket b in a STATIC table; counts the k-mer if present
__device__ __noinline__ void walk_count(uint64_t kmer, uint64_t b, const TableRef& T) {
  for (uint32_t probes = 0; probes < (1u << 20); ++probes) {
    ulonglong2 ka, kb;
    ld_bucket_keys(&T.tab[b], ka, kb);
    const int pos = match_pos(ka, kb, kmer);
    if (pos >= 0) {
      atomicAdd(&T.tab[b].cnt[pos], 1u);
      return;
    }
    if (kb.y == EMPTY_KEY) return;  // positions fill left to right: an empty last position ends the search
    b = b + 1 == T.nbuckets ? 0 : b + 1;
  }
  atomicOr(T.scalars + SC_ERROR, (unsigned long long)ERR_PROBE);
}

// block-wide: resolves the parked lookups, one per thread.  Contains barriers: call from uniform control flow.
__device__ __forceinline__ void drain_walks(WalkQueue* wq, const TableRef& T) {
  __syncthreads();
  const uint32_t n = min(wq->n, WQ_CAP);
  for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) walk_count(wq->kmer[j], wq->bkt[j], T);
  __syncthreads();
  if (threadIdx.x == 0) wq->n = 0;
  __syncthreads();
}

// Looks up / inserts (OP) the canonical k-mers cn[i], i in vm, and counts the hits.  Must be called by all 32 lanes
// of a warp (warp-aggregated increments).  Eight independent 128-bit key loads are in flight per thread before any
// is examined.  UPDATE: unresolved lookups are parked in `wq` (the caller drains it).
template <int OP, bool USEQ, int N>
__device__ __forceinline__ void probeN(const uint64_t (&cn)[N], uint32_t vm, const TableRef& T, uint32_t& inserted, WalkQueue* wq) {
  uint64_t bkt[N];
  ulonglong2 ka[N], kb[N];
  KmerBucket* tab = T.tab;
  const uint64_t nbuckets = T.nbuckets;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool v = (vm >> i) & 1u;
    bkt[i] = home_slot(cn[i], T.cap_q, T.cap_sh) >> 2;
    ka[i] = kb[i] = make_ulonglong2(0, 0);
    if (v) ld_bucket_keys(&tab[bkt[i]], ka[i], kb[i]);
  }
  uint32_t hitm = 0, need = 0;  // bit i: k-mer i found, its slot is in bkt[i] / continue in bucket bkt[i] (UPDATE) or CAS pending
  unsigned long long cas_old[N];
  uint32_t cas_pos = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if ((vm >> i) & 1u) {
      // a match in the snapshot is definitive; so is a miss while the table is static (UPDATE)
      const uint64_t kmer = cn[i];
      const int pos = match_pos(ka[i], kb[i], kmer);
      const bool full = kb[i].y != EMPTY_KEY;  // positions fill left to right
      if (pos >= 0) {
        hitm |= 1u << i;
        bkt[i] = 4 * bkt[i] + pos;
      } else if (OP == PG_OP_UPDATE) {
        if (full) {
          const uint64_t nxt = bkt[i] + 1 == nbuckets ? 0 : bkt[i] + 1;
          if (USEQ) {
            const uint32_t q = atomicAdd(&wq->n, 1u);
            if (q < WQ_CAP) {
              wq->kmer[q] = kmer;
              wq->bkt[q] = nxt;
            } else {
              walk_count(kmer, nxt, T);  // queue full (pathological collision chains): resolve in place
            }
          } else {
            need |= 1u << i;
            bkt[i] = nxt;
          }
        }
      } else if (full) {
        uint64_t slot = 0;
        if (resolve_insert<OP>(kmer, bkt[i] + 1 == nbuckets ? 0 : bkt[i] + 1, slot, tab, nbuckets, T.scalars, inserted, 0)) {
          hitm |= 1u << i;
          bkt[i] = slot;
        }
      } else {
        // claim the first empty position of the snapshot (everything before it holds other keys); the CAS of all the
        // k-mers of this round are in flight together, their outcome is examined below
        const int j0 = ka[i].x == EMPTY_KEY ? 0 : ka[i].y == EMPTY_KEY ? 1 : kb[i].x == EMPTY_KEY ? 2 : 3;
        cas_old[i] = atomicCAS(&tab[bkt[i]].key[j0], (unsigned long long)EMPTY_KEY, (unsigned long long)kmer);
        cas_pos |= (uint32_t)j0 << (2 * i);
        need |= 1u << i;
      }
    }
  }
  if (OP != PG_OP_UPDATE) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if ((need >> i) & 1u) {
        const int j0 = (int)((cas_pos >> (2 * i)) & 3u);
        const uint64_t kmer = cn[i];
        if (cas_old[i] == EMPTY_KEY) {
          ++inserted;
          hitm |= 1u << i;
          bkt[i] = 4 * bkt[i] + j0;
        } else if (cas_old[i] == kmer) {  // another thread inserted the same k-mer first
          hitm |= 1u << i;
          bkt[i] = 4 * bkt[i] + j0;
        } else {  // lost the position to a different key: continue behind it with fresh reads
          uint64_t slot = 0;
          const uint64_t b2 = j0 == 3 ? (bkt[i] + 1 == nbuckets ? 0 : bkt[i] + 1) : bkt[i];
          if (resolve_insert<OP>(kmer, b2, slot, tab, nbuckets, T.scalars, inserted, j0 == 3 ? 0 : j0 + 1)) {
            hitm |= 1u << i;
            bkt[i] = slot;
          }
        }
      }
    }
    need = 0;
  }
  if (OP == PG_OP_UPDATE && !USEQ) {
    // walk on in place, again with all pending loads of the thread in flight together
    uint32_t guard = 0;
    while (need) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        if ((need >> i) & 1u) ld_bucket_keys(&tab[bkt[i]], ka[i], kb[i]);
#pragma unroll
      for (int i = 0; i < N; ++i)
        if ((need >> i) & 1u) {
          const uint64_t kmer = cn[i];
          const int pos = match_pos(ka[i], kb[i], kmer);
          if (pos >= 0) {
            hitm |= 1u << i;
            need &= ~(1u << i);
            bkt[i] = 4 * bkt[i] + pos;
          } else if (kb[i].y == EMPTY_KEY) {
            need &= ~(1u << i);
          } else {
            bkt[i] = bkt[i] + 1 == nbuckets ? 0 : bkt[i] + 1;
          }
        }
      if (++guard > (1u << 20)) {
        atomicOr(T.scalars + SC_ERROR, (unsigned long long)ERR_PROBE);
        break;
      }
    }
  }
  if (OP != PG_OP_PRIME && (T.flags & 1u)) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if ((hitm >> i) & 1u) atomicAdd(&tab[bkt[i] >> 2].cnt[bkt[i] & 3], 1u);
  } else if (OP != PG_OP_PRIME) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      // warp-aggregated increment: lanes hitting the same slot elect one leader
      const bool hit = (hitm >> i) & 1u;
      const uint64_t slot = bkt[i];
      const unsigned long long tag = hit ? (unsigned long long)slot : (~0ull - lane);
      const unsigned peers = __match_any_sync(0xffffffffu, tag);
      if (hit && (__ffs(peers) - 1) == (int)lane) atomicAdd(&tab[slot >> 2].cnt[slot & 3], (uint32_t)__popc(peers));
    }
  }
}

// One k-mer, no warp cooperation (divergent callers: partition-region overflow).
template <int OP>
__device__ __noinline__ void probe1(uint64_t kmer, const TableRef& T, uint32_t& inserted) {
  uint64_t b = home_slot(kmer, T.cap_q, T.cap_sh) >> 2;
  if (OP != PG_OP_UPDATE) {
    uint64_t slot = 0;
    if (resolve_insert<OP>(kmer, b, slot, T.tab, T.nbuckets, T.scalars, inserted) && OP != PG_OP_PRIME)
      atomicAdd(&T.tab[slot >> 2].cnt[slot & 3], 1u);
    return;
  }
  walk_count(kmer, b, T);
}

// ---- partitioned counting -------------------------------------------------------------------------------------
// A table much larger than the L2 makes every probe a random DRAM burst (64 B read + 32 B write-back for an 8-byte
// key and a 4-byte count).  With partitioning the tile kernel does not probe: it appends each canonical k-mer to the
// buffer of the table partition its home bucket lies in (streaming 8-byte writes), and probe_parts_kernel then works
// through the partitions one after the other, so the slice of the table being hit (<= ~24 MB) stays L2-resident.
// Every partition is split into PART_REPL regions, one per residue class of the tile index: the region cursors are the only
// global atomics of the scatter pass, and with one cursor per partition ALL tiles hammered the same few L2 lines (measured
// on configs[2], 230 partitions: the pass ran at 113 GB/s of text against 267 GB/s with 23 partitions).  Cursors sit
// CURSOR_STRIDE words apart (one 128-byte line each) so they spread over the L2 slices.
constexpr int MAX_PARTS = 4095;
constexpr int PART_REPL = 16;
constexpr int CURSOR_STRIDE = 32;
struct PartArgs {
  uint64_t* buf;          // [n_parts * PART_REPL][region_cap] canonical k-mers
  uint32_t* cursor;       // [n_parts * PART_REPL * CURSOR_STRIDE] k-mers appended (may run past region_cap: the excess was probed directly)
  uint32_t* work;         // [n_parts] chunks of every partition handed out so far (probe pass)
  uint32_t region_cap;
  uint32_t n_parts;
};
__device__ __forceinline__ uint32_t part_of(uint64_t kmer, uint32_t n_parts) {
  return __umulhi((uint32_t)(hash_kmer(kmer) >> 32), n_parts);  // monotone in the home bucket index (home_slot)
}

constexpr int SCATTER_SUB = 8 * CT_THREADS;  // start positions per sub-round of the scatter pass (KPT per thread)
constexpr size_t SCATTER_SMEM_MAX = (size_t)SCATTER_SUB * 12 + (size_t)(MAX_PARTS + 1) * 8;  // staging + s_cnt + s_loc at the largest partition count
static inline size_t scatter_smem(uint32_t n_parts) { return (size_t)SCATTER_SUB * 12 + (size_t)(n_parts + 1) * 8; }
constexpr int PK_WORDS = CT_TILE / 16 + 4;  // packed codes: 16 symbols per word (+ read-ahead padding)
constexpr int NB_WORDS = CT_TILE / 32 + 4;  // not-a-base flags: 32 symbols per word

// Per tile: every thread classifies its 16 bytes (which lie in a sequence line?), the emitted symbols are compacted
// into a 2-bit packed stream in shared memory, and k-mer START positions are dealt to the threads round-robin, so all
// lanes stay busy whatever fraction of the text is sequence (FASTQ: ~48%).  A k-mer is one funnel-shift extraction from
// the packed stream (no rolling warm-up), canonicalised with a bit-reversal, then probed.
template <int OP, bool SCATTER>
__global__ void __launch_bounds__(CT_THREADS, 4)
count_tile_kernel(const char* __restrict__ text, uint64_t n, uint64_t n_avail, int is_fastq,
                  const uint32_t* __restrict__ tile_meta, uint32_t k, const TableRef T, const PartArgs pa) {
  __shared__ uint32_t s_pk[PK_WORDS];
  __shared__ uint32_t s_nb[NB_WORDS];
  __shared__ uint32_t s_warp[40];
  __shared__ int s_warp_i[32];
  // scatter mode (dynamic shared memory): the per-partition counts of the tile -> write cursors; each k-mer costs ONE
  // shared-memory atomic (it returns the rank within the partition)
  extern __shared__ __align__(16) unsigned char s_dyn[];
  unsigned long long* s_stage = reinterpret_cast<unsigned long long*>(s_dyn);                      // [SCATTER_SUB] staged k-mers
  uint32_t* s_dst = reinterpret_cast<uint32_t*>(s_stage + SCATTER_SUB);                            // [SCATTER_SUB] their destinations
  uint32_t* s_cnt = s_dst + SCATTER_SUB;                                                           // [n_parts] counts, then write cursors
  uint32_t* s_loc = s_cnt + pa.n_parts;                                                            // [n_parts] offsets in the staging array
  unsigned long long* scalars = T.scalars;
  if (SCATTER)
    for (uint32_t i = threadIdx.x; i < pa.n_parts; i += CT_THREADS) s_cnt[i] = 0;

  const int tid = threadIdx.x;
  const uint64_t base = (uint64_t)blockIdx.x * CT_ADV;
  const uint64_t owned_end = min(base + (uint64_t)CT_ADV, n);
  const uint64_t pos0 = base + (uint64_t)tid * 16;
  const uint32_t meta = tile_meta[blockIdx.x];
  for (int i = tid; i < PK_WORDS; i += CT_THREADS) s_pk[i] = 0;
  for (int i = tid; i < NB_WORDS; i += CT_THREADS) s_nb[i] = 0;  // (the scans below contain the barrier)

  // ---- 128-bit load of this thread's 16 bytes; bytes past the end read as 0 and are never emitted ----
  uint32_t w[4] = {0, 0, 0, 0};
  const uint32_t nbytes = pos0 >= n_avail ? 0u : (uint32_t)min((uint64_t)16, n_avail - pos0);
  if (nbytes == 16) {
    uint4 v = *reinterpret_cast<const uint4*>(text + pos0);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else if (nbytes) {
#pragma unroll 1
    for (uint32_t i = 0; i < nbytes; ++i) {
      const uint32_t b = (uint32_t)(uint8_t)text[pos0 + i] << (8 * (i & 3));
      if (i < 4) w[0] |= b; else if (i < 8) w[1] |= b; else if (i < 12) w[2] |= b; else w[3] |= b;
    }
  }
  // newline map of my bytes (bit i set <=> byte i is '\n'), SIMD compare
  uint32_t nlbits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t m = __vcmpeq4(w[i], 0x0a0a0a0au) & 0x01010101u;  // 0/1 per byte
    nlbits |= (((m * 0x01020408u) >> 24) & 0xfu) << (4 * i);         // gather the 4 flags: byte j -> bit j
  }
  const uint32_t bytemask = nbytes >= 16 ? 0xffffu : ((1u << nbytes) - 1u);
  nlbits &= bytemask;
  const uint32_t owned_bytes = pos0 >= owned_end ? 0u : (uint32_t)min((uint64_t)16, owned_end - pos0);
  const uint32_t ownedmask = owned_bytes >= 16 ? 0xffffu : ((1u << owned_bytes) - 1u);

  // ---- which of my bytes are emitted as symbols (bit mask E, bit i = byte i) ----
  // FASTQ: byte i is emitted iff it lies in a sequence line (line index mod 4 == 1); the newline that ends the
  // sequence line is emitted too: it is not a base, so no k-mer spans records.
  // FASTA: sequence-line bytes except newlines are emitted (lines of one record are joined); the '>' that opens a
  // HEADER line is emitted as the record separator (a non-base), the rest of the header is skipped - so a window is
  // closed as soon as the next record starts, however long its header is.
  uint32_t E = 0;
  if (is_fastq) {
    uint32_t tot;
    uint32_t line = (meta & 3u) + block_exscan_add(__popc(nlbits), s_warp, &tot);
    uint32_t rest = nlbits, start = 0;
    bool bad_layout = false;
#pragma unroll 1
    while (true) {
      const uint32_t p = rest ? (uint32_t)__ffs(rest) - 1u : 15u;  // segment [start, p] lies in line `line`
      if ((line & 3u) == 1u) E |= (0xffffu >> (15u - p)) & ~((1u << start) - 1u);
      if (!rest) break;
      rest &= rest - 1u;
      ++line;
      start = p + 1u;
      // 4-line layout check (jellyfish parses '@' header / sequence / '+' / quality; a wrapped sequence or a stray blank
      // line would silently shift the phase and count quality characters): the line after an owned newline must start
      // with '@' when it is a header line and with '+' when it is a separator line
      if (((ownedmask >> p) & 1u) && pos0 + start < n_avail && !(line & 1u)) {
        const uint32_t c0 = start <= 15u ? byte_of(w, start) : (uint32_t)(uint8_t)text[pos0 + start];
        bad_layout |= c0 != ((line & 2u) ? (uint32_t)'+' : (uint32_t)'@');
      }
      if (start > 15u) break;
    }
    if (bad_layout) atomicOr(scalars + SC_ERROR, (unsigned long long)ERR_FASTQ);
  } else {
    const int my_last = nlbits ? tid * 16 + (31 - __clz(nlbits)) : -1;
    const int last_before = block_exscan_max(my_last, s_warp_i);  // tile-relative index or -1
    uint32_t state;
    if (last_before >= 0) {
      if (last_before + 1 == tid * 16) state = LS_LINE_START;
      else state = (base + last_before + 1 < n_avail && text[base + last_before + 1] == '>') ? LS_HEADER : LS_SEQ;
    } else {
      // no newline before my bytes in this tile: I am still in the line the tile starts in.  If the tile starts exactly at a
      // line start, only thread 0 sees that line's first byte; everybody else takes the state that byte decides
      state = (meta >> 2) & 3u;
      if (state == LS_LINE_START && tid > 0) state = text[base] == '>' ? LS_HEADER : LS_SEQ;
    }
    uint32_t rest = nlbits, start = 0;
#pragma unroll 1
    while (start < nbytes) {
      if (state == LS_LINE_START) {
        state = byte_of(w, start) == '>' ? LS_HEADER : LS_SEQ;
        if (state == LS_HEADER) E |= 1u << start;                    // the '>' is the separator symbol
      }
      const uint32_t p = rest ? (uint32_t)__ffs(rest) - 1u : 16u;  // newline ending this line piece (16: none)
      const uint32_t last = p < 16u ? p : 15u;
      const uint32_t seg = (0xffffu >> (15u - last)) & ~((1u << start) - 1u);
      if (state == LS_SEQ) E |= seg & ~(p < 16u ? (1u << p) : 0u);  // sequence bytes, newline dropped
      if (p >= 16u) break;
      rest &= rest - 1u;
      state = LS_LINE_START;
      start = p + 1u;
    }
  }
  E &= bytemask;
  const uint32_t n_emit = __popc(E), n_emit_owned = __popc(E & ownedmask);

  // ---- compaction: my emitted symbols go to positions [my_off, my_off + n_emit) of the packed stream ----
  uint32_t tot_packed;
  const uint32_t off_packed = block_exscan_add(n_emit | (n_emit_owned << 16), s_warp, &tot_packed);
  const uint32_t n_syms = tot_packed & 0xffffu, n_owned_syms = tot_packed >> 16;
  if (E) {
    uint32_t codes, notbase;
    pack16(w, codes, notbase);
    uint32_t pos = off_packed & 0xffffu;
    uint32_t rest = E;
#pragma unroll 1
    while (rest) {  // one iteration per run of consecutive emitted bytes (normally a single run)
      const uint32_t a = (uint32_t)__ffs(rest) - 1u;
      const uint32_t inv = ~(rest >> a);
      const uint32_t len = inv ? (uint32_t)__ffs(inv) - 1u : 32u - a;  // <= 16
      stream_or(s_pk, 2u * pos, (codes << (2u * a)) >> (32u - 2u * len), 2u * len);
      const uint32_t nbv = ((notbase << (16u + a)) >> (32u - len));
      if (nbv) stream_or(s_nb, pos, nbv, len);
      pos += len;
      rest &= ~(((len >= 32u ? 0u : (1u << len)) - 1u) << a);
    }
  }
  __syncthreads();

  // halo sufficiency: an open window at the end of the look-ahead means a k-mer may have been cut
  // (only possible with pathological whitespace); report instead of silently miscounting.
  if (tid == 0 && n_avail >= owned_end + CT_HALO && n_syms > 0) {
    const uint32_t after = n_syms - n_owned_syms;
    if (after < k - 1 && n_owned_syms > 0) {
      bool open = true;
#pragma unroll 1
      for (uint32_t q = n_owned_syms - 1; q < n_syms; ++q)
        if ((s_nb[q >> 5] >> (31u - (q & 31u))) & 1u) open = false;
      if (open) atomicOr(scalars + SC_ERROR, (unsigned long long)ERR_HALO);
    }
  }

  // ---- k-mers: start position p = tid + 256 i, owned by this tile iff p < n_owned_syms. ----
  uint32_t inserted = 0, nk = 0;
  const uint32_t kshift = 64u - 2u * k, vshift = 32u - k;
  auto kmer_at = [&](uint32_t p, uint64_t& can) -> bool {
    const uint32_t wi = p >> 4, sh = 2u * (p & 15u);
    const uint32_t w0 = s_pk[wi], w1 = s_pk[wi + 1], w2 = s_pk[wi + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    const uint64_t fwd = (((uint64_t)hi << 32) | lo) >> kshift;
    const uint32_t m0 = s_nb[p >> 5], m1 = s_nb[(p >> 5) + 1];
    const uint32_t bad = __funnelshift_l(m1, m0, p & 31u) >> vshift;
    const uint64_t rc = revcomp_2bit(fwd, k);
    can = fwd < rc ? fwd : rc;
    return p < n_owned_syms && p + k <= n_syms && bad == 0;
  };
  if (!SCATTER) {
    // four k-mers per round
#pragma unroll 1
    for (uint32_t p0 = 0; p0 < n_owned_syms; p0 += 4 * CT_THREADS) {
      uint64_t cn[4];
      uint32_t vm = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) vm |= kmer_at(p0 + (uint32_t)i * CT_THREADS + (uint32_t)tid, cn[i]) ? 1u << i : 0u;
      nk += __popc(vm);
      probeN<OP, false, 4>(cn, vm, T, inserted, nullptr);
    }
  } else {
    // Every thread takes KPT CONSECUTIVE start positions: the packed stream and the not-a-base stream of its window are
    // loaded once (7 shared-memory loads for 8 k-mers instead of 5 per k-mer), the forward k-mers are constant-shift
    // extractions from that window, the reverse complements roll (the newest base of k-mer j is its last two bits), and the
    // canonical k-mers stay in registers between ranking and appending.  Per sub-round of KPT * 256 positions:
    // (1) rank my k-mers within their partition (the shared-memory atomic returns the rank), (2) reserve one contiguous
    // range per partition in this tile's replica region, (3) append.  A region that is full (skewed data) sends its
    // k-mers straight to the table.  A FASTQ tile (about half of its bytes are sequence) is one sub-round.
    constexpr int KPT = SCATTER_SUB / CT_THREADS;
    const uint32_t repl = blockIdx.x % PART_REPL;
    const int nq = (int)((pa.n_parts + CT_THREADS - 1) / CT_THREADS);
    constexpr int QMAX = (MAX_PARTS + CT_THREADS) / CT_THREADS;
    const uint32_t rc_shift = 2u * (k - 1u);
#pragma unroll 1
    for (uint32_t sub = 0; sub < n_owned_syms; sub += KPT * CT_THREADS) {
      if (sub) {  // second sub-round (FASTA tiles): fresh counts
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < pa.n_parts; i += CT_THREADS) s_cnt[i] = 0;
        __syncthreads();
      }
      const uint32_t p0 = sub + (uint32_t)tid * KPT;   // multiple of 8
      uint64_t can[KPT];
      uint32_t info[KPT];                              // partition << 12 | rank, 0xffffffff: no k-mer
      uint32_t okm = 0;
      if (p0 < n_owned_syms) {
        const uint32_t wi = p0 >> 4, sh = 2u * (p0 & 15u);          // sh is 0 or 16
        const uint32_t w0 = s_pk[wi], w1 = s_pk[wi + 1], w2 = s_pk[wi + 2], w3 = s_pk[wi + 3];
        // 128-bit window whose top symbol is position p0
        const uint64_t a_hi = ((uint64_t)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
        const uint64_t a_lo = ((uint64_t)__funnelshift_l(w3, w2, sh) << 32) | (uint64_t)(w3 << sh);
        const uint32_t ni = p0 >> 5, nsh = p0 & 31u;               // nsh is 0, 8, 16 or 24
        const uint32_t m0 = s_nb[ni], m1 = s_nb[ni + 1], m2 = s_nb[ni + 2];
        const uint64_t nbw = ((uint64_t)__funnelshift_l(m1, m0, nsh) << 32) | __funnelshift_l(m2, m1, nsh);  // bit 63 = position p0
        uint64_t rc = 0;
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
          const uint64_t top = j == 0 ? a_hi : ((a_hi << (2 * j)) | (a_lo >> (64 - 2 * j)));
          const uint64_t fwd = top >> kshift;
          rc = j == 0 ? revcomp_2bit(fwd, k) : ((rc >> 2) | ((3ull - (fwd & 3ull)) << rc_shift));
          can[j] = fwd < rc ? fwd : rc;
          const bool bad = ((nbw << j) >> (64u - k)) != 0ull;
          const uint32_t p = p0 + (uint32_t)j;
          okm |= (p < n_owned_syms && p + k <= n_syms && !bad) ? 1u << j : 0u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < KPT; ++j) can[j] = 0;
      }
      uint32_t part[KPT];
#pragma unroll
      for (int j = 0; j < KPT; ++j) part[j] = part_of(can[j], pa.n_parts);
#pragma unroll
      for (int j = 0; j < KPT; ++j) info[j] = ((okm >> j) & 1u) ? ((part[j] << 12) | atomicAdd(&s_cnt[part[j]], 1u)) : 0xffffffffu;
      if (T.flags & 16u) okm = 0;
      nk += __popc(okm);
      __syncthreads();
      uint32_t n_staged;
      {
        // per partition: offset of its k-mers in the tile's staging array (exclusive scan of the counts) and one
        // reservation in the replica region (all atomics of a thread in flight together)
        uint32_t cntv[QMAX], basev[QMAX], mine = 0;
#pragma unroll
        for (int u = 0; u < QMAX; ++u) {
          const uint32_t q = (uint32_t)tid + (uint32_t)u * CT_THREADS;
          cntv[u] = (u < nq && q < pa.n_parts) ? s_cnt[q] : 0u;
          mine += cntv[u];
        }
        uint32_t run = block_exscan_add(mine, s_warp, &n_staged);
#pragma unroll
        for (int u = 0; u < QMAX; ++u) {
          const uint32_t q = (uint32_t)tid + (uint32_t)u * CT_THREADS;
          basev[u] = 0;
          if (cntv[u]) basev[u] = atomicAdd(pa.cursor + (size_t)(q * PART_REPL + repl) * CURSOR_STRIDE, cntv[u]);
          if (u < nq && q < pa.n_parts) s_loc[q] = run;
          run += cntv[u];
        }
#pragma unroll
        for (int u = 0; u < QMAX; ++u) {
          const uint32_t q = (uint32_t)tid + (uint32_t)u * CT_THREADS;
          if (u < nq && q < pa.n_parts) s_cnt[q] = basev[u];
        }
      }
      __syncthreads();
      // stage: the k-mers of a partition become neighbours in shared memory, each with its destination element
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        if ((okm >> j) & 1u) {
          const uint32_t pt = info[j] >> 12, rk = info[j] & 0xfffu;
          const uint32_t i = s_loc[pt] + rk, pos = s_cnt[pt] + rk;
          if (pos < pa.region_cap) {
            s_stage[i] = can[j];
            s_dst[i] = (pt * PART_REPL + repl) * pa.region_cap + pos;   // < 2^32 elements (part_setup)
          } else {
            s_dst[i] = 0xffffffffu;
            probe1<OP>(can[j], T, inserted);
          }
        }
      }
      __syncthreads();
      // append: consecutive threads write consecutive staged k-mers = runs of consecutive addresses per partition (scattered
      // 8-byte stores straight from the registers were 48% of this kernel's time: 32 sectors per warp store)
      for (uint32_t i = (uint32_t)tid; i < n_staged; i += CT_THREADS) {
        const uint32_t d = s_dst[i];
        if (d != 0xffffffffu) __stcs(reinterpret_cast<unsigned long long*>(pa.buf + d), s_stage[i]);
      }
    }
  }
  // statistics: distinct keys inserted, k-mers processed
  for (int o = 16; o > 0; o >>= 1) {
    inserted += __shfl_xor_sync(0xffffffffu, inserted, o);
    nk += __shfl_xor_sync(0xffffffffu, nk, o);
  }
  if ((tid & 31) == 0) {
    if (inserted) atomicAdd(scalars + SC_DISTINCT, (unsigned long long)inserted);
    if (nk) atomicAdd(scalars + SC_KMERS, (unsigned long long)nk);
  }
}

// Probe pass over the filled partition buffers.  ALL warps of the (co-resident) grid march through the partitions together:
// within a partition the k-mers of its PART_REPL regions are handed out in chunks of PP_CHUNK from a per-partition work
// counter, and a warp moves on to the next partition only when the current one has nothing left to hand out - so at any
// moment the whole GPU probes one or two neighbouring table slices, and the working set that has to stay in the L2 is the
// slice size, whatever the number of k-mers per pass.  (Dealing work items statically - round-robin over the grid - lets
// the persistent CTAs drift apart over the thousands of items of a pass until they span hundreds of MB of table; measured
// at configs[2] with 16 MB slices: 234 GB of DRAM reads per pass of 1.8 G k-mers, L2 hit rate 22%, i.e. every probe a burst.
// With the hand-out: 32.5 GB, hit rate 58%.)
// Warps are autonomous - no CTA barrier anywhere (the per-chunk barrier of the first hand-out version was 24% of its stall
// samples): lane 0 takes the chunk ids, one chunk ahead; the k-mers of the next round are requested before the current
// round is probed; the counts are incremented with plain RED (a __match_any aggregation per k-mer was another 18%; a
// k-mer that is hot in the read set costs its partition's warps same-address REDs, which the L2 takes at about one per clock).
constexpr int PP_CHUNK = 1024;    // k-mers per chunk (8 rounds of a warp): ~14 k hand-outs per partition sweep at configs[2]
constexpr bool PP_QUEUE = false;  // (direct kernel experiments) park overflow walks in a shared-memory queue
constexpr int PP_BATCH = 4;       // k-mers per thread per round (cfg3s UPDATE pass: 2 -> 49.5 ms, 4 -> 42.8 ms, 8 -> 50.6 ms)
constexpr int PP_ROUND = PP_BATCH * 32;
static_assert(PP_CHUNK % PP_ROUND == 0, "a chunk is a whole number of warp rounds");

template <int OP>
__global__ void __launch_bounds__(256, PP_BATCH == 2 ? 5 : PP_BATCH == 4 ? 4 : 2) probe_parts_kernel(const PartArgs pa, TableRef T) {
  const uint32_t lane = threadIdx.x & 31u;
  T.flags |= 1u;  // plain atomics
  uint32_t inserted = 0;
  for (uint32_t q = 0; q < pa.n_parts; ++q) {
    // lane r < PART_REPL: fill of region (q, r) and the first chunk id of that region within the partition
    uint32_t fill = 0;
    if (lane < PART_REPL) fill = min(pa.cursor[(size_t)(q * PART_REPL + lane) * CURSOR_STRIDE], pa.region_cap);
    const uint32_t nch = (fill + PP_CHUNK - 1) / PP_CHUNK;
    uint32_t first = nch;  // inclusive scan over the lanes, then exclusive
#pragma unroll
    for (int o = 1; o < PART_REPL; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, first, o);
      if ((int)lane >= o) first += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, first, PART_REPL - 1);
    first -= nch;
    if (total == 0) continue;
    uint32_t c = 0, c_next = 0;
    if (lane == 0) {
      c = atomicAdd(pa.work + q, 1u);
      if (c < total) c_next = atomicAdd(pa.work + q, 1u);
    }
    c = __shfl_sync(0xffffffffu, c, 0);
    while (c < total) {
      // region of chunk c: the last region whose first chunk id is <= c and that has chunks at all
      const uint32_t owners = __ballot_sync(0xffffffffu, lane < PART_REPL && nch > 0 && first <= c);
      const int r = 31 - __clz((int)owners);
      const uint32_t fr = __shfl_sync(0xffffffffu, first, r), nr = __shfl_sync(0xffffffffu, fill, r);
      const uint32_t off = (c - fr) * PP_CHUNK;
      const uint64_t* src = pa.buf + (size_t)(q * PART_REPL + (uint32_t)r) * pa.region_cap + off;
      const uint32_t m = min((uint32_t)PP_CHUNK, nr - off);
      uint64_t nx[PP_BATCH];
#pragma unroll
      for (int i = 0; i < PP_BATCH; ++i) {
        const uint32_t j = (uint32_t)i * 32u + lane;
        nx[i] = j < m ? __ldcs(reinterpret_cast<const unsigned long long*>(src + j)) : 0ull;
      }
#pragma unroll 1
      for (uint32_t rd = 0; rd < m; rd += PP_ROUND) {
        uint64_t cn[PP_BATCH];
        uint32_t vm = 0;
#pragma unroll
        for (int i = 0; i < PP_BATCH; ++i) {
          const uint32_t j = rd + (uint32_t)i * 32u + lane;
          cn[i] = nx[i];
          vm |= j < m ? 1u << i : 0u;
          const uint32_t jn = j + PP_ROUND;
          nx[i] = jn < m ? __ldcs(reinterpret_cast<const unsigned long long*>(src + jn)) : 0ull;
        }
        probeN<OP, false, PP_BATCH>(cn, vm, T, inserted, nullptr);
      }
      // the id requested one chunk ago has long landed; request the one after it
      uint32_t c_after = 0;
      if (lane == 0 && c_next < total) c_after = atomicAdd(pa.work + q, 1u);
      c = __shfl_sync(0xffffffffu, c_next, 0);
      c_next = lane == 0 ? (c < total ? c_after : c) : 0u;
    }
  }
  for (int o = 16; o > 0; o >>= 1) inserted += __shfl_xor_sync(0xffffffffu, inserted, o);
  if (lane == 0 && inserted) atomicAdd(T.scalars + SC_DISTINCT, (unsigned long long)inserted);
}

// ------------------------------------------------------------------------------------------------
// lookups, histogram, sums
// ------------------------------------------------------------------------------------------------
__global__ void lookup_kernel(const uint64_t* __restrict__ codes, uint64_t n, uint32_t k,
                              const KmerBucket* __restrict__ slots, uint64_t cap,
                              uint32_t q, uint32_t sh, uint64_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    out[i] = table_lookup(codes[i], k, slots, cap, q, sh);
  }
}

constexpr uint32_t HIST_SMEM_BINS = 10240;
__global__ void __launch_bounds__(512) histogram_kernel(const KmerBucket* __restrict__ tab, uint64_t cap,
                                                        uint64_t max_count, unsigned long long* __restrict__ bins) {
  __shared__ uint32_t s_bins[HIST_SMEM_BINS];
  const bool use_smem = max_count < HIST_SMEM_BINS;
  if (use_smem) {
    for (uint32_t i = threadIdx.x; i <= max_count; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
  }
  const uint64_t nb = cap >> 2;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 c = *reinterpret_cast<const uint4*>(tab[b].cnt);  // empty positions hold 0: skipped like zero counts (:125)
    const uint32_t v[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (v[j] > 0 && v[j] <= max_count) {
        if (use_smem) atomicAdd(&s_bins[v[j]], 1u);
        else atomicAdd(bins + v[j], 1ull);
      }
  }
  if (use_smem) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= max_count; i += blockDim.x)
      if (s_bins[i]) atomicAdd(bins + i, (unsigned long long)s_bins[i]);
  }
}

__global__ void __launch_bounds__(512) count_sum_kernel(const KmerBucket* __restrict__ tab, uint64_t cap,
                                                        unsigned long long* __restrict__ scalars) {
  unsigned long long s = 0;
  const uint64_t nb = cap >> 2;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 c = *reinterpret_cast<const uint4*>(tab[b].cnt);
    s += (unsigned long long)c.x + c.y + c.z + c.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(scalars + SC_COUNT_SUM, s);
}

__global__ void fill_slots_kernel(KmerBucket* tab, uint64_t cap) {
  const uint64_t nq = cap;  // one 16-byte quarter per iteration: 4 quarters per bucket
  ulonglong2* q = reinterpret_cast<ulonglong2*>(tab);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * blockDim.x)
    q[i] = (i & 3) < 2 ? make_ulonglong2(EMPTY_KEY, EMPTY_KEY) : make_ulonglong2(0ull, 0ull);
}

// counts of buckets [b0, b0 + nb) <-> contiguous u32 array (the cross-GPU all-reduce runs on the contiguous copy)
__global__ void export_counts_kernel(const KmerBucket* __restrict__ tab, uint64_t b0, uint64_t nb, uint32_t* __restrict__ out) {
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(out)[b] = *reinterpret_cast<const uint4*>(tab[b0 + b].cnt);
}
__global__ void import_counts_kernel(KmerBucket* __restrict__ tab, uint64_t b0, uint64_t nb, const uint32_t* __restrict__ in) {
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x)
    *reinterpret_cast<uint4*>(tab[b0 + b].cnt) = reinterpret_cast<const uint4*>(in)[b];
}

// Canonical key layout after PRIME.  Which key sits where depends on the order the CAS insertions happened to win, so
// two GPUs that PRIME the same segment file hold the same SET of occupied positions (a property of linear probing) but
// not the same array.  A "run" is a maximal sequence of consecutive full buckets plus the bucket that ends it; its keys
// occupy consecutive positions starting at the first bucket, and none of them has its home bucket before the run.
// Sorting every run by (home bucket, key) gives a layout that depends on the key set only: every key still lies at or
// after its home with full buckets in between (the i-th key of a sorted run has at least 4*home keys before it), so
// lookups are unaffected, and the count arrays of all GPUs can be added position by position (ONE all-reduce, no broadcast
// of the table).  One thread per run start; runs are a single bucket in ~85% of the cases at load <= 0.6.
__device__ __forceinline__ bool key_before(uint64_t ha, uint64_t ka, uint64_t hb, uint64_t kb) { return ha < hb || (ha == hb && ka < kb); }
__global__ void __launch_bounds__(256) canonicalize_kernel(KmerBucket* __restrict__ tab, uint64_t nb, uint32_t q, uint32_t sh,
                                                            unsigned long long* __restrict__ scalars) {
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t prev = b == 0 ? nb - 1 : b - 1;
    if (tab[prev].key[3] != EMPTY_KEY) continue;  // the previous bucket is full: b continues its run
    // length of the run in positions: consecutive full buckets, then the occupied prefix of the first non-full one
    uint64_t n = 0, e = b;
    while (true) {
      const ulonglong2 kb2 = *reinterpret_cast<const ulonglong2*>(&tab[e].key[2]);
      if (kb2.y != EMPTY_KEY) {
        n += 4;
        e = e + 1 == nb ? 0 : e + 1;
        if (n >= 4 * nb) break;  // table completely full: cannot happen below the design load
        continue;
      }
      const ulonglong2 ka2 = *reinterpret_cast<const ulonglong2*>(&tab[e].key[0]);
      n += ka2.x == EMPTY_KEY ? 0 : ka2.y == EMPTY_KEY ? 1 : kb2.x == EMPTY_KEY ? 2 : 3;
      break;
    }
    if (n < 2) continue;
    // position i of the run = position (i & 3) of bucket (b + i / 4) mod nb; home distance from b is measured modulo nb
    auto slot_ptr = [&](uint64_t i) -> unsigned long long* {
      uint64_t bb = b + (i >> 2);
      if (bb >= nb) bb -= nb;
      return &tab[bb].key[i & 3];
    };
    auto home_rel = [&](uint64_t key) -> uint64_t {
      const uint64_t h = home_slot(key, q, sh) >> 2;
      return h >= b ? h - b : h + nb - b;
    };
    // insertion sort in place (runs are short; the keys of one run are touched by this thread only)
    for (uint64_t i = 1; i < n; ++i) {
      const uint64_t key = *slot_ptr(i);
      const uint64_t hk = home_rel(key);
      uint64_t j = i;
      while (j > 0) {
        const uint64_t other = *slot_ptr(j - 1);
        if (!key_before(hk, key, home_rel(other), other)) break;
        *slot_ptr(j) = other;
        --j;
      }
      if (j != i) *slot_ptr(j) = key;
      if (4 * hk > j) atomicOr(scalars + SC_ERROR, (unsigned long long)ERR_PROBE);  // would sit before its home: impossible
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ensure_stage(pg_counter* c, int ring, bool need_host) {
  for (int i = 0; i < ring; ++i) {
    if (!c->d_stage[i]) PG_CUDA(cudaMalloc((void**)&c->d_stage[i], STAGE_BYTES + 256));
    if (!c->stage_free[i]) PG_CUDA(cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming));
    if (!c->stage_ready[i]) PG_CUDA(cudaEventCreateWithFlags(&c->stage_ready[i], cudaEventDisableTiming));
    if (need_host && i < pg_counter::NSTAGE && !c->h_stage[i]) PG_CUDA(cudaMallocHost((void**)&c->h_stage[i], STAGE_BYTES + 256));
  }
  return PG_OK;
}

static int ensure_tile_meta(pg_counter* c, size_t n_tiles) {
  if (n_tiles <= c->tile_meta_cap) return PG_OK;
  if (c->d_tile_meta) cudaFree(c->d_tile_meta);
  c->d_tile_meta = nullptr;
  // layout: [meta u32 x T][nl_count u32 x T][last_nl i64 x T]
  size_t cap = n_tiles + 1024;
  PG_CUDA(cudaMalloc((void**)&c->d_tile_meta, cap * 16));
  c->tile_meta_cap = cap;
  return PG_OK;
}

static uint64_t env_u64(const char* name, uint64_t dflt) {
  const char* e = getenv(name);
  return e && *e ? (uint64_t)strtoull(e, nullptr, 10) : dflt;
}
static TableRef table_ref(const pg_counter* c) {
  TableRef T;
  T.tab = c->slots;
  T.nbuckets = c->capacity >> 2;
  T.cap_q = c->cap_q;
  T.cap_sh = c->cap_sh;
  T.scalars = c->d_scalars;
  T.flags = (env_u64("PG_COUNT_NOAGG", 0) ? 1u : 0u) | ((uint32_t)env_u64("PG_COUNT_DEBUG", 0) & 0xf0u);   // DEBUG: timing dissection only
  return T;
}

// `pa` non-null: scatter the k-mers of this chunk into the partition buffers instead of probing the table
static int launch_chunk(pg_counter* c, const char* d_text, uint64_t n, uint64_t n_avail, int is_fastq, int op, const PartArgs* pa) {
  const uint32_t n_tiles = (uint32_t)((n + CT_ADV - 1) / CT_ADV);
  if (n_tiles == 0) return PG_OK;
  PG_TRY(ensure_tile_meta(c, n_tiles));
  uint32_t* meta = c->d_tile_meta;
  uint32_t* nlc = meta + c->tile_meta_cap;
  int64_t* lnl = reinterpret_cast<int64_t*>(meta + 2 * c->tile_meta_cap);
  tile_lines_kernel<<<n_tiles, 256, 0, c->stream>>>(d_text, n, nlc, lnl);
  tile_scan_kernel<<<1, 1024, 0, c->stream>>>(d_text, n, n_tiles, nlc, lnl, is_fastq, meta, c->d_scalars);
  const TableRef T = table_ref(c);
  PartArgs none;
  memset(&none, 0, sizeof(none));
  if (pa) {
    static bool attr_set[64] = {};   // per device
    if (c->device < 64 && !attr_set[c->device]) {
      PG_CUDA(cudaFuncSetAttribute(count_tile_kernel<PG_OP_COUNT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM_MAX));
      PG_CUDA(cudaFuncSetAttribute(count_tile_kernel<PG_OP_UPDATE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM_MAX));
      attr_set[c->device] = true;
    }
    const size_t smem = scatter_smem(pa->n_parts);
    if (op == PG_OP_COUNT) count_tile_kernel<PG_OP_COUNT, true><<<n_tiles, CT_THREADS, smem, c->stream>>>(d_text, n, n_avail, is_fastq, meta, c->k, T, *pa);
    else count_tile_kernel<PG_OP_UPDATE, true><<<n_tiles, CT_THREADS, smem, c->stream>>>(d_text, n, n_avail, is_fastq, meta, c->k, T, *pa);
  } else {
    switch (op) {
      case PG_OP_COUNT: count_tile_kernel<PG_OP_COUNT, false><<<n_tiles, CT_THREADS, 0, c->stream>>>(d_text, n, n_avail, is_fastq, meta, c->k, T, none); break;
      case PG_OP_PRIME: count_tile_kernel<PG_OP_PRIME, false><<<n_tiles, CT_THREADS, 0, c->stream>>>(d_text, n, n_avail, is_fastq, meta, c->k, T, none); break;
      default: count_tile_kernel<PG_OP_UPDATE, false><<<n_tiles, CT_THREADS, 0, c->stream>>>(d_text, n, n_avail, is_fastq, meta, c->k, T, none);
    }
  }
  count_launch(3);
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

// ---- partitioned counting: geometry and the probe pass over the filled buffers ----
// Every probe pass sweeps the whole table through the L2 once (64 B read + 32 B write-back per touched bucket), so the
// more text is scattered before a pass the fewer sweeps a file costs: the super-chunk is as large as the k-mer buffers
// allow.  Knobs: PG_COUNT_SUPER_MB caps the text per pass (test knob), PG_COUNT_PART_BUF_MB the buffer memory (default: a
// third of the free HBM, at most 24 GiB), PG_COUNT_PART_KB = table bytes per partition in KiB (0 disables partitioning),
// PG_COUNT_PART_MIN_TEXT = smallest text (bytes) that is worth partitioning.
static uint64_t part_slice_bytes() { return env_u64("PG_COUNT_PART_KB", 96u << 10) << 10; }
constexpr size_t PART_CURSOR_WORDS = (size_t)MAX_PARTS * PART_REPL * CURSOR_STRIDE;   // followed by MAX_PARTS work counters

// decides whether this pass is partitioned; sizes the buffers; `super` = text bytes per probe pass
static int part_setup(pg_counter* c, uint64_t len, int op, bool resident, bool host_text, int is_fastq, PartArgs& pa, bool& use,
                      uint64_t& super) {
  use = false;
  // `resident`: device text, or pinned host text behind a deep staging ring (the PCIe stream runs on while a probe pass
  // occupies the compute stream).  Text the host has to copy first arrives too slowly for partitioning to matter.
  if (!resident && !env_u64("PG_COUNT_PART_STAGED", 0)) return PG_OK;
  const uint64_t slice = part_slice_bytes();
  const uint64_t table_bytes = (c->capacity >> 2) * sizeof(KmerBucket);
  if (op == PG_OP_PRIME || slice == 0 || table_bytes <= 2 * slice || len < env_u64("PG_COUNT_PART_MIN_TEXT", 4u << 20)) return PG_OK;
  const uint32_t n_parts = (uint32_t)std::min<uint64_t>(MAX_PARTS, (table_bytes + slice - 1) / slice);
  const uint64_t n_regions = (uint64_t)n_parts * PART_REPL;
  // k-mers per text byte: at most 1 (FASTA); a FASTQ record spends more than half of its bytes on header and qualities
  const double density = is_fastq ? 0.5 : 1.0, slack = 1.25;
  uint64_t budget = env_u64("PG_COUNT_PART_BUF_MB", 0) << 20;
  if (!budget) {
    size_t free_b = 0, total_b = 0;
    PG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    budget = std::min<uint64_t>(24ull << 30, (uint64_t)(free_b + c->part_buf_cap * 8) / 3);
  }
  uint64_t want_super = std::min<uint64_t>(len, env_u64("PG_COUNT_SUPER_MB", 1ull << 30) << 20);
  // Text that is still crossing PCIe: the LAST probe pass runs after the last byte has arrived and overlaps with nothing, so the
  // passes are kept short - as long as a sweep of the table is cheap next to a pass (measured on configs[2], 22 GB table: 36 /
  // 15 / 8 / 4 passes -> end to end 803 / 818 / 816 / 857 ms, resident probe time 333 / 240 / 229 / 228 ms).
  if (host_text && !getenv("PG_COUNT_SUPER_MB")) want_super = std::min<uint64_t>(want_super, std::max<uint64_t>(1ull << 30, table_bytes / 16));
  uint64_t kcap = (uint64_t)((double)want_super * density * slack) + n_regions * 4096;   // k-mers the buffers should hold
  if (c->part_buf_cap >= kcap) {
    kcap = c->part_buf_cap;                       // buffers of an earlier pass are large enough: keep them
  } else {
    kcap = std::min<uint64_t>(kcap, std::max<uint64_t>(budget / 8, c->part_buf_cap));
    kcap = std::max<uint64_t>(kcap, n_regions * 8192);
  }
  kcap = std::min<uint64_t>(kcap, 0xfff00000ull);   // destinations are 32-bit element indices in the scatter kernel
  const uint64_t region = std::min<uint64_t>((kcap / n_regions) & ~(uint64_t)15, 0xfffffff0ull);
  const uint64_t need = region * n_regions;
  if (c->part_buf_cap < need) {
    if (c->d_part_buf) cudaFree(c->d_part_buf);
    c->d_part_buf = nullptr;
    c->part_buf_cap = 0;
    PG_CUDA(cudaMalloc((void**)&c->d_part_buf, need * 8));
    c->part_buf_cap = need;
  }
  if (!c->d_part_cursor) {
    PG_CUDA(cudaMalloc((void**)&c->d_part_cursor, (PART_CURSOR_WORDS + MAX_PARTS + 1) * sizeof(uint32_t)));
    PG_CUDA(cudaMemsetAsync(c->d_part_cursor, 0, (PART_CURSOR_WORDS + MAX_PARTS + 1) * sizeof(uint32_t), c->stream));
  }
  pa.buf = reinterpret_cast<uint64_t*>(c->d_part_buf);
  pa.cursor = c->d_part_cursor;
  pa.work = c->d_part_cursor + PART_CURSOR_WORDS;
  pa.region_cap = (uint32_t)region;
  pa.n_parts = n_parts;
  // text per pass the regions are sized for (the excess of an over-full region is probed directly, so an underestimate
  // of the k-mer density costs speed, not correctness)
  super = std::max<uint64_t>((uint64_t)((double)(region * n_regions) / (density * slack)), 1ull << 20);
  super = std::min<uint64_t>(super, want_super);
  use = true;
  return PG_OK;
}

static int part_flush(pg_counter* c, const PartArgs& pa, int op) {
  const TableRef T = table_ref(c);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool timed = c->n_probe < pg_counter::MAX_PROBE_EV;
  if (timed) {
    while ((int)c->ev_probe.size() < 2 * (c->n_probe + 1)) {
      cudaEvent_t ev;
      PG_CUDA(cudaEventCreate(&ev));
      c->ev_probe.push_back(ev);
    }
    PG_CUDA(cudaEventRecord(c->ev_probe[2 * c->n_probe], c->stream));
  }
  // exactly the co-resident grid: every CTA is on the machine, so they move through the partitions together
  int occ = 0;
  if (op == PG_OP_COUNT) PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe_parts_kernel<PG_OP_COUNT>, 256, 0));
  else PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe_parts_kernel<PG_OP_UPDATE>, 256, 0));
  const int grid = sms * std::max(occ, 1);
  if (op == PG_OP_COUNT) probe_parts_kernel<PG_OP_COUNT><<<grid, 256, 0, c->stream>>>(pa, T);
  else probe_parts_kernel<PG_OP_UPDATE><<<grid, 256, 0, c->stream>>>(pa, T);
  count_launch();
  PG_CUDA(cudaGetLastError());
  if (timed) {
    PG_CUDA(cudaEventRecord(c->ev_probe[2 * c->n_probe + 1], c->stream));
    ++c->n_probe;
  }
  PG_CUDA(cudaMemsetAsync(c->d_part_cursor, 0, (size_t)pa.n_parts * PART_REPL * CURSOR_STRIDE * sizeof(uint32_t), c->stream));
  PG_CUDA(cudaMemsetAsync(pa.work, 0, (size_t)pa.n_parts * sizeof(uint32_t), c->stream));
  return PG_OK;
}

// Enqueues one pass over `src` (host or device text) without waiting for it: staged copies run on copy_stream ahead of
// the kernels on c->stream.  ev0/ev1 bracket the pass on c->stream.  feed_finish() waits and checks the error flags.
// `fd` >= 0: the text is the first `len` bytes of that file, read chunk by chunk into the pinned staging ring (pread), so a
// read set of any size streams file -> pinned ring -> device without ever being held in host memory as a whole.
static int feed_enqueue(pg_counter* c, const char* src, uint64_t len, int op, cudaEvent_t ev0, cudaEvent_t ev1,
                        const std::function<int()>* after_first_chunk = nullptr, int fd = -1) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  if (op < 0 || op > 2) return fail(PG_ERR_ARG, "invalid op");
  if (!src && len && fd < 0) return fail(PG_ERR_ARG, "null text");
  cudaPointerAttributes attr;
  memset(&attr, 0, sizeof(attr));
  cudaError_t pe = (len && fd < 0) ? cudaPointerGetAttributes(&attr, src) : cudaSuccess;
  if (pe != cudaSuccess) {
    cudaGetLastError();
    attr.type = cudaMemoryTypeUnregistered;
  }
  const bool on_device = fd < 0 && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  const bool pinned = fd < 0 && attr.type == cudaMemoryTypeHost;
  int is_fastq = 0;
  if (len) {
    char first = 0;
    if (fd >= 0) {
      if (pread(fd, &first, 1, 0) != 1) return fail(PG_ERR_IO, "cannot read the sequence file");
    } else if (on_device) PG_CUDA(cudaMemcpy(&first, src, 1, cudaMemcpyDeviceToHost));
    else first = src[0];
    if (first == '@') is_fastq = 1;
    else if (first == '>') is_fastq = 0;
    else return fail(PG_ERR_FORMAT, "unsupported sequence format: file must start with '>' (FASTA) or '@' (FASTQ)");
  }
  // reset per-feed scalars: carry := 0 (FASTQ: header line phase; FASTA: LS_LINE_START), k-mers of this pass := 0
  PG_CUDA(cudaMemsetAsync(c->d_scalars + SC_CARRY, 0, 2 * sizeof(unsigned long long), c->stream));
  static_assert(SC_KMERS == SC_CARRY + 1, "one memset clears both");
  if (op != PG_OP_PRIME) c->n_probe = 0;
  PG_CUDA(cudaEventRecord(ev0, c->stream));
  const bool direct = on_device && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  // ring depth: pinned host text may run far ahead of the kernels (PG_COUNT_RING caps it: test knob)
  int ring = pg_counter::NSTAGE;
  if (pinned) ring = (int)std::max<uint64_t>(pg_counter::NSTAGE, std::min<uint64_t>(std::min<uint64_t>(pg_counter::NSTAGE_DEEP, env_u64("PG_COUNT_RING", pg_counter::NSTAGE_DEEP)),
                                                                                  (len + STAGE_BYTES - 1) / STAGE_BYTES));
  if (!direct && len) PG_TRY(ensure_stage(c, ring, !on_device && !pinned));
  if (c->stage_next >= ring) c->stage_next = 0;
  uint64_t step = direct ? CHUNK_BYTES : STAGE_BYTES;
  PartArgs pa;
  memset(&pa, 0, sizeof(pa));
  bool parted = false;
  uint64_t super = 0;
  PG_TRY(part_setup(c, len, op, direct || (pinned && ring > pg_counter::NSTAGE), !on_device, is_fastq, pa, parted, super));
  if (parted) step = std::min<uint64_t>(step, super);  // a chunk never exceeds what the regions are sized for
  if (parted) PG_CUDA(cudaMemsetAsync(c->d_part_cursor, 0, (size_t)pa.n_parts * PART_REPL * CURSOR_STRIDE * sizeof(uint32_t), c->stream));
  if (parted) PG_CUDA(cudaMemsetAsync(pa.work, 0, (size_t)pa.n_parts * sizeof(uint32_t), c->stream));
  uint64_t scattered = 0;  // text bytes scattered into the partition buffers since the last flush
  for (uint64_t off = 0; off < len; off += step) {
    const uint64_t n = std::min<uint64_t>(step, len - off);
    const uint64_t nh = std::min<uint64_t>(CT_HALO, len - off - n);
    const char* d_text;
    int buf = 0;
    if (direct) {
      d_text = src + off;
    } else {
      buf = c->stage_next;
      c->stage_next = (c->stage_next + 1) % ring;
      // wait until the kernels that last read this staging buffer are done
      PG_CUDA(cudaEventSynchronize(c->stage_free[buf]));
      if (on_device) {
        PG_CUDA(cudaMemcpyAsync(c->d_stage[buf], src + off, n + nh, cudaMemcpyDeviceToDevice, c->copy_stream));
      } else if (pinned) {
        PG_CUDA(cudaMemcpyAsync(c->d_stage[buf], src + off, n + nh, cudaMemcpyHostToDevice, c->copy_stream));
      } else {
        if (fd >= 0) {
          uint64_t got = 0;
          while (got < n + nh) {
            const ssize_t r = pread(fd, c->h_stage[buf] + got, n + nh - got, (off_t)(off + got));
            if (r <= 0) return fail(PG_ERR_IO, "short read on the sequence file");
            got += (uint64_t)r;
          }
        } else {
          memcpy(c->h_stage[buf], src + off, n + nh);
        }
        PG_CUDA(cudaMemcpyAsync(c->d_stage[buf], c->h_stage[buf], n + nh, cudaMemcpyHostToDevice, c->copy_stream));
      }
      PG_CUDA(cudaEventRecord(c->stage_ready[buf], c->copy_stream));
      PG_CUDA(cudaStreamWaitEvent(c->stream, c->stage_ready[buf], 0));
      d_text = c->d_stage[buf];
    }
    PG_TRY(launch_chunk(c, d_text, n, n + nh, is_fastq, op, parted ? &pa : nullptr));
    if (!direct) PG_CUDA(cudaEventRecord(c->stage_free[buf], c->stream));
    if (after_first_chunk && off == 0) PG_TRY((*after_first_chunk)());  // host work that overlaps the copies in flight
    if (parted) {
      scattered += n;
      if (off + step >= len || scattered + step > super) {  // the buffers are sized for `super` bytes of text
        PG_TRY(part_flush(c, pa, op));
        scattered = 0;
      }
    }
  }
  PG_CUDA(cudaEventRecord(ev1, c->stream));
  return PG_OK;
}

static int feed_finish(pg_counter* c) {
  unsigned long long sc[SC_N];
  PG_CUDA(cudaMemcpyAsync(sc, c->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
  PG_CUDA(cudaStreamSynchronize(c->stream));
  c->kmers_seen = sc[SC_KMERS];
  c->last_probe_ms = 0.0;
  for (int i = 0; i < c->n_probe; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev_probe[2 * i], c->ev_probe[2 * i + 1]) == cudaSuccess) c->last_probe_ms += ms;
  }
  if (sc[SC_ERROR] & ERR_PROBE) return fail(PG_ERR_FULL, "k-mer table full: raise hash_size (-e)");
  if (sc[SC_ERROR] & ERR_HALO) return fail(PG_ERR_FORMAT, "sequence layout not supported: more than 97 line breaks inside one k-mer");
  if (sc[SC_ERROR] & ERR_FASTQ) return fail(PG_ERR_FORMAT, "FASTQ file is not in the 4-line layout (wrapped sequence lines or blank lines): every record must be '@' header / sequence / '+' / quality");
  if (sc[SC_DISTINCT] > c->max_distinct)
    return fail(PG_ERR_FULL, "k-mer table over its design load: " + std::to_string(sc[SC_DISTINCT]) + " distinct k-mers > " + std::to_string(c->max_distinct) + "; raise hash_size (-e)");
  return PG_OK;
}

static int feed_impl(pg_counter* c, const char* src, uint64_t len, int op) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  if (len == 0) return PG_OK;
  DeviceGuard g(c->device);
  PG_TRY(feed_enqueue(c, src, len, op, c->ev_t0, c->ev_t1));
  PG_TRY(feed_finish(c));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
  c->last_feed_ms = ms;
  return PG_OK;
}

// PRIME with the segments then UPDATE with the reads (src/jellyfishcounter.cpp:61-79), enqueued back to back: the first
// read chunks cross PCIe while the PRIME kernels still run.  Timings land in last_prime_ms / last_feed_ms.
int count_prime_update(pg_counter* c, const char* segments, uint64_t segments_len, const char* reads, uint64_t reads_len,
                       const std::function<int()>* overlap) {
  pg::NvtxRange nvtx_("pg: PRIME + UPDATE");
  if (!c) return fail(PG_ERR_ARG, "null counter");
  DeviceGuard g(c->device);
  PG_TRY(feed_enqueue(c, segments, segments_len, PG_OP_PRIME, c->ev_p0, c->ev_p1));
  PG_TRY(feed_enqueue(c, reads, reads_len, PG_OP_UPDATE, c->ev_t0, c->ev_t1, overlap));
  if (overlap && reads_len == 0) PG_TRY((*overlap)());
  PG_TRY(feed_finish(c));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev_p0, c->ev_p1);
  c->last_prime_ms = ms;
  cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
  c->last_feed_ms = ms;
  return PG_OK;
}

// opens a sequence file for streaming; returns the descriptor (or -1) and its size
static int open_sized(const char* path, uint64_t& size, std::string& err) {
  const int fd = open(path, O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
    if (fd >= 0) close(fd);
    err = std::string("File ") + path + " cannot be opened.";  // check_input_file, src/commands.cpp:42-49
    return -1;
  }
  size = (uint64_t)st.st_size;
  return fd;
}

static bool ends_with(const std::string& s, const std::string& e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; }


}  // namespace pg

using namespace pg;

extern "C" pg_counter* pg_count_new(uint32_t k, uint64_t max_distinct, int device) {
  clear_error();
  if (k < 1 || k > 32) {
    fail(PG_ERR_ARG, "k must be in [1,32]");
    return nullptr;
  }
  if (check_device(device) != PG_OK) return nullptr;
  DeviceGuard g(device);
  pg_counter* c = new pg_counter();
  c->device = device;
  c->k = k;
  c->max_distinct = std::max<uint64_t>(max_distinct, 1024);
  {
    uint64_t want = (uint64_t)((double)c->max_distinct / 0.6) + 1024;
    uint32_t sh = 0;
    while ((want >> sh) >= (1ull << 31)) ++sh;  // keep q below 2^31
    if (sh < 2) sh = 2;                          // capacity multiple of 4 (vectorised scans of counts[])
    const uint64_t q = (want + (1ull << sh) - 1) >> sh;
    c->cap_q = (uint32_t)q;
    c->cap_sh = sh;
    c->capacity = q << sh;
  }
  auto bail = [&](const char* what, cudaError_t e) {
    fail(PG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    pg_count_destroy(c);
    return (pg_counter*)nullptr;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaEventCreate(&c->ev_t0)) != cudaSuccess || (e = cudaEventCreate(&c->ev_t1)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev_p0)) != cudaSuccess || (e = cudaEventCreate(&c->ev_p1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaMalloc((void**)&c->slots, (c->capacity / 4) * sizeof(KmerBucket))) != cudaSuccess) return bail("cudaMalloc(slots)", e);
  if ((e = cudaMalloc((void**)&c->d_scalars, SC_N * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc(scalars)", e);
  fill_slots_kernel<<<1184, 512, 0, c->stream>>>(c->slots, c->capacity);
  count_launch();
  cudaMemsetAsync(c->d_scalars, 0, SC_N * sizeof(unsigned long long), c->stream);
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return bail("table init", e);
  return c;
}

extern "C" void pg_count_destroy(pg_counter* c) {
  if (!c) return;
  DeviceGuard g(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->slots) cudaFree(c->slots);
  if (c->d_counts_tmp) cudaFree(c->d_counts_tmp);
  if (c->d_scalars) cudaFree(c->d_scalars);
  if (c->d_tile_meta) cudaFree(c->d_tile_meta);
  if (c->d_bins) cudaFree(c->d_bins);
  if (c->d_part_buf) cudaFree(c->d_part_buf);
  if (c->d_part_cursor) cudaFree(c->d_part_cursor);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_p0) cudaEventDestroy(c->ev_p0);
  if (c->ev_p1) cudaEventDestroy(c->ev_p1);
  for (cudaEvent_t ev : c->ev_probe) cudaEventDestroy(ev);
  for (int i = 0; i < pg_counter::NSTAGE_DEEP; ++i) {
    if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    if (i < pg_counter::NSTAGE && c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
    if (c->stage_free[i]) cudaEventDestroy(c->stage_free[i]);
    if (c->stage_ready[i]) cudaEventDestroy(c->stage_ready[i]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

extern "C" int pg_count_feed(pg_counter* c, const char* text, uint64_t len, int op) {
  pg::NvtxRange nvtx_("pg_count_feed");
  clear_error();
  return feed_impl(c, text, len, op);
}

extern "C" int pg_count_feed_device(pg_counter* c, const char* d_text, uint64_t len, int op) {
  pg::NvtxRange nvtx_("pg_count_feed_device");
  clear_error();
  return feed_impl(c, d_text, len, op);
}

extern "C" pg_counter* pg_count_create_from_buffers(const char* reads, uint64_t reads_len, const char* segments,
                                                    uint64_t segments_len, uint32_t k, uint64_t hash_size, int device) {
  clear_error();
  if (!reads) {
    fail(PG_ERR_ARG, "null reads buffer");
    return nullptr;
  }
  // PRIME/UPDATE mode: every distinct key comes from the segment file, whose byte length bounds its
  // number of k-mer windows; count-all mode: hash_size is the caller's bound (jellyfish would grow).
  const uint64_t max_distinct = segments ? std::max<uint64_t>(segments_len, 1024)
                                         : std::max<uint64_t>(hash_size, 1024);
  pg_counter* c = pg_count_new(k, max_distinct, device);
  if (!c) return nullptr;
  int st;
  if (segments) {
    st = count_prime_update(c, segments, segments_len, reads, reads_len);
  } else {
    st = feed_impl(c, reads, reads_len, PG_OP_COUNT);
  }
  if (st != PG_OK) {
    pg_count_destroy(c);
    return nullptr;
  }
  return c;
}

extern "C" pg_counter* pg_count_create(const char* reads_path, const char* segments_path, uint32_t k,
                                       uint64_t hash_size, int device) {
  clear_error();
  if (!reads_path) {
    fail(PG_ERR_ARG, "null reads path");
    return nullptr;
  }
  // check_input_file (src/commands.cpp:42-56): must exist and must not be gzip-compressed
  for (const char* p : {reads_path, segments_path}) {
    if (p && ends_with(p, ".gz")) {
      fail(PG_ERR_IO, std::string("File ") + p + " seems to be gzip-compressed. PanGenie requires an uncompressed file.");
      return nullptr;
    }
  }
  // both files are streamed: file -> ring of pinned staging buffers (pread) -> device, the copies running ahead of the
  // counting kernels; nothing is held in host memory as a whole (a 30x human read set is ~185 GB, src/commands.cpp:829-833)
  std::string err;
  uint64_t reads_len = 0, segs_len = 0;
  const int rfd = open_sized(reads_path, reads_len, err);
  const int sfd = rfd >= 0 && segments_path ? open_sized(segments_path, segs_len, err) : -1;
  if (rfd < 0 || (segments_path && sfd < 0)) {
    if (rfd >= 0) close(rfd);
    fail(PG_ERR_IO, err);
    return nullptr;
  }
  const uint64_t max_distinct = segments_path ? std::max<uint64_t>(segs_len, 1024) : std::max<uint64_t>(hash_size, 1024);
  pg_counter* c = pg_count_new(k, max_distinct, device);
  int st = c ? PG_OK : last_code();
  if (c) {
    DeviceGuard g(c->device);
    if (segments_path) st = feed_enqueue(c, nullptr, segs_len, PG_OP_PRIME, c->ev_p0, c->ev_p1, nullptr, sfd);
    if (st == PG_OK) st = feed_enqueue(c, nullptr, reads_len, segments_path ? PG_OP_UPDATE : PG_OP_COUNT, c->ev_t0, c->ev_t1, nullptr, rfd);
    if (st == PG_OK) st = feed_finish(c);
    if (st == PG_OK) {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
      c->last_feed_ms = ms;
    }
  }
  close(rfd);
  if (sfd >= 0) close(sfd);
  if (st != PG_OK) {
    if (c) {
      const std::string msg = pg_last_error();
      pg_count_destroy(c);
      fail(st, msg);
    }
    return nullptr;
  }
  return c;
}

static int lookup_codes(const pg_counter* c, const uint64_t* h_codes, uint64_t n, uint64_t* out) {
  if (!c) return fail(PG_ERR_ARG, "null counter");
  if (n == 0) return PG_OK;
  DeviceGuard g(c->device);
  uint64_t *d_codes = nullptr, *d_out = nullptr;
  PG_CUDA(cudaMalloc((void**)&d_codes, n * 8));
  cudaError_t e = cudaMalloc((void**)&d_out, n * 8);
  if (e != cudaSuccess) {
    cudaFree(d_codes);
    return fail(PG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  // read-only on the table: a private stream keeps concurrent host threads independent
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaMemcpyAsync(d_codes, h_codes, n * 8, cudaMemcpyHostToDevice, s);
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
  lookup_kernel<<<grid, 256, 0, s>>>(d_codes, n, c->k, c->slots, c->capacity, c->cap_q, c->cap_sh, d_out);
  count_launch();
  cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, s);
  e = cudaStreamSynchronize(s);
  cudaStreamDestroy(s);
  cudaFree(d_codes);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(PG_ERR_CUDA, std::string("lookup: ") + cudaGetErrorString(e));
  return PG_OK;
}

extern "C" int pg_count_lookup(const pg_counter* c, const uint64_t* kmers, uint64_t n, uint64_t* out) {
  clear_error();
  if (!c) return fail(PG_ERR_ARG, "null counter");
  const uint64_t mask = kmer_mask(c->k);
  for (uint64_t i = 0; i < n; ++i)
    if (kmers[i] & ~mask) return fail(PG_ERR_ARG, "k-mer code out of range for k");
  return lookup_codes(c, kmers, n, out);
}

extern "C" int pg_count_lookup_ascii(const pg_counter* c, const char* kmers, uint64_t n, uint64_t* out) {
  clear_error();
  if (!c) return fail(PG_ERR_ARG, "null counter");
  std::vector<uint64_t> codes(n);
  std::vector<uint8_t> bad(n, 0);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v = 0;
    bool ok = true;
    for (uint32_t j = 0; j < c->k; ++j) {
      char ch = kmers[i * c->k + j];
      uint64_t code;
      switch (ch) {
        case 'A': case 'a': code = 0; break;
        case 'C': case 'c': code = 1; break;
        case 'G': case 'g': code = 2; break;
        case 'T': case 't': code = 3; break;
        default: code = 0; ok = false;
      }
      v = (v << 2) | code;
    }
    codes[i] = ok ? v : 0;
    bad[i] = !ok;
  }
  PG_TRY(lookup_codes(c, codes.data(), n, out));
  for (uint64_t i = 0; i < n; ++i)
    if (bad[i]) out[i] = 0;  // not a DNA k-mer: absent
  return PG_OK;
}

extern "C" int pg_count_histogram(const pg_counter* c, uint64_t max_count, uint64_t* bins) {
  clear_error();
  if (!c || !bins) return fail(PG_ERR_ARG, "null argument");
  DeviceGuard g(c->device);
  pg_counter* m = const_cast<pg_counter*>(c);  // scratch only; the table is not modified
  std::lock_guard<std::mutex> lock(m->scratch_mutex);  // d_bins and the stream are shared by concurrent callers
  if (m->d_bins_cap < max_count + 1) {
    if (m->d_bins) cudaFree(m->d_bins);
    m->d_bins = nullptr;
    m->d_bins_cap = 0;
    PG_CUDA(cudaMalloc((void**)&m->d_bins, (max_count + 1) * 8));
    m->d_bins_cap = max_count + 1;
  }
  unsigned long long* d_bins = m->d_bins;
  cudaMemsetAsync(d_bins, 0, (max_count + 1) * 8, c->stream);
  histogram_kernel<<<148 * 2, 512, 0, c->stream>>>(c->slots, c->capacity, max_count, d_bins);
  count_launch();
  cudaMemcpyAsync(bins, d_bins, (max_count + 1) * 8, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return fail(PG_ERR_CUDA, std::string("histogram: ") + cudaGetErrorString(e));
  bins[0] = 0;
  return PG_OK;
}

extern "C" int pg_count_kmer_coverage(const pg_counter* c, uint64_t genome_kmers, uint64_t* out) {
  clear_error();
  if (!c || !out || genome_kmers == 0) return fail(PG_ERR_ARG, "invalid argument");
  DeviceGuard g(c->device);
  std::lock_guard<std::mutex> lock(const_cast<pg_counter*>(c)->scratch_mutex);  // SC_COUNT_SUM is shared scratch
  cudaMemsetAsync(c->d_scalars + SC_COUNT_SUM, 0, 8, c->stream);
  count_sum_kernel<<<148 * 2, 512, 0, c->stream>>>(c->slots, c->capacity, c->d_scalars);
  count_launch();
  unsigned long long s = 0;
  cudaMemcpyAsync(&s, c->d_scalars + SC_COUNT_SUM, 8, cudaMemcpyDeviceToHost, c->stream);
  PG_CUDA(cudaStreamSynchronize(c->stream));
  // the reference accumulates count/genome in long double and takes the ceiling (jellyfishcounter.cpp:106-117)
  long double r = (long double)s / (long double)genome_kmers;
  uint64_t fl = (uint64_t)r;
  *out = ((long double)fl < r) ? fl + 1 : fl;
  return PG_OK;
}

extern "C" int pg_count_compute_histogram(const pg_counter* c, uint64_t max_count, int largest_peak,
                                          const char* filename, uint64_t* peak) {
  pg::NvtxRange nvtx_("pg: histogram peak");
  clear_error();
  if (!peak) return fail(PG_ERR_ARG, "null peak");
  std::vector<uint64_t> bins(max_count + 1);
  PG_TRY(pg_count_histogram(c, max_count, bins.data()));
  if (filename && *filename) {  // Histogram::write_to_file (src/histogram.cpp:32-39)
    std::ofstream f(filename);
    if (!f.good()) return fail(PG_ERR_IO, std::string("JellyfishCounter::computeHistogram: File ") + filename + " cannot be created.");
    for (uint64_t i = 0; i <= max_count; ++i) f << i << '\t' << bins[i] << '\n';
  }
  PG_TRY(pg_histogram_peak(bins.data(), bins.size(), largest_peak, peak));
  if (filename && *filename) {  // src/jellyfishcounter.cpp:142-150
    std::ofstream f(filename, std::ios::app);
    f << "parameters\t" << *peak / 2.0 << '\t' << *peak << std::endl;
  }
  return PG_OK;
}

extern "C" uint64_t pg_count_distinct(const pg_counter* c) {
  if (!c) return 0;
  DeviceGuard g(c->device);
  unsigned long long d = 0;
  cudaMemcpy(&d, c->d_scalars + SC_DISTINCT, 8, cudaMemcpyDeviceToHost);
  return d;
}

extern "C" uint64_t pg_count_capacity(const pg_counter* c) { return c ? c->capacity : 0; }

extern "C" int pg_count_device_arrays(const pg_counter* c, uint64_t* slots_addr, uint64_t* counts_addr, uint64_t* capacity) {
  clear_error();
  if (!c) return fail(PG_ERR_ARG, "null counter");
  DeviceGuard g(c->device);
  pg_counter* m = const_cast<pg_counter*>(c);
  if (counts_addr && m->counts_tmp_cap < c->capacity) {
    if (m->d_counts_tmp) cudaFree(m->d_counts_tmp);
    m->d_counts_tmp = nullptr;
    m->counts_tmp_cap = 0;
    PG_CUDA(cudaMalloc((void**)&m->d_counts_tmp, c->capacity * sizeof(uint32_t)));
    m->counts_tmp_cap = c->capacity;
  }
  if (slots_addr) *slots_addr = (uint64_t)(uintptr_t)c->slots;
  if (counts_addr) *counts_addr = (uint64_t)(uintptr_t)m->d_counts_tmp;
  if (capacity) *capacity = c->capacity;
  return PG_OK;
}

extern "C" int pg_count_export_counts(pg_counter* c) {
  clear_error();
  if (!c || !c->d_counts_tmp || c->counts_tmp_cap < c->capacity) return fail(PG_ERR_ARG, "call pg_count_device_arrays first");
  return pg_count_export_range(c, 0, c->capacity);
}

extern "C" int pg_count_import_counts(pg_counter* c) {
  clear_error();
  if (!c || !c->d_counts_tmp || c->counts_tmp_cap < c->capacity) return fail(PG_ERR_ARG, "call pg_count_device_arrays first");
  return pg_count_import_range(c, 0, c->capacity);
}

extern "C" int pg_count_exchange_buffer(pg_counter* c, uint64_t n_slots, uint64_t* addr) {
  clear_error();
  if (!c || !addr || n_slots == 0 || (n_slots & 3)) return fail(PG_ERR_ARG, "invalid exchange buffer request");
  DeviceGuard g(c->device);
  if (c->counts_tmp_cap < n_slots) {
    if (c->d_counts_tmp) cudaFree(c->d_counts_tmp);
    c->d_counts_tmp = nullptr;
    c->counts_tmp_cap = 0;
    PG_CUDA(cudaMalloc((void**)&c->d_counts_tmp, n_slots * sizeof(uint32_t)));
    c->counts_tmp_cap = n_slots;
  }
  *addr = (uint64_t)(uintptr_t)c->d_counts_tmp;
  return PG_OK;
}

static int range_args(const pg_counter* c, uint64_t first_slot, uint64_t n_slots) {
  if (!c || !c->d_counts_tmp) return fail(PG_ERR_ARG, "call pg_count_exchange_buffer first");
  if ((first_slot & 3) || (n_slots & 3) || first_slot + n_slots > c->capacity || n_slots > c->counts_tmp_cap)
    return fail(PG_ERR_ARG, "slot range must be bucket aligned, inside the table and not larger than the exchange buffer");
  return PG_OK;
}

extern "C" int pg_count_export_range(pg_counter* c, uint64_t first_slot, uint64_t n_slots) {
  clear_error();
  PG_TRY(range_args(c, first_slot, n_slots));
  if (n_slots == 0) return PG_OK;
  DeviceGuard g(c->device);
  export_counts_kernel<<<148 * 4, 512, 0, c->stream>>>(c->slots, first_slot >> 2, n_slots >> 2, c->d_counts_tmp);
  count_launch();
  PG_CUDA(cudaStreamSynchronize(c->stream));
  return PG_OK;
}

extern "C" int pg_count_import_range(pg_counter* c, uint64_t first_slot, uint64_t n_slots) {
  clear_error();
  PG_TRY(range_args(c, first_slot, n_slots));
  if (n_slots == 0) return PG_OK;
  DeviceGuard g(c->device);
  import_counts_kernel<<<148 * 4, 512, 0, c->stream>>>(c->slots, first_slot >> 2, n_slots >> 2, c->d_counts_tmp);
  count_launch();
  PG_CUDA(cudaStreamSynchronize(c->stream));
  return PG_OK;
}

extern "C" int pg_count_canonicalize(pg_counter* c) {
  clear_error();
  if (!c) return fail(PG_ERR_ARG, "null counter");
  DeviceGuard g(c->device);
  canonicalize_kernel<<<148 * 8, 256, 0, c->stream>>>(c->slots, c->capacity >> 2, c->cap_q, c->cap_sh, c->d_scalars);
  count_launch();
  PG_CUDA(cudaGetLastError());
  unsigned long long err = 0;
  PG_CUDA(cudaMemcpyAsync(&err, c->d_scalars + SC_ERROR, 8, cudaMemcpyDeviceToHost, c->stream));
  PG_CUDA(cudaStreamSynchronize(c->stream));
  if (err & ERR_PROBE) return fail(PG_ERR_ARG, "canonicalize: the table is not a valid linear-probing layout");
  return PG_OK;
}

extern "C" uint64_t pg_count_kmers_seen(const pg_counter* c) { return c ? c->kmers_seen : 0; }
extern "C" double pg_count_last_ms(const pg_counter* c) { return c ? c->last_feed_ms : 0.0; }
extern "C" double pg_count_last_probe_ms(const pg_counter* c, uint32_t* n_passes) {
  if (n_passes) *n_passes = c ? (uint32_t)c->n_probe : 0u;
  return c ? c->last_probe_ms : 0.0;
}

extern "C" int pg_count_clear(pg_counter* c) {
  clear_error();
  if (!c) return fail(PG_ERR_ARG, "null counter");
  DeviceGuard g(c->device);
  fill_slots_kernel<<<1184, 512, 0, c->stream>>>(c->slots, c->capacity);
  count_launch();
  PG_CUDA(cudaMemsetAsync(c->d_scalars, 0, SC_N * sizeof(unsigned long long), c->stream));
  PG_CUDA(cudaStreamSynchronize(c->stream));
  c->kmers_seen = 0;
  return PG_OK;
}
