// Shared host/device helpers of libpangenie_b200 (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pangenie_b200.h"

namespace pg {

int fail(int code, const std::string& msg);  // records the thread-local error string, returns code
void clear_error();
int last_code();  // status code of the last fail() on this thread

#define PG_CUDA(call)                                                                           \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return pg::fail(PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

#define PG_TRY(call)              \
  do {                            \
    int st_ = (call);             \
    if (st_ != PG_OK) return st_; \
  } while (0)

/** NVTX range around a host-side stage of the path (header-only NVTX 3: a no-op unless a profiler is attached), so that
 *  nsys / ncu timelines show count / histogram / fill / emission+HMM / fetch of every call. */
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int d) {
    cudaGetDevice(&prev);
    cudaSetDevice(d);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

/** Owning device allocation (freed on destruction); grows on demand, never shrinks. */
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  int reserve(size_t n) {
    if (n <= cap) return PG_OK;
    release();
    if (n == 0) return PG_OK;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(PG_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(n * sizeof(T)) + " B): " + cudaGetErrorString(e));
    }
    cap = n;
    return PG_OK;
  }
};

/** Verifies that `device` exists and is a Blackwell-class (sm_100+) part. */
int check_device(int device);

/** Counts kernel launches per host thread so pg_timings.kernel_launches is a measured number. */
extern thread_local uint64_t g_launches;
inline void count_launch(uint64_t n = 1) { g_launches += n; }

// ---- 2-bit k-mer arithmetic shared by the counting and the lookup kernels -------------------------
__host__ __device__ inline uint64_t kmer_mask(uint32_t k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL); }

__device__ __forceinline__ uint64_t revcomp_2bit(uint64_t x, uint32_t k) {
  // complement every base (3 - c == ~c & 3), reverse the order of the 2-bit groups
  uint64_t y = __brevll(~x);
  y = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
  return y >> (64 - 2 * k);
}

__device__ __forceinline__ uint64_t hash_kmer(uint64_t h) {
  h *= 0x9E3779B97F4A7C15ULL;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ULL;
  h ^= h >> 32;
  return h;
}

/** Home slot in a table of capacity q << sh (q < 2^32): high hash word scaled by q, low hash bits below. */
__device__ __forceinline__ uint64_t home_slot(uint64_t kmer, uint32_t q, uint32_t sh) {
  const uint64_t h = hash_kmer(kmer);
  const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
  return ((uint64_t)__umulhi(hi, q) << sh) | (uint64_t)(lo & ((1u << sh) - 1u));
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

constexpr uint64_t EMPTY_KEY = ~0ULL;  // never a canonical k-mer for k <= 32 (all-T canonicalises to all-A)

// The table is an array of 64-byte buckets = one DRAM burst: 4 keys (32 B, two 128-bit loads), their 4 counts
// (16 B) and 16 B of padding.  A key lives in the first free position of its home bucket, else of the next bucket
// (bucket-linear probing); positions fill left to right and nothing is ever deleted, so an EMPTY position ends a
// search.  "slot" s denotes position s & 3 of bucket s >> 2.
struct __align__(64) KmerBucket {
  unsigned long long key[4];
  uint32_t cnt[4];
  uint32_t pad[4];
};

/** The four keys of a bucket with ONE 256-bit load (LDG.E.256, new with sm_100: the key sector is 32 bytes, so a probe is a
 *  single request to the L2 instead of two 128-bit ones). */
__device__ __forceinline__ void ld_bucket_keys(const KmerBucket* b, ulonglong2& ka, ulonglong2& kb) {
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(ka.x), "=l"(ka.y), "=l"(kb.x), "=l"(kb.y) : "l"(b->key));
}

/** canonicalise + look up; 0 if absent (getKmerAbundance, reference src/jellyfishcounter.cpp:87-104) */
__device__ __forceinline__ uint32_t table_lookup(uint64_t code, uint32_t k, const KmerBucket* __restrict__ tab, uint64_t cap,
                                                 uint32_t q, uint32_t sh) {
  const uint64_t rc = revcomp_2bit(code, k);
  const uint64_t can = code < rc ? code : rc;
  uint64_t b = home_slot(can, q, sh) >> 2;
  const uint64_t nb = cap >> 2;
  for (uint32_t probes = 0; probes < (1u << 20); ++probes) {
    ulonglong2 k01, k23;
    ld_bucket_keys(&tab[b], k01, k23);
    if (k01.x == can) return tab[b].cnt[0];
    if (k01.x == EMPTY_KEY) return 0;
    if (k01.y == can) return tab[b].cnt[1];
    if (k01.y == EMPTY_KEY) return 0;
    if (k23.x == can) return tab[b].cnt[2];
    if (k23.x == EMPTY_KEY) return 0;
    if (k23.y == can) return tab[b].cnt[3];
    if (k23.y == EMPTY_KEY) return 0;
    b = b + 1 == nb ? 0 : b + 1;
  }
  return 0;
}

}  // namespace pg

struct pg_counter;
namespace pg {
/** PRIME (segments) + UPDATE (reads) enqueued back to back, one wait at the end (kmer_count.cu).  `overlap`, if given,
 *  is host work executed once the first read chunk is on its way (e.g. the panel upload of pg_genotype_run). */
int count_prime_update(pg_counter* c, const char* segments, uint64_t segments_len, const char* reads, uint64_t reads_len,
                       const std::function<int()>* overlap = nullptr);
}

/** Device k-mer table + streaming state (definition shared by kmer_count.cu and pipeline.cu). */
struct pg_counter {
  int device = 0;
  uint32_t k = 0;
  uint64_t capacity = 0;      // slots = cap_q << cap_sh
  uint32_t cap_q = 0, cap_sh = 0;
  uint64_t max_distinct = 0;  // keys the caller asked room for
  // [capacity / 4] 64-byte buckets (KmerBucket): a lookup normally costs ONE DRAM burst and one latency round
  pg::KmerBucket* slots = nullptr;
  uint32_t* d_counts_tmp = nullptr;  // contiguous copy of (a range of) the counts for the cross-GPU all-reduce
  uint64_t counts_tmp_cap = 0;       // in slots
  std::mutex scratch_mutex;          // histogram bins / count-sum scratch are shared by concurrent read-only callers
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // streaming scratch
  // staging ring: copies run ahead of the counting kernels.  NSTAGE buffers for sources the host has to touch (pageable
  // memory, files: the host copy is the bottleneck anyway); up to NSTAGE_DEEP for pinned host text, so the PCIe stream keeps
  // running while a probe pass of the partitioned mode occupies the compute stream (buffers are allocated on first use)
  static constexpr int NSTAGE = 4;
  static constexpr int NSTAGE_DEEP = 192;           // x 16 MiB = 3 GiB of device staging at most
  char* d_stage[NSTAGE_DEEP] = {};
  char* h_stage[NSTAGE] = {};
  cudaEvent_t stage_free[NSTAGE_DEEP] = {};   // kernel finished reading d_stage[i]
  cudaEvent_t stage_ready[NSTAGE_DEEP] = {};  // H2D into d_stage[i] finished
  int stage_next = 0;                    // ring position (persists across feeds so consecutive feeds overlap)
  uint32_t* d_tile_meta = nullptr;                  // per-tile scan scratch
  size_t tile_meta_cap = 0;
  unsigned long long* d_scalars = nullptr;          // [0]=distinct, [1]=error flags, [2]=carry state, [3]=kmers seen
  uint64_t kmers_seen = 0;
  double last_feed_ms = 0.0;
  // kept across calls: steady-state calls make no cudaMalloc / cudaFree / event creation (those serialise on the
  // driver's resource-manager lock, e.g. behind a concurrent NVML query)
  unsigned long long* d_bins = nullptr;  // histogram bins
  size_t d_bins_cap = 0;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // timing events of the last feed / histogram
  // partitioned counting (kmer_count.cu): k-mer buffers of the table partitions, their cursors + the work counter
  unsigned long long* d_part_buf = nullptr;
  size_t part_buf_cap = 0;  // in k-mers
  uint32_t* d_part_cursor = nullptr;  // [256]: cursors [0..255), work counter at [255]
  cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;  // timing events of a PRIME pass enqueued together with its UPDATE pass
  double last_prime_ms = 0.0;
  // probe passes of the last partitioned feed: event pairs around every probe_parts_kernel launch
  static constexpr int MAX_PROBE_EV = 256;
  std::vector<cudaEvent_t> ev_probe;  // 2 per pass, created on demand, reused
  int n_probe = 0;
  double last_probe_ms = 0.0;
};
