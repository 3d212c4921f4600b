// pg_result_layout for binaries that link the oracle but not libpangenie_b200 (test infrastructure only).
#include <algorithm>
#include "../include/pangenie_b200.h"
extern "C" int pg_result_layout(const pg_panel* p, uint64_t* offsets) {
  uint64_t off = 0;
  for (uint32_t v = 0; v < p->n_variants; ++v) {
    offsets[v] = off;
    uint64_t maxa = 0;
    for (uint32_t a = p->allele_offsets[v]; a < p->allele_offsets[v + 1]; ++a) maxa = std::max<uint64_t>(maxa, p->allele_ids[a]);
    off += (maxa + 1) * (maxa + 2) / 2;
  }
  offsets[p->n_variants] = off;
  return 0;
}
