// A pinned host buffer set of the file-driven batch pipeline (batch.cuh): 7 arrays of `cap` floats in one allocation.
// The context keeps them between calls: pinning 2.8 GB per set costs more than reconstructing the catalog that goes
// into it.
#pragma once
#include <stdint.h>

namespace baorec {
struct HostSlot {
  float* base = nullptr;
  int64_t cap = 0;
};
}  // namespace baorec
