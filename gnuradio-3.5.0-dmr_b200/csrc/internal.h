// Non-ABI accessors between the translation units of libgr_cuda (hidden visibility).
#pragma once
#include <cstddef>
#include "../../include/gr_cuda.h"

namespace grb {
void* mm_state_ptr(grcuda_mm* h);
size_t mm_state_bytes(grcuda_mm* h);
void* corr_state_ptr(grcuda_corr* h);
size_t corr_state_bytes(grcuda_corr* h);
}  // namespace grb
