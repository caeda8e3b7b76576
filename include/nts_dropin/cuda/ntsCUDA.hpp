// Include-path shadow of the reference's cuda/ntsCUDA.hpp: put `-I<this repo>/include/nts_dropin` (and
// `-I<this repo>/include`) BEFORE the reference root and link libnts_b200.so instead of libcuda_propagate.a.
// The whole C++ surface lives in nts_cuda_compat.hpp, implemented inline on top of the C ABI (nts_b200.h).
#pragma once
#include "cuda_type.h"
#include "nts_cuda_compat.hpp"
