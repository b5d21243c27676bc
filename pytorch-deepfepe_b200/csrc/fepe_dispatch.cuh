// Kernel-variant selection that the library normally derives from the problem size.  The values are process-wide,
// set through the C ABI (fepe_set_dispatch, include/fepe_b200.h) -- an explicit call, not an environment variable
// read on the hot path -- so that the tests can run every path at small sizes.
#pragma once

#include "../../include/fepe_b200.h"

namespace fepe {
int dispatch_get(int which);      // fepe_fit.cu
}  // namespace fepe
