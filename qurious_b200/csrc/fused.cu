// Fused scan -> predicate -> (group-by) aggregate pipeline kernels (filled in below).
#include "plan.h"

namespace qgpu {
bool try_fused_scan_aggregate(PlanNode& agg, View* out) {
  (void)agg;
  (void)out;
  return false;
}
}  // namespace qgpu
