// Internal interface of the warp-MMA fused bottleneck (bottleneck_thin_sm100.cu, algo 1 of vsb_bottleneck_*).
#pragma once
#include "../../include/vidsitu_b200.h"

namespace vsb {
struct ThinPlan;
int thin_plan_create(const vsb_bottleneck_desc* d, ThinPlan** out_plan);
int thin_run(const ThinPlan* plan, void* stream);
void thin_plan_destroy(ThinPlan* plan);
void thin_plan_info(const ThinPlan* plan, long long* out8);
}  // namespace vsb
