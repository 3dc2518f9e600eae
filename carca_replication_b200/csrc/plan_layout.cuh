// Layout (float offsets) of the fp32 inference plan built by carca_eval_prepare; shared by api.cu and rows.cu.
#pragma once
#include "../../include/carca_b200.h"

namespace carca {
struct PlanLayout {
  long long tfold, mc, blocks, cross, tc_blocks, tc_cross, tq, tw, mcq, mcw, total;
};
constexpr long long kTcPacked = 2 * 18 * 64 * 4;   // floats of one packed tensor-core weight (hi | lo)
inline PlanLayout plan_layout(const carca_model_params* m) {
  PlanLayout p;
  const long long d = m->embed.d, n = m->embed.n_items;
  p.tfold = 0;
  p.mc = p.tfold + n * d;
  p.blocks = p.mc + d * 8;
  p.cross = p.blocks + (long long)m->n_blocks * 5 * d * d;
  p.tc_blocks = p.cross + (m->decoder_kind == 1 ? 3 * d * d : 0);
  const bool tc = d == 64;   // packed operands of the tcgen05 kernel (fused_eval_tc.cuh)
  const bool tcx = tc && m->decoder_kind == 1;
  p.tc_cross = p.tc_blocks + (tc ? (long long)m->n_blocks * 5 * kTcPacked : 0);
  p.tq = p.tc_cross + (tcx ? 3 * kTcPacked : 0);
  p.tw = p.tq + (tcx ? n * 64 : 0);
  p.mcq = p.tw + (tcx ? (n + 3) / 4 * 4 : 0);
  p.mcw = p.mcq + (tcx ? 64 * 8 : 0);
  p.total = p.mcw + (tcx ? 8 : 0);
  return p;
}
}  // namespace carca
