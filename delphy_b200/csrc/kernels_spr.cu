// kernels_spr.cu -- SPR regraft study (placeholder while the kernels are being written)
#include "dphy_internal.h"
using namespace dphy;
extern "C" {
int dphy_spr_study_batch(dphy_ctx* ctx, dphy_forest*, int32_t, const dphy_spr_request*, dphy_spr_batch**) { return set_error(ctx, DPHY_ERR_INTERNAL, "spr: not built yet"); }
void dphy_spr_batch_destroy(dphy_ctx*, dphy_spr_batch*) {}
int dphy_spr_batch_get_summaries(dphy_ctx* ctx, dphy_spr_batch*, dphy_spr_summary*) { return set_error(ctx, DPHY_ERR_INTERNAL, "spr: not built yet"); }
int64_t dphy_spr_batch_total_regions(dphy_ctx*, dphy_spr_batch*) { return -1; }
int64_t dphy_spr_batch_get_regions(dphy_ctx*, dphy_spr_batch*, int32_t, dphy_candidate_region*, int64_t) { return -1; }
int dphy_spr_batch_pick_nexus_regions(dphy_ctx* ctx, dphy_spr_batch*, const double*, int32_t*) { return set_error(ctx, DPHY_ERR_INTERNAL, "spr: not built yet"); }
int dphy_spr_batch_find_region(dphy_ctx* ctx, dphy_spr_batch*, int32_t, int32_t, double, int32_t*) { return set_error(ctx, DPHY_ERR_INTERNAL, "spr: not built yet"); }
}
