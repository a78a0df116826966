// placeholder until the ORB kernels land
#include "common.cuh"
void gb_orb_state_free(gb_ctx* ctx) { (void)ctx; }
extern "C" {
void gb_orb_cfg_default(gb_orb_cfg* c) {
  if (!c) return;
  c->nfeatures = 500; c->scale_factor = 1.2f; c->nlevels = 8; c->edge_threshold = 31; c->first_level = 0; c->wta_k = 2;
  c->score_type = 0; c->patch_size = 31; c->fast_threshold = 20;
}
int gb_orb_extract(gb_ctx* ctx, const uint8_t*, int, int, const gb_orb_cfg*, gb_keypoint*, uint8_t*, int*) { gb_set_error(ctx, "not built yet"); return GB_ERR_INVALID; }
int gb_orb_extract_to(gb_ctx* ctx, const uint8_t*, int, int, int, int, const gb_orb_cfg*, gb_features*) { gb_set_error(ctx, "not built yet"); return GB_ERR_INVALID; }
}
