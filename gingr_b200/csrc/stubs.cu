// stubs.cu -- TEMPORARY: entry points not implemented yet return GINGR_ERR_UNSUPPORTED (never a CPU result).
#include "common.cuh"
#define STUB(name, ...) int32_t name(__VA_ARGS__) { return gingr_fail(nullptr, GINGR_ERR_UNSUPPORTED, #name ": not implemented yet"); }
extern "C" {
STUB(gingr_model_upload, gingr_ctx*, int32_t, int32_t, const double*, const double*, const double*, int64_t, const double*, const int32_t*, int32_t, gingr_model**)
STUB(gingr_model_destroy, gingr_model*)
STUB(gingr_posterior_mean, gingr_ctx*, const gingr_model*, const double*, const double*, int32_t, const int32_t*, const double*, int32_t, const double*, double*, double*)
STUB(gingr_coefficients, gingr_ctx*, const gingr_model*, const double*, const double*, const double*, double*)
STUB(gingr_model_instance, gingr_ctx*, const gingr_model*, const gingr_state*, const double*, double*)
STUB(gingr_registration_create, gingr_ctx*, const gingr_model*, const gingr_target*, const gingr_config*, gingr_registration**)
STUB(gingr_registration_destroy, gingr_registration*)
STUB(gingr_registration_set_landmarks, gingr_registration*, int32_t, const int32_t*, const double*, const double*)
STUB(gingr_initialize_state, gingr_registration*, gingr_state*, const double*, double*)
STUB(gingr_update, gingr_registration*, const gingr_state*, const double*, int32_t, uint64_t, gingr_state*, double*, double*)
STUB(gingr_update_chain, gingr_registration*, int32_t)
STUB(gingr_state_download, gingr_registration*, gingr_state*, double*, double*)
}
