// TEMPORARY stubs (replaced as the kernels land)
#include "common.cuh"
extern "C" {
int32_t nimfm_ffm_create(nimfm_ctx *ctx, int32_t nComponents, int64_t nFields, int64_t nFeatures, int32_t fitLinear, int32_t fitIntercept, nimfm_ffm **out) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_create: not implemented yet"); }
int32_t nimfm_ffm_set_params(nimfm_ctx *ctx, nimfm_ffm *m, const double *P, const double *w, double intercept) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_set_params: not implemented yet"); }
int32_t nimfm_ffm_get_params(nimfm_ctx *ctx, nimfm_ffm *m, double *P, double *w, double *intercept) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_get_params: not implemented yet"); }
int32_t nimfm_ffm_free(nimfm_ctx *ctx, nimfm_ffm *m) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_free: not implemented yet"); }
int32_t nimfm_ffm_decision_function(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_dataset *X, double *out) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_decision_function: not implemented yet"); }
int32_t nimfm_ffm_loss_grad(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_dataset *X, int32_t loss, double huberThreshold, int64_t rowBegin, int64_t nRows, const int64_t *rowIdx, int64_t miniBatchSize, int32_t zeroGrads, int32_t allreduce, double *lossSum) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_loss_grad: not implemented yet"); }
int32_t nimfm_ffm_get_grads(nimfm_ctx *ctx, nimfm_ffm *m, double *gP, double *gw, double *gb) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_get_grads: not implemented yet"); }
int32_t nimfm_ffm_adagrad_init(nimfm_ctx *ctx, nimfm_ffm *m, double eps, int32_t reset) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_adagrad_init: not implemented yet"); }
int32_t nimfm_ffm_adagrad_epoch(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_dataset *X, const nimfm_adagrad_cfg *cfg, int64_t *it, const int64_t *perm, int64_t nRows, double *viol, double *lossSum) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_adagrad_epoch: not implemented yet"); }
int32_t nimfm_ffm_adagrad_finalize(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_adagrad_cfg *cfg, int64_t it) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_adagrad_finalize: not implemented yet"); }
int32_t nimfm_ffm_time_loss_grad(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_dataset *X, int32_t loss, int64_t nRows, int64_t miniBatchSize, int32_t reps, int32_t gradToo, float *msPerLaunch) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_time_loss_grad: not implemented yet"); }
}
