// TEMPORARY stubs (replaced as the kernels land)
#include "common.cuh"
extern "C" {
int32_t nimfm_fm_sgd_begin(nimfm_ctx *ctx, nimfm_fm *fm) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_sgd_begin: not implemented yet"); }
int32_t nimfm_fm_sgd_epoch(nimfm_ctx *ctx, nimfm_fm *fm, const nimfm_dataset *X, const nimfm_sgd_cfg *cfg, int64_t *it, const int64_t *perm, int64_t nRows, double *viol, double *lossSum) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_sgd_epoch: not implemented yet"); }
int32_t nimfm_fm_sgd_end(nimfm_ctx *ctx, nimfm_fm *fm) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_sgd_end: not implemented yet"); }
int32_t nimfm_ffm_sgd_begin(nimfm_ctx *ctx, nimfm_ffm *m) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_sgd_begin: not implemented yet"); }
int32_t nimfm_ffm_sgd_epoch(nimfm_ctx *ctx, nimfm_ffm *m, const nimfm_dataset *X, const nimfm_sgd_cfg *cfg, int64_t *it, const int64_t *perm, int64_t nRows, double *viol, double *lossSum) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_sgd_epoch: not implemented yet"); }
int32_t nimfm_ffm_sgd_end(nimfm_ctx *ctx, nimfm_ffm *m) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_ffm_sgd_end: not implemented yet"); }
}
