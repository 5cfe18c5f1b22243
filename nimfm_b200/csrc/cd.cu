// TEMPORARY stubs (replaced as the kernels land)
#include "common.cuh"
extern "C" {
int32_t nimfm_fm_cd_begin(nimfm_ctx *ctx, nimfm_fm *fm, const nimfm_dataset *Xcsc, const nimfm_cd_cfg *cfg) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_cd_begin: not implemented yet"); }
int32_t nimfm_fm_cd_epoch(nimfm_ctx *ctx, nimfm_fm *fm, const nimfm_dataset *Xcsc, const nimfm_cd_cfg *cfg, double *viol, double *lossMean, double *regOverN) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_cd_epoch: not implemented yet"); }
int32_t nimfm_fm_cd_get_ypred(nimfm_ctx *ctx, nimfm_fm *fm, double *yPred) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_cd_get_ypred: not implemented yet"); }
int32_t nimfm_fm_cd_end(nimfm_ctx *ctx, nimfm_fm *fm) { return nimfm_fail(ctx, NIMFM_ERR_UNSUPPORTED, "nimfm_fm_cd_end: not implemented yet"); }
}
