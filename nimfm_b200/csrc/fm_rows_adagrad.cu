// fm_rows_adagrad.cu -- instantiations of fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD, KT> (one TU per mode so
// the three families compile in parallel).  KT = 8/16/32 specialisations exist for degree 2 and 3
// (the BASELINE configs); everything else runs the generic runtime-k kernel (KT = 0).
#include "fm_rows.cuh"

typedef void (*RowKernel)(const RowArgs);

template <int DEGREE, bool EXPLICIT>
static RowKernel pick_kt(int k) {
  switch (k) {
    case 8: return fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD, 8>;
    case 16: return fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD, 16>;
    case 32: return fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD, 32>;
    default: return fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD, 0>;
  }
}

RowKernel nimfm_row_kernel_adagrad(int degree, bool explicitLower, int k) {
  switch (degree) {
    case 2: return pick_kt<2, false>(k);
    case 3: return explicitLower ? pick_kt<3, true>(k) : pick_kt<3, false>(k);
    case 4: return explicitLower ? fm_rows_kernel<4, true, MODE_ADAGRAD, 0> : fm_rows_kernel<4, false, MODE_ADAGRAD, 0>;
    case 5: return explicitLower ? fm_rows_kernel<5, true, MODE_ADAGRAD, 0> : fm_rows_kernel<5, false, MODE_ADAGRAD, 0>;
    case 6: return explicitLower ? fm_rows_kernel<6, true, MODE_ADAGRAD, 0> : fm_rows_kernel<6, false, MODE_ADAGRAD, 0>;
    default: return nullptr;
  }
}
