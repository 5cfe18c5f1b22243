// fm_rows_adagrad.cu -- instantiations of fm_rows_kernel<DEGREE, EXPLICIT, MODE_ADAGRAD> (one TU per mode so the
// three families compile in parallel).
#include "fm_rows.cuh"

typedef void (*RowKernel)(const RowArgs);

RowKernel nimfm_row_kernel_adagrad(int degree, bool explicitLower) {
  switch (degree) {
    case 2: return fm_rows_kernel<2, false, MODE_ADAGRAD>;
    case 3: return explicitLower ? fm_rows_kernel<3, true, MODE_ADAGRAD> : fm_rows_kernel<3, false, MODE_ADAGRAD>;
    case 4: return explicitLower ? fm_rows_kernel<4, true, MODE_ADAGRAD> : fm_rows_kernel<4, false, MODE_ADAGRAD>;
    case 5: return explicitLower ? fm_rows_kernel<5, true, MODE_ADAGRAD> : fm_rows_kernel<5, false, MODE_ADAGRAD>;
    case 6: return explicitLower ? fm_rows_kernel<6, true, MODE_ADAGRAD> : fm_rows_kernel<6, false, MODE_ADAGRAD>;
    default: return nullptr;
  }
}
