// Instantiations of the shared-factor local-Kriging kernel: 1-D / 2-D coordinates, 1 block column(s).
#define GSM_SHR_DIM3 0
#define GSM_SHR_NBC 1
#define GSM_SHR_ENTRY gsm_shr_entry_d0_b1
#include "krige_local_shr.cuh"
