// Instantiations of the pipelined local-Kriging kernel: 1-D / 2-D coordinates, 4 block column(s).
#define GSM_PIPE_DIM3 0
#define GSM_PIPE_NBC 4
#define GSM_PIPE_ENTRY gsm_pipe_entry_d0_b4
#include "krige_local_pipe.cuh"
