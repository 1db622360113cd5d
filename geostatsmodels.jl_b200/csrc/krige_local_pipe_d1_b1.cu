// Instantiations of the pipelined local-Kriging kernel: 3-D coordinates, 1 block column(s).
#define GSM_PIPE_DIM3 1
#define GSM_PIPE_NBC 1
#define GSM_PIPE_ENTRY gsm_pipe_entry_d1_b1
#include "krige_local_pipe.cuh"
