// Instantiations of the pipelined local-Kriging kernel: 3-D coordinates, 4 block column(s).
#define GSM_PIPE_DIM3 1
#define GSM_PIPE_NBC 4
#define GSM_PIPE_ENTRY gsm_pipe_entry_d1_b4
#include "krige_local_pipe.cuh"
