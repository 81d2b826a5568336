// operand / epilogue descriptors of the tensor-core GEMMs — defined once, in the public C header
#pragma once
#include "../../include/repsurf_b200.h"

// second-generation (TMA-fed) launchers, mlp_tc2.cu: 0 = launched, -1 = not eligible (use the first generation), > 0 = error
int rsb_gemm_rows2_launch(long rows, int N, const rsb_opnd_t *A, const float *Wp, const rsb_epi_t *E, cudaStream_t stream);
int rsb_gemm_wgrad2_launch(long rows, const rsb_opnd_t *G, const rsb_opnd_t *X, float *dW, int ldw, cudaStream_t stream);
