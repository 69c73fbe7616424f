// Internal (not exported) interface of csrc/conv_halo.cu, used by the conv entry points in gemm_tc.cu.
#pragma once
#include <cuda_runtime.h>

// bn = N tile (64 | 128 | 256), sub = sub-tiles per CTA, cluster = 1 (lone CTA) | 2 (CTA pair, cta_group::2)
bool scda_conv_halo_plan(int NB, int H, int W, int Cred, int Nout, bool dgrad, int *bn, int *sub, int *cluster);
// a: [NB,H,W,Cred] bf16; forward: w_krsc = [Nout][3][3][Cred]; dgrad: the FORWARD weights
// [Cred][3][3][Nout] (read as an MN-major operand with the tap mirrored).  flags as in gemm_tc.cu.
int scda_conv_halo_launch(int NB, int H, int W, int Cred, int Nout, const void *a, const void *w_krsc,
                          const float *bias, void *out, int flags, const void *mask_src, bool dgrad, int bn,
                          int sub, int cluster, cudaStream_t stream);

// stride-2 3x3 convolution (padding 1) / its data gradient; wd = the [Cout][3][3][4C] layout of scda_conv_s2_weights
int scda_conv_halo_s2_launch(int NB, int Ho, int Wo, int C, int Cout, const void *a, const void *wd,
                             const float *bias, void *out, int flags, const void *mask_src, float slope,
                             bool dgrad, cudaStream_t stream);
