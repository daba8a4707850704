// The GELU + gelu' (second output) instantiations of gemm_tc.cu as their own translation unit with 16 epilogue warps
// (kernel, launcher and epilogue are the code of gemm_tc.cu; see the note above gemm_tc_launch_c2g there).
#define PFN_GEMM_EPI_WARPS 16
#define PFN_GEMM_TC_C2G_TU 1
#include "gemm_tc.cu"
