// fused_w, translation unit 1 of 5: see fused_w_impl.cuh (LOLB_W_PART selects the kernels instantiated here)
#define LOLB_W_PART 1
#include "fused_w_impl.cuh"
