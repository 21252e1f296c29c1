// fused_w, translation unit 2 of 5: see fused_w_impl.cuh (LOLB_W_PART selects the kernels instantiated here)
#define LOLB_W_PART 2
#include "fused_w_impl.cuh"
