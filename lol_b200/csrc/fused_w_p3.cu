// fused_w, translation unit 3 of 5: see fused_w_impl.cuh (LOLB_W_PART selects the kernels instantiated here)
#define LOLB_W_PART 3
#include "fused_w_impl.cuh"
