/*
 * bgw_fastk.cu -- the instantiations of the specialised team-battle step kernel (bgw_fast.cuh): run-time shapes with 8- and
 * 16-bit list heads, and the compile-time shapes of BASELINE configs[4] and configs[1].
 */
#include "bgw_dev.cuh"
#include "bgw_fast.cuh"

const void *bgw_fast_step_fn(int shape, int head_elem)
{
    if (shape == 1) return (const void *)bgw_step_fast_kernel<FastStaticC5, uint8_t>;
    if (shape == 2) return (const void *)bgw_step_fast_kernel<FastStaticC2, uint8_t>;
    if (head_elem == 1) return (const void *)bgw_step_fast_kernel<FastDynamic, uint8_t>;
    return (const void *)bgw_step_fast_kernel<FastDynamic, uint16_t>;
}
