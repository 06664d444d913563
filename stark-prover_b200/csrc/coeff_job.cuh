// coeff_job.cuh — device side of CoeffJob (kernels.hpp): the coefficient-space fold c'_j = c_2j + beta * c_2j+1 with the
// exact degree of the result (reference src/fri/fri_commit.rs:32-50, src/polynomial/ops.rs:19-37), run by `job.ctas` CTAs of
// a launch that otherwise hashes.  Block max -> atomic max -> the last job CTA publishes degree + 1 and re-arms the scratch.
#pragma once
#include "kernels.hpp"

namespace starkb200 {

__device__ __forceinline__ void coeff_job_run(const CoeffJob& job, unsigned cta, const FieldParams& fp) {
    int mine = 0;
    for (uint32_t j = cta * blockDim.x + threadIdx.x; j < job.out_len; j += job.ctas * blockDim.x) {
        const uint32_t e = job.c[2 * j];
        const uint32_t o = (2 * j + 1 < job.len) ? mont_mul(job.c[2 * j + 1], job.beta_m, fp) : 0u;
        const uint32_t v = fadd(o, e, fp);
        job.out[j] = v;
        if (v != 0) mine = (int)(j + 1);             // j grows along the loop: the last non-zero wins
    }
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, d));
    if ((threadIdx.x & 31) == 0 && mine) atomicMax(&s_max, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max) atomicMax(&job.scratch->maxv, s_max);
        __threadfence();
        const unsigned t = atomicAdd(&job.scratch->ticket, 1u);
        if (t == job.ctas - 1) {
            __threadfence();
            job.result->degree_plus1 = atomicExch(&job.scratch->maxv, 0);
            job.scratch->ticket = 0;
            if (job.top) {                                     // early hand-over: the degree is out before its flag
                __threadfence_system();
                *reinterpret_cast<volatile uint32_t*>(&job.top->deg_seq) = job.seq;
            }
        }
    }
}

}  // namespace starkb200
