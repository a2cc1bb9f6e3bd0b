// common.cuh -- shared helpers for the c3poa_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define C3_WARP 32
#define C3_FULL 0xffffffffu

// base codes: A,C,G,T -> 0..3, anything else -> 4
__device__ __forceinline__ uint8_t c3_encode_base(uint8_t c)
{
    uint8_t u = c | 0x20;
    return u == 'a' ? 0 : u == 'c' ? 1 : u == 'g' ? 2 : u == 't' ? 3 : 4;
}

// ASCII -> codes, grid-stride, 16 bytes per thread per step
__global__ void c3_encode_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int64_t n)
{
    int64_t n16 = n >> 4;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint4 *in4 = reinterpret_cast<const uint4 *>(in);
    uint4 *out4 = reinterpret_cast<uint4 *>(out);
    for (int64_t i = tid; i < n16; i += stride) {
        uint4 v = in4[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = w[k], r = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) r |= (uint32_t)c3_encode_base((x >> (8 * b)) & 0xff) << (8 * b);
            w[k] = r;
        }
        out4[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int64_t i = (n16 << 4) + tid; i < n; i += stride) out[i] = c3_encode_base(in[i]);
}

__device__ __forceinline__ int c3_lane() { return threadIdx.x & 31; }
