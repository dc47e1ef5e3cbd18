// Throughput of the warp-collective / XU instructions the mask kernel leans on: cycles per warp instruction per SM
// sub-partition with 8 warps per sub-partition resident (1024 threads per SM), dependent chains of 4 independent streams.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(uint32_t* out, int iters, long long* cyc) {
    uint32_t a = threadIdx.x * 2654435761u, b = a ^ 0x9E3779B9u, c = a + 77u, d = b + 1234567u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a = __reduce_add_sync(0xffffffffu, a); b = __reduce_add_sync(0xffffffffu, b); c = __reduce_add_sync(0xffffffffu, c); d = __reduce_add_sync(0xffffffffu, d); }
        if (OP == 1) { a = __popc(a) + i; b = __popc(b) + i; c = __popc(c) + i; d = __popc(d) + i; }
        if (OP == 2) { a = __ffs(a | 1u) + i; b = __ffs(b | 1u) + i; c = __ffs(c | 1u) + i; d = __ffs(d | 1u) + i; }
        if (OP == 3) { a = __ballot_sync(0xffffffffu, a & 1u) + i; b = __ballot_sync(0xffffffffu, b & 2u) + i; c = __ballot_sync(0xffffffffu, c & 4u) + i; d = __ballot_sync(0xffffffffu, d & 8u) + i; }
        if (OP == 4) { a = __shfl_sync(0xffffffffu, a, i & 31); b = __shfl_sync(0xffffffffu, b, (i + 1) & 31); c = __shfl_sync(0xffffffffu, c, (i + 2) & 31); d = __shfl_sync(0xffffffffu, d, (i + 3) & 31); }
        if (OP == 5) { a = (a & b) | (c << 1); b = (b ^ c) | (d >> 1); c = (c & d) ^ a; d = (d | a) & b; }
        if (OP == 6) { a = __reduce_or_sync(0xffffffffu, a); b = __reduce_or_sync(0xffffffffu, b ^ i); c = __reduce_or_sync(0xffffffffu, c + i); d = __reduce_or_sync(0xffffffffu, d); }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
    const char* nm[7] = {"redux.add", "popc", "ffs (brev+flo)", "ballot", "shfl.idx", "lop3/shf mix", "redux.or"};
    const int iters = 2000;
    for (int warps = 8; warps <= 32; warps *= 2) {
        for (int op = 0; op < 7; ++op) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k<0><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 1: k<1><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 2: k<2><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 3: k<3><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 4: k<4><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 5: k<5><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 6: k<6><<<148, warps * 32>>>(out, iters, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            // 4 ops per iteration per warp; warps / 4 warps per sub-partition
            printf("%2d warps/SM  %-16s %.2f cycles per warp instruction per sub-partition (%.1f cycles per op in one warp)\n", warps, nm[op],
                   (double)*cyc / (iters * 4.0 * (warps / 4.0)), (double)*cyc / (iters * 4.0));
        }
    }
    return 0;
}
