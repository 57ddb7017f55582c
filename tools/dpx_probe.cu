#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned a, unsigned b, unsigned *out) {
    bool ph, pl;
    unsigned r = __vibmax_s16x2(a, b, &ph, &pl);
    out[0] = r; out[1] = ph; out[2] = pl;
    out[3] = __viaddmax_s16x2(a, b, 0x8AD08AD0u);
    out[4] = __vimax3_s16x2(a, b, 0x00050005u);
    out[5] = __byte_perm(a, b, 0x5432);
    out[6] = __vmins2(a, b);
}
int main() {
    unsigned *d, h[8];
    cudaMalloc(&d, 32);
    unsigned tests[][2] = {{0x00030002u, 0x00030004u}, {0xfffe0005u, 0xffff0005u}, {0x8AD07000u, 0xfff07000u}};
    for (auto &t : tests) {
        k<<<1,1>>>(t[0], t[1], d); cudaMemcpy(h, d, 28, cudaMemcpyDeviceToHost);
        printf("a=%08x b=%08x: vibmax=%08x ph=%u pl=%u viaddmax=%08x vimax3=%08x prmt5432=%08x vimin=%08x\n", t[0], t[1], h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return 0;
}
__global__ void k2(const unsigned *in, unsigned *out) {
    unsigned a = in[threadIdx.x], b = in[threadIdx.x + 32];
    out[threadIdx.x] = __vmaxs2(a, b);
    out[threadIdx.x + 32] = __vcmpeq2(a, b);
    out[threadIdx.x + 64] = __vadd2(a, b);
    out[threadIdx.x + 96] = __vcmpges2(a, b);
}
