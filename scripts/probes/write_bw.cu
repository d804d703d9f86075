// Probe: achievable write-only HBM bandwidth on this GPU (the RoIAlign output is a pure 822 MB write stream).
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void wr(float4 *p, size_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) p[i] = v;
        else if (MODE == 1) __stcs(p + i, v);
        else __stcg(p + i, v);
    }
}
__global__ void wr1(float *p, size_t n) {   // 4-byte coalesced stores (128 B per warp instruction)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, 1.f);
}
__global__ void cp(const float4 *a, float4 *b, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
    const size_t bytes = 822083584;   // [1024,1024,14,14] fp32
    float4 *p, *q;
    cudaMalloc(&p, bytes); cudaMalloc(&q, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 5; ++mode)
        for (int blocks : {148 * 4, 148 * 16, 148 * 64}) {
            float best = 1e9;
            for (int it = 0; it < 6; ++it) {
                cudaEventRecord(a);
                if (mode == 0) wr<0><<<blocks, 256>>>(p, bytes / 16);
                if (mode == 1) wr<1><<<blocks, 256>>>(p, bytes / 16);
                if (mode == 2) wr<2><<<blocks, 256>>>(p, bytes / 16);
                if (mode == 3) wr1<<<blocks, 256>>>((float *)p, bytes / 4);
                if (mode == 4) cp<<<blocks, 256>>>(p, q, bytes / 16);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (it > 0 && ms < best) best = ms;
            }
            printf("mode %d (%s) blocks %5d: %.1f us  %.0f GB/s%s\n", mode,
                   mode == 0 ? "st.v4" : mode == 1 ? "st.cs.v4" : mode == 2 ? "st.cg.v4" : mode == 3 ? "st.cs.f32" : "copy v4", blocks, best * 1e3,
                   (mode == 4 ? 2.0 : 1.0) * bytes / best / 1e6, mode == 4 ? " (read+write)" : "");
        }
    return 0;
}
