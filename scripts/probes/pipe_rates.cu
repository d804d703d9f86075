// Developer probe: issue / pipe throughput of the instructions the LSM epilogue is made of, on one SM, with 1, 2 and 4 warps per
// scheduler (4, 8, 16 warps per CTA).  Prints cycles per warp-instruction per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 256
#define UNROLL 16

template <int OP>
__global__ void k(float *out, long long *cyc, float seed) {
    __shared__ float sm[4096];
    float a[UNROLL], b[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; b[i] = seed * 0.5f + i; }
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }                                   // MUFU.EX2, independent chains
            if (OP == 1) { a[i] = fmaf(a[i], 1.0001f, b[i]); }                                                            // FFMA
            if (OP == 2) {                                                                                                // FFMA2 (packed)
                float2 x = make_float2(a[i], b[i]);
                x = __ffma2_rn(x, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
                a[i] = x.x; b[i] = x.y;
            }
            if (OP == 3) { float2 x = __fadd2_rn(make_float2(a[i], b[i]), make_float2(0.5f, 0.25f)); a[i] = x.x; b[i] = x.y; }   // FADD2
            if (OP == 4) { a[i] = fmaxf(fmaxf(a[i], b[i]), a[(i + 1) % UNROLL]); }                                      // FMNMX3 (hopefully)
            if (OP == 5) { a[i] += sm[(threadIdx.x * 4 + i * 132 + it) & 4095]; }                                        // LDS.32 (conflict-free-ish)
            if (OP == 6) { sm[(threadIdx.x + i * 128 + it) & 4095] = a[i]; }                                             // STS.32 conflict-free
            if (OP == 7) {                                                                                                // LDS.128
                const float4 v = *reinterpret_cast<const float4 *>(&sm[((threadIdx.x * 132 + i * 4 + it * 4) & 4095) & ~3]);
                a[i] += v.x + v.w;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < UNROLL; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, float *out, long long *cyc) {
    for (int warps : {4, 8, 16}) {
        k<OP><<<1, warps * 32>>>(out, cyc, 0.001f);
        cudaDeviceSynchronize();
        k<OP><<<1, warps * 32>>>(out, cyc, 0.001f);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double per_sched_instr = (double)N_ITER * UNROLL * (warps / 4);
        printf("%-10s %2d warps/SM (%d per scheduler): %8lld cycles, %.2f cycles per warp-instruction per scheduler\n", name, warps, warps / 4, c,
               c / per_sched_instr);
    }
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    run<0>("MUFU.EX2", out, cyc);
    run<1>("FFMA", out, cyc);
    run<2>("FFMA2", out, cyc);
    run<3>("FADD2", out, cyc);
    run<4>("FMNMX3", out, cyc);
    run<5>("LDS.32", out, cyc);
    run<6>("STS.32", out, cyc);
    run<7>("LDS.128", out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
