// Throughput of packed FP32 (FFMA2 / FMUL2 / FADD2, PTX *.f32x2, new in sm_100) against scalar FFMA on one GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float const m = 1.0000001f, c = 1e-9f;
    unsigned long long p0 = pack(a0, a1), p1 = pack(a2, a3), p2 = pack(a4, a5), p3 = pack(a6, a7);
    unsigned long long q0 = pack(a1, a0), q1 = pack(a3, a2), q2 = pack(a5, a4), q3 = pack(a7, a6);
    unsigned long long const pm = pack(m, m), pc = pack(c, c);
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { // 8 independent scalar FFMA
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        } else { // 8 independent packed FFMA2 (16 flops-pairs)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q0) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q1) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q2) : "l"(pm), "l"(pc));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q3) : "l"(pm), "l"(pc));
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    unsigned long long x = p0 ^ p1 ^ p2 ^ p3 ^ q0 ^ q1 ^ q2 ^ q3;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + static_cast<float>(x & 0xff);
}

int main() {
    int const blocks = 148 * 8, iters = 1 << 16;
    float *out;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, iters, 1.0f);
            else k<1><<<blocks, 256>>>(out, iters, 1.0f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double const inst = double(blocks) * 256 / 32 * iters * 8;          // warp instructions
        double const lanes_ops = inst * 32 * (mode == 0 ? 1 : 2);           // FMA lane-operations
        std::printf("%s: %.3f ms, %.2f T warp-lane FMA/s (%.1f Tflop/s counting 2 per FMA), %.1f warp instr/clk/SM\n",
                    mode == 0 ? "FFMA " : "FFMA2", ms, lanes_ops / ms / 1e9, 2 * lanes_ops / ms / 1e9,
                    inst / (ms * 1e-3) / 1.965e9 / 148);
    }
    return cudaGetLastError() != cudaSuccess;
}
