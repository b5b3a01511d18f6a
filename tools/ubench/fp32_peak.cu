// Measured FP32 instruction peak of one GPU: the denominator of bench.py's roofline for the intersect / shade kernels.
//
// The parity build issues one IEEE operation per instruction (--fmad=false), so the relevant ceiling is the rate of
// scalar FADD / FMUL warp instructions, not the FMA-counted "Tflop/s" of the data sheet.  Modes: FADD, FMUL, an
// alternating FMUL/FADD stream, FFMA, and packed FFMA2 (fma.rn.f32x2, new in sm_100).  Every mode keeps 8 independent
// dependency chains per thread, 8 resident CTAs of 256 threads per SM.  Writes ONE JSON object to stdout:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu && ./fp32_peak > fp32_peak.json
// Rates are per second from CUDA events (best of 3 after a warm-up launch); per clock they are quoted against the
// device's maximum SM clock (cudaDevAttrClockRate).  clock64() is reported too, as ticks per second: on this part it does
// not advance at the SM clock, so it is not used for the per-clock figures.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

#define OP8(stmt)                                                                                                      \
    stmt(a0) stmt(a1) stmt(a2) stmt(a3) stmt(a4) stmt(a5) stmt(a6) stmt(a7)
#define FADD(x) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c));
#define FMUL(x) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(m));
#define FFMA(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(m), "f"(c));
#define FFMA2(x) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(pm), "l"(pc));

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, long long *cycles, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float const m = 1.0000001f, c = 1e-9f;
    unsigned long long p0 = pack(a0, a1), p1 = pack(a2, a3), p2 = pack(a4, a5), p3 = pack(a6, a7);
    unsigned long long p4 = pack(a1, a0), p5 = pack(a3, a2), p6 = pack(a5, a4), p7 = pack(a7, a6);
    unsigned long long const pm = pack(m, m), pc = pack(c, c);
    long long const t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { OP8(FADD) }
        if (MODE == 1) { OP8(FMUL) }
        if (MODE == 2) { FMUL(a0) FADD(a1) FMUL(a2) FADD(a3) FMUL(a4) FADD(a5) FMUL(a6) FADD(a7) }
        if (MODE == 3) { OP8(FFMA) }
        if (MODE == 4) { FFMA2(p0) FFMA2(p1) FFMA2(p2) FFMA2(p3) FFMA2(p4) FFMA2(p5) FFMA2(p6) FFMA2(p7) }
    }
    long long const t1 = clock64();
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    unsigned long long x = p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + static_cast<float>(x & 0xff);
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess)
        return 1;
    int const sms = prop.multiProcessorCount, perSM = 8, blocks = sms * perSM, iters = 1 << 16;
    int clockKHz = 0;
    cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0);
    float *out;
    long long *cycles, *hostCycles = new long long[blocks];
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaMalloc(&cycles, blocks * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const char *names[5] = {"fadd", "fmul", "fmul_fadd_mix", "ffma", "ffma2"};
    std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.1f, \"threads_per_sm\": %d, \"chains_per_thread\": 8, "
                "\"modes\": {", prop.name, sms, clockKHz / 1e3, perSM * 256);
    for (int mode = 0; mode < 5; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            switch (mode) {
            case 0: k<0><<<blocks, 256>>>(out, cycles, iters, 1.0f); break;
            case 1: k<1><<<blocks, 256>>>(out, cycles, iters, 1.0f); break;
            case 2: k<2><<<blocks, 256>>>(out, cycles, iters, 1.0f); break;
            case 3: k<3><<<blocks, 256>>>(out, cycles, iters, 1.0f); break;
            default: k<4><<<blocks, 256>>>(out, cycles, iters, 1.0f); break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best)
                best = ms;
        }
        cudaMemcpy(hostCycles, cycles, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        double meanCycles = 0;
        for (int b = 0; b < blocks; b++)
            meanCycles += double(hostCycles[b]) / blocks;
        double const warpInstPerSM = double(perSM) * 256 / 32 * iters * 8; // per SM
        double const perSecond = warpInstPerSM * sms / (best * 1e-3);        // warp instructions / s, whole GPU
        int const opsPerLane = mode == 4 ? 2 : 1;
        std::printf("%s\"%s\": {\"ms\": %.4f, \"warp_inst_per_clk_per_sm\": %.4f, \"t_warp_inst_per_s\": %.4f, "
                    "\"t_lane_ops_per_s\": %.4f, \"clock64_ticks_mhz\": %.1f}",
                    mode ? ", " : "", names[mode], best, perSecond / sms / (clockKHz * 1e3), perSecond / 1e12,
                    perSecond * 32 * opsPerLane / 1e12, meanCycles / (best * 1e-3) / 1e6);
    }
    std::printf("}}\n");
    return cudaGetLastError() != cudaSuccess;
}
