// tools/peaks.cu -- measured FP64 roofline denominators on this B200: DFMA issue rate, SHFL and LDS.64 rates, DMMA rate.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/peaks tools/peaks.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void shfl_kernel(double* out, int iters) {
    int x = threadIdx.x, y = threadIdx.x * 3, z = threadIdx.x * 5, w = threadIdx.x * 7;
    for (int it = 0; it < iters; it++) {
        x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        y = __shfl_sync(0xffffffffu, y, (threadIdx.x + 2) & 31);
        z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 3) & 31);
        w = __shfl_sync(0xffffffffu, w, (threadIdx.x + 5) & 31);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z + w;
}
__global__ void lds_kernel(double* out, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double a = 0, b = 0, c = 0, d = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        a += sm[idx]; b += sm[(idx + 32) & 2047]; c += sm[(idx + 64) & 2047]; d += sm[(idx + 96) & 2047];
        idx = (idx + 128) & 2047;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
// DMMA m8n8k4: 8*8*4*2 = 512 flop per warp instruction
__global__ void dmma_kernel(double* out, int iters) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; it++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

template <typename F> float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // DFMA: warps per SM x ILP sweep
    int wps[] = {4, 8, 16, 32};
    double best_tf = 0;
    for (int w : wps) {
        float ms8 = time_ms([&] { dfma_kernel<8><<<sms, w * 32>>>(out, iters, 1.0000001, 1e-9); });
        double tf = 2.0 * 8 * iters * (double)sms * w * 32 / (ms8 * 1e-3) / 1e12;
        printf(", \"dfma_tflops_w%d_ilp8\": %.2f", w, tf);
        if (tf > best_tf) best_tf = tf;
    }
    {   // one warp per SMSP, ILP sweep: dependent-issue latency
        float ms1 = time_ms([&] { dfma_kernel<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
        float ms2 = time_ms([&] { dfma_kernel<2><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
        float ms4 = time_ms([&] { dfma_kernel<4><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
        printf(", \"dfma_ns_per_dependent_fma_1warp_per_smsp\": %.3f, \"ilp2\": %.3f, \"ilp4\": %.3f", ms1 * 1e6 / iters, ms2 * 1e6 / iters / 2, ms4 * 1e6 / iters / 4);
    }
    printf(", \"dfma_peak_tflops\": %.2f", best_tf);
    {
        float ms = time_ms([&] { shfl_kernel<<<sms, 1024>>>(out, iters); });
        printf(", \"shfl32_warp_instr_per_clk_per_sm\": %.3f", 4.0 * iters * 32 / (ms * 1e-3) / (p.clockRate * 1e3));
        ms = time_ms([&] { lds_kernel<<<sms, 1024>>>(out, iters); });
        printf(", \"lds64_warp_instr_per_clk_per_sm\": %.3f", 4.0 * iters * 32 / (ms * 1e-3) / (p.clockRate * 1e3));
    }
    {
        float ms = time_ms([&] { dmma_kernel<<<sms, 1024>>>(out, iters); });
        printf(", \"dmma_m8n8k4_tflops\": %.2f", 512.0 * 4 * iters * 32 * sms / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { dmma_kernel<<<sms, 256>>>(out, iters); });
        printf(", \"dmma_m8n8k4_tflops_8warps\": %.2f", 512.0 * 4 * iters * 8 * sms / (ms * 1e-3) / 1e12);
    }
    printf(", \"clock_khz\": %d}\n", p.clockRate);
    return 0;
}
