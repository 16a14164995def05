// peaks.cu — B200 microbenchmarks behind the rooflines in DESIGN.md: FP64 pipe rates, launch floor, in-place stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE, int ILP> __global__ void __launch_bounds__(128) fp64_kernel(double* out, double a, double b, int iters) {
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 1e-9 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) v[j] = fma(v[j], a, b);
            else if (MODE == 1) v[j] = __dadd_rn(__dmul_rn(v[j], a), b);
            else v[j] = __dadd_rn(v[j], b);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 1.2345) out[0] = s;
}

// FP64 tensor core: mma.sync m8n8k4, NACC independent accumulator pairs per warp
template <int NACC> __global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    if (s == 1.2345) out[0] = s;
}

struct Big { double pad[340]; };
__global__ void empty_small(int) {}
__global__ void empty_big(const __grid_constant__ Big) {}

__global__ void __launch_bounds__(256) stream_inplace(double* x, int64_t n, double k) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n / 2; i += stride) {
        double2 v = reinterpret_cast<double2*>(x)[i];
        v.x = v.x * k, v.y = v.y * k;
        reinterpret_cast<double2*>(x)[i] = v;
    }
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
    double* out;
    CK(cudaMalloc(&out, 8));
    const int sms = p.multiProcessorCount;
    const int iters = 4096;
    for (int bps : {1, 2, 4, 8}) {
        const int grid = sms * bps;
        auto report = [&](const char* name, float ms, int ilp, int ops_per) {
            const double ops = (double)grid * 128 * iters * ilp * ops_per;
            printf("%-28s blocks/SM %d ILP %d: %.3f ms  %.2f Tinstr/s  = %.1f lanes/clk/SM @%.3f GHz\n", name, bps, ilp, ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (clk * 1e3), clk * 1e-6);
        };
        report("DFMA", time_ms([&] { fp64_kernel<0, 8><<<grid, 128>>>(out, 1.0000001, 1e-9, iters); }, 5), 8, 1);
        report("DMUL+DADD (dependent pair)", time_ms([&] { fp64_kernel<1, 8><<<grid, 128>>>(out, 1.0000001, 1e-9, iters); }, 5), 8, 2);
        report("DADD", time_ms([&] { fp64_kernel<2, 8><<<grid, 128>>>(out, 1.0000001, 1e-9, iters); }, 5), 8, 1);
        report("DFMA ILP2", time_ms([&] { fp64_kernel<0, 2><<<grid, 128>>>(out, 1.0000001, 1e-9, iters); }, 5), 2, 1);
        report("DFMA ILP1", time_ms([&] { fp64_kernel<0, 1><<<grid, 128>>>(out, 1.0000001, 1e-9, iters); }, 5), 1, 1);
    }
    for (int bps : {1, 2, 4}) {
        const int grid = sms * bps, it2 = 8192;
        float m8 = time_ms([&] { dmma_kernel<8><<<grid, 256>>>(out, it2); }, 5);
        float m2 = time_ms([&] { dmma_kernel<2><<<grid, 256>>>(out, it2); }, 5);
        printf("DMMA m8n8k4  CTAs/SM %d (8 warps each): 8 chains %.2f TFLOP/s, 2 chains %.2f TFLOP/s\n", bps,
               (double)grid * 8 * it2 * 8 * 512 / (m8 * 1e-3) / 1e12, (double)grid * 8 * it2 * 2 * 512 / (m2 * 1e-3) / 1e12);
    }
    printf("empty kernel, 4 B params, grid 1117x128: %.2f us/launch\n", 1e3 * time_ms([&] { empty_small<<<1117, 128>>>(0); }, 2000));
    Big big{};
    printf("empty kernel, 2720 B params, grid 1117x128: %.2f us/launch\n", 1e3 * time_ms([&] { empty_big<<<1117, 128>>>(big); }, 2000));
    printf("empty kernel, 4 B params, grid 1x32: %.2f us/launch\n", 1e3 * time_ms([&] { empty_small<<<1, 32>>>(0); }, 2000));
    // in-place stream over rotating buffers (working set 16 buffers)
    for (int64_t n : {3000000LL, 6000000LL, 12000000LL, 24000000LL, 96000000LL}) {
        const int nb = n <= 24000000 ? 16 : 4;
        std::vector<double*> bufs(nb);
        for (auto& b : bufs) { CK(cudaMalloc(&b, n * 8)); CK(cudaMemset(b, 0, n * 8)); }
        int r = 0;
        const int grid = sms * 8;
        float ms = time_ms([&] { stream_inplace<<<grid, 256>>>(bufs[r++ % nb], n, 1.0000001); }, 800);
        printf("in-place stream %lld doubles (%.0f MB r + w): %.2f us/launch = %.0f GB/s\n", (long long)n, n * 8 / 1e6, ms * 1e3, 2.0 * n * 8 / (ms * 1e-3) / 1e9);
        for (auto& b : bufs) cudaFree(b);
    }
    return 0;
}
