// Shared-memory / shuffle throughput microbenchmark for the tile kernel's design choices (run on one B200):
// cycles per warp-instruction at 16 warps per SM for LDS.64 / LDS.128, broadcast vs distinct addresses, STS.64, SHFL, DADD.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu && ./lds_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(512) k(double *out, long long *cyc, int stride) {
    __shared__ double sm[4096 + 64];
    for (int i = threadIdx.x; i < 4096 + 64; i += blockDim.x) sm[i] = i * 0.5;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // stride = 0: all lanes one address; 1: consecutive; 2: two cells (lanes 0-15 / 16-31)
    int base = warp * 64;
    if (stride == 1) base += lane;
    if (stride == 2) base += (lane >> 4);
    if (stride == 3) base += 2 * lane;  // LDS.128 distinct
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int o = (base + 2 * u + (it & 64)) & 4094;
            if (MODE == 0) {  // LDS.64
                a0 += sm[o];
            } else if (MODE == 1) {  // LDS.128
                const double2 v = *reinterpret_cast<const double2 *>(&sm[o & ~1]);
                a0 += v.x; a1 += v.y;
            } else if (MODE == 2) {  // STS.64
                sm[o] = a0 + u;
            } else if (MODE == 3) {  // SHFL x2 (one double)
                a0 += __shfl_xor_sync(0xffffffffu, a1 + u, 1 + (u & 15));
            } else if (MODE == 4) {  // DADD only
                a0 += 1.0 + u; a1 += 2.0; a2 += 3.0; a3 += 4.0;
            } else if (MODE == 5) {  // LDS.32 x2
                const float *f = reinterpret_cast<const float *>(sm);
                a0 += f[2 * o]; a1 += f[2 * o + 1];
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int stride, double *out, long long *cyc) {
    k<MODE><<<148, 512>>>(out, cyc, stride);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148;
    // 16 warps per SM, ITER instructions each (MODE 1/3/5: see name)
    printf("%-44s stride %d: %8.3f SM-cycles per warp-instruction (16 warps resident)\n", name, stride, avg / (ITER * 16.0));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 512);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    for (int s = 0; s < 3; s++) run<0>("LDS.64 (+DADD)", s, out, cyc);
    run<1>("LDS.128 (+2 DADD)", 0, out, cyc);
    run<1>("LDS.128 (+2 DADD)", 2, out, cyc);
    run<1>("LDS.128 (+2 DADD)", 3, out, cyc);
    run<2>("STS.64", 1, out, cyc);
    run<3>("SHFL.64 = 2 SHFL (+DADD)", 0, out, cyc);
    run<4>("4 DADD", 0, out, cyc);
    run<5>("2 LDS.32 (+cvt, DADD)", 0, out, cyc);
    run<5>("2 LDS.32 (+cvt, DADD)", 1, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
