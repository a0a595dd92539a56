// ubench.cu -- micro-benchmarks that steer kernel design (shared atomics, match.any, global atomics).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void k_smem(int iters, int nbins, unsigned long long* sink) {
    extern __shared__ unsigned char sm[];
    uint32_t* su = (uint32_t*)sm;
    double* sd = (double*)sm;
    unsigned long long* sl = (unsigned long long*)sm;
    for (int i = threadIdx.x; i < nbins * 2; i += blockDim.x) su[i] = 0;
    __syncthreads();
    uint32_t r = hash32(blockIdx.x * 1024 + threadIdx.x);
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        r = r * 1664525u + 1013904223u;
        uint32_t b = (r >> 8) % nbins;
        if (MODE == 0) atomicAdd(&su[b], 1u);                     // u32 add, no return
        if (MODE == 1) atomicAdd(&sd[b], 1.0);                    // f64 add
        if (MODE == 2) acc += __match_any_sync(0xffffffffu, b);   // match
        if (MODE == 3) acc += atomicAdd(&su[b], 1u);              // u32 add with return
        if (MODE == 4) atomicMax(&sl[b], (unsigned long long)r);  // u64 max
        if (MODE == 5) { su[b] = r; acc += su[(b * 7 + 1) % nbins]; }   // plain store + load
        if (MODE == 6) acc += atomicCAS(&sl[b], 0ull, (unsigned long long)r);
        if (MODE == 7) atomicAdd(&sl[b], (unsigned long long)r);  // u64 add
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = acc + su[0];
}

template <int MODE>
__global__ void k_gmem(int iters, uint32_t nbins, double* gd, unsigned long long* gl, unsigned long long* sink) {
    uint32_t r = hash32(blockIdx.x * 1024 + threadIdx.x);
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        r = hash32(r + it);
        uint32_t b = r % nbins;
        if (MODE == 0) atomicAdd(&gd[b], 1.0);
        if (MODE == 1) acc += atomicCAS(&gl[b], 0ull, (unsigned long long)r + 1);
        if (MODE == 2) atomicMax(&gl[b], (unsigned long long)r);
        if (MODE == 3) acc += gl[b];
    }
    if (threadIdx.x == 0 && acc == 12345) sink[blockIdx.x] = acc;
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    unsigned long long* sink; CK(cudaMalloc(&sink, 1 << 20));
    const int grid = 148 * 4, block = 512, iters = 4096;
    const char* names[] = {"smem u32 add (RED)", "smem f64 add", "match_any", "smem u32 add ret", "smem u64 max", "smem st+ld", "smem u64 CAS", "smem u64 add"};
    int binsv[] = {1024, 4096, 64};
    for (int bi = 0; bi < 3; ++bi) {
        int nb = binsv[bi];
        size_t smem = nb * 8;
        float ms[8];
        ms[0] = timeit([&] { k_smem<0><<<grid, block, smem>>>(iters, nb, sink); });
        ms[1] = timeit([&] { k_smem<1><<<grid, block, smem>>>(iters, nb, sink); });
        ms[2] = timeit([&] { k_smem<2><<<grid, block, smem>>>(iters, nb, sink); });
        ms[3] = timeit([&] { k_smem<3><<<grid, block, smem>>>(iters, nb, sink); });
        ms[4] = timeit([&] { k_smem<4><<<grid, block, smem>>>(iters, nb, sink); });
        ms[5] = timeit([&] { k_smem<5><<<grid, block, smem>>>(iters, nb, sink); });
        ms[6] = timeit([&] { k_smem<6><<<grid, block, smem>>>(iters, nb, sink); });
        ms[7] = timeit([&] { k_smem<7><<<grid, block, smem>>>(iters, nb, sink); });
        double ops = (double)grid * block * iters;
        for (int m = 0; m < 8; ++m)
            printf("bins=%5d %-22s %8.3f ms  %7.1f Gops/s  %.2f ops/clk/SM(@1.9GHz)\n", nb, names[m], ms[m], ops / ms[m] / 1e6, ops / ms[m] / 1e6 / 148 / 1.9);
    }
    const char* gnames[] = {"gmem f64 RED", "gmem u64 CAS", "gmem u64 max", "gmem u64 load"};
    uint32_t gbins[] = {1024, 1u << 16, 1u << 20, 1u << 24, 1u << 27};
    double* gd; unsigned long long* gl;
    CK(cudaMalloc(&gd, (size_t)(1u << 27) * 8)); CK(cudaMalloc(&gl, (size_t)(1u << 27) * 8));
    for (int bi = 0; bi < 5; ++bi) {
        uint32_t nb = gbins[bi];
        cudaMemset(gd, 0, (size_t)nb * 8); cudaMemset(gl, 0, (size_t)nb * 8);
        int git = 256;
        float ms[4];
        ms[0] = timeit([&] { k_gmem<0><<<grid, block>>>(git, nb, gd, gl, sink); });
        ms[1] = timeit([&] { k_gmem<1><<<grid, block>>>(git, nb, gd, gl, sink); });
        ms[2] = timeit([&] { k_gmem<2><<<grid, block>>>(git, nb, gd, gl, sink); });
        ms[3] = timeit([&] { k_gmem<3><<<grid, block>>>(git, nb, gd, gl, sink); });
        double ops = (double)grid * block * git;
        for (int m = 0; m < 4; ++m)
            printf("gbins=%9u %-16s %8.3f ms  %7.1f Gops/s\n", nb, gnames[m], ms[m], ops / ms[m] / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
