// Cost of one tcgen05.mma instruction on a B200 SM as a function of N, operand source and accumulator dependence.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I diff-sampler_b200/csrc -o profiles/bin/mma_microbench profiles/mma_microbench.cu
//   profiles/bin/mma_microbench            (prints a table; one CTA on one SM, plus an all-SM run for the clock under load)
//
// One thread issues COUNT MMAs (M = 128, K = 16 for kind::f16, K = 32 for kind::f8f6f4) round-robin over `chains` accumulators, commits,
// and waits; cycles = clock64 around issue + drain.  Operand contents are irrelevant for timing (shared memory is zero-filled).
// Questions it answers: is a small-N MMA bound by a per-instruction floor, by the read-modify-write latency of its
// accumulator (then independent chains overlap), or by the shared-memory read of A (then A from TMEM removes it)?
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace dsb;

struct Case {
    int n;          // MMA N
    int chains;     // accumulators used round-robin (1 = fully dependent chain)
    int a_tmem;     // A operand from tensor memory
    int f8;         // kind::f8f6f4 (K = 32) instead of kind::f16 (K = 16)
    int count;      // MMAs issued
};

__global__ void __launch_bounds__(128, 1) mma_bench(Case c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_f16(c.n);
        const uint64_t da = umma_desc_sw128(smem_u32(smem));
        const uint64_t db = umma_desc_sw128(smem_u32(smem + 16384));
        // accumulators: chain i at column i * n (chains * n <= 448); A in TMEM at columns 448 .. 511
        __syncwarp();
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < c.count; ++i) {
                const uint32_t d = tmem + (i % c.chains) * c.n;
                const int k = i & 3;
                if (c.f8) umma_f8(d, da + 2 * k, db + 2 * k, idesc, 1u);
                else if (c.a_tmem) umma_f16_ts(d, tmem + 448 + 8 * k, db + 2 * k, idesc, 1u);
                else umma_f16(d, da + 2 * k, db + 2 * k, idesc, 1u);
            }
            umma_commit(&bar);
            t1 = clock64();
            mbar_wait(&bar, 0);
            t2 = clock64();
            out[blockIdx.x * 2 + 0] = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    const int count = argc > 1 ? atoi(argv[1]) : 2048;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* out = nullptr;
    cudaMalloc(&out, sizeof(long long) * 2 * sms);
    std::vector<long long> host(2 * sms);
    printf("# tcgen05.mma M=128 cost per instruction, %d instructions per measurement, %d SMs\n", count, sms);
    printf("# kind  N    A-operand chains | grid=1: cycles/MMA (issue only) | grid=%d: cycles/MMA (min..max over SMs) | FLOP/clk/SM\n", sms);
    const int ns[] = {16, 32, 64, 96, 128, 192, 256};
    for (int f8 = 0; f8 < 2; ++f8)
        for (int a_tmem = 0; a_tmem < (f8 ? 1 : 2); ++a_tmem)
            for (int n : ns)
                for (int chains : {1, 2, 4}) {
                    if (chains * n > 448) continue;
                    Case c{n, chains, a_tmem, f8, count};
                    double one = 0, one_issue = 0, lo = 1e30, hi = 0;
                    for (int grid : {1, sms}) {
                        for (int rep = 0; rep < 3; ++rep) {       // last repetition counts (warm)
                            mma_bench<<<grid, 128, smem>>>(c, out);
                            if (cudaDeviceSynchronize() != cudaSuccess) {
                                printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
                                return 1;
                            }
                        }
                        cudaMemcpy(host.data(), out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
                        if (grid == 1) {
                            one_issue = double(host[0]) / count;
                            one = double(host[1]) / count;
                        } else {
                            for (int b = 0; b < grid; ++b) {
                                const double v = double(host[2 * b + 1]) / count;
                                lo = v < lo ? v : lo;
                                hi = v > hi ? v : hi;
                            }
                        }
                    }
                    const double flop = 2.0 * 128 * n * (f8 ? 32 : 16);
                    printf("%-5s N=%-3d %-5s chains=%d | %7.1f (%6.1f) | %7.1f .. %7.1f | %7.0f\n", f8 ? "f8" : "f16", n, a_tmem ? "tmem" : "smem", chains, one,
                           one_issue, lo, hi, flop / one);
                }
    cudaFree(out);
    return 0;
}
