// profiles/microbench_stream.cu -- how fast can a B200 stream 170 MB when every WARP owns a private block?
//
// ks_strings (w-fsa_b200/csrc/kernels_seg.cuh) reads one ~5.7 KB block of 128-byte rows per warp and group,
// and tops out near 3.5 TB/s.  This micro-benchmark separates the access pattern from the arithmetic:
//   linear   : classic grid-stride, warp i reads rows i, i+GW, ...           (the STREAM-like upper bound)
//   blocked  : warp owns BLOCK_ROWS consecutive rows, 8 rows per step, double buffered (the ks_strings pattern)
//   inter    : same, but the blocks of the 16 warps of a CTA are interleaved chunk by chunk (8 rows = 1 KB)
// plus the blocked pattern with 1, 2, 4 chunks in flight per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_stream microbench_stream.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int BLOCK_ROWS = 48;      // rows of 32 words per block (6 KB)

template <int MODE, int DEPTH>
__global__ void __launch_bounds__(512, 2) k_stream(const uint32_t* __restrict__ w, long long n_blocks, unsigned long long* out,
                                                     unsigned int* counter)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long acc = 0;
    if (MODE == 0) {
        const long long rows = n_blocks * BLOCK_ROWS;
        const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, GW = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long r = gw; r < rows; r += GW * 8) {
            uint32_t a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = r + j * GW < rows ? __ldcs(w + (r + j * GW) * 32 + lane) : 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += a[j];
        }
    } else {
        __shared__ long long s_sg;
        for (;;) {
            long long b;
            if (MODE == 1) {                                   // warp-level dynamic blocks
                long long g = 0;
                if (lane == 0) g = (long long)atomicAdd(counter, 1u);
                b = __shfl_sync(0xffffffffu, g, 0);
                if (b >= n_blocks) break;
            } else {                                           // CTA-level dynamic super-blocks of 16 blocks
                __syncthreads();
                if (threadIdx.x == 0) s_sg = (long long)atomicAdd(counter, 1u);
                __syncthreads();
                b = s_sg * 16 + warp;
                if (s_sg * 16 >= n_blocks) break;
            }
            // row i of block b: MODE 1 -> (b*BLOCK_ROWS + i); MODE 2 -> chunk-interleaved inside the super-block
            auto row_ptr = [&](int i) -> const uint32_t* {
                if (MODE == 1) return w + ((size_t)b * BLOCK_ROWS + i) * 32 + lane;
                const size_t sg = (size_t)(b >> 4), wi = (size_t)(b & 15);
                return w + ((sg * (BLOCK_ROWS / 8) + (size_t)(i >> 3)) * 16 + wi) * 256 + (size_t)(i & 7) * 32 + lane;
            };
            uint32_t buf[DEPTH][8];
#pragma unroll
            for (int d = 0; d < DEPTH - 1; ++d)
#pragma unroll
                for (int j = 0; j < 8; ++j) buf[d][j] = __ldcs(row_ptr(d * 8 + j));
#pragma unroll
            for (int c = 0; c < BLOCK_ROWS / 8; ++c) {
                const int pc = c + DEPTH - 1;
                if (pc < BLOCK_ROWS / 8)
#pragma unroll
                    for (int j = 0; j < 8; ++j) buf[pc % DEPTH][j] = __ldcs(row_ptr(pc * 8 + j));
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += buf[c % DEPTH][j];
            }
        }
    }
    if (acc == 0x123456789abcdefull) out[0] = acc;
}

template <int MODE, int DEPTH>
static void run(const char* name, const uint32_t* d_w, long long n_blocks, unsigned long long* d_out, unsigned int* d_cnt, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        cudaMemset(d_cnt, 0, 4);
        cudaEventRecord(e0);
        k_stream<MODE, DEPTH><<<sms * 2, 512>>>(d_w, n_blocks, d_out, d_cnt);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    const double bytes = (double)n_blocks * BLOCK_ROWS * 128.0;
    printf("%-28s %8.1f us  %7.1f GB/s  (%s)\n", name, best * 1e3, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const long long n_blocks = 29600;                          // x 6 KB = 182 MB (> 126 MB L2), a multiple of 16
    const size_t words = (size_t)n_blocks * BLOCK_ROWS * 32;
    uint32_t* d_w; unsigned long long* d_out; unsigned int* d_cnt;
    cudaMalloc(&d_w, words * 4); cudaMalloc(&d_out, 8); cudaMalloc(&d_cnt, 4);
    cudaMemset(d_w, 1, words * 4);
    printf("%s, %d SMs, %.0f MB per pass\n", prop.name, prop.multiProcessorCount, words * 4e-6);
    run<0, 1>("linear grid-stride", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<1, 1>("blocked, 1 chunk in flight", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<1, 2>("blocked, 2 chunks in flight", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<1, 3>("blocked, 3 chunks in flight", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<1, 6>("blocked, whole block in flight", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<2, 2>("interleaved, 2 chunks", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<2, 3>("interleaved, 3 chunks", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    run<2, 6>("interleaved, whole block", d_w, n_blocks, d_out, d_cnt, prop.multiProcessorCount);
    return 0;
}
