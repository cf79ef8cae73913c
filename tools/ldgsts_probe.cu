// How many shared-memory wavefronts does one LDGSTS.128 warp instruction cost, depending on the alignment of
// its global source and its shared destination?  (ncu source page: "L1 Wavefronts Shared" / instruction.)
// nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ldgsts_probe tools/ldgsts_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int SRC_OFF, int DST_OFF, int ROW>   // byte offsets (multiples of 16); ROW = contiguous bytes per source row
__global__ void __launch_bounds__(512) probe(const char* __restrict__ src, double* __restrict__ sink, int iters)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // each warp copies 512 B per instruction: lane -> 16 B piece; pieces come from rows of ROW bytes that are 16 KB apart
    const int piece = lane * 16;
    const int row = piece / ROW, col = piece % ROW;
    const char* g = src + (size_t)blockIdx.x * (1 << 20) + (size_t)warp * (64 << 10) + (size_t)row * 16384 + col + SRC_OFF;
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem + warp * 1024 + DST_OFF + piece);
    for (int it = 0; it < iters; it++) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + (size_t)it * 2048));
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = *reinterpret_cast<double*>(smem + 8 * (iters & 7));
}

template <int S, int D, int R>
void run(const char* src, double* sink, const char* name)
{
    cudaFuncSetAttribute(probe<S, D, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<S, D, R><<<296, 512, 32768>>>(src, sink, 448);
    cudaEventRecord(e0);
    probe<S, D, R><<<296, 512, 32768>>>(src, sink, 448);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s src+%d dst+%d row %d: %.3f ms (%s)\n", name, S, D, R, ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    char* src; double* sink;
    cudaMalloc(&src, (size_t)297 << 20);
    cudaMemset(src, 1, (size_t)297 << 20);
    cudaMalloc(&sink, 4096);
    run<0, 0, 512>(src, sink, "A contiguous, all 128B aligned");
    run<32, 32, 512>(src, sink, "B contiguous, both +32");
    run<16, 16, 512>(src, sink, "C contiguous, both +16");
    run<32, 16, 512>(src, sink, "D contiguous, src +32 dst +16");
    run<16, 32, 512>(src, sink, "E contiguous, src +16 dst +32");
    run<0, 0, 128>(src, sink, "F rows of 128 B, aligned");
    run<32, 32, 128>(src, sink, "G rows of 128 B, both +32");
    run<32, 16, 128>(src, sink, "H rows of 128 B, src +32 dst +16");
    run<0, 0, 96>(src, sink, "I rows of 96 B, aligned");
    run<16, 0, 96>(src, sink, "J rows of 96 B, src +16");
    cudaDeviceSynchronize();
    return 0;
}
