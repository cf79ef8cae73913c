// tma_probe.cu -- development probe (not part of the product): can the TMA engine stream the apply's
// footprints as many small 2D boxes {w cells, 16 frames} and write tiles back as {16 cols, 16 frames, 4 rows}
// boxes at HBM speed?  Copy-only (no compute).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

constexpr int T = 16, NROWS = 26;

template <int WBOX>
__global__ void __launch_bounds__(32) k_probe(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                                             int w_in, int h_in, int tiles_x, int n_frames, int do_store)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int BUF = NROWS * WBOX * T * 8;
    constexpr int BUFA = (BUF + 1023) / 1024 * 1024;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * BUFA);
    const int lane = threadIdx.x;
    const int tile = blockIdx.x, ty = tile / tiles_x, tx = tile % tiles_x;
    // emulated footprint: NROWS input rows, each a span of WBOX cells shifting by ~2.35 cells per row
    const int r0 = min(max(0, ty * 4 + (tx * 32) * 5 / 10 - 8), h_in - NROWS);
    const int c0 = min(max(0, tx * 32 - 6), w_in - 80);
    const int my_start = ((r0 + lane) * w_in + c0 + (lane * 235) / 100) & ~1;  // TMA: 16-byte aligned box rows
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    const int nsub = (n_frames + T - 1) / T;
    auto load = [&](int s, int buf) {
        if (lane == 0) mbar_expect_tx(&full[buf], NROWS * WBOX * T * 8);
        __syncwarp();
        if (lane < NROWS) tma_load_2d(smem + buf * BUFA + lane * (WBOX * T * 8), &map_in, my_start, s * T, &full[buf]);
    };
    load(0, 0);
    if (nsub > 1) load(1, 1);
    for (int s = 0; s < nsub; s++) {
        const int buf = s & 1;
        mbar_wait(&full[buf], (s >> 1) & 1);
        if (do_store) {
            if (lane < 2) {
                tma_store_3d(&map_out, smem + buf * BUFA + lane * 8192, tx * 32 + lane * 16, s * T, ty * 4);
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            }
            __syncwarp();
        }
        if (s + 2 < nsub) load(s + 2, buf);
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

template <int WBOX>
static void run(EncodeFn enc, double* vin, double* vout, int H, int W, int F, int do_store, int ctas_hint)
{
    CUtensorMap mi, mo;
    {
        cuuint64_t dims[2] = {(cuuint64_t)H * W, (cuuint64_t)F};
        cuuint64_t strides[1] = {(cuuint64_t)H * W * 8};
        cuuint32_t box[2] = {WBOX, T};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, vin, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode in failed %d\n", (int)r); exit(1); }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)F, (cuuint64_t)H};
        cuuint64_t strides[2] = {(cuuint64_t)H * W * 8, (cuuint64_t)W * 8};
        cuuint32_t box[3] = {16, T, 4};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, vout, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode out failed %d\n", (int)r); exit(1); }
    }
    const int tiles_x = W / 32, tiles = tiles_x * (H / 4);
    constexpr int BUF = NROWS * WBOX * T * 8;
    constexpr int BUFA = (BUF + 1023) / 1024 * 1024;
    size_t smem = 2 * BUFA + 64 + (size_t)ctas_hint;  // ctas_hint: extra bytes to limit CTAs per SM
    CK(cudaFuncSetAttribute(k_probe<WBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) k_probe<WBOX><<<tiles, 32, smem>>>(mi, mo, W, H, tiles_x, F, do_store);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int rep = 0; rep < reps; rep++) k_probe<WBOX><<<tiles, 32, smem>>>(mi, mo, W, H, tiles_x, F, do_store);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double staged = (double)tiles * NROWS * WBOX * 8.0 * F, outb = (double)H * W * 8.0 * F;
    printf("w=%2d store=%d smem/CTA=%6zu: %.3f ms  staged-in %.2f TB/s (unique in %.2f TB/s)  +out: %.2f TB/s algorithmic\n", WBOX,
           do_store, smem, ms, staged / ms / 1e9, outb / ms / 1e9, (do_store ? 2 : 1) * outb / ms / 1e9);
}

int main(int argc, char** argv)
{
    const int H = 2048, W = 2048, F = argc > 1 ? atoi(argv[1]) : 256;
    double *vin, *vout;
    CK(cudaMalloc(&vin, (size_t)H * W * F * 8));
    CK(cudaMalloc(&vout, (size_t)H * W * F * 8));
    CK(cudaMemset(vin, 0, (size_t)H * W * F * 8));
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int st = 0; st < 2; st++) {
        run<14>(enc, vin, vout, H, W, F, st, 0);
        run<14>(enc, vin, vout, H, W, F, st, 20000);
        run<14>(enc, vin, vout, H, W, F, st, 60000);
        run<10>(enc, vin, vout, H, W, F, st, 0);
        run<18>(enc, vin, vout, H, W, F, st, 0);
    }
    return 0;
}
