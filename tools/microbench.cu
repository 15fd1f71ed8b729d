// microbench.cu -- write-bandwidth ceilings on B200 for the observation-row store pattern (not product code).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o tools/libmicrobench.so tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store(void *g, const void *s, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}

// every CTA writes `rows_per_cta` rows of row_bytes; each of the first 64 threads owns rows i, i+64, ...
// split_bytes > 0: the first split_bytes of a row go in their own bulk op (the per-env state run)
__global__ void __launch_bounds__(128) k_tma_rows(char *dst, int row_bytes, int rows_per_cta, int split_bytes, int lsu_split) {
    extern __shared__ __align__(128) char img[];
    for (int j = threadIdx.x; j < row_bytes / 8; j += blockDim.x) ((double *)img)[j] = (double)j;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    char *base = dst + (size_t)blockIdx.x * rows_per_cta * row_bytes;
    if (threadIdx.x < 64) {
        for (int r = threadIdx.x; r < rows_per_cta; r += 64) {
            char *row = base + (size_t)r * row_bytes;
            if (split_bytes > 0) {
                if (lsu_split) {
                    for (int b = 0; b < split_bytes; b += 16) *(double2 *)(row + b) = make_double2(1.0, 2.0);
                } else {
                    tma_store(row, img, split_bytes);
                }
                tma_store(row + split_bytes, img + split_bytes, row_bytes - split_bytes);
            } else {
                tma_store(row, img, row_bytes);
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// one bulk op per CTA covering rows_per_cta * row_bytes contiguous bytes (image replicated in smem)
__global__ void __launch_bounds__(128) k_tma_block(char *dst, int bytes_per_cta) {
    extern __shared__ __align__(128) char img[];
    for (int j = threadIdx.x; j < bytes_per_cta / 8; j += blockDim.x) ((double *)img)[j] = (double)j;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store(dst + (size_t)blockIdx.x * bytes_per_cta, img, bytes_per_cta);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) k_lsu(double2 *dst, size_t n16, int streaming) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) {
        if (streaming) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + i), "d"(1.0), "d"(2.0) : "memory");
        else dst[i] = make_double2(1.0, 2.0);
    }
}

extern "C" int mb_tma_rows(void *dst, long long n_rows, int row_bytes, int rows_per_cta, int split_bytes, int lsu_split, void *stream) {
    int ctas = (int)(n_rows / rows_per_cta);
    cudaFuncSetAttribute(k_tma_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k_tma_rows<<<ctas, 128, row_bytes + 128, (cudaStream_t)stream>>>((char *)dst, row_bytes, rows_per_cta, split_bytes, lsu_split);
    return (int)cudaGetLastError();
}
extern "C" int mb_tma_block(void *dst, long long total_bytes, int bytes_per_cta, void *stream) {
    cudaFuncSetAttribute(k_tma_block, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    k_tma_block<<<(int)(total_bytes / bytes_per_cta), 128, bytes_per_cta, (cudaStream_t)stream>>>((char *)dst, bytes_per_cta);
    return (int)cudaGetLastError();
}
extern "C" int mb_lsu(void *dst, long long total_bytes, int streaming, int ctas, void *stream) {
    k_lsu<<<ctas, 256, 0, (cudaStream_t)stream>>>((double2 *)dst, (size_t)total_bytes / 16, streaming);
    return (int)cudaGetLastError();
}
