// microbench_store.cu -- how fast can one B200 write [n_rows, 150] f64 observation rows (1200 B, 16-byte aligned starts)?
// Design data for the row emitters of pymgrid_b200/csrc/mg_engine.cu; standalone (no torch):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_store tools/microbench_store.cu
// Every variant runs as a persistent grid of 64-row tiles (128 threads, a warp owns 16 consecutive rows = one contiguous
// 19.2 KB chunk of the output per step) for `steps` steps, output slot = step % 4 (4 x 157 MB > L2).
//   lsu        lanes store their 3 pairs per row straight from registers (st.global.cs.v2.f64)           -- the round-1 emitter
//   tma R B    warp image of R rows in shared memory, B buffers; one cp.async.bulk S2G per R rows, image never rewritten
//   fill R B   as tma, but every row is first written into the image with STS.128 from registers (3 pairs per lane)
//   hyb R B    as fill, but every other chunk of R rows is written with per-lane 16-byte stores instead (both store paths at once)
//   real R B   as fill, but the row is gathered the way a per-env row is: 48 grid pairs by LDG.128 from an L2-resident table
//              (row-dependent offset), 2 x 24 window values from a shared-memory ring (LDS.64), 6 state values
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define D 150
#define ROW_BYTES (D * 8)
#define TILE 64
#define THREADS 128

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_cs(double *p, double a, double b) {
#if defined(STORE_NA)        // -DSTORE_NA: no L1 allocation instead of evict-first
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
#elif defined(STORE_PLAIN)
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
#else
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
#endif
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(THREADS) k_lsu(double *out, size_t slot_stride, int n_rows, int steps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * TILE + warp * 16;
    double v0 = lane, v1 = lane + 0.5;
    for (int s = 0; s < steps; ++s) {
        double *o = out + (size_t)(s & 3) * slot_stride + (size_t)r0 * D + 2 * lane;
        for (int r = 0; r < 16 && r0 + r < n_rows; ++r, o += D) {
            st_cs(o, v0, v1);
            st_cs(o + 64, v1, v0);
            if (lane < 11) st_cs(o + 128, v0, v0);
            v0 += 1.0;
        }
    }
}

// MODE 1 tma, 2 fill, 3 real
template <int R, int B, int MODE, int EW>
__global__ void __launch_bounds__(THREADS) k_img(double *out, size_t slot_stride, int n_rows, int steps, const double *__restrict__ table,
                                               int table_rows, int ragged) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RW = TILE / EW;   // rows per emitting warp and step (EW = 2: the other two warps idle, like the owner warps of the split kernels)
    double *img = reinterpret_cast<double *>(smem) + (size_t)(warp % EW) * B * R * D;
    double *ring = reinterpret_cast<double *>(smem) + (size_t)EW * B * R * D;   // [2][TILE][26] (MODE 3)
    const int r0 = blockIdx.x * TILE + (warp % EW) * RW;
    if (MODE == 3)
        for (int i = threadIdx.x; i < 2 * TILE * 26; i += THREADS) ring[i] = i * 0.25;
    for (int i = lane; i < B * R * D; i += 32) img[i] = i;
    __syncthreads();
    if (warp >= EW) return;
    fence_async();
    double v0 = lane, v1 = lane + 0.5;
    int buf = 0;
    for (int s = 0; s < steps; ++s) {
        double *o = out + (size_t)(s & 3) * slot_stride + (size_t)r0 * D;
#pragma unroll 1
        for (int c = 0; c < RW / R; ++c) {
            double *im = img + (size_t)buf * R * D;
            if (MODE == 4 && (c & 1)) {      // hybrid: this chunk leaves through the LSU
                double *ol = o + (size_t)c * R * D + 2 * lane;
#pragma unroll
                for (int r = 0; r < R; ++r, ol += D) {
                    st_cs(ol, v0, v1);
                    st_cs(ol + 64, v1, v0);
                    if (lane < 11) st_cs(ol + 128, v0, v0);
                    v0 += 1.0;
                }
                continue;
            }
            if (MODE >= 2) {
                if (lane == 0) bulk_wait_read<B - 1>();   // the store that last read this buffer has drained it
                __syncwarp();
                if (MODE == 2 || MODE == 4) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double2 *row = reinterpret_cast<double2 *>(im + r * D);
                        row[lane] = make_double2(v0, v1);
                        row[lane + 32] = make_double2(v1, v0);
                        if (lane < 11) row[lane + 64] = make_double2(v0, v0);
                        v0 += 1.0;
                    }
                } else {
                    // layout PV-first: pv 0..23 | bat, genset 24..29 | grid 30..125 | load 126..149
                    double2 g[R][2];
                    double w[R][2];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int row_id = r0 + c * R + r;
                        const int t = ragged ? (int)(((unsigned)row_id * 2654435761u) % (unsigned)(table_rows - 32)) : (s % (table_rows - 32));
                        const double2 *gs = reinterpret_cast<const double2 *>(table + ((size_t)(row_id & 3) * table_rows + t) * 4);
                        g[r][0] = __ldg(gs + lane);
                        g[r][1] = lane < 16 ? __ldg(gs + 32 + lane) : make_double2(0, 0);
                        const int e = (warp % EW) * RW + c * R + r;
                        int slot = (s % 26) + lane;
                        if (slot >= 26) slot -= 26;
                        w[r][0] = lane < 24 ? ring[e * 26 + slot] : ring[e * 26 + (lane - 24)];
                        w[r][1] = lane < 24 ? ring[(TILE + e) * 26 + slot] : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double *row = im + r * D;
                        reinterpret_cast<double2 *>(row + 30)[lane] = g[r][0];
                        if (lane < 16) reinterpret_cast<double2 *>(row + 30)[32 + lane] = g[r][1];
                        if (lane < 24) { row[lane] = w[r][0]; row[126 + lane] = w[r][1]; }
                        else if (lane < 30) row[lane] = w[r][0];
                    }
                }
                __syncwarp();
                fence_async();
            }
            if (lane == 0) {
                if (r0 + c * R < n_rows) bulk_store(o + (size_t)c * R * D, im, R * ROW_BYTES);
                bulk_commit();
            }
            buf = buf + 1 == B ? 0 : buf + 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef void (*ImgKernel)(double *, size_t, int, int, const double *, int, int);

static float time_kernel(void (*launch)(int), int steps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(20);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(steps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    return ms;
}

static double *g_out, *g_table;
static size_t g_slot;
static int g_rows, g_tiles, g_table_rows = 8784, g_ragged = 0, g_pad = 0;
static ImgKernel g_k;
static size_t g_smem;

static void launch_lsu(int steps) { k_lsu<<<g_tiles, THREADS>>>(g_out, g_slot, g_rows, steps); }
static void launch_img(int steps) { g_k<<<g_tiles, THREADS, g_smem + g_pad>>>(g_out, g_slot, g_rows, steps, g_table, g_table_rows, g_ragged); }

template <int R, int B, int MODE, int EW = 4>
static void run_img(const char *name, int steps, int pad_kb) {
    g_k = k_img<R, B, MODE, EW>;
    g_smem = (size_t)EW * B * R * ROW_BYTES + (MODE == 3 ? 2 * TILE * 26 * 8 : 0);
    g_pad = pad_kb * 1024;
    cudaFuncSetAttribute(g_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(g_smem + g_pad));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, g_k, THREADS, g_smem + g_pad);
    const float ms = time_kernel(launch_img, steps);
    const double us = 1e3 * ms / steps;
    printf("%-5s R=%2d B=%d EW=%d ragged=%d smem=%6.1f KB ctas/sm=%2d : %7.2f us/step  %6.3f TB/s\n", name, R, B, EW, g_ragged, (g_smem + g_pad) / 1024.0, occ, us,
           (double)g_rows * ROW_BYTES / us / 1e6);
}

int main(int argc, char **argv) {
    g_rows = argc > 1 ? atoi(argv[1]) : 131072;
    const int steps = argc > 2 ? atoi(argv[2]) : 200;
    g_tiles = (g_rows + TILE - 1) / TILE;
    g_slot = (size_t)g_rows * D;
    cudaMalloc(&g_out, 4 * g_slot * sizeof(double));
    cudaMalloc(&g_table, (size_t)4 * g_table_rows * 4 * sizeof(double));
    cudaMemset(g_table, 0, (size_t)4 * g_table_rows * 4 * sizeof(double));
    cudaMemset(g_out, 0, 4 * g_slot * sizeof(double));
    printf("rows=%d (%.1f MB per step), steps=%d\n", g_rows, g_rows * (double)ROW_BYTES / 1e6, steps);
    {
        const float ms = time_kernel(launch_lsu, steps);
        const double us = 1e3 * ms / steps;
        printf("lsu                                             : %7.2f us/step  %6.3f TB/s\n", us, (double)g_rows * ROW_BYTES / us / 1e6);
    }
    run_img<1, 2, 1>("tma", steps, 0);
    run_img<2, 2, 1>("tma", steps, 0);
    run_img<4, 2, 1>("tma", steps, 0);
    run_img<8, 2, 1>("tma", steps, 0);
    run_img<16, 1, 1>("tma", steps, 0);
    run_img<1, 2, 2>("fill", steps, 0);
    run_img<2, 2, 2>("fill", steps, 0);
    run_img<4, 2, 2>("fill", steps, 0);
    run_img<8, 2, 2>("fill", steps, 0);
    run_img<2, 4, 2>("fill", steps, 0);
    run_img<4, 3, 2>("fill", steps, 0);
    for (g_ragged = 0; g_ragged < 2; ++g_ragged) {
        run_img<1, 2, 3>("real", steps, 0);
        run_img<2, 2, 3>("real", steps, 0);
        run_img<4, 2, 3>("real", steps, 0);
        run_img<2, 3, 3>("real", steps, 0);
        run_img<2, 4, 3>("real", steps, 0);
        run_img<4, 3, 3>("real", steps, 0);
        run_img<8, 2, 3>("real", steps, 0);
    }
    g_ragged = 0;
    // hybrid: TMA bulk stores and per-lane stores alternate chunk by chunk
    run_img<4, 2, 4, 4>("hyb", steps, 0);
    run_img<4, 2, 4, 2>("hyb", steps, 0);
    run_img<4, 2, 4, 2>("hyb", steps, 12);
    run_img<2, 2, 4, 2>("hyb", steps, 0);
    run_img<4, 1, 4, 2>("hyb", steps, 0);
    run_img<8, 1, 4, 2>("hyb", steps, 0);
    // two emitting warps per CTA (the warp-specialised kernels), with enough padding to cap the CTAs at 7 per SM
    run_img<4, 2, 1, 2>("tma", steps, 0);
    run_img<4, 2, 2, 2>("fill", steps, 0);
    run_img<4, 2, 3, 2>("real", steps, 0);
    run_img<2, 2, 3, 2>("real", steps, 0);
    run_img<4, 4, 3, 2>("real", steps, 0);
    run_img<4, 2, 2, 2>("fill", steps, 12);
    run_img<4, 2, 3, 2>("real", steps, 8);
    // occupancy sensitivity of the best candidates (pad shared memory to cap the resident CTAs)
    run_img<2, 2, 3>("real", steps, 16);
    run_img<2, 2, 3>("real", steps, 32);
    run_img<4, 2, 3>("real", steps, 16);
    cudaFree(g_out);
    cudaFree(g_table);
    return 0;
}
