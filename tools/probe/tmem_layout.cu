// Probe: which (lane, column) of tensor memory does each register of tcgen05.ld.16x256b.x4 hold?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout tmem_layout.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t *out) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_ptr)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr;
    const uint32_t row = warp * 32 + lane;
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = (row << 16) | c;
    const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                    "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int h = 0; h < 2; ++h) {
        uint32_t r[16];
        const uint32_t a = base + ((uint32_t)(warp * 32 + 16 * h) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(a));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[((warp * 2 + h) * 32 + lane) * 16 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(64u) : "memory");
}
int main() {
    uint32_t *d, h[4 * 2 * 32 * 16];
    cudaMalloc(&d, sizeof h);
    probe<<<1, 128>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int w = 0; w < 4; ++w) for (int hh = 0; hh < 2; ++hh) for (int l = 0; l < 32; ++l) for (int i = 0; i < 16; ++i) {
        const uint32_t x = h[((w * 2 + hh) * 32 + l) * 16 + i];
        const int row = x >> 16, col = x & 0xffff;
        // hypothesis: reg i = 4k + 2g + e -> row = 32w + 16h + l/4 + 8g, col = 8k + 2(l%4) + e
        const int k = i >> 2, g = (i >> 1) & 1, e = i & 1;
        const int erow = 32 * w + 16 * hh + l / 4 + 8 * g, ecol = 8 * k + 2 * (l % 4) + e;
        if (row != erow || col != ecol) { if (bad < 40) printf("w%d h%d lane%2d reg%2d: row %3d col %2d (expected %3d %2d)\n", w, hh, l, i, row, col, erow, ecol); ++bad; }
    }
    printf("mismatches vs hypothesis: %d\n", bad);
    for (int l = 0; l < 8; ++l) { printf("lane %d:", l); for (int i = 0; i < 16; ++i) { uint32_t x = h[l * 16 + i]; printf(" (%d,%d)", x >> 16, x & 0xffff); } printf("\n"); }
    return 0;
}
