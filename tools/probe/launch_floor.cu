// How long does a kernel with the all-pairs kernel's launch shape take when it does nothing?  (136 CTAs x 896 threads, one CTA
// per SM, ~200 KB of dynamic shared memory, a TMEM allocation.)  Period of back-to-back launches on one stream = duration + gap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/launch_floor tools/probe/launch_floor.cu && tools/probe/launch_floor
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(896, 1) shape_kernel(unsigned *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned tmem_base;
    if (MODE >= 2) {
        if (threadIdx.x < 32) {
            unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
    if (MODE >= 1 && threadIdx.x == 0) smem[0] = 1;
    if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1;
}

template <int MODE>
static void run(const char *name, size_t dyn, int threads) {
    cudaFuncSetAttribute(shape_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, st);
        for (int i = 0; i < 500; ++i) shape_kernel<MODE><<<136, threads, dyn, st>>>(nullptr);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %6.2f us per launch (%s)\n", name, ms * 1000 / 500, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0>("136 x 128 threads, no smem", 0, 128);
    run<0>("136 x 896 threads, no smem", 0, 896);
    run<1>("136 x 896 threads, 200 KB dynamic smem", 200 * 1024, 896);
    run<2>("136 x 896 threads, 200 KB smem, TMEM alloc", 200 * 1024, 896);
    return 0;
}
