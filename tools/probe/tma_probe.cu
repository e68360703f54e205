// TMA constraint probe: tma_probe nx ny nz bx by c0 c1 c2
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../nyles_b200/csrc/ny_tma.cuh"
void ny_set_error(const char* fmt, ...) {}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, double* out, int n, int c0, int c1, int c2, int off)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    if (threadIdx.x == 0) { nytma::mbar_init(bar, 1); nytma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { nytma::mbar_expect_tx(bar, n * 8); nytma::load_3d(smem + off, &m, c0, c1, c2, bar); }
    nytma::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<double*>(smem + off)[i];
}
int main(int argc, char** argv)
{
    int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]);
    int c0 = atoi(argv[6]), c1 = atoi(argv[7]), c2 = atoi(argv[8]);
    int off = argc > 9 ? atoi(argv[9]) : 0;
    size_t n = (size_t)nx * ny * nz;
    double* h = (double*)malloc(n * 8);
    for (size_t i = 0; i < n; i++) h[i] = (double)i + 1;
    double *d, *o;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, bx * by * 8);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz}, str[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d ", (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
    k<<<1, 128, 66000>>>(m, o, bx * by, c0, c1, c2, off);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        double* ho = (double*)malloc(bx * by * 8);
        cudaMemcpy(ho, o, bx * by * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < by; j++) for (int i = 0; i < bx; i++) {
            int gi = c0 + i, gj = c1 + j, gk = c2;
            double ref = (gi >= 0 && gi < nx && gj >= 0 && gj < ny && gk >= 0 && gk < nz) ? h[((size_t)gk * ny + gj) * nx + gi] : 0.0;
            if (ho[j * bx + i] != ref) bad++;
        }
        printf("mismatches %d", bad);
    }
    printf("\n");
    return 0;
}
