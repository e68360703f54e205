/* Minimal C program on the C ABI of libnyles_b200.so (include/nyles_b200.h): one multigrid solve of a point-source
 * pair on a 64^3 closed box, device memory through the CUDA runtime only -- no Python, no torch.
 *
 *   gcc -std=c99 -I include examples/c_abi_example.c -L nyles_b200 -lnyles_b200 -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,$PWD/nyles_b200 -o c_abi_example && ./c_abi_example
 *
 * This is what a non-Python host (or the reference's own ctypes layer, core/mgfordriver.py:14-24) binds to. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nyles_b200.h"

/* the four CUDA runtime calls used here, declared by hand so that the example needs no CUDA headers */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);   /* 1: host->device, 2: device->host */
extern int cudaDeviceSynchronize(void);

#define CHECK(call) do { int r_ = (call); if (r_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, r_, ny_last_error()); return 1; } } while (0)

int main(void)
{
    const int n = 64, nh = 3;
    ny_ctx* ctx = NULL;
    ny_mg* mg = NULL;
    CHECK(ny_init(0, &ctx));
    CHECK(ny_mg_create(ctx, n, n, n, 1 /* closed, mg_enums.f90:5-7 */, &mg));
    int shape[3];
    CHECK(ny_mg_shape(mg, 1, shape));
    const size_t cells = (size_t)shape[0] * shape[1] * shape[2];
    double* h = (double*)calloc(cells, sizeof(double));
    const size_t sj = (size_t)shape[2], sk = sj * shape[1];
    h[(nh + n / 4) * sk + (nh + n / 4) * sj + (nh + n / 4)] = 1.0;             /* core/mgfor/tests.f90:50-57 */
    h[(nh + 3 * n / 4) * sk + (nh + 3 * n / 4) * sj + (nh + 3 * n / 4)] = -1.0;
    double* d = NULL;
    if (cudaMalloc((void**)&d, cells * sizeof(double)) != 0) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemcpy(d, h, cells * sizeof(double), 1);
    CHECK(ny_mg_set_array(mg, 1, 2 /* b */, d, NULL));
    ny_mg_stats st;
    memset(&st, 0, sizeof(st));
    CHECK(ny_mg_solve(mg, &st, NULL));
    CHECK(ny_mg_get_array(mg, 1, 1 /* x */, d, NULL));
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, cells * sizeof(double), 2);
    printf("V-cycles %d, ||r||^2/||b||^2 = %.3e, ||b||^2 = %g, x at the source = %.6f\n", st.nite, st.res, st.normb,
           h[(nh + n / 4) * sk + (nh + n / 4) * sj + (nh + n / 4)]);
    cudaFree(d);
    free(h);
    ny_mg_destroy(mg);
    ny_free(ctx);
    return st.res < 1e-6 ? 0 : 2;
}
