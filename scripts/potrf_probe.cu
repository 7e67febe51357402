// potrf_probe.cu — standalone timing probe (not part of the library): cuSOLVER Dpotrf UPPER / LOWER, Xpotrf, Dgetrf and a
// blocked right-looking Cholesky on cuBLAS (potrf of the diagonal block + trsm + syrk) at the repair's size.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/potrf_probe.cu -o /tmp/potrf_probe -lcusolver -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { auto _e = (x); if (_e != 0) { printf("FAIL %s = %d at line %d\n", #x, (int)_e, __LINE__); exit(1); } } while (0)

__global__ void fill_spd(double *A, long n, unsigned long long seed) {
    // A = small symmetric noise + dominant diagonal (positive definite)
    long j = blockIdx.y;
    for (long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long a = i < j ? i : j, b = i < j ? j : i;
        unsigned long long h = (a * 1000003ull + b) * 0x9E3779B97F4A7C15ull + seed;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        double v = (double)(h & 0xFFFFF) / 1048576.0 - 0.5;
        A[i + j * n] = (i == j) ? (double)n * 0.3 + v : v;
    }
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 20000;
    long fail_at = argc > 2 ? atol(argv[2]) : -1;  // make the leading minor of this order non-positive (0-based pivot)
    double *A, *W;
    CK(cudaMalloc(&A, n * n * 8));
    CK(cudaMalloc(&W, n * n * 8));
    fill_spd<<<dim3(64, n), 256>>>(A, n, 7);
    CK(cudaDeviceSynchronize());
    if (fail_at >= 0) {
        const double neg = -1.0;
        CK(cudaMemcpy(A + fail_at + fail_at * n, &neg, 8, cudaMemcpyHostToDevice));
        printf("pivot %ld made negative: every Cholesky below must report info = %ld\n", fail_at, fail_at + 1);
    }
    cusolverDnHandle_t h; cublasHandle_t bh;
    CK(cusolverDnCreate(&h)); CK(cublasCreate(&bh));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    CK(cusolverDnSetStream(h, st)); CK(cublasSetStream(bh, st));
    int lw1, lw2, lw3;
    CK(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_UPPER, n, W, n, &lw1));
    CK(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, W, n, &lw2));
    CK(cusolverDnDgetrf_bufferSize(h, n, n, W, n, &lw3));
    int lw = std::max(lw1, std::max(lw2, lw3));
    double *work; int *ipiv, *info;
    CK(cudaMalloc(&work, (size_t)lw * 8)); CK(cudaMalloc(&ipiv, n * 4)); CK(cudaMalloc(&info, 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, double flop, auto &&fn) {
        float best = 1e30f; int hinfo = 0;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemcpyAsync(W, A, n * n * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaEventRecord(e0, st));
            fn();
            CK(cudaEventRecord(e1, st));
            CK(cudaMemcpyAsync(&hinfo, info, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("%-34s %8.1f ms  %6.1f TFLOP/s  info=%d\n", name, best, flop / (best * 1e-3) / 1e12, hinfo);
    };
    const double n3 = (double)n * n * n;
    timeit("Dpotrf UPPER", n3 / 3, [&] { CK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, n, W, n, work, lw, info)); });
    timeit("Dpotrf LOWER", n3 / 3, [&] { CK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, n, W, n, work, lw, info)); });
    timeit("Dgetrf", 2 * n3 / 3, [&] { CK(cusolverDnDgetrf(h, n, n, W, n, work, ipiv, info)); });
    {
        cusolverDnParams_t pr; CK(cusolverDnCreateParams(&pr));
        size_t wd, wh;
        CK(cusolverDnXpotrf_bufferSize(h, pr, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, W, n, CUDA_R_64F, &wd, &wh));
        void *xd; CK(cudaMalloc(&xd, wd)); std::vector<char> xh(wh + 1);
        timeit("Xpotrf UPPER", n3 / 3, [&] { CK(cusolverDnXpotrf(h, pr, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, W, n, CUDA_R_64F, xd, wd, xh.data(), wh, info)); });
    }
    const double one = 1.0, mone = -1.0;
    for (int nb : {512, 1024, 2048, 4096}) {
        char name[64];
        snprintf(name, sizeof name, "blocked UPPER nb=%d trsm+syrk", nb);
        timeit(name, n3 / 3, [&] {
            for (long k = 0; k < n; k += nb) {
                long kb = std::min<long>(nb, n - k);
                CK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, kb, W + k + k * n, n, work, lw, info));
                int hi = 0;
                CK(cudaMemcpyAsync(&hi, info, 4, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (hi != 0) break;  // early exit: the panel that holds the non-positive pivot ends the factorisation
                long m2 = n - k - kb;
                if (m2 > 0) {
                    CK(cublasDtrsm(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, kb, m2, &one,
                                   W + k + k * n, n, W + k + (k + kb) * n, n));
                    CK(cublasDsyrk(bh, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, m2, kb, &mone, W + k + (k + kb) * n, n, &one,
                                   W + (k + kb) + (k + kb) * n, n));
                }
            }
        });
        snprintf(name, sizeof name, "blocked UPPER nb=%d trsm+gemm", nb);
        timeit(name, n3 / 3, [&] {
            // trailing update by block columns of width 2048 with dgemm (computes a bit more than the triangle)
            for (long k = 0; k < n; k += nb) {
                long kb = std::min<long>(nb, n - k);
                CK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, kb, W + k + k * n, n, work, lw, info));
                long m2 = n - k - kb;
                if (m2 > 0) {
                    CK(cublasDtrsm(bh, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, kb, m2, &one,
                                   W + k + k * n, n, W + k + (k + kb) * n, n));
                    const long cw = 2048;
                    for (long c = 0; c < m2; c += cw) {
                        long cb = std::min(cw, m2 - c);
                        // rows 0 .. c+cb of block column c: A22[0:c+cb, c:c+cb] -= A12[:, 0:c+cb]^T A12[:, c:c+cb]
                        CK(cublasDgemm(bh, CUBLAS_OP_T, CUBLAS_OP_N, c + cb, cb, kb, &mone, W + k + (k + kb) * n, n,
                                       W + k + (k + kb + c) * n, n, &one, W + (k + kb) + (k + kb + c) * n, n));
                    }
                }
            }
        });
    }
    timeit("dgemm n x n x 2048 (reference rate)", 2.0 * n * n * 2048, [&] {
        CK(cublasDgemm(bh, CUBLAS_OP_T, CUBLAS_OP_N, n, n, 2048, &mone, A, n, A, n, &one, W, n));
    });
    timeit("dsyrk n x 2048 (reference rate)", 1.0 * n * n * 2048, [&] {
        CK(cublasDsyrk(bh, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, n, 2048, &mone, A, n, &one, W, n));
    });
    return 0;
}
