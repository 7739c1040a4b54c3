// Prototypes only: the cblas calls in array/util/gemm.h are instantiated solely for DistrArrayFile (out of scope).
#ifndef ITSOLV_B200_SHIM_MOLPRO_CBLAS_H
#define ITSOLV_B200_SHIM_MOLPRO_CBLAS_H
extern "C" {
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, double alpha, const double* A,
                 int lda, const double* B, int ldb, double beta, double* C, int ldc);
void cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int M, int N, double alpha, const double* A, int lda, const double* X,
                 int incX, double beta, double* Y, int incY);
}
#endif
