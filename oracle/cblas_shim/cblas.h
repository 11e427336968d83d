/* Minimal cblas.h for building the reference's OpenBLAS backend
 * (/root/reference/OpenBLAS/gemm.hh:4, gemv.hh:4) in an image that ships no
 * cblas.h and no system libopenblas.  The only OpenBLAS here is the one inside
 * the scipy wheel, whose symbols carry a scipy_ prefix (LP64: int is 32-bit).
 * TEST INFRASTRUCTURE ONLY -- used by oracle/Makefile, never by the product. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {
  CblasNoTrans = 111,
  CblasTrans = 112,
  CblasConjTrans = 113,
  CblasConjNoTrans = 114
} CBLAS_TRANSPOSE;
typedef CBLAS_ORDER CBLAS_LAYOUT;

#define cblas_sgemm scipy_cblas_sgemm
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_sgemv scipy_cblas_sgemv
#define cblas_dgemv scipy_cblas_dgemv
#define openblas_get_config scipy_openblas_get_config
#define openblas_get_corename scipy_openblas_get_corename
#define openblas_get_num_threads scipy_openblas_get_num_threads
#define openblas_set_num_threads scipy_openblas_set_num_threads

void cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE transa, CBLAS_TRANSPOSE transb,
                 int m, int n, int k, float alpha, const float *a, int lda,
                 const float *b, int ldb, float beta, float *c, int ldc);
void cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE transa, CBLAS_TRANSPOSE transb,
                 int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc);
void cblas_sgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int m, int n,
                 float alpha, const float *a, int lda, const float *x, int incx,
                 float beta, float *y, int incy);
void cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int m, int n,
                 double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy);
char *openblas_get_config(void);
char *openblas_get_corename(void);
int openblas_get_num_threads(void);
void openblas_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
