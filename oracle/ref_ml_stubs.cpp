// ref_ml_stubs.cpp -- TEST INFRASTRUCTURE.  The reference's ML sources (ML/regression.cpp, ML/lda.cpp,
// ML/naive_bayes.cpp) are compiled unmodified into oracle/_ref for their *_impute functions (the predict side of the
// MICE write-back).  The same files hold the trainers, which call BLAS / LAPACK; neither is installed
// here and the trainers are out of scope:
//   dgemv   is needed by LDA_impute itself (lda.cpp:561): the textbook column-major y = alpha * op(A) x + beta * y,
//           written here (the reference links a system BLAS);
//   dgemm, dgelsd   are only reached from the trainers: they abort loudly.
#include <cstdio>
#include <cstdlib>

extern "C" void dgemv(char *trans, int *m, int *n, double *alpha, double *a, int *lda, double *x, int *incx, double *beta,
                      double *y, int *incy) {
  const bool t = *trans == 'T' || *trans == 't' || *trans == 'C' || *trans == 'c';
  const int rows = *m, cols = *n, leny = t ? cols : rows, lenx = t ? rows : cols;
  for (int i = 0; i < leny; i++) {
    double acc = 0.0;
    for (int j = 0; j < lenx; j++) acc += (t ? a[j + (long)i * *lda] : a[i + (long)j * *lda]) * x[(long)j * *incx];
    double &out = y[(long)i * *incy];
    out = *alpha * acc + (*beta == 0.0 ? 0.0 : *beta * out);
  }
}

#define CFB_REF_STUB(name)                                                                           \
  extern "C" void name(...) {                                                                        \
    fprintf(stderr, "oracle/_ref: " #name " called -- the reference's trainers are not available\n"); \
    abort();                                                                                         \
  }
CFB_REF_STUB(dgemm)
CFB_REF_STUB(dgelsd)
