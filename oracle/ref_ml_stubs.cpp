// ref_ml_stubs.cpp -- TEST INFRASTRUCTURE.  The reference's ML sources (ML/regression.cpp, ML/lda.cpp,
// ML/naive_bayes.cpp) are compiled unmodified into oracle/_ref for their *_impute functions (the predict side of the
// MICE write-back).  The same files hold the trainers, which call BLAS / LAPACK; neither is installed
// here and the trainers are out of scope:
//   dgemv   is needed by LDA_impute itself (lda.cpp:561): the textbook column-major y = alpha * op(A) x + beta * y,
//           written here (the reference links a system BLAS);
//   dgemm, dgelsd   are only reached from lda_train: forwarded to scipy's bundled OpenBLAS when the test harness
//           names it (below), else they abort loudly.
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

// dgelsd / dgemm (lda_train, lda.cpp:294-316): forwarded to the LAPACK that ships inside scipy's wheel
// (scipy.libs/libscipy_openblas*.so exports the Fortran symbols with a scipy_ prefix); oracle/ref_replay.py puts the
// path into CFB_REF_LAPACK.  Without it the trainers abort loudly, as before.
#include <dlfcn.h>

namespace {
void *lapack_symbol(const char *name) {
  static void *handle = [] {
    const char *path = getenv("CFB_REF_LAPACK");
    return path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
  }();
  void *sym = handle ? dlsym(handle, name) : nullptr;
  if (!sym) {
    fprintf(stderr, "oracle/_ref: %s is not available (set CFB_REF_LAPACK to scipy's libscipy_openblas) -- the reference's LDA trainer cannot run\n", name);
    abort();
  }
  return sym;
}
}  // namespace

extern "C" void dgelsd(int *m, int *n, int *nrhs, double *a, int *lda, double *b, int *ldb, double *s, double *rcond, int *rank,
                       double *work, int *lwork, int *iwork, int *info) {
  using fn = void (*)(int *, int *, int *, double *, int *, double *, int *, double *, double *, int *, double *, int *, int *, int *);
  static fn f = (fn)lapack_symbol("scipy_dgelsd_");
  f(m, n, nrhs, a, lda, b, ldb, s, rcond, rank, work, lwork, iwork, info);
}
extern "C" void dgemm(char *ta, char *tb, int *m, int *n, int *k, double *alpha, double *a, int *lda, double *b, int *ldb,
                      double *beta, double *c, int *ldc) {
  using fn = void (*)(char *, char *, int *, int *, int *, double *, double *, int *, double *, int *, double *, double *, int *);
  static fn f = (fn)lapack_symbol("scipy_dgemm_");
  f(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
// qda_train (ML/qda.cpp:27-330) additionally calls dgesvd and dscal
extern "C" void dgesvd(char *jobu, char *jobvt, int *m, int *n, double *a, int *lda, double *s, double *u, int *ldu, double *vt,
                       int *ldvt, double *work, int *lwork, int *info) {
  using fn = void (*)(char *, char *, int *, int *, double *, int *, double *, double *, int *, double *, int *, double *, int *, int *);
  static fn f = (fn)lapack_symbol("scipy_dgesvd_");
  f(jobu, jobvt, m, n, a, lda, s, u, ldu, vt, ldvt, work, lwork, info);
}
extern "C" void dscal(int *n, double *da, double *dx, int *incx) {
  for (int i = 0; i < *n; i++) dx[(long)i * *incx] *= *da;
}
