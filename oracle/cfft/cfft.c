/* cfft.c -- TEST INFRASTRUCTURE: plain-C restatement of the FFTW3 transforms P3DFFT++ uses.
 *
 * FFTW3 is the third-party dependency that holds the reference's 1D arithmetic (call sites: reference
 * build/init.C:1104-1607 plan/execute thunks, build/templ.C:1283-1366 planning, build/exec.C:565-1832 execution).
 * It is not vendored in /root/reference and not installed here, so this file restates its PUBLISHED
 * definitions (FFTW 3 manual, section "What FFTW Really Computes"):
 *   c2c   Y_k = sum_j X_j exp(sign * 2 pi i j k / n), sign = -1 FFTW_FORWARD, +1 FFTW_BACKWARD, unnormalised
 *   r2c   the k = 0..n/2 outputs of the forward transform of real data
 *   c2r   real Y_j = sum_k X_k exp(+2 pi i j k / n) over the Hermitian extension of the n/2+1 inputs
 *   REDFT00  Y_k = X_0 + (-1)^k X_{n-1} + 2 sum_{j=1}^{n-2} X_j cos(pi j k / (n-1))
 *   REDFT10  Y_k = 2 sum_j X_j cos(pi (j+1/2) k / n)          REDFT01  Y_k = X_0 + 2 sum_{j>=1} X_j cos(pi j (k+1/2) / n)
 *   REDFT11  Y_k = 2 sum_j X_j cos(pi (j+1/2)(k+1/2) / n)
 *   RODFT00  Y_k = 2 sum_j X_j sin(pi (j+1)(k+1) / (n+1))     RODFT10  Y_k = 2 sum_j X_j sin(pi (j+1/2)(k+1) / n)
 *   RODFT01  Y_k = (-1)^k X_{n-1} + 2 sum_{j<n-1} X_j sin(pi (j+1)(k+1/2) / n)
 *   RODFT11  Y_k = 2 sum_j X_j sin(pi (j+1/2)(k+1/2) / n)
 * behind the FFTW "plan_many" interface (rank 1 only; howmany, stride, dist honoured; in-place allowed).
 * The complex DFT is a Stockham autosort with one pass per prime factor (radix 4/2 specialised); the r2r kinds
 * are evaluated directly from the sums above (O(n^2), exact angle reduction) so that they are independent of
 * every fast algorithm used elsewhere in this repository.  Single precision computes in double and rounds.
 * Used by: oracle/_ref (the reference's host code on shims), tests (via ctypes), bench.py's CPU arms.
 * Never linked into the product library. */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

enum { K_C2C = 0, K_R2C = 1, K_C2R = 2, K_R2R = 3 };

struct cfft_plan_s {
  int kind, n, howmany, istride, idist, ostride, odist, sign, r2r, single;
  int nfac, fac[64];
  cpx *tw;      /* exp(-2 pi i m / n), m < n */
  double *trig; /* r2r: cos(pi m / D) and sin(pi m / D), m < 2D, interleaved */
  int D;
  cpx *a, *b;   /* work pencils */
};

static const long double PI_L = 3.14159265358979323846264338327950288L;

static void factorize(int n, int *fac, int *nfac) {
  int k = 0, p;
  while (n % 4 == 0) { fac[k++] = 4; n /= 4; }
  while (n % 2 == 0) { fac[k++] = 2; n /= 2; }
  for (p = 3; (long)p * p <= n; p += 2)
    while (n % p == 0) { fac[k++] = p; n /= p; }
  if (n > 1) fac[k++] = n;
  *nfac = k;
}

static struct cfft_plan_s *new_plan(int kind, int n, int howmany, int istride, int idist, int ostride, int odist, int single) {
  struct cfft_plan_s *p = (struct cfft_plan_s *)calloc(1, sizeof *p);
  int m;
  p->kind = kind; p->n = n; p->howmany = howmany; p->istride = istride; p->idist = idist; p->ostride = ostride;
  p->odist = odist; p->single = single;
  if (kind != K_R2R) {
    factorize(n, p->fac, &p->nfac);
    p->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    for (m = 0; m < n; m++) {
      long double ang = -2.0L * PI_L * (long double)m / (long double)n;
      p->tw[m].re = (double)cosl(ang);
      p->tw[m].im = (double)sinl(ang);
    }
  }
  p->a = (cpx *)malloc(sizeof(cpx) * (size_t)(n + 2));
  p->b = (cpx *)malloc(sizeof(cpx) * (size_t)(n + 2));
  return p;
}

/* forward (exp(-i...)) DFT of a[0..n) -> returns pointer to the buffer holding the result (a or b) */
static cpx *dft_forward(const struct cfft_plan_s *p, cpx *a, cpx *b) {
  const int n = p->n;
  const cpx *tw = p->tw;
  int Ns = 1, f;
  for (f = 0; f < p->nfac; f++) {
    const int r = p->fac[f], Lr = n / r, tws = n / (Ns * r);
    int j;
    if (r == 2) {
      for (j = 0; j < Lr; j++) {
        int k = j % Ns, base = (j - k) * 2 + k;
        cpx x0 = a[j], x1 = a[j + Lr], w = tw[k * tws], t;
        t.re = x1.re * w.re - x1.im * w.im; t.im = x1.re * w.im + x1.im * w.re;
        b[base].re = x0.re + t.re; b[base].im = x0.im + t.im;
        b[base + Ns].re = x0.re - t.re; b[base + Ns].im = x0.im - t.im;
      }
    } else if (r == 4) {
      for (j = 0; j < Lr; j++) {
        int k = j % Ns, base = (j - k) * 4 + k, e = k * tws;
        cpx x0 = a[j], x1 = a[j + Lr], x2 = a[j + 2 * Lr], x3 = a[j + 3 * Lr], w1 = tw[e], w2 = tw[2 * e], w3 = tw[3 * e];
        cpx y1, y2, y3, s0, s1, d0, d1;
        y1.re = x1.re * w1.re - x1.im * w1.im; y1.im = x1.re * w1.im + x1.im * w1.re;
        y2.re = x2.re * w2.re - x2.im * w2.im; y2.im = x2.re * w2.im + x2.im * w2.re;
        y3.re = x3.re * w3.re - x3.im * w3.im; y3.im = x3.re * w3.im + x3.im * w3.re;
        s0.re = x0.re + y2.re; s0.im = x0.im + y2.im; d0.re = x0.re - y2.re; d0.im = x0.im - y2.im;
        s1.re = y1.re + y3.re; s1.im = y1.im + y3.im;
        d1.re = y1.im - y3.im; d1.im = -(y1.re - y3.re); /* -i (y1 - y3) */
        b[base].re = s0.re + s1.re; b[base].im = s0.im + s1.im;
        b[base + Ns].re = d0.re + d1.re; b[base + Ns].im = d0.im + d1.im;
        b[base + 2 * Ns].re = s0.re - s1.re; b[base + 2 * Ns].im = s0.im - s1.im;
        b[base + 3 * Ns].re = d0.re - d1.re; b[base + 3 * Ns].im = d0.im - d1.im;
      }
    } else {
      for (j = 0; j < Lr; j++) {
        int k = j % Ns, base = (j - k) * r + k, q2;
        for (q2 = 0; q2 < r; q2++) {
          int step = (k * tws + q2 * Lr) % n, e = 0, q;
          double sr = a[j].re, si = a[j].im;
          for (q = 1; q < r; q++) {
            cpx x = a[j + q * Lr], w;
            e += step; if (e >= n) e -= n;
            w = tw[e];
            sr += x.re * w.re - x.im * w.im;
            si += x.re * w.im + x.im * w.re;
          }
          b[base + q2 * Ns].re = sr; b[base + q2 * Ns].im = si;
        }
      }
    }
    { cpx *t = a; a = b; b = t; }
    Ns *= r;
  }
  return a;
}

#define LOAD_R(ptr, idx) (p->single ? (double)((const float *)(ptr))[idx] : ((const double *)(ptr))[idx])
#define STORE_R(ptr, idx, v) do { if (p->single) ((float *)(ptr))[idx] = (float)(v); else ((double *)(ptr))[idx] = (v); } while (0)

static void exec_fft(const struct cfft_plan_s *p, const void *in, void *out) {
  const int n = p->n, h = n / 2 + 1;
  int bch, j;
  for (bch = 0; bch < p->howmany; bch++) {
    cpx *a = p->a, *res;
    if (p->kind == K_C2C) {
      const long ib = (long)bch * p->idist, ob = (long)bch * p->odist;
      for (j = 0; j < n; j++) {
        long i = 2 * (ib + (long)j * p->istride);
        a[j].re = LOAD_R(in, i);
        a[j].im = (p->sign > 0 ? -1.0 : 1.0) * LOAD_R(in, i + 1); /* backward = conj(F(conj x)) */
      }
      res = dft_forward(p, a, p->b);
      for (j = 0; j < n; j++) {
        long o = 2 * (ob + (long)j * p->ostride);
        STORE_R(out, o, res[j].re);
        STORE_R(out, o + 1, p->sign > 0 ? -res[j].im : res[j].im);
      }
    } else if (p->kind == K_R2C) {
      const long ib = (long)bch * p->idist, ob = (long)bch * p->odist;
      for (j = 0; j < n; j++) { a[j].re = LOAD_R(in, ib + (long)j * p->istride); a[j].im = 0.0; }
      res = dft_forward(p, a, p->b);
      for (j = 0; j < h; j++) {
        long o = 2 * (ob + (long)j * p->ostride);
        STORE_R(out, o, res[j].re);
        STORE_R(out, o + 1, res[j].im);
      }
    } else { /* K_C2R: Y = sum_k X_k e^{+...} = conj(F(conj X)) of the Hermitian extension; output is its real part */
      const long ib = (long)bch * p->idist, ob = (long)bch * p->odist;
      for (j = 0; j < h; j++) {
        long i = 2 * (ib + (long)j * p->istride);
        double re = LOAD_R(in, i), im = LOAD_R(in, i + 1);
        if (j == 0 || 2 * j == n) im = 0.0;
        a[j].re = re; a[j].im = -im;
        if (j > 0 && 2 * j < n) { a[n - j].re = re; a[n - j].im = im; }
      }
      res = dft_forward(p, a, p->b);
      for (j = 0; j < n; j++) STORE_R(out, ob + (long)j * p->ostride, res[j].re);
    }
  }
}

/* cos / sin of pi * m / D with m reduced modulo 2D */
static double tcos(const struct cfft_plan_s *p, long m) { return p->trig[2 * (m % (2L * p->D))]; }
static double tsin(const struct cfft_plan_s *p, long m) { return p->trig[2 * (m % (2L * p->D)) + 1]; }

static void exec_r2r_plan(const struct cfft_plan_s *p, const void *in, void *out) {
  const int n = p->n;
  double *x = (double *)p->a, *y = (double *)p->b;
  int bch, j, k;
  for (bch = 0; bch < p->howmany; bch++) {
    const long ib = (long)bch * p->idist, ob = (long)bch * p->odist;
    for (j = 0; j < n; j++) x[j] = LOAD_R(in, ib + (long)j * p->istride);
    for (k = 0; k < n; k++) {
      long double s = 0.0L;
      switch (p->r2r) {
        case FFTW_REDFT00: /* D = n-1 */
          s = x[0] + ((k & 1) ? -x[n - 1] : x[n - 1]);
          for (j = 1; j < n - 1; j++) s += 2.0L * x[j] * tcos(p, (long)j * k);
          break;
        case FFTW_REDFT10: /* D = 2n, angle pi (2j+1) k / (2n) */
          for (j = 0; j < n; j++) s += 2.0L * x[j] * tcos(p, (long)(2 * j + 1) * k);
          break;
        case FFTW_REDFT01: /* D = 2n, angle pi j (2k+1) / (2n) */
          s = x[0];
          for (j = 1; j < n; j++) s += 2.0L * x[j] * tcos(p, (long)j * (2 * k + 1));
          break;
        case FFTW_REDFT11: /* D = 4n, angle pi (2j+1)(2k+1) / (4n) */
          for (j = 0; j < n; j++) s += 2.0L * x[j] * tcos(p, (long)(2 * j + 1) * (2 * k + 1));
          break;
        case FFTW_RODFT00: /* D = n+1 */
          for (j = 0; j < n; j++) s += 2.0L * x[j] * tsin(p, (long)(j + 1) * (k + 1));
          break;
        case FFTW_RODFT10: /* D = 2n, angle pi (2j+1)(k+1) / (2n) */
          for (j = 0; j < n; j++) s += 2.0L * x[j] * tsin(p, (long)(2 * j + 1) * (k + 1));
          break;
        case FFTW_RODFT01: /* D = 2n, angle pi (j+1)(2k+1) / (2n) */
          s = (k & 1) ? -x[n - 1] : x[n - 1];
          for (j = 0; j < n - 1; j++) s += 2.0L * x[j] * tsin(p, (long)(j + 1) * (2 * k + 1));
          break;
        case FFTW_RODFT11: /* D = 4n */
          for (j = 0; j < n; j++) s += 2.0L * x[j] * tsin(p, (long)(2 * j + 1) * (2 * k + 1));
          break;
        default: abort();
      }
      y[k] = (double)s;
    }
    for (k = 0; k < n; k++) STORE_R(out, ob + (long)k * p->ostride, y[k]);
  }
}

static struct cfft_plan_s *plan_r2r(int n, int howmany, int istride, int idist, int ostride, int odist, int kind, int single) {
  struct cfft_plan_s *p = new_plan(K_R2R, n, howmany, istride, idist, ostride, odist, single);
  long m;
  p->r2r = kind;
  switch (kind) {
    case FFTW_REDFT00: p->D = n - 1; break;
    case FFTW_RODFT00: p->D = n + 1; break;
    case FFTW_REDFT11: case FFTW_RODFT11: p->D = 4 * n; break;
    default: p->D = 2 * n;
  }
  if (p->D < 1) p->D = 1;
  p->trig = (double *)malloc(sizeof(double) * 4 * (size_t)p->D);
  for (m = 0; m < 2L * p->D; m++) {
    long double ang = PI_L * (long double)m / (long double)p->D;
    p->trig[2 * m] = (double)cosl(ang);
    p->trig[2 * m + 1] = (double)sinl(ang);
  }
  return p;
}

static void destroy(struct cfft_plan_s *p) {
  if (!p) return;
  free(p->tw); free(p->trig); free(p->a); free(p->b); free(p);
}

#define CFFT_IMPL(P, R, C, SINGLE)                                                                                           \
  P##_plan P##_plan_many_dft(int rank, const int *n, int howmany, C *in, const int *inembed, int istride, int idist, C *out, \
                             const int *onembed, int ostride, int odist, int sign, unsigned flags) {                         \
    struct cfft_plan_s *p;                                                                                                   \
    (void)in; (void)out; (void)inembed; (void)onembed; (void)flags;                                                          \
    if (rank != 1) return NULL;                                                                                              \
    p = new_plan(K_C2C, n[0], howmany, istride, idist, ostride, odist, SINGLE);                                              \
    p->sign = sign;                                                                                                          \
    return p;                                                                                                                \
  }                                                                                                                          \
  P##_plan P##_plan_many_dft_r2c(int rank, const int *n, int howmany, R *in, const int *inembed, int istride, int idist,     \
                                 C *out, const int *onembed, int ostride, int odist, unsigned flags) {                       \
    (void)in; (void)out; (void)inembed; (void)onembed; (void)flags;                                                          \
    if (rank != 1) return NULL;                                                                                              \
    return new_plan(K_R2C, n[0], howmany, istride, idist, ostride, odist, SINGLE);                                           \
  }                                                                                                                          \
  P##_plan P##_plan_many_dft_c2r(int rank, const int *n, int howmany, C *in, const int *inembed, int istride, int idist,     \
                                 R *out, const int *onembed, int ostride, int odist, unsigned flags) {                       \
    (void)in; (void)out; (void)inembed; (void)onembed; (void)flags;                                                          \
    if (rank != 1) return NULL;                                                                                              \
    return new_plan(K_C2R, n[0], howmany, istride, idist, ostride, odist, SINGLE);                                           \
  }                                                                                                                          \
  P##_plan P##_plan_many_r2r(int rank, const int *n, int howmany, R *in, const int *inembed, int istride, int idist, R *out, \
                             const int *onembed, int ostride, int odist, const P##_r2r_kind *kind, unsigned flags) {         \
    (void)in; (void)out; (void)inembed; (void)onembed; (void)flags;                                                          \
    if (rank != 1) return NULL;                                                                                              \
    return plan_r2r(n[0], howmany, istride, idist, ostride, odist, (int)kind[0], SINGLE);                                    \
  }                                                                                                                          \
  void P##_execute_dft(const P##_plan p, C *in, C *out) { exec_fft(p, in, out); }                                            \
  void P##_execute_dft_r2c(const P##_plan p, R *in, C *out) { exec_fft(p, in, out); }                                        \
  void P##_execute_dft_c2r(const P##_plan p, C *in, R *out) { exec_fft(p, in, out); }                                        \
  void P##_execute_r2r(const P##_plan p, R *in, R *out) { exec_r2r_plan(p, in, out); }                                       \
  void P##_destroy_plan(P##_plan p) { destroy(p); }                                                                          \
  void P##_cleanup(void) {}                                                                                                  \
  void *P##_malloc(size_t n) { void *q = NULL; return posix_memalign(&q, 64, n ? n : 64) ? NULL : q; }                       \
  void P##_free(void *p) { free(p); }

CFFT_IMPL(fftw, double, fftw_complex, 0)
CFFT_IMPL(fftwf, float, fftwf_complex, 1)
