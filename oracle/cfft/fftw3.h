/* fftw3.h -- TEST INFRASTRUCTURE.  Stand-in for the FFTW3 header, declaring only the entry points the
 * P3DFFT++ reference calls (reference build/init.C:1104-1607, build/templ.C:1283-1366, include/p3dfft.h:93-101,
 * 275-280).  FFTW3 itself is a third-party dependency that is neither vendored in /root/reference nor
 * installed in this image (version unpinned: configure.ac:129-153 only probes for fftw_execute), so
 * oracle/cfft/cfft.c restates FFTW's PUBLISHED transform definitions (FFTW 3 manual, "What FFTW Really
 * Computes") in plain C.  With this header + cfft.c the reference's own host code (planner, reorder, pack,
 * unpack) compiles unmodified into oracle/_ref and serves as the executable layout oracle.
 * Never linked into the product library. */
#ifndef ORACLE_CFFT_FFTW3_H
#define ORACLE_CFFT_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct cfft_plan_s *fftw_plan;
typedef struct cfft_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_PATIENT (1U << 5)

typedef enum {
  FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2,
  FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6,
  FFTW_RODFT00 = 7, FFTW_RODFT01 = 8, FFTW_RODFT10 = 9, FFTW_RODFT11 = 10
} fftw_r2r_kind;
typedef fftw_r2r_kind fftwf_r2r_kind;

#define CFFT_DECL(P, R, C)                                                                                                   \
  P##_plan P##_plan_many_dft(int rank, const int *n, int howmany, C *in, const int *inembed, int istride, int idist, C *out, \
                             const int *onembed, int ostride, int odist, int sign, unsigned flags);                          \
  P##_plan P##_plan_many_dft_r2c(int rank, const int *n, int howmany, R *in, const int *inembed, int istride, int idist,     \
                                 C *out, const int *onembed, int ostride, int odist, unsigned flags);                        \
  P##_plan P##_plan_many_dft_c2r(int rank, const int *n, int howmany, C *in, const int *inembed, int istride, int idist,     \
                                 R *out, const int *onembed, int ostride, int odist, unsigned flags);                        \
  P##_plan P##_plan_many_r2r(int rank, const int *n, int howmany, R *in, const int *inembed, int istride, int idist, R *out, \
                             const int *onembed, int ostride, int odist, const P##_r2r_kind *kind, unsigned flags);          \
  void P##_execute_dft(const P##_plan p, C *in, C *out);                                                                     \
  void P##_execute_dft_r2c(const P##_plan p, R *in, C *out);                                                                 \
  void P##_execute_dft_c2r(const P##_plan p, C *in, R *out);                                                                 \
  void P##_execute_r2r(const P##_plan p, R *in, R *out);                                                                     \
  void P##_destroy_plan(P##_plan p);                                                                                         \
  void P##_cleanup(void);                                                                                                    \
  void *P##_malloc(size_t n);                                                                                                \
  void P##_free(void *p);

CFFT_DECL(fftw, double, fftw_complex)
CFFT_DECL(fftwf, float, fftwf_complex)
#undef CFFT_DECL

#ifdef __cplusplus
}
#endif
#endif
