// ref_driver.C -- TEST INFRASTRUCTURE (our code, not the reference's): runs ONE transform case through the
// reference's own C API (reference include/Cwrap.h:101-121) when linked against the reference's unmodified host
// code (oracle/Makefile -> oracle/_ref/ref_driver).  Each rank reads its local input array from
// <in>.<rank>.bin, executes, and writes <out>.<rank>.bin plus <out>.<rank>.meta (Ldims/GlobStart of both grids).
// tests/golden/make_golden.py drives it to produce the golden vectors committed under tests/golden/.
//
// case file (whitespace separated key/value lines):
//   mode 3d|1d|deriv      procdims a b c      grid1 g0 g1 g2 cs d0 d1 d2 m0 m1 m2      grid2 ...(same)
//   types T0 T1 T2 (3d) | type T dim d (1d)   idir i   prec 4|8   dtout 1|2   ow 0|1   in PREFIX   out PREFIX
#include "p3dfft.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

static std::map<std::string, int> type_table() {
  std::map<std::string, int> t;
#define T(n) t[#n] = P3DFFT_##n;
  T(EMPTY_TYPE_SINGLE) T(EMPTY_TYPE_DOUBLE) T(EMPTY_TYPE_SINGLE_COMPLEX) T(EMPTY_TYPE_DOUBLE_COMPLEX)
  T(R2CFFT_S) T(R2CFFT_D) T(C2RFFT_S) T(C2RFFT_D) T(CFFT_FORWARD_S) T(CFFT_FORWARD_D) T(CFFT_BACKWARD_S) T(CFFT_BACKWARD_D)
#define R(k) T(k##_REAL_S) T(k##_REAL_D) T(k##_COMPLEX_S) T(k##_COMPLEX_D)
  R(DCT1) R(DST1) R(DCT2) R(DST2) R(DCT3) R(DST3) R(DCT4) R(DST4)
#undef R
#undef T
  return t;
}

struct GridSpec { int g[3], cs, dmap[3], mo[3]; };

static std::vector<char> read_file(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "ref_driver: cannot open %s\n", path.c_str()); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> b(n);
  if (n && fread(b.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "ref_driver: short read %s\n", path.c_str()); exit(2); }
  fclose(f);
  return b;
}

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  if (argc < 2) { fprintf(stderr, "usage: ref_driver <case file>\n"); return 2; }
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  std::string mode = "3d", tnames[3], inp, outp;
  int pd[3] = {1, 1, 1}, dim = 0, idir = -1, prec = 8, dtout = 2, ow = 0;
  GridSpec gs[2];
  FILE *cf = fopen(argv[1], "r");
  if (!cf) { fprintf(stderr, "ref_driver: cannot open case file\n"); return 2; }
  char key[64], buf[512];
  while (fscanf(cf, "%63s", key) == 1) {
    std::string k = key;
    if (k == "mode") { if (fscanf(cf, "%511s", buf) == 1) mode = buf; }
    else if (k == "procdims") { if (fscanf(cf, "%d %d %d", pd, pd + 1, pd + 2) != 3) return 2; }
    else if (k == "grid1" || k == "grid2") {
      GridSpec &s = gs[k == "grid2"];
      if (fscanf(cf, "%d %d %d %d %d %d %d %d %d %d", s.g, s.g + 1, s.g + 2, &s.cs, s.dmap, s.dmap + 1, s.dmap + 2, s.mo, s.mo + 1, s.mo + 2) != 10) return 2;
    } else if (k == "types") { for (int i = 0; i < 3; i++) { if (fscanf(cf, "%511s", buf) != 1) return 2; tnames[i] = buf; } }
    else if (k == "type") { if (fscanf(cf, "%511s", buf) != 1) return 2; tnames[0] = buf; }
    else if (k == "dim") { if (fscanf(cf, "%d", &dim) != 1) return 2; }
    else if (k == "idir") { if (fscanf(cf, "%d", &idir) != 1) return 2; }
    else if (k == "prec") { if (fscanf(cf, "%d", &prec) != 1) return 2; }
    else if (k == "dtout") { if (fscanf(cf, "%d", &dtout) != 1) return 2; }
    else if (k == "ow") { if (fscanf(cf, "%d", &ow) != 1) return 2; }
    else if (k == "in") { if (fscanf(cf, "%511s", buf) != 1) return 2; inp = buf; }
    else if (k == "out") { if (fscanf(cf, "%511s", buf) != 1) return 2; outp = buf; }
    else { fprintf(stderr, "ref_driver: unknown key %s\n", key); return 2; }
  }
  fclose(cf);
  if (pd[0] * pd[1] * pd[2] != size) { fprintf(stderr, "ref_driver: procdims do not match the number of ranks\n"); return 2; }

  p3dfft_setup();
  std::map<std::string, int> tt = type_table();
  int pg = p3dfft_init_proc_grid(pd, MPI_COMM_WORLD);
  Grid *g1 = p3dfft_init_data_grid(gs[0].g, gs[0].cs, pg, gs[0].dmap, gs[0].mo);
  Grid *g2 = p3dfft_init_data_grid(gs[1].g, gs[1].cs, pg, gs[1].dmap, gs[1].mo);
  long n1 = (long)g1->Ldims[0] * g1->Ldims[1] * g1->Ldims[2], n2 = (long)g2->Ldims[0] * g2->Ldims[1] * g2->Ldims[2];
  std::vector<char> in = read_file(inp + "." + std::to_string(rank) + ".bin");
  size_t cap = (size_t)(n1 > n2 ? n1 : n2) * 2 * prec + 64;
  std::vector<char> out(cap, 0);
  if (in.size() < cap) in.resize(cap, 0);

  if (mode == "3d") {
    int ids[3];
    for (int i = 0; i < 3; i++) {
      if (!tt.count(tnames[i])) { fprintf(stderr, "ref_driver: unknown type %s\n", tnames[i].c_str()); return 2; }
      ids[i] = tt[tnames[i]];
    }
    Type3D t3 = p3dfft_init_3Dtype(ids);
    Plan3D plan = p3dfft_plan_3Dtrans(g1, g2, t3);
    if (prec == 8) {
      if (idir >= 0) p3dfft_exec_3Dderiv_double(plan, (double *)in.data(), (double *)out.data(), idir, ow);
      else p3dfft_exec_3Dtrans_double(plan, (double *)in.data(), (double *)out.data(), ow);
    } else {
      if (idir >= 0) p3dfft_exec_3Dderiv_single(plan, (float *)in.data(), (float *)out.data(), idir, ow);
      else p3dfft_exec_3Dtrans_single(plan, (float *)in.data(), (float *)out.data(), ow);
    }
  } else if (mode == "1d") {
    if (!tt.count(tnames[0])) { fprintf(stderr, "ref_driver: unknown type %s\n", tnames[0].c_str()); return 2; }
    int plan = p3dfft_plan_1Dtrans(g1, g2, tt[tnames[0]], dim);
    if (prec == 8) p3dfft_exec_1Dtrans_double(plan, (double *)in.data(), (double *)out.data(), ow);
    else p3dfft_exec_1Dtrans_single(plan, (float *)in.data(), (float *)out.data(), ow);
  } else if (mode == "deriv") {
    if (prec == 8) p3dfft_compute_deriv_double((double *)in.data(), (double *)out.data(), g1, idir);
    else p3dfft_compute_deriv_single((float *)in.data(), (float *)out.data(), g1, idir);
    n2 = n1;
  } else { fprintf(stderr, "ref_driver: unknown mode\n"); return 2; }

  std::string ob = outp + "." + std::to_string(rank);
  FILE *f = fopen((ob + ".bin").c_str(), "wb");
  fwrite(out.data(), 1, (size_t)n2 * dtout * prec, f);
  fclose(f);
  f = fopen((ob + ".meta").c_str(), "w");
  fprintf(f, "%d %d %d %d %d %d %d %d %d %d %d %d\n", g1->Ldims[0], g1->Ldims[1], g1->Ldims[2], g1->GlobStart[0], g1->GlobStart[1],
          g1->GlobStart[2], g2->Ldims[0], g2->Ldims[1], g2->Ldims[2], g2->GlobStart[0], g2->GlobStart[1], g2->GlobStart[2]);
  fclose(f);
  MPI_Barrier(MPI_COMM_WORLD);
  p3dfft_cleanup();
  MPI_Finalize();
  return 0;
}
