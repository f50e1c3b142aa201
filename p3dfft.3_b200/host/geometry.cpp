// geometry.cpp -- processor grid, data grid (block distribution) and small permutation helpers.
// Integer-exact restatement of the reference's rules:
//   ProcGrid   build/init.C:1631-1696  (row-major Cartesian ranks, one sub-communicator per dimension)
//   DataGrid   build/init.C:1699-1863  (Pdims[i] = ProcDims[Dmap[i]]; first P - N%P blocks get floor(N/P),
//                                       the rest one more; Ldims / GlobStart from this rank's position)
//   inv_mo     build/exec.C:2959-2967, rel_change build/exec.C:2270-2296, arcmp build/exec.C:721-729
#include <cstdio>

#include "plan.h"

namespace p3dfft {

int arcmp(int *A, int *B, int N) {
  for (int i = 0; i < N; i++)
    if (A[i] != B[i]) return 1;
  return 0;
}

void inv_mo(int mo[3], int imo[3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      if (mo[i] == j) imo[j] = i;
}

// mc[r] = output storage rank of the logical dimension found at input storage rank r
void rel_change(int *imo1, int *imo2, int *mc) {
  for (int r = 0; r < 3; r++)
    for (int s = 0; s < 3; s++)
      if (imo2[s] == imo1[r]) mc[r] = s;
}

ProcGrid::ProcGrid(int procdims[3], MPI_Comm mpi_comm_init) {
  nd = 0;
  MPI_Comm_dup(mpi_comm_init, &mpi_comm_glob);
  for (int i = 0; i < 3; i++) {
    ProcDims[i] = procdims[i];
    if (ProcDims[i] > 1) nd++;
  }
  if (nd == 0) nd = 1;
  MPI_Comm_rank(mpi_comm_glob, &taskid);
  MPI_Comm_size(mpi_comm_glob, &numtasks);
  if (ProcDims[0] * ProcDims[1] * ProcDims[2] != numtasks)
    printf("Error in ProcGrid: processor grid %d x %d x %d does not match %d tasks\n", ProcDims[0], ProcDims[1], ProcDims[2],
           numtasks);
  int periodic[3] = {1, 1, 1};
  MPI_Cart_create(mpi_comm_glob, 3, ProcDims, periodic, 0, &mpi_comm_cart);
  MPI_Cart_coords(mpi_comm_cart, taskid, 3, grid_id_cart);
  for (int i = 0; i < 3; i++) {
    int remain[3] = {0, 0, 0};
    remain[i] = 1;
    MPI_Cart_sub(mpi_comm_cart, remain, &mpicomm[i]);
    MPI_Comm_rank(mpicomm[i], &grid_id_cart[i]);
  }
}

ProcGrid::ProcGrid(const ProcGrid &rhs) {
  nd = rhs.nd;
  taskid = rhs.taskid;
  numtasks = rhs.numtasks;
  MPI_Comm_dup(rhs.mpi_comm_glob, &mpi_comm_glob);
  MPI_Comm_dup(rhs.mpi_comm_cart, &mpi_comm_cart);
  for (int i = 0; i < 3; i++) {
    MPI_Comm_dup(rhs.mpicomm[i], &mpicomm[i]);
    ProcDims[i] = rhs.ProcDims[i];
    grid_id_cart[i] = rhs.grid_id_cart[i];
  }
}

ProcGrid::~ProcGrid() {
  for (int i = 0; i < 3; i++) MPI_Comm_free(&mpicomm[i]);
  MPI_Comm_free(&mpi_comm_cart);
  MPI_Comm_free(&mpi_comm_glob);
}

bool ProcGrid::operator==(const ProcGrid &P) const {
  if (nd != P.nd || taskid != P.taskid || numtasks != P.numtasks) return false;
  int res;
  MPI_Comm_compare(mpi_comm_glob, P.mpi_comm_glob, &res);
  if (res != MPI_IDENT && res != MPI_CONGRUENT) return false;
  for (int i = 0; i < 3; i++)
    if (ProcDims[i] != P.ProcDims[i] || grid_id_cart[i] != P.grid_id_cart[i]) return false;
  return true;
}

DataGrid::DataGrid(int *gdims_, int dim_conj_sym_, ProcGrid *pgrid, int *dmap, int *mem_order) {
  dim_conj_sym = dim_conj_sym_;
  Pgrid = pgrid;
  nd = pgrid->nd;
  for (int i = 0; i < 3; i++) {
    Gdims[i] = gdims_[i];
    Dmap[i] = dmap[i];
    Pdims[i] = pgrid->ProcDims[dmap[i]];
    MemOrder[i] = mem_order[i];
    grid_id[i] = pgrid->grid_id_cart[dmap[i]];
  }
  InitPencil();
  is_set = true;
}

DataGrid::DataGrid(const DataGrid &rhs) {
  is_set = rhs.is_set;
  if (!is_set) return;
  nd = rhs.nd;
  dim_conj_sym = rhs.dim_conj_sym;
  Pgrid = rhs.Pgrid;
  for (int i = 0; i < 3; i++) {
    Gdims[i] = rhs.Gdims[i];
    Ldims[i] = rhs.Ldims[i];
    Pdims[i] = rhs.Pdims[i];
    MemOrder[i] = rhs.MemOrder[i];
    L[i] = rhs.L[i];
    D[i] = rhs.D[i];
    Dmap[i] = rhs.Dmap[i];
    grid_id[i] = rhs.grid_id[i];
    GlobStart[i] = rhs.GlobStart[i];
    st_[i] = rhs.st_[i];
    sz_[i] = rhs.sz_[i];
  }
}

void DataGrid::InitPencil() {
  int nloc = 0, ndist = 0;
  for (int k = 0; k < 3; k++) L[k] = D[k] = -1;
  for (int k = 0; k < 3; k++) {
    if (Pdims[k] == 1) L[nloc++] = k;
    else D[ndist++] = k;
  }
  for (int i = 0; i < 3; i++) {
    const int n = Gdims[i], p = Pdims[i];
    const int base = n / p, nlow = p - n % p;  // blocks [0,nlow) hold `base` points, the rest base+1
    st_[i].assign(p, 0);
    sz_[i].assign(p, 0);
    int pos = 0;
    for (int b = 0; b < p; b++) {
      st_[i][b] = pos;
      sz_[i][b] = b < nlow ? base : base + 1;
      pos += sz_[i][b];
    }
    Ldims[i] = sz_[i][grid_id[i]];
    GlobStart[i] = st_[i][grid_id[i]];
  }
}

namespace b200 {

void Layout::set(const int ld[3], const int mo_[3], int pad_elems) {
  int imo[3];
  for (int i = 0; i < 3; i++) {
    ldims[i] = ld[i];
    mo[i] = mo_[i];
    imo[mo_[i]] = i;
  }
  long long s = 1;
  for (int r = 0; r < 3; r++) {
    stride[imo[r]] = s;
    long long ext = ld[imo[r]];
    // pad rows of at least 4 pad units (waste <= 25 %) that are not already whole units
    if (r == 0 && pad_elems > 1 && ext >= 4LL * pad_elems && ext % pad_elems) ext = (ext + pad_elems - 1) / pad_elems * pad_elems;
    s *= ext;
  }
  span = s;
}

}  // namespace b200
}  // namespace p3dfft
