/* Fwrap.h -- entry points bound by the Fortran module p3dfft_plus_plus (all arguments by reference,
 * grids and plans are integer handles, idir is 1-based).  Mirrors the reference's include/Fwrap.h:81-93
 * and build/wrap.C:575-791. */
#ifndef P3DFFT_B200_FWRAP_H
#define P3DFFT_B200_FWRAP_H

void p3dfft_init_3Dtype_f(int *type, int types[3]);
void p3dfft_plan_1Dtrans_f(int *plan, int *grid1, int *grid2, int *type_id, int *dim);
void p3dfft_plan_3Dtrans_f(int *plan, int *grid1, int *grid2, Type3D *type);
int p3dfft_init_proc_grid_f(int *pdims, int *mpicomm);
void p3dfft_init_data_grid_f(int *mygrid, int *ldims, int *glob_start, int *gdims, int *dim_conj_sym, int *pgrid_id,
                             int *dmap, int *mem_order);
void p3dfft_exec_1Dtrans_double_f(int *plan, double *in, double *out, int *OW);
void p3dfft_exec_1Dtrans_single_f(int *plan, float *in, float *out, int *OW);
void p3dfft_exec_3Dtrans_double_f(Plan3D *plan, double *in, double *out, int *OW);
void p3dfft_exec_3Dtrans_single_f(Plan3D *plan, float *in, float *out, int *OW);
void p3dfft_exec_3Dderiv_double_f(Plan3D *plan, double *in, double *out, int *idir, int *OW);
void p3dfft_exec_3Dderiv_single_f(Plan3D *plan, float *in, float *out, int *idir, int *OW);
void p3dfft_compute_deriv_single_f(float *in, float *out, int *grid, int *idir);
void p3dfft_compute_deriv_double_f(double *in, double *out, int *grid, int *idir);

#endif
