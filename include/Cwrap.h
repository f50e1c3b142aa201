/* Cwrap.h -- C interface of the B200-native P3DFFT++ transform path.
 * Same symbols, argument order and Grid layout as the reference's include/Cwrap.h:82-121, so C programs
 * written for P3DFFT++ (the programs under sample/C) compile and link unchanged.  Included through p3dfft.h. */
#ifndef P3DFFT_B200_CWRAP_H
#define P3DFFT_B200_CWRAP_H

/* Plain-C mirror of a DataGrid (reference Cwrap.h:82-95).  Returned by p3dfft_init_data_grid, owned by the
 * caller, released with p3dfft_free_data_grid. */
struct CDataGrid_struct {
  int nd;            /* number of distributed dimensions (1 = slabs, 2 = pencils) */
  int Gdims[3];      /* global grid size */
  int dim_conj_sym;  /* dimension stored as N/2+1 after R2C, or -1 */
  int MemOrder[3];   /* storage rank of each logical dimension (0 = unit stride) */
  int Ldims[3];      /* local size on this rank */
  int Dmap[3];       /* processor-grid dimension each logical dimension is spread over */
  int pgrid;         /* handle from p3dfft_init_proc_grid */
  int grid_id[3];    /* this rank's block index along each logical dimension */
  int GlobStart[3];  /* global index of the first local point */
  int taskid, numtasks;
  int ProcDims[3];
  MPI_Comm mpi_comm_glob;
};
typedef struct CDataGrid_struct CDataGrid;
typedef struct CDataGrid_struct Grid;

void p3dfft_setup();                                      /* wrap.C:88-139 */
void p3dfft_cleanup();                                    /* wrap.C:141-143 */
Type3D p3dfft_init_3Dtype(int types[3]);                  /* wrap.C:147-153 */
int p3dfft_plan_1Dtrans(Grid *, Grid *, int type_id, int dim);   /* wrap.C:235-277 */
Plan3D p3dfft_plan_3Dtrans(Grid *, Grid *, Type3D);       /* wrap.C:157-233 */
int find_grid(int gdims[3], int pgrid, int *dmap, int mem_order[3]);   /* wrap.C:279-294 */
int p3dfft_init_proc_grid(int pdims[3], MPI_Comm comm);   /* wrap.C:296-302 */
Grid *p3dfft_init_data_grid(int gdims[3], int dim_conj_sym, int pgrid, int dmap[3], int mem_order[3]); /* wrap.C:304-324 */
void p3dfft_free_data_grid(Grid *gr);                     /* wrap.C:334-337 */
void p3dfft_free_proc_grid(int pgrid);                    /* wrap.C:339-344 */
void p3dfft_inv_mo(int mo[3], int imo[3]);                /* wrap.C:346-348 */
void p3dfft_write_buf(double *, char *, int dims[3], int imo[3]); /* declared Cwrap.h:113, undefined in the reference */
void p3dfft_exec_1Dtrans_double(int plan, double *in, double *out, int OW);   /* wrap.C:484-516 */
void p3dfft_exec_1Dtrans_single(int plan, float *in, float *out, int OW);     /* wrap.C:518-548 */
void p3dfft_exec_3Dtrans_double(Plan3D, double *in, double *out, int OW);     /* wrap.C:387-417 */
void p3dfft_exec_3Dtrans_single(Plan3D, float *in, float *out, int OW);       /* wrap.C:354-385 */
void p3dfft_exec_3Dderiv_double(Plan3D, double *in, double *out, int idir, int OW);  /* wrap.C:451-482 */
void p3dfft_exec_3Dderiv_single(Plan3D, float *in, float *out, int idir, int OW);    /* wrap.C:419-449 */
void p3dfft_compute_deriv_single(float *in, float *out, Grid *, int idir);    /* wrap.C:551-559 */
void p3dfft_compute_deriv_double(double *in, double *out, Grid *, int idir);  /* wrap.C:561-569 */

#endif
