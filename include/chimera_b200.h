/* chimera_b200.h -- C ABI of libchimera_b200.so: the B200 (sm_100a) replacement for the f2py
 * Fortran module `chimera.moduls.fimera` of hightower8083/chimera (PIC-cycle hot path only).
 *
 * One entry point per Fortran subroutine the reference driver calls through f2py.  Arguments
 * are the Fortran dummy arguments in their original order; scalars by value; arrays as plain
 * pointers to Fortran-ordered (column-major) HOST memory, complex128 as interleaved (re,im)
 * doubles; the f2py `intent(hide)` dimensions are explicit and passed LAST, as the numpy
 * shape extents:
 *     np   particles            nxn  x nodes  (Fortran nx+1)     nm   azimuthal-mode slots
 *     nrn  r nodes incl. ghost (Fortran nr+1)                     nkx, nkr spectral extents
 * Every function returns 0 on success; on failure a non-zero status and a message retrievable
 * with chimera_last_error().  The calls are synchronous: host buffers are valid on return.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Reference citations are file:line in the reference checkout.
 */
#ifndef CHIMERA_B200_H
#define CHIMERA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef long long chb_i64;

/* ---- library / device management ---------------------------------------------------------- */
const char* chimera_last_error(void);
const char* chimera_version(void);
int chimera_device_count(int* n);
int chimera_set_device(int device);          /* default: current CUDA device */
int chimera_sync(void);
int chimera_kernel_launches(chb_i64* n);     /* kernels launched by this library so far */

/* ---- f90/particle_tools.f90 ---------------------------------------------------------------- */
/* push_velocs :18   momenta(3,np) inout, Fld(6,np) */
int chimera_push_velocs(double* momenta, const double* Fld, double dt, chb_i64 np);
/* push_coords :58   coord(3,np) inout, coord_cntr(3,np) inout */
int chimera_push_coords(double* coord, const double* momenta, double* coord_cntr, double dt, chb_i64 np);
/* genparts :84      coord(4,np) inout, indPart out, RandPackO(nx,nr), PackO complex(PPC) */
int chimera_genparts(double* coord, int* indPart, const double* Xgrid, const double* Rgrid,
                     const double* RandPackO, const double* PackX, const double* PackR, const double* PackO,
                     chb_i64 np, chb_i64 nx, chb_i64 nr, chb_i64 ppc);
/* sortpartsout :130 indx2stay(np) out (0-based), num2stay out */
int chimera_sortpartsout(int* indx2stay, int* num2stay, const double* coord, const double* lims, chb_i64 np);
/* chunk_coords_boundaries :155  chunked_indx int8(np), IndInChnk int32(nchnk+1), GoOut; nxg = len(Xgrid) */
int chimera_chunk_coords_boundaries(int8_t* chunked_indx, int* IndInChnk, int* GoOut, const double* coord,
                                    const double* lims, const double* Xgrid, int nchnk, chb_i64 np, chb_i64 nxg);
/* align_data_vec :270 / align_data_scl :298   dat(3,np0)|dat(np0) inout, idx int64(np) 0-based */
int chimera_align_data_vec(double* dat, const chb_i64* chunked_indx, chb_i64 np, chb_i64 np0);
int chimera_align_data_scl(double* dat, const chb_i64* chunked_indx, chb_i64 np, chb_i64 np0);
/* sortoutghosts :326 */
int chimera_sortoutghosts(int* indx2stay, int* num2stay, const double* coord, chb_i64 np);

/* ---- f90/grid_deps.f90, grid_deps_chnk.f90, grid_deps_env.f90, grid_deps_env_chnk.f90 ------- */
/* grids: curr(nxn,nrn,nm,3), dens(nxn,nrn,nm), Fld(nxn,nrn,nm,6) complex; Rgrid(nrn) */
int chimera_dep_curr(const double* coord, const double* momenta, const double* wghts, double* curr,
                     double leftX, const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn,
                     chb_i64 nrn, chb_i64 nm);                                    /* grid_deps.f90:18 */
int chimera_dep_dens(const double* coord, const double* wghts, double* dens, double leftX, const double* Rgrid,
                     double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm); /* :89 */
int chimera_proj_fld(const double* coord, const double* wghts, const double* Fld, double* Fld_tot, double leftX,
                     const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn,
                     chb_i64 nm);                                                 /* :149 */
int chimera_eb_correction(double* eb_spc, chb_i64 nxn, chb_i64 nrn, chb_i64 nm); /* :219 */
int chimera_dep_curr_chnk(const double* coord, const double* momenta, const double* wghts, double* curr,
                          const int* IndInChunk, int guards, double leftX, const double* Rgrid, double dx_inv,
                          double dr_inv, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm,
                          chb_i64 nchnk);                                         /* grid_deps_chnk.f90:18 */
int chimera_dep_dens_chnk(const double* coord, const double* wghts, double* dens, const int* IndInChunk,
                          int guards, double leftX, const double* Rgrid, double dx_inv, double dr_inv, chb_i64 np,
                          chb_i64 nxn, chb_i64 nrn, chb_i64 nm, chb_i64 nchnk);   /* :132 */
int chimera_dep_curr_env(const double* coord, const double* momenta, const double* wghts, double* curr,
                         double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0, chb_i64 np,
                         chb_i64 nxn, chb_i64 nrn, chb_i64 nm);                   /* grid_deps_env.f90:18 */
int chimera_dep_dens_env(const double* coord, const double* wghts, double* dens, double leftX,
                         const double* Rgrid, double dx_inv, double dr_inv, double kx0, chb_i64 np, chb_i64 nxn,
                         chb_i64 nrn, chb_i64 nm);                                /* :97 */
int chimera_proj_fld_env(const double* coord, const double* wghts, const double* Fld, double* Fld_tot,
                         double leftX, const double* Rgrid, double dx_inv, double dr_inv, double kx0, chb_i64 np,
                         chb_i64 nxn, chb_i64 nrn, chb_i64 nm);                   /* :164 */
int chimera_eb_correction_env(double* eb_spc, chb_i64 nxn, chb_i64 nrn, chb_i64 nm); /* :240 */
int chimera_dep_curr_env_chnk(const double* coord, const double* momenta, const double* wghts, double* curr,
                              const int* IndInChunk, int guards, double leftX, const double* Rgrid,
                              double dx_inv, double dr_inv, double kx0, chb_i64 np, chb_i64 nxn, chb_i64 nrn,
                              chb_i64 nm, chb_i64 nchnk);                         /* grid_deps_env_chnk.f90:18 */
int chimera_dep_dens_env_chnk(const double* coord, const double* wghts, double* dens, const int* IndInChunk,
                              int guards, double leftX, const double* Rgrid, double dx_inv, double dr_inv,
                              double kx0, chb_i64 np, chb_i64 nxn, chb_i64 nrn, chb_i64 nm,
                              chb_i64 nchnk);                                     /* :147 */

/* ---- f90/fb_io.f90 : DHT (r) + FFT (x) ------------------------------------------------------ */
/* vec(nkx,nrn,nm,3) <-> vec_fb(nkx,nkr,nm,3);  In(nrn-1,nkr,nm), Out(nkr,nrn-1,nm) real */
int chimera_fb_vec_in(double* vec_fb, const double* vec, double leftX, const double* kx, const double* In,
                      chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr);         /* :18 */
int chimera_fb_scl_in(double* scl_fb, const double* scl, double leftX, const double* kx, const double* In,
                      chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr);         /* :61 */
int chimera_fb_vec_out(double* vec, const double* vec_fb, double leftX, const double* kx, const double* Out,
                       chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr);        /* :100 */
int chimera_fb_scl_out(double* scl, const double* scl_fb, double leftX, const double* kx, const double* Out,
                       chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr);        /* :142 */
/* e_fb(nkx,nkr,nm,6) (first 3 comps used), b_fb(nkx,nkr,nm,3) -> eb_spc(nkx,nrn,nm,6) */
int chimera_fb_eb_out(double* eb_spc, const double* e_fb, const double* b_fb, double leftX, const double* kx,
                      const double* Out, chb_i64 nkx, chb_i64 nrn, chb_i64 nm, chb_i64 nkr); /* :182 */
int chimera_fb_filtr(double* vec, double leftX, const double* kx, const double* filtr, int modefilt, chb_i64 nkx,
                     chb_i64 nkr, chb_i64 nm, chb_i64 nxfilt);                    /* :230 */

/* ---- f90/fb_math.f90 (modes 0..nm-1; D stacks have nm+1 slots) ------------------------------ */
int chimera_fb_rot(double* vec_fb_loc, const double* vec_fb, const double* DpS2S, const double* DmS2S,
                   const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);   /* :18 */
int chimera_fb_grad(double* vec_fb_loc, const double* scl_fb, const double* DpS2S, const double* DmS2S,
                    const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);  /* :96 */
int chimera_fb_div(double* scl_fb_loc, const double* vec_fb, const double* DpS2S, const double* DmS2S,
                   const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);   /* :151 */
int chimera_fb_graddiv(double* vec_fb, const double* DpS2S, const double* DmS2S, const double* kx, chb_i64 nkx,
                       chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);                              /* :201 */
/* ---- f90/fb_math_env.f90 (modes -nko..nko, nm = 2nko+1; D stacks have nm+2 slots) ----------- */
int chimera_fb_grad_env(double* vec_fb_loc, const double* scl_fb, const double* DpS2S, const double* DmS2S,
                        const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);  /* :18 */
int chimera_fb_div_env(double* scl_fb_loc, const double* vec_fb, const double* DpS2S, const double* DmS2S,
                       const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);   /* :63 */
int chimera_fb_rot_env(double* vec_fb_loc, const double* vec_fb, const double* DpS2S, const double* DmS2S,
                       const double* kx, chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);   /* :106 */
int chimera_fb_graddiv_env(double* vec_fb, const double* DpS2S, const double* DmS2S, const double* kx,
                           chb_i64 nkx, chb_i64 nkr, chb_i64 nm, chb_i64 nkr_loc);                 /* :164 */

/* ---- f90/maxwell_solvers.f90 ---------------------------------------------------------------- */
int chimera_maxwell_push_with_spchrg(double* EG_fb, const double* j_fb, const double* grad_rho_n_fb,
                                     const double* grad_rho_np1_fb, const double* C1, const double* C2,
                                     chb_i64 nkx, chb_i64 nkr, chb_i64 nm);       /* :18  C real (..,5) */
int chimera_maxwell_push_wo_spchrg(double* EG_fb, const double* j_fb, const double* C1, const double* C2,
                                   chb_i64 nkx, chb_i64 nkr, chb_i64 nm);         /* :62  C complex (..,3) */
int chimera_maxwell_init_push(double* EG_fb, const double* j_fb, const double* grad_rho_n_fb, const double* C1,
                              const double* C2, chb_i64 nkx, chb_i64 nkr, chb_i64 nm); /* :98 C complex (..,2) */
int chimera_poiss_corr(double* j_fb, const double* grad_div_j_fb, const double* grad_rho_n_fb,
                       const double* grad_rho_np1_fb, double dt_inv, const double* w2_inv, chb_i64 nkx,
                       chb_i64 nkr, chb_i64 nm);                                  /* :131 */
int chimera_poiss_corr_stat(double* j_fb, const double* grad_div_j_fb, const double* grad_rho_n_fb,
                            const double* DT, const double* w2_inv, chb_i64 nkx, chb_i64 nkr,
                            chb_i64 nm);                                          /* :166 DT complex(nkx) */
int chimera_field_drift(double* EG_fb, const double* kx, double beta0, double dt, chb_i64 nkx, chb_i64 nkr,
                        chb_i64 nm);                                              /* :199 */
int chimera_omp_mult_vec(double* vec_fb, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm); /* :228 */
int chimera_omp_mult_scl(double* scl_fb, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm); /* :252 */
int chimera_omp_add_vec(double* vec_fb, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm);  /* :274 */
int chimera_omp_add_scl(double* scl_fb, const double* A, chb_i64 nkx, chb_i64 nkr, chb_i64 nm);  /* :298 */

/* ---- f90/devices.f90 (SURVEY section 8f NEXT-1) ---------------------------------------------- */
int chimera_undul_analytic(const double* coord, double* Fld, double t, const double* params, chb_i64 np); /* :162 */

/* ---- microbenchmark hook: the DHT contraction alone on device-resident random data ----------- */
/* C[2nkx x N] = A[2nkx x K] . B[K x N], `batch` independent problems, `iters` timed launches;
 * returns the mean milliseconds per launch (CUDA events) in *ms. */
int chimera_bench_gemm(chb_i64 nkx, chb_i64 K, chb_i64 N, int batch, int iters, double* ms);

#ifdef __cplusplus
}
#endif
#endif
