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
#include <stddef.h>
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

/* ---- resident mode of the per-function drop-in (chimera_b200/resident.py) -------------------------------------
 * The reference's driver owns every array as a numpy array and mutates them in Python between calls
 * (moduls/chimera_main.py:110-190, species.py:246-256, solvers.py:318).  With numpy's data allocator (NEP 49) backed
 * by these functions the arrays live in CUDA managed memory: every entry point above then takes the caller's pointer
 * as it is (no staging copy; the CUDA driver keeps host accesses coherent), and the driver's whole-array statements
 * run on the device.  Pointers that are plain host memory keep the staged-copy path. */
void* chimera_managed_alloc(size_t bytes, int zero);
void* chimera_managed_realloc(void* p, size_t new_bytes);
void chimera_managed_free(void* p);                          /* the block goes to a size-bucketed cache (<= 8 GB)       */
void chimera_managed_trim(void);                             /* release the cached blocks                               */
void chimera_managed_touched(const void* p);                 /* the host wrote into the block holding p: prefetch it to
                                                                the device at its next use (blocks are otherwise
                                                                prefetched once per (re)allocation;
                                                                CHIMERA_B200_PREFETCH=always|first|never)               */
int chimera_managed_owns(const void* p, size_t* bytes);      /* 1 when p came from chimera_managed_alloc            */
int chimera_is_device_accessible(const void* p);             /* 0 host, 1 managed, 2 device                          */
int chimera_fill(double* y, chb_i64 n, double value_re, double value_im, int is_complex); /* y[:] = value          */
int chimera_copy(void* dst, const void* src, chb_i64 bytes);                              /* dst[:] = src          */
int chimera_add_inplace(double* y, const double* x, chb_i64 n);                           /* y += x (n doubles)    */
/* bytes copied host->device / device->host by the host-buffer entry points below since the last reset */
int chimera_host_traffic(chb_i64* h2d, chb_i64* d2h, int reset);

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
int chimera_undul_analytic_taper(const double* coord, double* Fld, double t, const double* params, chb_i64 np); /* :117 */
/* a0(2,nx): tabulated on-axis field, node k (1-based) at Xleft + k dx; Q12: the reference reads node 0 (out of
 * bounds) for Xleft+dx <= x < Xleft+1.5dx -- taken as zero here */
int chimera_undul_mapped(const double* coord, double* Fld, double t, const double* a0, const double* params,
                         chb_i64 np, chb_i64 nx); /* :18 */
int chimera_undul_mapped_tap(const double* coord, double* Fld, double t, const double* a0, const double* params,
                             chb_i64 np, chb_i64 nx); /* :64 */
int chimera_planewave(const double* coord, double* Fld, double t, const double* params, chb_i64 np); /* :205 */
int chimera_gaussbeam(const double* coord, double* Fld, double time, double a0, const double* params, chb_i64 np); /* :254 */

/* ---- f90/SR.f90: synchrotron-radiation spectra from stored tracks (moduls/SR.py:165-215) ----------- */
/* spect(nom,n1,n2) inout (accumulated into); coords/momenta (3,nt,np); comp = 1,2,3 picks x,y,z (SR.py:168).
 * far field  :18 / :139   angles theta (n1 = nth) x phi (n2 = nph) given as sines and cosines
 * near field :352 / :256  Cartesian screen Xgrid(nx) x Ygrid(ny) at z_scr
 * near field :546 / :449  polar screen Rgrid(nr) x phi(nph) at z_scr */
int chimera_sr_calc_far_tot(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                            const double* wghts, double dt, const double* omega, const double* SinTh,
                            const double* CosTh, const double* SinPh, const double* CosPh, chb_i64 nt, chb_i64 np,
                            chb_i64 nom, chb_i64 nth, chb_i64 nph);
int chimera_sr_calc_far_comp(double* spect, const double* coords, const double* momenta_prv, const double* momenta_nxt,
                             const double* wghts, int comp, double dt, const double* omega, const double* SinTh,
                             const double* CosTh, const double* SinPh, const double* CosPh, chb_i64 nt, chb_i64 np,
                             chb_i64 nom, chb_i64 nth, chb_i64 nph);
int chimera_sr_calc_near_tot(double* spect, const double* coords, const double* momenta, const double* wghts, double dt,
                             const double* omega, const double* Xgrid, const double* Ygrid, double z_scr, chb_i64 nt,
                             chb_i64 np, chb_i64 nom, chb_i64 nx, chb_i64 ny);
int chimera_sr_calc_near_comp(double* spect, const double* coords, const double* momenta, const double* wghts, int comp,
                              double dt, const double* omega, const double* Xgrid, const double* Ygrid, double z_scr,
                              chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nx, chb_i64 ny);
int chimera_sr_calc_nearcirc_tot(double* spect, const double* coords, const double* momenta, const double* wghts,
                                 double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                 const double* CosPh, double z_scr, chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nr,
                                 chb_i64 nph);
int chimera_sr_calc_nearcirc_comp(double* spect, const double* coords, const double* momenta, const double* wghts,
                                  int comp, double dt, const double* omega, const double* Rgrid, const double* SinPh,
                                  const double* CosPh, double z_scr, chb_i64 nt, chb_i64 np, chb_i64 nom, chb_i64 nr,
                                  chb_i64 nph);

/* ---- f90/utils.f90: diagnostics helpers the driver reaches through fimera -------------------------- */
/* intens_profO :18  PWR_RO(NO, nrn-1) out; Fld(nxn,nrn,nm,3) complex, nm = 2 nkO + 1 (diagnostics.py:136,147,158) */
int chimera_intens_profo(double* PWR_RO, const double* Fld, int NO, chb_i64 nxn, chb_i64 nrn, chb_i64 nm);
/* DENSITY_2x :210   dens(bins_x+5, bins_y+5) out; grid = (xmin, xmax, ymin, ymax) */
int chimera_density_2x(const double* x, const double* y, const double* wght, const double* grid, int bins_x,
                       int bins_y, double* dens, chb_i64 n_part);

/* ---- microbenchmark hook: the DHT contraction alone on device-resident random data ----------- */
/* C[2nkx x N] = A[2nkx x K] . B[K x N], `batch` independent problems, `iters` timed launches;
 * returns the mean milliseconds per launch (CUDA events) in *ms. */
int chimera_bench_gemm(chb_i64 nkx, chb_i64 K, chb_i64 N, int batch, int iters, double* ms);
/* per-launch CUDA-event timing of the contraction kernel wherever it is launched (engine or host API):
 * accumulated milliseconds, flop (2*M*N*K per problem) and launches since the last reset */
int chimera_gemm_profile(int on);
int chimera_gemm_profile_read(double* ms, double* flops, chb_i64* launches, int reset);
/* per-stage clock profile of the fused particle kernel (particles_fused.cu): SM cycles summed over CTAs for
 * stages A (records + histogram), B+C (scan, counting sort), D (gather), E (push), F (deposit), G (cell
 * changers) in cycles8[0..5], CTA count in cycles8[7].  Enabling adds a barrier per stage: diagnosis only. */
int chimera_fused_profile(int on);
int chimera_fused_profile_read(unsigned long long* cycles8);

/* ---- device-resident PIC engine ------------------------------------------------------------ */
/* The per-function entry points above take HOST buffers and copy per call (the f2py drop-in).
 * The engine keeps particles (structure of arrays), grids, spectral state and operator tables in
 * HBM and runs the reference's step sequence (moduls/chimera_main.py:82-92 make_step, :61-80
 * make_halfstep) on the device.  Stage order and per-stage semantics are those of the reference
 * wrappers cited per phase below. */
typedef struct chimera_engine_config {
  int env;           /* 1: envelope solver (solver dict has KxShift, solvers.py:75)                  */
  int space_charge;  /* 'SpaceCharge' feature: rho deposition + 5-coefficient PSATD (solvers.py:244) */
  int poisson_iters; /* Poisson-correction iterations (solvers.py:301: 3); 0 = 'NoPoissonCorrection'  */
  int coef_complex;  /* PSATD_E/G tables are complex128 (KxShift) instead of float64                  */
  int chunked;       /* deposit with the *_chnk edge semantics (species dict has Xchunked)            */
  int nchnk, guards; /* Xchunked = (nchnk, guards)                                                    */
  int sort_every;    /* re-bin particles every this many steps (chimera_main.py:310: guards+1); 0: never */
  int undulator;     /* add the analytic undulator device between gather and push (species.py:258)   */
  chb_i64 nx, nrn, nkr, nm; /* x nodes, r nodes incl. ghost, radial modes, azimuthal-mode slots       */
  double leftX, rightX, dx, dr, dt, kx0;
  double rcull2;     /* particles with y^2+z^2 > rcull2 are removed at re-binning (SimDom[3])         */
  double chunk_len;  /* Xgrid[Nx/nchnk]-Xgrid[0] (nchnk>1) or Xgrid[Nx-1]-Xgrid[0] (particle_tools.f90:171) */
  double und_a0, und_lambda, und_X0, und_Lx; /* devices.f90:162 parameters                           */
  /* kx-slab sharding of the spectral solve (multi-GPU): this engine holds nx_slab of the nx kx rows (a set of
   * mirror pairs, uploaded as "slab_rows"); 0 = all rows.  mirror_shift: partner of local row j is
   * (nx_slab - j - mirror_shift) mod nx_slab (reference f90/fb_math.f90:35-36 in slab-local form).           */
  chb_i64 nx_slab;
  int mirror_shift;
  /* 'StaticKick' feature (chimera_main.py:106-125, 186): quasi-static field of each species' mean momentum, rebuilt
   * from zero every step (poiss_corr_stat, maxwell_solver_stat, field_drift); rho is deposited on coords_halfstep.
   * Needs the table "w" (nx,nkr,nm float64, solvers.py w) uploaded. */
  int static_kick;
} chimera_engine_config;

typedef struct chimera_engine chimera_engine;

int chimera_engine_create(const chimera_engine_config* cfg, chimera_engine** out);
int chimera_engine_destroy(chimera_engine* e);
/* named arrays: grids "J" "Rho" "BckGrndRho" "EB"; spectral "EG_fb" "J_fb" "B_fb" "Rho_fb"
 * "gradRho_fb_prv" "gradRho_fb_nxt" "vec_fb"; tables "InCurr" "Out" "DpS2S" "DmS2S" "kx" "kx_base"
 * "DepFact" "PoissFact" "PSATD_E" "PSATD_G" "CPSATD1" "CPSATD2" "Rgrid"; kx-slab engines also "slab_rows"
 * (int64 global row of each local row), "gather_map" (int64, nx entries: rank * nx_slab + local row of each
 * global row), "EB_slab", "EB_gath".  Shapes as in solvers.py:160-212 (spectral arrays: nx_slab rows);
 * `src`/`dst` may be host or device pointers; nbytes must equal the array size. */
int chimera_engine_upload(chimera_engine* e, const char* name, const void* src, chb_i64 nbytes);
int chimera_engine_download(chimera_engine* e, const char* name, void* dst, chb_i64 nbytes);
int chimera_engine_array(chimera_engine* e, const char* name, void** dev_ptr, chb_i64* nbytes);
/* particles: (3,np) Fortran-ordered coords / coords_halfstep / momenta and weights(np), host or device;
 * push_fact = 2 pi Charge / Mass (species.py:64); still != 0: never pushed nor deposited as current */
int chimera_engine_add_species(chimera_engine* e, const double* coords, const double* coords_half,
                               const double* momenta, const double* weights, chb_i64 np, double push_fact,
                               int still, chb_i64 capacity, int* id);
/* external-field device of a species (species.py:55 Args['Devices'], applied between gather and push as in
 * species.py:258-277): kind 1 undul_analytic, 2 undul_analytic_taper, 3 undul_mapped, 4 undul_mapped_tap,
 * 5 planewave, 6 gaussbeam; params as the Fortran `params` array, a0 / map only where the routine has them;
 * species -1 = every non-still species; at most 4 devices per species */
int chimera_engine_add_device(chimera_engine* e, int species, int kind, double a0, const double* params, int nparams,
                              const double* map, chb_i64 nx);
/* time seen by time-dependent devices (i_step * TimeStep) in phases driven through chimera_engine_run;
 * chimera_engine_step / _step_host set it themselves from their istep argument */
int chimera_engine_set_time(chimera_engine* e, double t);
/* ---- moving window on the device (chimera_main.py:250-304 frame_act, stage 1) ----
 * damp_field: solvers.py:619 (x-space window `filtr` of nxfilt points on E and G; mode 0 left, 1 right, 2 both;
 *             filtr may be a host or device pointer); not available on kx-slab engines (needs the full x-FFT)
 * move_window: chimera_main.py:286 (leftX, rightX += shiftX)
 * append_particles: species.py:218 add_particles ((3,n) Fortran-ordered coords / momenta, weights(n); host or
 *             device pointers; coords_halfstep = coords; the species grows as needed); call _sort before depositing
 * sort: species.py:351 chunk_and_damp with SimDom = [leftX + left_margin, rightX, 0, upperR^2] */
int chimera_engine_damp_field(chimera_engine* e, const double* filtr, chb_i64 nxfilt, int mode);
/* Solver.damp_field on a kx-slab engine (the x-space window needs every kx row): damp_prepare allocates "EG_gath",
 * "EG_full" and "kx_full" (upload the full kx); the caller all-gathers the "EG_fb" slabs of all ranks into "EG_gath"
 * ([rank][(nx_slab, nkr, nm, 6)]); damp_field_slab rebuilds the full rows, applies fb_filtr and keeps this rank's rows */
int chimera_engine_damp_prepare(chimera_engine* e);
int chimera_engine_damp_field_slab(chimera_engine* e, const double* filtr, chb_i64 nxfilt, int mode);
/* A window that moves EVERY step (MovingFrame 'Steps': 1, e.g. the FEL runs, doc/tests/fel-testrun.py:61-63) inside
 * chimera_engine_step: the grid origin advances by shift_stage1 before push_coords (ChimeraRun.frame_act stage 1,
 * chimera_main.py:83,292-302) and by shift_stage2 between dep_curr and dep_dens (stage 2 of a 'Staged' frame,
 * :87,303).  init_Moving_Frames (:40-51): 'Staged' -> both = Velocity*TimeStep/2, otherwise (Velocity*TimeStep, 0).
 * The fused particle kernel deposits on the moved grids, so multi-step calls stay fused.  (0, 0) switches it off. */
int chimera_engine_set_window(chimera_engine* e, double shift_stage1, double shift_stage2);
int chimera_engine_move_window(chimera_engine* e, double shiftX);
int chimera_engine_append_particles(chimera_engine* e, int species, const double* coords, const double* momenta,
                                    const double* weights, chb_i64 n);
int chimera_engine_sort(chimera_engine* e, int on_halfstep, double left_margin);
/* chunk_and_damp called by a moving window (chimera_main.py:258-260): radial limit upper_r2 = the species' upperR^2
 * (species.py:92,376), one radial cell inside the solver's Rgrid.max() when both use the same Grid */
int chimera_engine_sort_window(chimera_engine* e, int on_halfstep, double left_margin, double upper_r2);
/* ---- integrated diagnostics on the device (moduls/diagnostics.py) ----
 * field_energy: nrg_out (diagnostics.py:109) before its roll: out[kx] = sum_{kr,m} EnergyFact[kx,kr,m] *
 *               sum_{c<3} |EG_fb[kx,kr,m,c]|^2; energy_fact (nx,nkr,nm) float64 host/device on the first call, NULL later
 * beam_moments: the sums behind get_beam_envelops (diagnostics.py:174) on coords_halfstep / momenta:
 *               out[0] = sum w; per axis c: out[1+5c..5+5c] = sum w x, w x^2, w p^2, w x p, w p
 * spectrum:     weighted histogram (numpy.histogram rule) of gamma (quantity 0) or p_x (quantity 1)
 * lineout:      out[ix] = A[ix, ir, m, l] (complex, nx values) of a named grid or spectral array */
int chimera_engine_field_energy(chimera_engine* e, const double* energy_fact, double* out);
int chimera_engine_beam_moments(chimera_engine* e, int species, double* out16);
int chimera_engine_spectrum(chimera_engine* e, int species, int quantity, double lo, double hi, chb_i64 nbins, double* hist);
int chimera_engine_lineout(chimera_engine* e, const char* name, chb_i64 ir, chb_i64 m, chb_i64 l, double* out);
int chimera_engine_species_count(chimera_engine* e, int id, chb_i64* np);
int chimera_engine_get_species(chimera_engine* e, int id, double* coords, double* coords_half, double* momenta,
                               double* weights);
/* IndInChunk(0:nchnk) of the last re-binning (particle_tools.f90:155) */
int chimera_engine_get_chunks(chimera_engine* e, int id, int* ind);
enum chimera_engine_phase {
  CHB_PUSH_COORDS = 0, /* species.py:300 push_coords                                              */
  CHB_SORT = 1,        /* species.py:351 chunk_and_damp; arg: 0 bin on coords, 1 on coords_halfstep */
  CHB_DEPOSIT_J = 2,   /* chimera_main.py:153 dep_curr (J zeroed first, ghost row folded)         */
  CHB_DEPOSIT_RHO = 3, /* chimera_main.py:183 dep_dens; arg: 1 = start from BckGrndRho, 0 = from 0 */
  CHB_DEPOSIT_BG = 4,  /* chimera_main.py:220 dep_bg: still species -> BckGrndRho                 */
  CHB_FB_IN_J = 5,     /* solvers.py:407 fb_curr_in                                               */
  CHB_FB_IN_RHO = 6,   /* chimera_main.py:110 + solvers.py:422 fb_dens_in + :517 FBGradDens       */
  CHB_POISSON = 7,     /* solvers.py:301 poiss_corr                                               */
  CHB_MAXWELL = 8,     /* solvers.py:281 maxwell_solver                                           */
  CHB_INIT_PUSH = 9,   /* solvers.py:333 maxwell_solver_stat with the uploaded CPSATD1/2           */
  CHB_FIELDS_OUT = 10, /* solvers.py:536 G2B_FBRot + :450 fb_fld_out                              */
  CHB_GATHER_PUSH = 11,/* chimera_main.py:139 proj_fld + devices + species.py:279 push_velocs; arg: dt_frac */
  CHB_ADD_BG = 12,     /* Rho += BckGrndRho (for ranks that deposited from zero before an all-reduce) */
  CHB_FIELDS_OUT_A = 13, /* kx-slab mode: G2B_FBRot + backward DHT of this rank's rows -> "EB_slab"; arg 0: E and B,
                            1: the E half only, 2: the B half only (each half of EB_slab is contiguous)        */
  CHB_FIELDS_OUT_B = 14, /* kx-slab mode: rows of the all-gathered "EB_gath" -> EB, inverse x-FFT, eb_correction; arg 0:
                            EB_gath = [rank][(nx_slab,Nr,M,6)]; 1 | 2: one half of EB_gath = [half][rank][(..,3)]  */
  CHB_PARTICLES_FUSED = 15, /* gather + device + push_velocs of one step and push_coords + dep_curr + dep_dens of the
                             next in one kernel (the per-particle work between two field solves); arg as DEPOSIT_RHO */
  CHB_STATIC_FIELDS = 16, /* chimera_main.py:118-125 update_fields with 'StaticKick' (needs every kx row)        */
  CHB_WINDOW = 17,        /* one stage of a window that moves every step (chimera_main.py:286); arg: 1 | 2       */
  CHB_GATHER_PUSH_COORDS = 18, /* gather + push of step k and push_coords of step k+1, no deposit (before a sort);
                                  applies window stage 1                                                          */
  CHB_DEPOSIT_FUSED = 19, /* dep_curr + dep_dens from the stored x_half / x / p in one kernel (after a sort); arg != 0:
                             rho starts from BckGrndRho; applies window stage 2                                   */
  /* column-block dataflow of the multi-rank solve (chimera_engine_set_colflow) */
  CHB_COL_FWD = 20,       /* x-FFT of the reduce-scattered column block + rows sorted by destination rank; arg 0 J, 1 Rho */
  CHB_FB_IN_COL = 21,     /* forward DHT of the slab received through the all-to-all (+ fb_grad for Rho); arg 0 J, 1 Rho  */
  CHB_COL_BWD = 22,       /* received (rank, column, row) blocks -> own column block of EB, normalised, inverse x-FFT      */
  CHB_EB_FINISH = 23,     /* ghost rows of eb_correction on the all-gathered EB                                           */
  CHB_NPHASES = 24
};
int chimera_engine_run(chimera_engine* e, int phase, double arg);
/* nsteps x make_step on the engine's stream; istep0 = index of the first step (re-binning cadence) */
int chimera_engine_step(chimera_engine* e, chb_i64 istep0, chb_i64 nsteps);
int chimera_engine_sync(chimera_engine* e);
/* multi-step calls fuse the particle work between two field solves into one kernel (default on) */
int chimera_engine_set_fuse(chimera_engine* e, int on);
/* replay the fused step (particle kernel + spectral update) as a CUDA graph between two re-binnings (default on; used when
   no window moves every step and no device field depends on time) */
int chimera_engine_set_graph(chimera_engine* e, int on);
int chimera_engine_graph_info(chimera_engine* e, int* ngraphs, int* state); /* cached graphs; state 1 warm, -1 capture failed */
/* chimera_engine_step leaves the gather + push_velocs that closes its last step pending, and the next chimera_engine_step
   runs it inside its first fused kernel (a loop of one-step calls then runs the kernels of one long call).  Every other
   engine entry point that reads or changes state (download, array, sync, run, diagnostics, ...) completes it first, so
   the deferral is not observable through the API; device pointers from chimera_engine_array are current after
   chimera_engine_sync.  Default on; 0 completes every call eagerly. */
int chimera_engine_set_lazy_tail(chimera_engine* e, int on);
int chimera_engine_set_static_px(chimera_engine* e, const double* px, int n); /* 'StaticKick' across ranks: PXmean per species */
int chimera_engine_set_colflow(chimera_engine* e, int rank, int world); /* column-block dataflow buffers (kx-slab engines) */
/* One make_step (chimera_main.py:82-92) with the PIC state in HOST buffers, the reference's calling model:
 * coords/momenta (3,np) Fortran-ordered in-out, coords_half (3,np) out, weights (np) in (rewritten in the
 * new particle order on re-binning steps), EG_fb (nx,nkr,nm,6) and gradRho_fb_nxt (nx,nkr,nm,3) complex
 * in-out (either may be NULL: the engine-resident copy is used and nothing is copied).  Operators, tables,
 * BckGrndRho and still species stay resident.  Copies run on two extra streams and overlap the kernels;
 * pass page-locked buffers (chimera_host_register) for full PCIe speed.  The call is synchronous.
 * coords_half may be NULL: the centred positions are then not copied back (a caller that leaves deposit and re-binning
 * to the engine never reads them; saves 24 B per particle on the device->host link).
 * rebin != 0 forces a re-binning in this step (required after the caller reordered or added particles);
 * *np_out = particles kept (re-binning culls the ones that left the domain). */
int chimera_engine_step_host(chimera_engine* e, int species, double* coords, double* coords_half, double* momenta,
                             double* weights, chb_i64 np, chb_i64* np_out, double* EG_fb, double* gradRho_fb_nxt,
                             chb_i64 istep, int rebin);
/* the same in two halves, for multi-GPU runs: _begin copies the particles in, pushes coordinates and
 * deposits J / Rho (with from_bg = 0 the background charge is left out: ranks > 0 before an all-reduce,
 * set with chimera_engine_set_rho_from_bg); the caller all-reduces the grids on the engine stream; _end
 * runs the field update, gather + push and the copies out, and synchronises. */
int chimera_engine_step_host_begin(chimera_engine* e, int species, double* coords, double* coords_half,
                                   double* momenta, double* weights, chb_i64 np, double* EG_fb,
                                   double* gradRho_fb_nxt, chb_i64 istep, int rebin);
/* kx-slab engines: _begin, all-reduce, _mid, all-gather "EB_slab" -> "EB_gath", _end */
int chimera_engine_step_host_mid(chimera_engine* e);
int chimera_engine_step_host_end(chimera_engine* e, chb_i64* np_out);
int chimera_engine_set_rho_from_bg(chimera_engine* e, int from_bg);
/* page-lock / unlock a host buffer the caller owns (cudaHostRegister) */
int chimera_host_register(void* ptr, chb_i64 nbytes);
int chimera_host_unregister(void* ptr);
int chimera_engine_set_stream(chimera_engine* e, void* cuda_stream);
/* per-phase device time (CUDA events on the engine stream), accumulated since the last reset */
int chimera_engine_profile(chimera_engine* e, int on);
int chimera_engine_timings(chimera_engine* e, double* ms /* CHB_NPHASES */, chb_i64* calls /* CHB_NPHASES */, int reset);

#ifdef __cplusplus
}
#endif
#endif
