/*
 * nrsb200.h -- C ABI of the B200-native pressure-Poisson path (drop-in for nekRS v23.0's
 * elliptic / oogs / linAlg device path).  Plain C: pointers and sizes only.
 *
 * Conventions
 *   - Every function returns NRSB_OK (0) or a negative NRSB_ERR_* code; never exits/aborts
 *     (the reference aborts through nrsAbort/MPI_Abort, nrssys.hpp:86-108; the adapter maps codes).
 *     nrsb_last_error_string() describes the last failure on the calling thread.
 *   - Pointers named d_* / documented "device" are CUDA device pointers (what occa::memory::ptr()
 *     returns); "host" pointers are ordinary memory.  `stream` is a cudaStream_t passed as void*.
 *   - Types as in nrssys.hpp: dfloat=double, pfloat=float, dlong=int32, hlong=int64.
 *   - `precision` is sizeof(element): 8 (dfloat) or 4 (pfloat) -- the reference selects the same
 *     two instances by kernel-name suffix ("..._<N>pfloat", registerEllipticKernels.cpp:166-195).
 *   - One handle <-> one host thread <-> one device.  No hidden device-wide synchronisation except
 *     where a scalar is returned to the host.
 *
 * Each entry point cites the reference interface it replaces (file:line under /root/reference).
 */
#ifndef NRSB200_H
#define NRSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRSB_OK 0
#define NRSB_ERR_INVALID (-1) /* bad argument / unsupported configuration */
#define NRSB_ERR_CUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define NRSB_ERR_NOMEM (-3)
#define NRSB_ERR_DIVERGED (-4) /* NaN residual (PCG.cpp:193-195) */

typedef int32_t nrsb_dlong;
typedef int64_t nrsb_hlong;

const char* nrsb_last_error_string(void);
const char* nrsb_version(void);
/* sizeof of the structs of this header as the library was built ("nrsb_elliptic_config", "nrsb_shared_topology"; 0 for an
 * unknown name): lets a binding written in another language check its mirror of the layout */
int nrsb_sizeof(const char* struct_name);

/* ---- device plumbing (platform->device.malloc / occa::memory::copyFrom/copyTo) ------------- */
int nrsb_device_count(int* count);
int nrsb_set_device(int device);
int nrsb_malloc(void** d_ptr, size_t bytes);
int nrsb_free(void* d_ptr);
int nrsb_malloc_host(void** h_ptr, size_t bytes); /* pinned */
int nrsb_free_host(void* h_ptr);
int nrsb_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int nrsb_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int nrsb_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
int nrsb_memset(void* d_dst, int value, size_t bytes, void* stream);
int nrsb_stream_synchronize(void* stream);
int nrsb_device_synchronize(void);
int nrsb_l2_flush(void* stream); /* writes, then reads back, a 2xL2 scratch buffer: cold AND clean L2 (benchmark hygiene) */
/* CUDA events for device-side timing on the launching stream (timer::tic/toc, timer.cpp:199-233) */
int nrsb_event_create(void** event);
int nrsb_event_destroy(void* event);
int nrsb_event_record(void* event, void* stream);
int nrsb_event_synchronize(void* event);
int nrsb_event_elapsed_ms(void* start, void* stop, float* ms);
/* cudaProfilerStart/Stop, for `ncu --profile-from-start off` */
int nrsb_profiler_start(void);
int nrsb_profiler_stop(void);
int nrsb_stream_create(void** stream);
int nrsb_stream_destroy(void* stream);

/* ---- kernel-level entry points (one per reference kernel, reference argument order) -------- */

/* ellipticPartialAxCoeffHex3D  (kernels/elliptic/ellipticPartialAxCoeffHex3D.okl:1..1444, serial
 * twin .c:2-12; called from ellipticAx, ellipticOperator.cpp:83-93).
 * Nq = N+1 in 2..12.  D_host: HOST pointer to the Nq*Nq row-major derivative matrix
 * D[i][m] = l_m'(r_i) (mesh->D); the reference's extra `S = D^T` argument is derived from it.
 * poisson != 0 : p_poisson build (lambda1 ignored); lambda_field != 0 : p_lambda=1 (per-node
 * coefficients, `ELLIPTIC COEFF FIELD`), else lambda0[0]/lambda1[0] are scalars on the device.
 * variant: 0 slab kernel, 1 pencil kernel, -1 library default for (Nq, precision). */
int nrsb_ellipticPartialAxCoeffHex3D(int Nq, int precision, int variant, nrsb_dlong Nelements, nrsb_dlong offset,
                                     nrsb_dlong loffset, const nrsb_dlong* d_elementList, const void* d_ggeo,
                                     const void* D_host, const void* d_lambda0, const void* d_lambda1, int poisson,
                                     int lambda_field, const void* d_q, void* d_Aq, void* stream);

/* mask  (kernels/core/mask.okl; ellipticApplyMask.cpp:3-27) */
int nrsb_mask(int precision, nrsb_dlong Nmasked, const nrsb_dlong* d_maskIds, void* d_q, void* stream);

/* gatherScatterMany_{float,double}Add  (3rd_party/gslib/ogs/okl/gatherScatterMany.okl) */
int nrsb_gatherScatterMany_add(int precision, nrsb_dlong Ngather, int Nentries, nrsb_dlong stride,
                               const nrsb_dlong* d_gatherStarts, const nrsb_dlong* d_gatherIds, void* d_q,
                               void* stream);

/* linAlg (src/linAlg/linAlg.hpp:33-282; kernels/linAlg/*.okl).  Scalars by value. */
int nrsb_fill(int precision, nrsb_dlong N, double alpha, void* d_a, void* stream);
int nrsb_axpby(int precision, nrsb_dlong N, double alpha, const void* d_x, double beta, void* d_y, void* stream);
int nrsb_axpbyMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong offset, double alpha, const void* d_x,
                   double beta, void* d_y, void* stream);
int nrsb_axmyz(int precision, nrsb_dlong N, double alpha, const void* d_x, const void* d_y, void* d_z, void* stream);
int nrsb_scale(int precision, nrsb_dlong N, double alpha, void* d_x, void* stream);
/* the rest of the family the elliptic path and its callers use (linAlg.hpp:70-142): a *= alpha per field,
 * a += alpha, |a|, y = alpha x y (axmy / paxmy; Many: mode 1 x per field, mode 0 one x for all fields),
 * z = alpha x y per field (axmyzMany / paxmyzMany), y = alpha / y (ady / adyMany / padyMany), y = alpha x / y,
 * z = alpha x + beta y per field (axpbyzMany) */
int nrsb_scaleMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong fieldOffset, double alpha, void* d_a,
                   void* stream);
int nrsb_add(int precision, nrsb_dlong N, double alpha, void* d_a, void* stream);
int nrsb_abs(int precision, nrsb_dlong N, void* d_a, void* stream);
int nrsb_axmy(int precision, nrsb_dlong N, double alpha, const void* d_x, void* d_y, void* stream);
int nrsb_axmyMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong offset, int mode, double alpha, const void* d_x,
                  void* d_y, void* stream);
int nrsb_axmyzMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong offset, double alpha, const void* d_x,
                   const void* d_y, void* d_z, void* stream);
int nrsb_adyMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong offset, double alpha, void* d_y, void* stream);
int nrsb_axdy(int precision, nrsb_dlong N, double alpha, const void* d_x, void* d_y, void* stream);
int nrsb_axpbyzMany(int precision, nrsb_dlong N, int Nfields, nrsb_dlong offset, double alpha, const void* d_x,
                    double beta, const void* d_y, void* d_z, void* stream);
/* ellipticBlockPartialAxCoeffHex3D (kernels/elliptic/ellipticBlockPartialAxCoeffHex3D.okl, serial twin .c:2-152;
 * registered for the velocity solve, registerEllipticKernels.cpp): three Helmholtz operators sharing ggeo,
 * Aq[id + f*offset] = (D^T lambda0_f G D + lambda1_f GwJ) q[id + f*offset], f = 0..2; coefficients at
 * lambda[p_lambda*id + f*loffset] (lambdaField = p_lambda).  The fields of an element run as neighbouring blocks, so
 * the geometric factors cross HBM once. */
int nrsb_ellipticBlockPartialAxCoeffHex3D(int Nq, int precision, nrsb_dlong Nelements, nrsb_dlong offset,
                                          nrsb_dlong loffset, const nrsb_dlong* d_elementList, const void* d_ggeo,
                                          const void* D_host, const void* d_lambda0, const void* d_lambda1,
                                          int lambdaField, const void* d_q, void* d_Aq, void* stream);
/* ellipticStressPartialAxCoeffHex3D (kernels/elliptic/ellipticStressPartialAxCoeffHex3D.okl, serial twin .c:1-169;
 * the AxKernel of a block solver with stressForm, ellipticSetup.cpp:240-249): the coupled viscous-stress operator
 * A (u,v,w) = -div(lambda0 (grad q + grad q^T)) + lambda1 q in weak form.  d_vgeo: 12 planes per element in the
 * reference's order (rx,ry,rz,sx,sy,sz,tx,ty,tz,J,JW,1/JW: mesh3D.h:82-93); fields `offset` apart, coefficients at
 * lambda[p_lambda*id + f*loffset]. */
int nrsb_ellipticStressPartialAxCoeffHex3D(int Nq, int precision, nrsb_dlong Nelements, nrsb_dlong offset,
                                           nrsb_dlong loffset, const nrsb_dlong* d_elementList, const void* d_vgeo,
                                           const void* D_host, const void* d_lambda0, const void* d_lambda1,
                                           int lambdaField, const void* d_q, void* d_Aq, void* stream);
/* ellipticBlockBuildDiagonalHex3D (kernels/elliptic/ellipticBlockBuildDiagonalHex3D.okl; ellipticUpdateJacobi.cpp:
 * 38-47,66-75): Aq[id + l*offset] = diag(D^T lambda0 G D)[id] (+ lambda1 GwJ), l < Nfields.  lambdaField = 1:
 * lambda0/lambda1 are per-node fields read at id + l*loffset (the reference's layout), 0: one value each.
 * D_host: HOST, row-major D[i][m] in the precision of the call. */
int nrsb_ellipticBlockBuildDiagonalHex3D(int Nq, int precision, nrsb_dlong Nelements, int Nfields, nrsb_dlong offset,
                                         nrsb_dlong loffset, const void* d_ggeo, const void* D_host,
                                         const void* d_lambda0, const void* d_lambda1, int poisson, int lambdaField,
                                         void* d_Aq, void* stream);
int nrsb_copyDfloatToPfloat(nrsb_dlong N, const double* d_x, float* d_y, void* stream);
int nrsb_copyPfloatToDfloat(nrsb_dlong N, const float* d_x, double* d_y, void* stream);
/* reductions return the (rank-local) value to the host: they synchronise the stream. */
int nrsb_weightedInnerProdMany(nrsb_dlong N, int Nfields, nrsb_dlong offset, const double* d_w, const double* d_x,
                               const double* d_y, double* result, void* stream);
int nrsb_weightedNorm2Many(nrsb_dlong N, int Nfields, nrsb_dlong offset, const double* d_w, const double* d_x,
                           double* result, void* stream);
int nrsb_weightedInnerProdMulti(nrsb_dlong N, int NVec, nrsb_dlong offset, const double* d_w, const double* d_x,
                                const double* d_y, double* results, void* stream);
int nrsb_sum(int precision, nrsb_dlong N, const void* d_x, double* result, void* stream);
/* ellipticBlockUpdatePCG (kernels/elliptic/ellipticBlockUpdatePCG.okl:26-111, PCG.cpp:33-83):
 * r -= alpha Ap, returns sum w r^2; if d_x and d_p are non-NULL also x += alpha p (fused). */
int nrsb_ellipticBlockUpdatePCG(nrsb_dlong N, nrsb_dlong offset, const double* d_invDegree, const double* d_Ap,
                                double alpha, double* d_r, const double* d_p, double* d_x, double* rdotr,
                                void* stream);
int nrsb_updateChebyshev(nrsb_dlong N, float dCoeff, float rCoeff, const float* d_SAd, float* d_d, float* d_r,
                         float* d_x, void* stream);
int nrsb_updateFourthKindChebyshev(nrsb_dlong N, float beta, const float* d_Ad, const float* d_d, float* d_r,
                                   float* d_x, void* stream);
int nrsb_gramSchmidtOrthogonalization(nrsb_dlong N, nrsb_dlong offset, int gmresSize, const double* d_weights,
                                      const double* d_y, const double* d_V, double* d_w, double* result,
                                      void* stream);
int nrsb_updatePGMRESSolution(nrsb_dlong N, nrsb_dlong offset, int gmresSize, const double* d_y, const double* d_Z,
                              double* d_x, void* stream);
int nrsb_fusedResidualAndNorm(nrsb_dlong N, nrsb_dlong offset, const double* d_weights, const double* d_b,
                              const double* d_Ax, double* d_r, double* result, void* stream);

/* Schwarz / FDM  (kernels/elliptic/{preFDM,fusedFDM,postFDM}.okl; ellipticMultiGridSchwarz.cpp:1056-1156).
 * Nq is the ELEMENT Nq; the extended size is Nq+2.  pfloat only (production precision). */
int nrsb_preFDM(int Nq, nrsb_dlong Nelements, const float* d_u, float* d_work1, void* stream);
int nrsb_fusedFDM(int Nq, int restrict_, nrsb_dlong Nelements, const nrsb_dlong* d_elementList, float* d_Su,
                  const float* d_Sx, const float* d_Sy, const float* d_Sz, const float* d_invL, const float* d_wts,
                  float* d_u, void* stream);
/* fusedFDM kernel variant (the reference autotunes fusedFDM_v0..v4, benchmarkFDM.cpp:38-296):
 * 0 = one pencil per thread, 1 = register-blocked pencil pairs (default, even extended sizes) */
int nrsb_set_fdm_variant(int variant);
/* coarse-grid solve variant (stands where coarseLevel_t::solve calls BoomerAMG, MG/coarseLevel.cpp:182-222):
 * 1 = whole PCG solve in one thread-block-cluster kernel when the coarse grid fits (default, one rank),
 * 0 = two launches per iteration,
 * 2 = as 1 but with the SpMV input in an L2-resident global buffer instead of distributed shared memory
 *     (what large coarse grids fall back to; selectable so that tests can cover it; set before setup) */
int nrsb_set_coarse_variant(int variant);
int nrsb_postFDM(int Nq, nrsb_dlong Nelements, float* d_work1, float* d_work2, float* d_Su, const float* d_wts,
                 void* stream);

/* p-multigrid transfers (kernels/elliptic/ellipticPrecon{Coarsen,Prolongate}Hex3D.okl).
 * R_host: HOST pointer, R[NqC][NqF] pfloat (ellipticMultiGridLevelSetup.cpp:238-258). */
int nrsb_ellipticPreconCoarsenHex3D(int NqF, int NqC, nrsb_dlong Nelements, const float* R_host, const float* d_qf,
                                    float* d_qc, void* stream);
int nrsb_ellipticPreconProlongateHex3D(int NqF, int NqC, nrsb_dlong Nelements, const float* R_host,
                                       const float* d_qc, float* d_qN, void* stream);

/* geometricFactorsHex3D (kernels/mesh/geometricFactorsHex3D.okl:26-142): ggeo[E][7][Np], host D/gllw */
int nrsb_geometricFactorsHex3D(int Nq, nrsb_dlong Nelements, const double* D_host, const double* gllw_host,
                               const double* d_x, const double* d_y, const double* d_z, double* d_ggeo,
                               double* d_jacobian, void* stream);

/* ---- gather-scatter handles  (ogs_t / oogs_t, 3rd_party/gslib/ogs/ogs.hpp:145-295) ------------ */
typedef struct nrsb_ogs* nrsb_ogs_t;

/* topology of ids shared with other ranks, discovered by the bootstrap layer (the reference gets it
 * from gslib gs_setup over MPI, ogsSetup.cpp:150-175).  NULL / nranks==1 for a single rank. */
typedef struct {
  int rank, nranks;
  int64_t nShared;
  const nrsb_hlong* sharedIds; /* ascending global ids present on this rank AND another one */
  const int32_t* sharerOffsets; /* nShared+1 */
  const int32_t* sharerRanks;   /* ascending ranks sharing each id, this rank included */
} nrsb_shared_topology;

/* ogsSetup (ogsSetup.cpp:111-397).  ids: HOST, 0 = ignored node. */
int nrsb_ogs_setup(nrsb_dlong N, const nrsb_hlong* ids_host, const nrsb_shared_topology* topo, nrsb_ogs_t* out);
int nrsb_ogs_destroy(nrsb_ogs_t ogs);
/* sizes: {N, Nlocal, NlocalGather, Nhalo, NhaloGather, nPairs, nQuads, nOcts, nGen} */
int nrsb_ogs_sizes(nrsb_ogs_t ogs, int64_t sizes[9]);
/* copies of the reference-layout maps for parity tests (HOST buffers, caller sized via nrsb_ogs_sizes) */
int nrsb_ogs_get_local_maps(nrsb_ogs_t ogs, nrsb_dlong* gatherOffsets, nrsb_dlong* gatherIds);
int nrsb_ogs_get_halo_maps(nrsb_ogs_t ogs, nrsb_dlong* gatherOffsets, nrsb_dlong* gatherIds, nrsb_hlong* baseIds);
int nrsb_ogs_get_inv_degree(nrsb_ogs_t ogs, double* invDegree_host);
int nrsb_ogs_inv_degree_device(nrsb_ogs_t ogs, const double** d_invDegree, const float** d_invDegreePfloat);
/* oogs::startFinish (oogs.cpp:823-837) with ogsAdd on one rank: in-place v <- Q Q^T v, k fields.
 * d_maskIds may be NULL; when given the ids are zeroed in the same launch (they must not belong to
 * any gather row, i.e. the handle was built from masked ids as in ellipticOgs.cpp:126-131). */
int nrsb_ogs_gather_scatter(nrsb_ogs_t ogs, int precision, int k, nrsb_dlong stride, nrsb_dlong Nmasked,
                            const nrsb_dlong* d_maskIds, void* d_v, void* stream);

/* ---- communicator (platform->comm + MPI on this path) ---------------------------------------- */
/* One process per GPU.  The bootstrap layer (torch.distributed, MPI, ...) supplies two host
 * collectives used ONLY at setup time (IPC handle / table exchange); every data-path exchange
 * afterwards is device-initiated over NVLink (oogs halo stores, one-shot scalar all-reduce).
 * allgather(buf, bytes_per_rank, user): in place, buf holds nranks blocks, block `rank` is valid. */
typedef struct nrsb_comm* nrsb_comm_t;
typedef void (*nrsb_allgather_fn)(void* buf, size_t bytes_per_rank, void* user);
typedef void (*nrsb_barrier_fn)(void* user);
int nrsb_comm_create(int rank, int nranks, nrsb_allgather_fn allgather, nrsb_barrier_fn barrier, void* user,
                     nrsb_comm_t* out);
int nrsb_comm_destroy(nrsb_comm_t comm);
/* MPI_Allreduce(SUM) of n <= 16 doubles through the device one-shot path (testing / setup) */
int nrsb_comm_allreduce_sum(nrsb_comm_t comm, int n, double* values_host);

/* multi-rank oogs (oogs::setup / startFinish, oogs.cpp:336-837): exchange over NVLink peer windows */
typedef struct nrsb_oogs* nrsb_oogs_t;
int nrsb_oogs_setup(nrsb_ogs_t ogs, nrsb_comm_t comm, int maxFields, nrsb_oogs_t* out);
int nrsb_oogs_destroy(nrsb_oogs_t oogs);
int nrsb_oogs_start(nrsb_oogs_t oogs, int precision, int k, nrsb_dlong stride, void* d_v, void* stream);
int nrsb_oogs_finish(nrsb_oogs_t oogs, int precision, int k, nrsb_dlong stride, void* d_v, void* stream);

/* ---- elliptic solver handle  (elliptic_t, elliptic.h:73-238) ------------------------------------ */
typedef struct nrsb_elliptic* nrsb_elliptic_t;

typedef struct {
  int N;                        /* polynomial order of the solve (mesh->N) */
  nrsb_dlong Nelements;         /* elements on this rank */
  const double *x, *y, *z;      /* HOST node coordinates [Nelements*Np] (mesh->x/y/z) */
  const nrsb_hlong* globalIds;  /* HOST C0 numbering, >= 1 (mesh->globalIds) */
  const int* EToB;              /* HOST [Nelements*6] boundary flag per face: 0 interior, 1 DIRICHLET, 4 NEUMANN */
  const nrsb_shared_topology* topo; /* sharing of globalIds across ranks; NULL on one rank */
  /* coarser p-multigrid meshes: numbering (and sharing) at each level order (createMeshMG) */
  int nLevels;
  const int* levelOrders;
  const nrsb_hlong* const* levelGlobalIds;
  const nrsb_shared_topology* const* levelTopo; /* NULL on one rank */
  /* setupAide options of the solver section, "KEY=VALUE" lines, keys as in the reference
   * (SOLVER, PRECONDITIONER, MULTIGRID SMOOTHER, MAXIMUM ITERATIONS, SOLVER TOLERANCE, ...) */
  const char* options;
  int poisson;                  /* elliptic->poisson */
  double lambda0, lambda1;      /* constant coefficients (o_lambda0/o_lambda1) */
  nrsb_comm_t comm;             /* NULL on one rank */
  const char* name;             /* "pressure" */
  /* block solver (elliptic->Nfields, elliptic->stressForm; ellipticSetup.cpp:81-131,209-249).  Nfields = 0 or 1: scalar
   * solve.  Nfields = 3: three fields `fieldOffset` apart in every vector, EToB then holds Nfields*Nelements*6 flags
   * (field-major, as elliptic->EToB), PRECONDITIONER = JACOBI or NONE, SOLVER = PCG; stressForm selects
   * ellipticStressPartialAxCoeffHex3D instead of ellipticBlockPartialAxCoeffHex3D.  blockLambda0/1: HOST, one constant
   * per field (NULL: lambda0/lambda1 for every field); per-node fields go through nrsb_elliptic_set_coeff_field with
   * loffset = fieldOffset. */
  int Nfields;
  int stressForm;
  const double* blockLambda0;
  const double* blockLambda1;
} nrsb_elliptic_config;

/* ellipticSolveSetup (ellipticSetup.cpp:116-327) */
int nrsb_elliptic_setup(const nrsb_elliptic_config* cfg, nrsb_elliptic_t* out);
int nrsb_elliptic_destroy(nrsb_elliptic_t h);
/* ellipticSolve (ellipticSolve.cpp:32-190): d_r = rhs (overwritten), d_x = initial guess in / solution out;
 * both fieldOffset doubles on the device.  Returns Niter and the three norms of elliptic.h:89-90. */
int nrsb_elliptic_solve(nrsb_elliptic_t h, double* d_r, double* d_x, int* Niter, double* res00Norm, double* res0Norm,
                        double* resNorm);
/* same with HOST vectors of Nlocal doubles (copies in and out included) */
int nrsb_elliptic_solve_host(nrsb_elliptic_t h, const double* rhs_host, double* x_host, int* Niter, double* res00Norm,
                             double* res0Norm, double* resNorm);
/* ellipticOperator (ellipticOperator.cpp:117-172): Aq = Q Q^T mask (A q); level = multigrid level index
 * (0 = the fp64 solver itself when precision = 8; fp32 instances live on the MG levels) */
int nrsb_elliptic_operator(nrsb_elliptic_t h, int level, int precision, const void* d_q, void* d_Aq, int masked);
int nrsb_elliptic_operator_host(nrsb_elliptic_t h, const double* q_host, double* Aq_host);
/* fp64 operator + q^T A q in one call: the p^T A p of PCG.cpp:150-157 (weightedInnerProdMany of p and Ap with
 * invDegree), here taken from the axhelm launch itself when FUSED DOT AX is on (*fromAxLaunch = 1) */
int nrsb_elliptic_operator_dot(nrsb_elliptic_t h, const double* d_q, double* d_Aq, int masked, double* qAq,
                               int* fromAxLaunch);
/* the same, queued: upload, operator and download of consecutive calls overlap (two staging slots, both copy
 * engines); q_host / Aq_host should be pinned and must stay valid until nrsb_elliptic_host_wait returns */
int nrsb_elliptic_operator_host_async(nrsb_elliptic_t h, const double* q_host, double* Aq_host);
int nrsb_elliptic_host_wait(nrsb_elliptic_t h);
/* device-side rendezvous of all ranks on the handle's stream (one scalar all-reduce through the peer windows, no host
 * synchronisation): work queued behind it starts within an NVLink round trip on every GPU (MPI_Barrier leaves the
 * ranks tens of microseconds apart: timeEllipticOperator, ellipticSetup.cpp:255-271, pays that once per sample) */
int nrsb_elliptic_device_barrier(nrsb_elliptic_t h);
/* ellipticAx on the full element list */
int nrsb_elliptic_ax(nrsb_elliptic_t h, int level, int precision, const void* d_q, void* d_Aq);
/* ellipticPreconditioner (ellipticPreconditioner.cpp:33-84) */
/* gather-scatter (+ Dirichlet mask) half of ellipticOperator alone: oogs::startFinish(o_Aq, Nfields, fieldOffset,
 * ogsDfloat, ogsAdd, oogs) as called at ellipticOperator.cpp:164-168, preceded by ellipticApplyMask (:158) */
int nrsb_elliptic_gather_scatter(nrsb_elliptic_t h, int level, int precision, void* d_v, int masked);
int nrsb_elliptic_preconditioner(nrsb_elliptic_t h, double* d_r, double* d_z);
/* pMGLevel::smoothSchwarz / smooth / coarsen / prolongate on level `level` (pfloat vectors) */
int nrsb_elliptic_level_op(nrsb_elliptic_t h, int level, const char* op, float* d_in, float* d_out);
/* integer / real properties: "Nlocal","fieldOffset","Nmasked","nLevels","overlap","allNeumann",
 * "NglobalGatherElements","NlocalGatherElements","coarseIterations"; per level (key "level<k>:<name>"):
 * "N","Nlocal","Nmasked","lambda1","lambda0","maxEig","downDegree","upDegree" ; "volume" */
int nrsb_elliptic_get_int(nrsb_elliptic_t h, const char* key, int64_t* value);
int nrsb_elliptic_get_real(nrsb_elliptic_t h, const char* key, double* value);
/* "level<k>:maxEig": overwrite that level's lambda_max(S A) estimate (the Chebyshev bounds follow) */
int nrsb_elliptic_set_real(nrsb_elliptic_t h, const char* key, double value);
/* arrays (HOST out): "maskIds" (int32), "invDegree" (double), "resHistory" (double, length Niter),
 * "level<k>:invDegree", "level<k>:maskIds", "level<k>:Sx|Sy|Sz|invL|wts" (float) ; returns count */
int nrsb_elliptic_get_array(nrsb_elliptic_t h, const char* key, void* out_host, int64_t capacity, int64_t* count);
int nrsb_elliptic_set_option(nrsb_elliptic_t h, const char* key, const char* value); /* before re-setup of precon */
/* variable coefficients (ELLIPTIC COEFF FIELD; p_lambda = 1 in ellipticPartialAxCoeffHex3D.okl): per-node lambda0 /
 * lambda1 (device, fp64, caller-owned like the reference's o_lambda0 / o_lambda1).  Refreshes the multigrid levels'
 * copies (ellipticMultiGridUpdateLambda, MG/ellipticMultiGridUpdateLambda.cpp) and the inverse diagonals
 * (ellipticUpdateJacobi, ellipticUpdateJacobi.cpp:87-115).  With the option ELLIPTIC PRECO COEFF FIELD = TRUE both
 * refreshes also run at the top of every nrsb_elliptic_solve (ellipticSolve.cpp:79-88). */
int nrsb_elliptic_set_coeff_field(nrsb_elliptic_t h, const double* d_lambda0, const double* d_lambda1);
int nrsb_elliptic_set_coefficients(nrsb_elliptic_t h, double lambda0, double lambda1); /* constant coefficients */
int nrsb_elliptic_update_jacobi(nrsb_elliptic_t h);  /* ellipticUpdateJacobi(elliptic) */
int nrsb_elliptic_update_lambda(nrsb_elliptic_t h);  /* ellipticMultiGridUpdateLambda(elliptic) */
int nrsb_elliptic_set_ax_variant(nrsb_elliptic_t h, int precision, int variant);
int nrsb_elliptic_set_stream(nrsb_elliptic_t h, void* stream);
/* kernel-variant autotuning as in benchmarkAx (src/bench/axHelm/benchmarkAx.cpp:140-146,289-305) */
int nrsb_elliptic_autotune(nrsb_elliptic_t h, int* variant_fp64, int* variant_fp32);

/* determineMGLevels (MG/determineMGLevels.cpp:58-95): level orders for solver order N under the given options
 * ("KEY=VALUE" lines: MULTIGRID SMOOTHER, MULTIGRID SCHEDULE); count = number of levels */
int nrsb_mg_levels(int N, const char* options, int* levels_out, int capacity, int* count);
/* 1 when every coarsen/prolongate pair and Schwarz size of that schedule is instantiated in this library */
int nrsb_mg_schedule_supported(int N, const char* options);

/* setup helpers exposed for parity tests */
int nrsb_gll(int N, double* z_host, double* w_host, double* D_host);
int nrsb_sym_generalized_eig(int n, double* A_host, double* B_host, double* lam_host);
int nrsb_spectral_radius(int n, const double* H_host, double* rho);

#ifdef __cplusplus
}
#endif
#endif /* NRSB200_H */
