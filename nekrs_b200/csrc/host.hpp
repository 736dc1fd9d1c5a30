// host.hpp -- host-side objects of the elliptic path.  Same names and roles as the reference's
// mesh_t (src/mesh/mesh.h:40-255), oogs_t (ogs.hpp:256-278), elliptic_t (elliptic.h:73-181),
// precon_t, MGSolver_t (MGSolver.hpp:42-140) and pMGLevel (ellipticMultiGrid.h:53-167); the control
// flow that uses them (ellipticSolve.cpp, PCG.cpp, PGMRES.cpp, MGSolver.cpp,
// ellipticMultiGridLevel.cpp, ellipticMultiGridSchwarz.cpp) is restated in elliptic.cpp /
// multigrid.cpp.  Device work goes only through the CUDA launchers of this library.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "gs.hpp"
#include "kernels.hpp"
#include "linalg.hpp"

namespace nrsb {

// ---------------------------------------------------------------- options (setupAide)
class options_t {
 public:
  std::map<std::string, std::string> kv;
  void setArgs(const std::string& key, const std::string& value) { kv[key] = value; }
  bool has(const std::string& key) const { return kv.count(key) != 0; }
  std::string getArgs(const std::string& key) const
  {
    auto it = kv.find(key);
    return it == kv.end() ? std::string() : it->second;
  }
  bool getArgs(const std::string& key, int& v) const;
  bool getArgs(const std::string& key, double& v) const;
  // setupAide::compareArgs (setupAide.cpp:86-97): exact match or substring
  bool compareArgs(const std::string& key, const std::string& token) const
  {
    auto it = kv.find(key);
    return it != kv.end() && it->second.find(token) != std::string::npos;
  }
};

// ---------------------------------------------------------------- device buffer
template <typename T>
struct dbuf {
  T* p = nullptr;
  size_t n = 0;
  dbuf() = default;
  dbuf(const dbuf&) = delete;
  dbuf& operator=(const dbuf&) = delete;
  ~dbuf() { cudaFree(p); }
  int alloc(size_t count, bool zero = true);
  int upload(const std::vector<T>& h);
  int upload(const T* h, size_t count);
  int download(std::vector<T>& h) const;
  void release()
  {
    cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// ---------------------------------------------------------------- communicator (comm.cu)
// One process per GPU.  Peer windows are opened once through CUDA IPC handles exchanged by the
// bootstrap layer; afterwards all data-path communication is device-initiated NVLink stores.
class comm_t {
 public:
  int rank = 0, nranks = 1;
  // scalar all-reduce windows (see linalg.cu peer_allreduce)
  dbuf<double> redSlots;                  // [2][nranks][kMaxRed]
  dbuf<unsigned long long> redFlags;      // [nranks]
  dbuf<unsigned long long> redEpoch;      // [1]
  std::vector<double*> peerRedSlots;
  std::vector<unsigned long long*> peerRedFlags;
  dbuf<double*> d_peerRedSlots;
  dbuf<unsigned long long*> d_peerRedFlags;
  PeerReduce peerReduce() const;
  // generic host collectives used at setup time, provided by the bootstrap layer
  std::function<void(void*, size_t)> allgather_bytes;  // in-place: buffer holds nranks blocks
  std::function<void()> barrier;
  // device-visible error word (pinned, mapped): a kernel that gave up waiting for a peer (halo flags, all-reduce
  // slots) sets it instead of hanging; the host turns it into NRSB_ERR_CUDA at its next synchronisation point
  int* h_err = nullptr;
  int* d_err = nullptr;
  ~comm_t()
  {
    if (h_err) cudaFreeHost(h_err);
  }
  bool peer_timeout() const { return h_err && *h_err != 0; }
};

int comm_setup_reduce(comm_t* c);

// ---------------------------------------------------------------- oogs
enum class gs_op { add, min, max };

class oogs_t {
 public:
  ogs_t* ogs = nullptr;  // not owned
  comm_t* comm = nullptr;
  // halo exchange state (allocated when ogs->NhaloGather > 0)
  struct Peer {
    int rank;
    int nSend;                // shared rows with this peer
    std::vector<int> rows;    // halo row index for each slot (ascending global id)
    size_t recvOffset;        // offset of this peer's block in my receive window (in slots)
    size_t remoteOffset;      // offset of MY block in the peer's receive window (in slots)
  };
  std::vector<Peer> peers;
  size_t windowSlots = 0;
  int maxFields = 1;
  dbuf<double> window[2];             // receive windows, parity double-buffered, sized in doubles
  dbuf<unsigned long long> flags;     // [nranks] arrival epochs
  std::vector<double*> peerWindow[2];
  std::vector<unsigned long long*> peerFlags;
  dbuf<double*> d_peerWindow[2];
  dbuf<unsigned long long*> d_peerFlags;
  unsigned long long epoch = 0;
  // device CSR describing the exchange
  dbuf<int> d_sendStarts, d_sendPeer, d_sendSlot, d_recvStarts, d_recvSrc, d_peerRanks;
  dbuf<double> d_partial;  // per halo row partial sums (k fields)
  cudaStream_t commStream = nullptr;
  cudaEvent_t evStart = nullptr, evDone = nullptr;
  // fused axhelm + halo push (axhelm_tma.cu): finished-halo-element counter, never reset
  dbuf<unsigned long long> fusedCounter;
  unsigned long long fusedTarget = 0;

  ~oogs_t();
  int setup(ogs_t* ogs, comm_t* comm, int maxFields);
  // oogs::start / finish / startFinish (oogs.cpp:682-837).  Between start and finish the caller may
  // launch independent work on `stream` (interior-element Ax), exactly as ellipticOperator does.
  template <typename T>
  int start(T* v, int k, dlong stride, gs_op op, cudaStream_t stream);
  template <typename T>
  int finish(T* v, int k, dlong stride, gs_op op, dlong Nmasked, const dlong* maskIds, cudaStream_t stream);
  // start() without the pack launch: advances the epoch and describes the exchange to a kernel that
  // performs the pack itself once `NhaloElements` more elements have been counted (struct FusedHalo, halo.cuh)
  int begin_fused(struct FusedHalo* F, dlong NhaloElements, dlong stride);
  // one-launch exchange with flag-in-data windows (one field, ogsAdd); returns 1 when the case is not covered
  template <typename T>
  int exchange_ll(T* v, int k, gs_op op, dlong Nmasked, const dlong* maskIds, cudaStream_t stream);
  template <typename T>
  int startFinish(T* v, int k, dlong stride, gs_op op, dlong Nmasked, const dlong* maskIds, cudaStream_t stream)
  {
    int rc = exchange_ll<T>(v, k, op, Nmasked, maskIds, stream);
    if (rc != 1) return rc;
    rc = start(v, k, stride, op, stream);
    if (rc) return rc;
    return finish(v, k, stride, op, Nmasked, maskIds, stream);
  }
};

// ---------------------------------------------------------------- mesh
class mesh_t {
 public:
  int N = 0, Nq = 0, Np = 0;
  dlong Nelements = 0, Nlocal = 0;
  hlong NelementsGlobal = 0;
  std::vector<double> gllz, gllw, D;  // D row-major D[i][m]
  std::vector<float> Dpfloat;
  std::vector<double> x, y, z;        // host node coordinates [E*Np]
  std::vector<hlong> globalIds;
  std::vector<int> EToB;              // [E*6]
  double volume = 0;                  // global
  // device
  dbuf<double> o_ggeo;
  dbuf<float> o_ggeoPfloat;
  dbuf<double> o_vgeo;  // [E][12][Np] rx..tz, J, JW, 1/JW (mesh3D.h:82-93): built on demand (stress-form block solves)
  int ensure_vgeo();
  dbuf<dlong> o_elementList, o_globalGatherElementList, o_localGatherElementList;
  dbuf<dlong> o_haloFirstElementList;  // globalGather elements followed by localGather elements
  dlong NglobalGatherElements = 0, NlocalGatherElements = 0;
  std::vector<dlong> globalGatherElementList, localGatherElementList;
  std::unique_ptr<ogs_t> ogs;    // unmasked
  std::unique_ptr<oogs_t> oogs;  // unmasked exchange (setup-time sums)
  SharedTopology topo;           // points into topoStore
  std::vector<hlong> topoIds;
  std::vector<int> topoOffsets, topoRanks;
  comm_t* comm = nullptr;

  // meshLoadReferenceNodesHex3D + geometric factors + meshParallelGatherScatterSetup
  int setup(int N, dlong Nelements, const double* x, const double* y, const double* z, const hlong* globalIds,
            const int* EToB, comm_t* comm, const SharedTopology* topo, bool keepFp64Geo);
  static void gll(int N, std::vector<double>& z, std::vector<double>& w);
  static void dmatrix(const std::vector<double>& z, std::vector<double>& D);
  static void interp_matrix(const std::vector<double>& zin, const std::vector<double>& zout, std::vector<double>& I);
};

class elliptic_t;

// ---------------------------------------------------------------- multigrid
enum class SmootherType { CHEBYSHEV, OPT_FOURTH_CHEBYSHEV, FOURTH_CHEBYSHEV, ASM, RAS, JACOBI };
enum class ChebyshevSmootherType { ASM, RAS, JACOBI };

class MGSolver_t;

class pMGLevel {
 public:
  elliptic_t* elliptic = nullptr;   // level solver (owned by MGSolver_t::ellipticLevels)
  elliptic_t* ellipticBase = nullptr;
  mesh_t* mesh = nullptr;
  options_t options;
  int degree = 0;
  bool isCoarse = false;
  dlong Nrows = 0;
  float *o_x = nullptr, *o_rhs = nullptr;  // level vectors (level 0: caller's)
  dbuf<float> x_store, rhs_store, o_res;
  // transfer
  int NqF = 0;
  dlong NpF = 0;
  std::vector<float> R;                 // [NqC][NqF]
  const float* o_invDegreeFine = nullptr;
  // smoother
  SmootherType smootherType = SmootherType::CHEBYSHEV;
  ChebyshevSmootherType chebySmootherType = ChebyshevSmootherType::ASM;
  bool hasSmoother = false;
  double lambda0 = 0, lambda1 = 0, maxEig = 0;
  int UpLegChebyshevDegree = 3, DownLegChebyshevDegree = 3;
  std::vector<float> UpLegBetas, DownLegBetas;
  dbuf<float> o_invDiagA;
  // Schwarz
  dbuf<float> o_Sx, o_Sy, o_Sz, o_invL, o_work1, o_work2, o_wts;
  std::unique_ptr<ogs_t> ogsExtData;
  std::unique_ptr<oogs_t> ogsExt;
  // scratch shared by the smoothers of this level
  dbuf<float> o_smootherResidual, o_smootherResidual2, o_smootherUpdate;

  int Ax(const float* x, float* Ax);
  int residual(const float* rhs, const float* x, float* res);
  int coarsen(float* x, float* Rx);       // x is scaled in place by invDegreeFine (paxmy), as the reference
  int prolongate(const float* x, float* Px);
  int smooth(const float* rhs, float* x, bool xIsZero);
  int smoother(const float* x, float* Sx, bool xIsZero);
  int smoothChebyshev(const float* r, float* x, bool xIsZero);
  int smoothFourthKindChebyshev(const float* r, float* x, bool xIsZero);
  int smoothJacobi(const float* r, float* x, bool xIsZero);
  int smoothSchwarz(const float* u, float* Su, bool xIsZero);
  int setupSmoother();
  int buildSchwarz();
  int generate_weights();
  int maxEigSmoothAx(double* rho);
};

class MGSolver_t {
 public:
  std::vector<std::unique_ptr<pMGLevel>> levels;
  std::vector<std::unique_ptr<elliptic_t>> ellipticLevels;  // level 0 fine copy + coarser levels
  std::vector<std::unique_ptr<mesh_t>> meshLevels;
  int baseLevel = 0;
  std::function<int(float* rhs, float* x)> coarseSolve;
  // MGSOLVER CYCLE = VCYCLE+ADDITIVE (MGSolver.cpp:99-124): all levels smoothed from zero on the restricted
  // right-hand side, corrections added; only legal without Chebyshev acceleration
  bool additive = false;
  int Run(float* o_rhs, float* o_x);
  int runVcycle(int k);
  int runAdditiveVcycle();
};

// coarse-grid solver stand-in for BoomerAMG: Jacobi-preconditioned CG on the
// assembled N=1 operator, fixed relative tolerance, all scalars device resident.
class coarseSolver_t {
 public:
  pMGLevel* level = nullptr;
  int NT = 0;  // unique unmasked coarse nodes on this rank (T-vector)
  int ellWidth = 0;
  bool multiRank = false;
  dbuf<int> d_rowStarts, d_cols, d_rowNode, d_tIndex;
  dbuf<float> d_vals, d_weight, invDiag, x, r, u, p, s, w;
  dbuf<double> scal;
  std::unique_ptr<ogs_t> ogsT;
  std::unique_ptr<oogs_t> oogsT;
  int maxIter = 200;
  double tol = 1e-3;
  int checkEvery = 8;  // convergence is tested every checkEvery iterations (both solve paths)
  int lastIter = 0;
  bool iterOnDevice = false;
  // single-kernel cluster path (coarse_cluster.cu); clusterSize == 0 -> multi-launch path
  int clusterSize = 0, clusterRPC = 0, clusterRmax = 0, clusterMatInSmem = 0, clusterUGlobal = 0;
  size_t clusterSmem = 0;
  dbuf<float> uScratch;
  // single-kernel grid path (coarse_grid.cu: all SMs, grid barrier); gridSize == 0 -> not available
  int gridSize = 0, gridRPC = 0, gridRmax = 0, gridMatInSmem = 0;
  size_t gridSmem = 0;
  dbuf<float> gridU;
  dbuf<double> gridRed;
  dbuf<unsigned> gridBar;
  // several ranks: replicated global coarse problem for the cluster kernel (coarse_cluster.cu header)
  bool replicated = false;
  int NTg = 0, gEllWidth = 0, nOwn = 0;
  dbuf<int> g_cols, g_tIndex, d_ownG, d_ownNode;
  dbuf<float> g_vals, g_invDiag;
  float* rhsWindow = nullptr;  // float[2][NTgpad] + u64 flags[nranks], peer-mapped
  std::vector<void*> peerWinBase;
  dbuf<float*> d_peerWin;
  dbuf<unsigned long long*> d_peerWinFlags;
  unsigned long long winEpoch = 0;
  int setup_replicated(const std::vector<hlong>& idsT, const std::vector<int>& rowNode,
                       const std::vector<int>& tIndex, const std::vector<std::map<int, double>>& rows);
  ~coarseSolver_t();
  // 1 (default): one-kernel solve -- the grid kernel for large systems (> kGridRows rows or replicated), else the
  // cluster kernel; 0: always the multi-launch path; 2: cluster kernel with the SpMV input through L2 (tests);
  // 3: grid kernel whenever it is available; 4: cluster kernel whenever it is available
  static int variant;
  static constexpr int kGridRows = 16384;
  int setup(pMGLevel* lvl, int maxIter, double tol);
  int solve(float* rhs, float* x);
  int spmv_dots(bool first);
  int plan_cluster();
  int solve_cluster(float* rhs, float* x);
  int plan_grid();
  int solve_grid(float* rhs, float* x);
  int iterations();
};

class precon_t {
 public:
  std::unique_ptr<MGSolver_t> MGSolver;
  std::unique_ptr<coarseSolver_t> coarse;
  dbuf<double> o_invDiagA;  // JACOBI
};

// ---------------------------------------------------------------- projection (ellipticSolutionProjection.cpp)
class SolutionProjection;

// ---------------------------------------------------------------- elliptic
class elliptic_t {
 public:
  std::string name = "pressure";
  options_t options;
  mesh_t* mesh = nullptr;
  std::unique_ptr<mesh_t> ownedMesh;
  comm_t* comm = nullptr;
  bool mgLevel = false;
  int Nfields = 1;
  dlong fieldOffset = 0, loffset = 0;
  bool poisson = true;
  int allNeumann = 0;
  double lambda0Value = 1.0, lambda1Value = 0.0;
  dbuf<double> o_lambda0, o_lambda1;            // scalars (device), fp64 solver
  dbuf<float> o_lambda0Pfloat, o_lambda1Pfloat;  // MG levels
  // variable coefficients (ELLIPTIC COEFF FIELD, p_lambda = 1): per-node lambda0 / lambda1.  The fp64 fields of the
  // solver belong to the caller (like the reference's o_lambda0 / o_lambda1 handles); the fp32 copies -- level 0 a
  // cast, coarser levels interpolated -- are refreshed by ellipticMultiGridUpdateLambda.
  // block solver (Nfields > 1; ellipticSetup.cpp:81-131,209-249): the unmasked mesh numbering serves all fields, the
  // Dirichlet mask is per field (ids n + fld * fieldOffset), the operator is the block Helmholtz or, with stressForm,
  // the coupled stress operator.  Constant coefficients: one value per field, loffset = 1.
  bool stressForm = false;
  std::vector<double> blockLambda0, blockLambda1;  // per-field constants given at setup (empty: lambda0/1Value)
  dbuf<double> o_weightBlock;                      // invDegree replicated per field, zero in the padding
  // vector length and weights of every reduction / streaming op of the Krylov loop
  long Nvec() const { return Nfields == 1 ? (long)mesh->Nlocal : (long)Nfields * fieldOffset; }
  const double* o_weight() const { return Nfields == 1 ? o_invDegree : o_weightBlock.p; }
  bool lambdaField = false;
  const double* o_lambda0Field = nullptr;
  const double* o_lambda1Field = nullptr;
  dbuf<float> o_lambda0FieldPfloat, o_lambda1FieldPfloat;
  std::vector<float> interpFromFine;  // MG level: [Nq][NqFine] nodal interpolation from the next finer level
  std::vector<int> EToB;
  // masked gather-scatter
  std::unique_ptr<ogs_t> ogs;
  std::unique_ptr<oogs_t> oogs;
  bool overlap = false;  // oogsAx != oogs in the reference: split Ax into halo / interior elements
  // which form ellipticOperator takes on several ranks: the reference's split (Ax halo, oogs::start, Ax interior,
  // oogs::finish) or Ax on all elements + the one-launch exchange.  ENABLE GS COMM OVERLAP = SPLIT / UNSPLIT fix it,
  // TIMED measures both at setup and keeps the faster (ellipticSetup.cpp:255-302, ellipticMultiGridSetup.cpp:123)
  bool splitOverlap = false;
  double overlapTimes[2] = {0.0, 0.0};  // TIMED: seconds per operator {unsplit, split}, max over ranks
  dbuf<double> o_resHist;  // device-side residual history of the current PCG solve
  bool fusedHaloAx = true;  // overlap through ONE launch (axhelm + in-kernel halo push) when Nq == 8
  // mask + on-rank gather-scatter as phase 2 of the axhelm launch when Nq == 8 (struct FusedRows, gs.hpp)
  bool fusedGsAx = false;
  dbuf<unsigned long long> fusedArrive;  // arrival counter of the axhelm CTAs, never reset
  // PCG: p^T A p from the axhelm launch (energy form) instead of a separate weighted-inner-product pass
  bool fusedDotAx = true;
  dbuf<double> o_dotPartials;
  unsigned long long fusedArriveTarget = 0;
  dlong Nmasked = 0, NmaskedLocal = 0, NmaskedGlobal = 0;
  dbuf<dlong> o_maskIds, o_maskIdsLocal, o_maskIdsGlobal;
  std::vector<dlong> maskIds;
  const double* o_invDegree = nullptr;
  const float* o_invDegreePfloat = nullptr;
  // workspace (ellipticUpdateWorkspace, elliptic.h:240-262)
  dbuf<double> o_p, o_z, o_Ap, o_x0, o_rtmp;
  dbuf<float> o_rPfloat, o_zPfloat;
  // reductions
  dbuf<double> redPartials;
  dbuf<unsigned> redTicket;
  ReduceWs ws;
  dbuf<double> o_scal;       // device scalars
  double* h_scal = nullptr;  // pinned mirror
  // host-mapped error word of this handle: a bounded device-side wait that gave up sets it (fused launches whose
  // grid was not co-resident); read_scalars and the solve entry points turn it into NRSB_ERR_CUDA
  int* h_err = nullptr;
  int* d_err = nullptr;
  // Krylov
  int ax_variant[2] = {-1, -1};  // [0] fp64, [1] fp32
  std::unique_ptr<precon_t> precon;
  // GMRES
  int nRestartVectors = 15;
  dbuf<double> o_V, o_Z, o_y;
  std::vector<double> gmres_H, gmres_sn, gmres_cs, gmres_s, gmres_y;
  // results (elliptic.h:89-90)
  int Niter = 0;
  double res00Norm = 0, res0Norm = 0, resNorm = 0, resNormFactor = 0;
  std::vector<double> resHistory;  // per-iteration residual norms of the last solve
  std::unique_ptr<SolutionProjection> solutionProjection;
  cudaStream_t stream = nullptr;
  // global numbering (and cross-rank sharing) of the coarser p-multigrid meshes, keyed by order.
  // The reference obtains them from nek5000 for every level mesh (createMeshMG -> meshGlobalIds,
  // meshSetup.cpp:348); here the caller supplies them.
  std::map<int, std::vector<hlong>> levelGlobalIds;
  struct TopoStore {
    std::vector<hlong> ids;
    std::vector<int> offsets, ranks;
  };
  std::map<int, TopoStore> levelTopoStore;
  std::map<int, SharedTopology> levelTopology;

  elliptic_t();
  ~elliptic_t();
  int read_scalars(int first, int count, double* out);  // device -> host, synchronises `stream`
};

// elliptic.cpp
int ellipticSolveSetup(elliptic_t* elliptic);
int ellipticChooseOverlap(elliptic_t* elliptic, int precision);
int ellipticKrylovWorkspace(elliptic_t* elliptic);  // (re)sizes the PGMRES buffers for the current SOLVER options
int ellipticSolve(elliptic_t* elliptic, double* o_r, double* o_x);
template <typename T>
int ellipticAx(elliptic_t* elliptic, dlong NelementsList, const dlong* o_elementList, const T* o_q, T* o_Aq);
// dot != nullptr: ask the axhelm launch for q^T A q (per-CTA partials, kernels.hpp AxDot); dot->n == 0 on return
// means the launch path taken could not provide it
template <typename T>
int ellipticOperator(elliptic_t* elliptic, const T* o_q, T* o_Aq, bool masked = true, AxDot* dot = nullptr);
template <typename T>
int ellipticApplyMask(elliptic_t* elliptic, T* o_x);
int ellipticZeroMean(elliptic_t* elliptic, double* o_q);
int ellipticPreconditioner(elliptic_t* elliptic, double* o_r, double* o_z);
int ellipticPreconditionerSetup(elliptic_t* elliptic);
int ellipticOgs(mesh_t* mesh, const std::vector<int>& EToB, elliptic_t* elliptic);
int pcg(elliptic_t* elliptic, double* o_r, double* o_x, double tol, int MAXIT, double& rdotr);
int pgmres(elliptic_t* elliptic, double* o_r, double* o_x, double tol, int MAXIT, double& rdotr);
template <typename T>
int ellipticBuildDiagonal(elliptic_t* elliptic, T* o_invDiagA);  // ellipticUpdateJacobi(elliptic, o_invDiagA)
int ellipticUpdateJacobi(elliptic_t* ellipticBase);             // ellipticUpdateJacobi.cpp:87-115
int ellipticMultiGridUpdateLambda(elliptic_t* elliptic);        // MG/ellipticMultiGridUpdateLambda.cpp
int ellipticSetCoeffField(elliptic_t* elliptic, const double* o_lambda0, const double* o_lambda1);
// multigrid.cpp
int ellipticMultiGridSetup(elliptic_t* elliptic, precon_t* precon);
std::vector<int> determineMGLevels(const options_t& options, int N);
// dense.cpp
int sym_generalized_eig(int n, std::vector<double>& A, std::vector<double>& B, std::vector<double>& lam);
double hessenberg_spectral_radius(int n, std::vector<double> H);

uint64_t splitmix64(uint64_t x);
inline double id_uniform(hlong id) { return (double)(splitmix64((uint64_t)id) >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace nrsb
