// mesh.cpp -- mesh_t setup: reference nodes (GLL, D, interpolation), geometric factors, unmasked
// gather-scatter, halo/interior element lists.
//
// Restates meshLoadReferenceNodesHex3D.cpp:33-120, meshBasis1D.cpp (JacobiGLL :237-266, Dmatrix1D
// :99-119, InterpolationMatrix1D :131-150), mesh_t::geometricFactors (meshGeometricFactorsHex3D.cpp:33-73,
// kernel geometricFactorsHex3D.okl) and meshParallelGatherScatterSetup.cpp:35-166.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "host.hpp"

namespace nrsb {

bool options_t::getArgs(const std::string& key, int& v) const
{
  auto it = kv.find(key);
  if (it == kv.end() || it->second.empty()) return false;
  v = std::atoi(it->second.c_str());
  return true;
}
bool options_t::getArgs(const std::string& key, double& v) const
{
  auto it = kv.find(key);
  if (it == kv.end() || it->second.empty()) return false;
  v = std::atof(it->second.c_str());
  return true;
}

template <typename T>
int dbuf<T>::alloc(size_t count, bool zero)
{
  release();
  n = count;
  if (!count) return NRSB_OK;
  cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    p = nullptr;
    n = 0;
    set_last_error("out of device memory");
    return NRSB_ERR_NOMEM;
  }
  NRSB_CUDA(e);
  if (zero) {
    // cudaMemset on device memory is asynchronous and runs on the legacy default stream, which is NOT ordered
    // against cudaStreamNonBlocking streams (what nrsb_stream_create hands out): wait for it here, so that work
    // queued on any stream right after a (lazy) allocation cannot be overtaken by the zero-fill.
    NRSB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), cudaStreamLegacy));
    NRSB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  }
  return NRSB_OK;
}
template <typename T>
int dbuf<T>::upload(const T* h, size_t count)
{
  int rc = alloc(count, false);
  if (rc) return rc;
  if (count) NRSB_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
  return NRSB_OK;
}
template <typename T>
int dbuf<T>::upload(const std::vector<T>& h)
{
  return upload(h.data(), h.size());
}
template <typename T>
int dbuf<T>::download(std::vector<T>& h) const
{
  h.resize(n);
  if (n) NRSB_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
  return NRSB_OK;
}
template struct dbuf<double>;
template struct dbuf<float>;
template struct dbuf<int>;
template struct dbuf<unsigned>;
template struct dbuf<unsigned long long>;
template struct dbuf<double*>;
template struct dbuf<float*>;
template struct dbuf<int4>;
template struct dbuf<int2>;
template struct dbuf<unsigned long long*>;
template struct dbuf<void*>;
template struct dbuf<long>;

// ------------------------------------------------------------------------------------------
// GLL points: roots of (1-x^2) P_N'(x) by Newton on the Legendre recurrence; weights
// w_i = 2 / (N (N+1) P_N(x_i)^2)  (identical to the mass-lumped weights of JacobiGLL)
void mesh_t::gll(int N, std::vector<double>& z, std::vector<double>& w)
{
  const int n = N + 1;
  z.assign(n, 0.0);
  w.assign(n, 0.0);
  auto legendre = [&](double x, double& p, double& pm1) {
    p = x;
    pm1 = 1.0;
    for (int k = 2; k <= N; ++k) {
      const double pk = ((2 * k - 1) * x * p - (k - 1) * pm1) / k;
      pm1 = p;
      p = pk;
    }
  };
  if (N == 1) {
    z[0] = -1;
    z[1] = 1;
    w[0] = w[1] = 1;
    return;
  }
  for (int i = 0; i < n; ++i) {
    double x = -std::cos(M_PI * i / N);
    for (int it = 0; it < 100; ++it) {
      double p, pm1;
      legendre(x, p, pm1);
      const double dx = (x * p - pm1) / ((N + 1) * p);
      x -= dx;
      if (std::fabs(dx) < 1e-16) break;
    }
    z[i] = x;
  }
  z[0] = -1.0;
  z[N] = 1.0;
  for (int i = 0; i < n / 2; ++i) {
    const double s = 0.5 * (z[N - i] - z[i]);
    z[i] = -s;
    z[N - i] = s;
  }
  if (n % 2) z[N / 2] = 0.0;
  for (int i = 0; i < n; ++i) {
    double p, pm1;
    legendre(z[i], p, pm1);
    w[i] = 2.0 / (N * (N + 1) * p * p);
  }
}

static void bary_weights(const std::vector<double>& z, std::vector<double>& bw)
{
  const int n = (int)z.size();
  bw.assign(n, 1.0);
  for (int i = 0; i < n; ++i) {
    double p = 1.0;
    for (int j = 0; j < n; ++j)
      if (j != i) p *= (z[i] - z[j]);
    bw[i] = 1.0 / p;
  }
}

void mesh_t::dmatrix(const std::vector<double>& z, std::vector<double>& D)
{
  const int n = (int)z.size();
  std::vector<double> bw;
  bary_weights(z, bw);
  D.assign(n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j)
      if (j != i) {
        D[i * n + j] = (bw[j] / bw[i]) / (z[i] - z[j]);
        s += D[i * n + j];
      }
    D[i * n + i] = -s;
  }
}

// I[o][i] = l_i(zout[o]), row-major [nout][nin]
void mesh_t::interp_matrix(const std::vector<double>& zin, const std::vector<double>& zout, std::vector<double>& I)
{
  const int ni = (int)zin.size(), no = (int)zout.size();
  std::vector<double> bw;
  bary_weights(zin, bw);
  I.assign((size_t)no * ni, 0.0);
  for (int o = 0; o < no; ++o) {
    int hit = -1;
    for (int i = 0; i < ni; ++i)
      if (std::fabs(zout[o] - zin[i]) < 1e-14) hit = i;
    if (hit >= 0) {
      I[(size_t)o * ni + hit] = 1.0;
      continue;
    }
    double s = 0.0;
    for (int i = 0; i < ni; ++i) {
      const double t = bw[i] / (zout[o] - zin[i]);
      I[(size_t)o * ni + i] = t;
      s += t;
    }
    for (int i = 0; i < ni; ++i) I[(size_t)o * ni + i] /= s;
  }
}

int mesh_t::ensure_vgeo()
{
  if (o_vgeo.p || Nlocal == 0) return NRSB_OK;
  int rc;
  dbuf<double> dx, dy, dz, dD, dw;
  if ((rc = dx.upload(x)) || (rc = dy.upload(y)) || (rc = dz.upload(z)) || (rc = dD.upload(D)) || (rc = dw.upload(gllw)))
    return rc;
  if ((rc = o_vgeo.alloc((size_t)Nlocal * 12, false))) return rc;
  if ((rc = geometric_factors_launch(Nq, Nelements, dD.p, dw.p, dx.p, dy.p, dz.p, nullptr, nullptr, nullptr, o_vgeo.p)))
    return rc;
  NRSB_CUDA(cudaDeviceSynchronize());
  return NRSB_OK;
}

int mesh_t::setup(int N_, dlong Nelements_, const double* x_, const double* y_, const double* z_,
                  const hlong* globalIds_, const int* EToB_, comm_t* comm_, const SharedTopology* topo_,
                  bool keepFp64Geo)
{
  NRSB_REQUIRE(N_ >= 1 && N_ + 1 <= 12, "polynomial order must be in 1..11");
  N = N_;
  Nq = N + 1;
  Np = Nq * Nq * Nq;
  Nelements = Nelements_;
  Nlocal = Nelements * Np;
  comm = comm_;
  gll(N, gllz, gllw);
  dmatrix(gllz, D);
  Dpfloat.assign(D.begin(), D.end());
  x.assign(x_, x_ + Nlocal);
  y.assign(y_, y_ + Nlocal);
  z.assign(z_, z_ + Nlocal);
  globalIds.assign(globalIds_, globalIds_ + Nlocal);
  EToB.assign(EToB_, EToB_ + (size_t)Nelements * 6);

  // topology (own copy)
  topo = SharedTopology();
  if (topo_ && topo_->nranks > 1) {
    topoIds.assign(topo_->sharedIds, topo_->sharedIds + topo_->nShared);
    topoOffsets.assign(topo_->sharerOffsets, topo_->sharerOffsets + topo_->nShared + 1);
    topoRanks.assign(topo_->sharerRanks, topo_->sharerRanks + topoOffsets.back());
    topo.rank = topo_->rank;
    topo.nranks = topo_->nranks;
    topo.nShared = topo_->nShared;
    topo.sharedIds = topoIds.data();
    topo.sharerOffsets = topoOffsets.data();
    topo.sharerRanks = topoRanks.data();
  }

  // geometric factors on the device
  int rc;
  {
    dbuf<double> dx, dy, dz, dD, dw, dJ;
    if ((rc = dx.upload(x))) return rc;
    if ((rc = dy.upload(y))) return rc;
    if ((rc = dz.upload(z))) return rc;
    if ((rc = dD.upload(D))) return rc;
    if ((rc = dw.upload(gllw))) return rc;
    if ((rc = dJ.alloc(Nlocal))) return rc;
    if ((rc = o_ggeo.alloc((size_t)Nlocal * 7, false))) return rc;
    if ((rc = geometric_factors_launch(Nq, Nelements, dD.p, dw.p, dx.p, dy.p, dz.p, o_ggeo.p, dJ.p, nullptr)))
      return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    // Jacobian positivity (meshGeometricFactorsHex3D.cpp:52-58) and volume = sum JW (meshSetup.cpp:444)
    std::vector<double> J;
    if ((rc = dJ.download(J))) return rc;
    double minJ = 1e300, maxJ = -1e300;
    for (double v : J) {
      minJ = std::min(minJ, v);
      maxJ = std::max(maxJ, v);
    }
    if (Nlocal && (minJ <= 0 || maxJ <= 0 || std::isnan(minJ))) {
      set_last_error("Jacobian < 0 or nan");
      return NRSB_ERR_INVALID;
    }
    double vol = 0.0;
    for (dlong e = 0; e < Nelements; ++e)
      for (int k = 0; k < Nq; ++k)
        for (int j = 0; j < Nq; ++j)
          for (int i = 0; i < Nq; ++i) vol += J[(size_t)e * Np + k * Nq * Nq + j * Nq + i] * gllw[i] * gllw[j] * gllw[k];
    volume = vol;
    if ((rc = o_ggeoPfloat.alloc((size_t)Nlocal * 7, false))) return rc;
    if ((rc = copy_d2f_launch((long)Nlocal * 7, o_ggeo.p, o_ggeoPfloat.p, nullptr))) return rc;
    NRSB_CUDA(cudaDeviceSynchronize());
    if (!keepFp64Geo) o_ggeo.release();
  }
  NelementsGlobal = Nelements;
  if (comm && comm->nranks > 1) {
    // volume and element count are global sums
    std::vector<double> buf(2 * comm->nranks, 0.0);
    buf[2 * comm->rank] = volume;
    buf[2 * comm->rank + 1] = (double)Nelements;
    comm->allgather_bytes(buf.data(), 2 * sizeof(double));
    volume = 0;
    double ne = 0;
    for (int r = 0; r < comm->nranks; ++r) {
      volume += buf[2 * r];
      ne += buf[2 * r + 1];
    }
    NelementsGlobal = (hlong)(ne + 0.5);
  }

  // unmasked gather-scatter
  ogs.reset(new ogs_t());
  if ((rc = ogs->setup(Nlocal, globalIds.data(), topo.nranks > 1 ? &topo : nullptr))) return rc;
  oogs.reset(new oogs_t());
  if ((rc = oogs->setup(ogs.get(), comm, 1))) return rc;

  // element lists (meshParallelGatherScatterSetup.cpp:72-134): an element is "global" when any of
  // its nodes is shared with another rank
  std::vector<char> isHaloNode(Nlocal, 0);
  for (dlong n : ogs->haloGatherIds) isHaloNode[n] = 1;
  std::vector<dlong> elist(Nelements);
  globalGatherElementList.clear();
  localGatherElementList.clear();
  for (dlong e = 0; e < Nelements; ++e) {
    elist[e] = e;
    bool halo = false;
    for (int n = 0; n < Np && !halo; ++n) halo = isHaloNode[(size_t)e * Np + n];
    (halo ? globalGatherElementList : localGatherElementList).push_back(e);
  }
  NglobalGatherElements = (dlong)globalGatherElementList.size();
  NlocalGatherElements = (dlong)localGatherElementList.size();
  if ((rc = o_elementList.upload(elist))) return rc;
  if ((rc = o_globalGatherElementList.upload(globalGatherElementList))) return rc;
  if ((rc = o_localGatherElementList.upload(localGatherElementList))) return rc;
  {
    std::vector<dlong> both(globalGatherElementList);
    both.insert(both.end(), localGatherElementList.begin(), localGatherElementList.end());
    if ((rc = o_haloFirstElementList.upload(both))) return rc;
  }
  return NRSB_OK;
}

}  // namespace nrsb
