"""Synthetic hex meshes for the pressure-Poisson path: box and kershaw.

This is *input generation*, not solver code.  nekRS obtains the same data from
the Fortran nek5000 interface (`src/nekInterface/`, not buildable here):

  * element vertices of an `nelx x nely x nelz` box, lexicographic element order
    with x fastest (genbox convention; `examples/kershaw/kershaw.box`),
  * the kershaw vertex map (`examples/kershaw/kershaw.usr:107-226`) followed by
    the shift to [-1/2,1/2]^3 (`kershaw.usr:88-90`),
  * GLL nodes of each (trilinear) element (`meshPhysicalNodesHex3D.cpp:33-64`),
  * the C0 global numbering `globalIds` (nek's `set_glo_num`,
    `meshGlobalIds.cpp:26-33`); for a structured box it is the lexicographic
    number of the lattice point, starting at 1 (0 means "masked", `ogs.hpp:42-44`),
  * `EToB` boundary flags per element face in nekRS face order
    (`meshBasisHex3D.cpp:68-80`: f0 t=-1, f1 s=-1, f2 r=+1, f3 s=+1, f4 r=-1, f5 t=+1),
  * the element -> rank partition.  parRSB needs MPI; we use a deterministic
    brick partition.

Only numpy is used.  Everything is vectorised so the 44^3 kershaw mesh (43.6 M
nodes) is generated in seconds.
"""
from __future__ import annotations

import dataclasses
import numpy as np

DIRICHLET = 1  # elliptic.h:46
NEUMANN = 4    # elliptic.h:49


def gll_nodes(N: int) -> np.ndarray:
    """Gauss-Lobatto-Legendre points on [-1,1] (roots of (1-x^2) P_N'(x))."""
    if N == 1:
        return np.array([-1.0, 1.0])
    # Chebyshev-Gauss-Lobatto start, Newton on the Legendre recurrence
    x = -np.cos(np.pi * np.arange(N + 1) / N)
    for _ in range(100):
        P = np.zeros((N + 1, N + 1))
        P[0] = 1.0
        P[1] = x
        for k in range(2, N + 1):
            P[k] = ((2 * k - 1) * x * P[k - 1] - (k - 1) * P[k - 2]) / k
        dx = (x * P[N] - P[N - 1]) / ((N + 1) * P[N])
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    return x


def _kershaw_right(eps, x):
    return np.where(x <= 0.5, (2.0 - eps) * x, 1.0 + eps * (x - 1.0))


def _kershaw_left(eps, x):
    return 1.0 - _kershaw_right(eps, 1.0 - x)


def _kershaw_step(a, b, x):
    return np.where(x <= 0.0, a, np.where(x >= 1.0, b, a + (b - a) * x))


def kershaw_map(eps: float, x, y, z):
    """Vertex map of kershaw.usr:142-183 (epsy = epsz = eps), inputs in [0,1]."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    layer = (x * 6.0).astype(np.int64)
    lam = (x - layer / 6.0) * 6.0
    out = []
    for c in (y, z):
        L = _kershaw_left(eps, c)
        R = _kershaw_right(eps, c)
        v14 = _kershaw_step(L, R, lam)
        v2 = _kershaw_step(R, L, lam / 2.0)
        v3 = _kershaw_step(R, L, (1.0 + lam) / 2.0)
        v = np.where(layer == 0, L,
            np.where((layer == 1) | (layer == 4), v14,
            np.where(layer == 2, v2,
            np.where(layer == 3, v3, R))))
        out.append(v)
    return x.copy(), out[0], out[1]


@dataclasses.dataclass
class HexMesh:
    """One rank's share of a structured hex mesh at polynomial order N."""
    N: int
    nel_global: tuple            # (nx, ny, nz) of the whole box
    brick_lo: tuple              # first element (ix,iy,iz) of this rank's brick
    brick_n: tuple               # elements per direction on this rank
    x: np.ndarray                # [E*Np] node coordinates, node index i + Nq*j + Nq^2*k
    y: np.ndarray
    z: np.ndarray
    global_ids: np.ndarray       # int64 [E*Np], >= 1
    EToB: np.ndarray             # int32 [E*6]
    vertices: np.ndarray         # [E, 8, 3] element corner coordinates (nek vertex order)
    rank: int = 0
    nranks: int = 1

    @property
    def Nelements(self) -> int:
        return int(np.prod(self.brick_n))

    @property
    def Nq(self) -> int:
        return self.N + 1

    @property
    def Np(self) -> int:
        return (self.N + 1) ** 3


def brick_partition(nranks: int) -> tuple:
    """(px,py,pz) process grid: 1->1x1x1, 2->2x1x1, 4->2x2x1, 8->2x2x2."""
    p = [1, 1, 1]
    d = 0
    n = nranks
    while n > 1:
        if n % 2:
            raise ValueError("nranks must be a power of two")
        p[d % 3] *= 2
        n //= 2
        d += 1
    return tuple(p)


def _split(n, p, r):
    base, rem = divmod(n, p)
    lo = r * base + min(r, rem)
    return lo, base + (1 if r < rem else 0)


def box_mesh(N: int, nel, *, kershaw_eps: float | None = None, rank: int = 0, nranks: int = 1,
             bc: int = DIRICHLET, lo=(-0.5, -0.5, -0.5), hi=(0.5, 0.5, 0.5),
             proc_grid=None, coords_at_order: int | None = None) -> HexMesh:
    """Box of nel=(nx,ny,nz) trilinear elements on [lo,hi]; optional kershaw(eps) map.

    `bc` is applied on all six box faces (kershaw: pressure DIRICHLET).
    """
    nx, ny, nz = (int(v) for v in nel)
    px, py, pz = proc_grid if proc_grid is not None else brick_partition(nranks)
    assert px * py * pz == nranks
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    x0, ex = _split(nx, px, rx)
    y0, ey = _split(ny, py, ry)
    z0, ez = _split(nz, pz, rz)
    E = ex * ey * ez

    # element (ix,iy,iz) in local lexicographic order, x fastest
    iz, iy, ix = np.meshgrid(np.arange(z0, z0 + ez), np.arange(y0, y0 + ey),
                             np.arange(x0, x0 + ex), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()

    # vertices in [0,1]^3: corner c = a + 2b + 4c  (a: r, b: s, c: t)
    ca = np.array([0, 1, 0, 1, 0, 1, 0, 1])
    cb = np.array([0, 0, 1, 1, 0, 0, 1, 1])
    cc = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    vx = (ix[:, None] + ca[None, :]) / nx
    vy = (iy[:, None] + cb[None, :]) / ny
    vz = (iz[:, None] + cc[None, :]) / nz
    if kershaw_eps is not None:
        vx, vy, vz = kershaw_map(kershaw_eps, vx, vy, vz)
    vx = lo[0] + (hi[0] - lo[0]) * vx
    vy = lo[1] + (hi[1] - lo[1]) * vy
    vz = lo[2] + (hi[2] - lo[2]) * vz
    verts = np.stack([vx, vy, vz], axis=-1)  # [E,8,3]

    # trilinear blend at GLL points
    Nq = N + 1
    g = gll_nodes(N)
    h0, h1 = 0.5 * (1 - g), 0.5 * (1 + g)
    H = np.stack([h0, h1])  # [2,Nq]
    # weight[c, k, j, i] = H[cc,k] H[cb,j] H[ca,i]
    W = (H[cc][:, :, None, None] * H[cb][:, None, :, None] * H[ca][:, None, None, :]).reshape(8, Nq ** 3)
    x = (vx @ W).ravel()
    y = (vy @ W).ravel()
    z = (vz @ W).ravel()

    # global lattice numbering
    NX, NY = nx * N + 1, ny * N + 1
    li = np.arange(Nq)
    kk, jj, ii = np.meshgrid(li, li, li, indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    gi = ix[:, None] * N + ii[None, :]
    gj = iy[:, None] * N + jj[None, :]
    gk = iz[:, None] * N + kk[None, :]
    gid = (1 + gi + NX * (gj + NY * gk)).astype(np.int64).ravel()

    EToB = np.zeros((E, 6), dtype=np.int32)
    EToB[iz == 0, 0] = bc
    EToB[iy == 0, 1] = bc
    EToB[ix == nx - 1, 2] = bc
    EToB[iy == ny - 1, 3] = bc
    EToB[ix == 0, 4] = bc
    EToB[iz == nz - 1, 5] = bc

    return HexMesh(N=N, nel_global=(nx, ny, nz), brick_lo=(x0, y0, z0), brick_n=(ex, ey, ez),
                   x=x, y=y, z=z, global_ids=gid, EToB=EToB.ravel(), vertices=verts,
                   rank=rank, nranks=nranks)


def kershaw_rhs(mesh: HexMesh) -> np.ndarray:
    """f = 3 pi^2 sin(pi x) sin(pi y) sin(pi z)   (kershaw.udf:20-23)."""
    return 3 * np.pi ** 2 * np.sin(np.pi * mesh.x) * np.sin(np.pi * mesh.y) * np.sin(np.pi * mesh.z)


def global_ids_at_order(mesh: HexMesh, Nc: int) -> np.ndarray:
    """C0 numbering of the same elements at polynomial order Nc (the numbering nek's set_glo_num
    gives the level mesh created by createMeshMG, meshSetup.cpp:293-348)."""
    nx, ny, nz = mesh.nel_global
    x0, y0, z0 = mesh.brick_lo
    ex, ey, ez = mesh.brick_n
    iz, iy, ix = np.meshgrid(np.arange(z0, z0 + ez), np.arange(y0, y0 + ey), np.arange(x0, x0 + ex), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    Nq = Nc + 1
    li = np.arange(Nq)
    kk, jj, ii = np.meshgrid(li, li, li, indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    NX, NY = nx * Nc + 1, ny * Nc + 1
    gi = ix[:, None] * Nc + ii[None, :]
    gj = iy[:, None] * Nc + jj[None, :]
    gk = iz[:, None] * Nc + kk[None, :]
    return (1 + gi + NX * (gj + NY * gk)).astype(np.int64).ravel()
