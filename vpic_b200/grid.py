"""Host-side grid description consumed by the device path.

The reference builds `grid_t` on the host (src/grid/partition.cc:35-89, src/grid/ops.cc:19-211) and the hot path
only READS it: geometry constants, the `neighbor[6*nv]` table with its negative particle-boundary codes, the voxel
range of this rank and the 27-entry field boundary table.  This module produces the same read-only data for tests
and bench.py when no reference host program is present (e.g. on the GPU box), following the same rules:
slab/box decomposition with periodic wrap of the rank index, `reflect_particles` on every local wall until a
particle boundary is set, neighbours expressed as GLOBAL voxel ids (range[rank] + local voxel).
"""
from dataclasses import dataclass, field
import numpy as np

REFLECT_PARTICLES = -1
ABSORB_PARTICLES = -2
PEC_FIELDS = -1
SYMMETRIC_FIELDS = -2
PMC_FIELDS = -3
ABSORB_FIELDS = -4

FACES = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1)]


def boundary_index(i, j, k):
    """BOUNDARY(i,j,k), grid.h:16."""
    return 13 + i + 3 * j + 9 * k


def voxel(x, y, z, nx, ny, nz):
    return x + (nx + 2) * (y + (ny + 2) * z)


@dataclass
class Grid:
    nx: int
    ny: int
    nz: int
    dt: float
    cvac: float
    eps0: float
    x0: float
    y0: float
    z0: float
    x1: float
    y1: float
    z1: float
    dx: float
    dy: float
    dz: float
    dV: float
    rdx: float
    rdy: float
    rdz: float
    r8V: float
    rank: int = 0
    world_size: int = 1
    bc: list = field(default_factory=lambda: [PEC_FIELDS] * 27)
    range: np.ndarray = None
    neighbor: np.ndarray = None   # int64 [nv, 6]
    step: int = 0

    @property
    def nv(self):
        return (self.nx + 2) * (self.ny + 2) * (self.nz + 2)

    @property
    def rangel(self):
        return int(self.range[self.rank])

    @property
    def rangeh(self):
        return int(self.range[self.rank + 1] - 1)

    def face_codes(self):
        """Per-face code for the device field kernels (vpb_field_args_t.face): periodic onto this same rank -> 0,
        another rank -> 1 (halo exchange), local field BC -> its negative code."""
        out = []
        for f in FACES:
            b = self.bc[boundary_index(*f)]
            if b < 0:
                out.append(int(b))
            elif b == self.rank:
                out.append(0)
            else:
                out.append(1)
        return out

    def set_fbc(self, face, code):
        self.bc[boundary_index(*FACES[face])] = code

    def set_pbc(self, face, code):
        """set_pbc, ops.cc:184-211: every voxel on that wall gets the particle boundary code."""
        nx, ny, nz = self.nx, self.ny, self.nz
        X = face % 3
        n = (nx, ny, nz)
        plane = 1 if face < 3 else n[X]
        rng = [np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1)]
        rng[X] = np.array([plane])
        xx, yy, zz = np.meshgrid(*rng, indexing="ij")
        self.neighbor[voxel(xx, yy, zz, nx, ny, nz).ravel(), face] = code


def _rank_to_index(rank, gpx, gpy, gpz):
    ix = rank % gpx
    iy = (rank // gpx) % gpy
    iz = rank // (gpx * gpy)
    return ix, iy, iz


def _index_to_rank(ix, iy, iz, gpx, gpy, gpz):
    return (ix % gpx) + gpx * ((iy % gpy) + gpy * (iz % gpz))


def partition_periodic_box(gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz,
                           rank=0, dt=0.0, cvac=1.0, eps0=1.0):
    """Same arithmetic as partition_periodic_box + size_grid + join_grid (all ranks have equal local sizes)."""
    world_size = gpx * gpy * gpz
    if gnx % gpx or gny % gpy or gnz % gpz:
        raise ValueError("Bad resolution for domain decomposition")
    f32 = np.float32
    px, py, pz = _rank_to_index(rank, gpx, gpy, gpz)
    nx, ny, nz = gnx // gpx, gny // gpy, gnz // gpz

    def lerp(a, b, f):
        return a * (1 - f) + b * f

    g = Grid(nx=nx, ny=ny, nz=nz, dt=float(f32(dt)), cvac=float(f32(cvac)), eps0=float(f32(eps0)),
             x0=float(f32(lerp(gx0, gx1, px / gpx))), y0=float(f32(lerp(gy0, gy1, py / gpy))),
             z0=float(f32(lerp(gz0, gz1, pz / gpz))),
             x1=float(f32(lerp(gx0, gx1, (px + 1) / gpx))), y1=float(f32(lerp(gy0, gy1, (py + 1) / gpy))),
             z1=float(f32(lerp(gz0, gz1, (pz + 1) / gpz))),
             dx=float(f32((gx1 - gx0) / gnx)), dy=float(f32((gy1 - gy0) / gny)), dz=float(f32((gz1 - gz0) / gnz)),
             dV=float(f32(((gx1 - gx0) / gnx) * ((gy1 - gy0) / gny) * ((gz1 - gz0) / gnz))),
             rdx=float(f32(gnx / (gx1 - gx0))), rdy=float(f32(gny / (gy1 - gy0))), rdz=float(f32(gnz / (gz1 - gz0))),
             r8V=float(f32((gnx / (gx1 - gx0)) * (gny / (gy1 - gy0)) * (gnz / (gz1 - gz0)) * 0.125)),
             rank=rank, world_size=world_size)
    nv = g.nv
    g.range = np.arange(world_size + 1, dtype=np.int64) * nv
    rangel = g.rangel
    g.bc = [PEC_FIELDS] * 27
    g.bc[boundary_index(0, 0, 0)] = rank

    # size_grid: local neighbours, reflecting walls, ghosts reflect everywhere
    x, y, z = np.meshgrid(np.arange(nx + 2), np.arange(ny + 2), np.arange(nz + 2), indexing="ij")
    v = voxel(x, y, z, nx, ny, nz)
    nb = np.empty((nv, 6), dtype=np.int64)
    sx, sy, sz = 1, nx + 2, (nx + 2) * (ny + 2)
    for fidx, off in enumerate((-sx, -sy, -sz, sx, sy, sz)):
        nb[v.ravel(), fidx] = rangel + v.ravel() + off
    ghost = (x == 0) | (x == nx + 1) | (y == 0) | (y == ny + 1) | (z == 0) | (z == nz + 1)
    walls = [(x == 1), (y == 1), (z == 1), (x == nx), (y == ny), (z == nz)]
    for fidx, w in enumerate(walls):
        nb[v[w].ravel(), fidx] = REFLECT_PARTICLES
    nb[v[ghost].ravel(), :] = REFLECT_PARTICLES
    g.neighbor = nb

    # join_grid on all six faces with the periodically wrapped neighbour rank
    n = (nx, ny, nz)
    for fidx, (i, j, k) in enumerate(FACES):
        r = _index_to_rank(px + i, py + j, pz + k, gpx, gpy, gpz)
        g.bc[boundary_index(i, j, k)] = r
        X = fidx % 3
        lplane = 1 if fidx < 3 else n[X]
        rplane = n[X] if fidx < 3 else 1          # remote sizes equal local sizes
        rng = [np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1)]
        rng[X] = np.array([lplane])
        lx, ly, lz = np.meshgrid(*rng, indexing="ij")
        rc = [lx, ly, lz]
        rc[X] = np.full_like(lx, rplane)
        g.neighbor[voxel(lx, ly, lz, nx, ny, nz).ravel(), fidx] = g.range[r] + voxel(rc[0], rc[1], rc[2], nx, ny, nz).ravel()
    return g


def courant_dt(dx, dy, dz, nx, ny, nz, cvac=1.0, frac=0.99):
    """dt = frac * courant length / c over the non-degenerate axes (test/unit/energy_comparison/3d_test.cc:112-113)."""
    inv = sum((1.0 / d) ** 2 for d, n in ((dx, nx), (dy, ny), (dz, nz)) if n > 1)
    return frac / (cvac * np.sqrt(inv))
