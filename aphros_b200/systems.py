"""Synthetic 7-point pressure systems (inputs only -- no solver code here).

Each generator returns the system in the reference's own row format
(`FieldCell<Expr>`, 8 doubles per cell ``[c, x-, x+, y-, y+, z-, z+, const]``,
src/geom/mesh.h:484-485; meaning ``e0*x[c] + sum_q e[1+q]*x[nb_q] + e7 = 0``,
src/linear/linear.h:34-44) as a C-contiguous array of shape ``(nz, ny, nx, 8)``
-- the shape the reference's ``--system_in`` HDF files use
(src/test/linear/main.cpp:182-189).

The formulas restate how the reference assembles such systems
(src/solver/proj.ipp:343-398: face coefficient ``a_f = A_f*dt/(rho_f*h)``,
harmonic ``rho_f``, zero coefficient through non-periodic domain faces;
src/test/linear/main.cpp:44-92: the ``t.linear`` resistivity system).
SURVEY.md section 8(d) names them S1..S5.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "projection_rows",
    "projection_inputs",
    "cell_volume",
    "tlinear_system",
    "density_poisson_system",
    "periodic_constant_system",
    "random_spheres",
    "sphere_density",
]


def cell_volume(shape) -> float:
    """h**3 with h = extent/max(n), extent = 1 (src/distr/native.ipp:98)."""
    h = 1.0 / max(shape)
    return h * h * h


def _centres(n, h):
    return (np.arange(n, dtype=np.float64) + 0.5) * h


def _exact_tlinear(nz, ny, nx, h):
    # src/test/linear/main.cpp:47-52
    x = _centres(nx, h)[None, None, :]
    y = _centres(ny, h)[None, :, None]
    z = _centres(nz, h)[:, None, None]
    return np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y ** 2) * np.sin(2 * np.pi * z ** 3)


def _assemble(a_lo, periodic, sign, shape):
    """Rows from lower-face coefficients.

    a_lo[d][k,j,i] >= 0 is the coefficient of the face between cell (k,j,i)
    and its lower neighbour in direction d (d=0:x, 1:y, 2:z).  The upper-face
    coefficient of a cell is the lower-face coefficient of its upper neighbour
    (bitwise the same number, as in the reference where one face expression is
    appended to both cells, src/geom/mesh.h:575-579).  sign=+1: diag>0, off<0
    (proj.ipp); sign=-1: the negated system (t.linear).
    """
    nz, ny, nx = shape
    sys = np.zeros((nz, ny, nx, 8), dtype=np.float64)
    axis_of = {0: 2, 1: 1, 2: 0}
    diag = np.zeros(shape, dtype=np.float64)
    for d in range(3):
        ax = axis_of[d]
        lo = a_lo[d].copy()
        if not periodic[d]:
            idx = [slice(None)] * 3
            idx[ax] = 0
            lo[tuple(idx)] = 0.0  # face on the domain boundary: Neumann
        hi = np.roll(lo, -1, axis=ax)
        sys[..., 1 + 2 * d] = -sign * lo
        sys[..., 2 + 2 * d] = -sign * hi
        # the reference appends faces q=0..5 in order (main.cpp:71-76)
        diag += lo
        diag += hi
    sys[..., 0] = sign * diag
    return sys


def _apply(sys, v, periodic):
    """A*v with the reference's accumulation order (linear.ipp:65-72)."""
    out = v * sys[..., 0]
    for q in range(6):
        d, up = divmod(q, 2)
        ax = {0: 2, 1: 1, 2: 0}[d]
        nb = np.roll(v, -1 if up else 1, axis=ax)
        if not periodic[d]:
            idx = [slice(None)] * 3
            idx[ax] = -1 if up else 0
            nb[tuple(idx)] = 0.0
        out = out + nb * sys[..., 1 + q]
    return out


def tlinear_system(n, rho_in=10.0, shape=None):
    """The reference unit test's system (src/test/linear/main.cpp:44-92).

    Periodic; resistivity ``rho_in`` on faces whose centre is within 0.2 of the
    domain centre, 1 elsewhere; negative-definite sign (diag<0); the constant
    term is ``-A*exact`` so that ``exact`` solves it up to a constant.
    Returns (system, exact).
    """
    nz, ny, nx = shape if shape is not None else (n, n, n)
    h = 1.0 / max(nx, ny, nz)
    length = np.array([nx * h, ny * h, nz * h])
    xc, yc, zc = _centres(nx, h), _centres(ny, h), _centres(nz, h)
    a_lo = []
    for d in range(3):
        # lower face centre: shift by -h/2 in direction d
        fx = (xc - (h / 2 if d == 0 else 0))[None, None, :]
        fy = (yc - (h / 2 if d == 1 else 0))[None, :, None]
        fz = (zc - (h / 2 if d == 2 else 0))[:, None, None]
        dist = np.sqrt((fx - length[0] / 2) ** 2 + (fy - length[1] / 2) ** 2
                       + (fz - length[2] / 2) ** 2)
        rho = np.where(dist < 0.2, float(rho_in), 1.0)
        # (1/h)/rho * area, area = h*h   (main.cpp:73)
        a_lo.append((1.0 / h) / rho * (h * h) * np.ones((nz, ny, nx)))
    per = (True, True, True)
    sys = _assemble(a_lo, per, -1.0, (nz, ny, nx))
    exact = _exact_tlinear(nz, ny, nx, h)
    sys[..., 7] = -_apply(sys, exact, per)
    return sys, exact


def random_spheres(count, seed):
    """Sphere list of S2..S4: centres U[0.1,0.9]^3, radii U[0.03,0.08]."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.1, 0.9, size=(count, 3))
    r = rng.uniform(0.03, 0.08, size=count)
    return np.concatenate([c, r[:, None]], axis=1)


def sphere_density(shape, spheres, rho_out=1.0, rho_in=1e-3, z0=0, nz_global=None):
    """Cell density: rho_in inside any sphere, rho_out outside.

    `shape` may be a z-slab (nz_local, ny, nx) starting at plane z0 of a
    domain with nz_global planes.
    """
    nz, ny, nx = shape
    nzg = nz_global if nz_global is not None else nz
    h = 1.0 / max(nx, ny, nzg)
    x = _centres(nx, h)[None, None, :]
    y = _centres(ny, h)[None, :, None]
    z = ((np.arange(nz, dtype=np.float64) + z0 + 0.5) * h)[:, None, None]
    inside = np.zeros(shape, dtype=bool)
    for cx, cy, cz, r in spheres:
        k0 = max(int(np.floor((cz - r) / h - z0)) - 1, 0)
        k1 = min(int(np.ceil((cz + r) / h - z0)) + 1, nz)
        if k1 <= k0:
            continue
        j0 = max(int(np.floor((cy - r) / h)) - 1, 0)
        j1 = min(int(np.ceil((cy + r) / h)) + 1, ny)
        i0 = max(int(np.floor((cx - r) / h)) - 1, 0)
        i1 = min(int(np.ceil((cx + r) / h)) + 1, nx)
        sub = ((x[:, :, i0:i1] - cx) ** 2 + (y[:, j0:j1, :] - cy) ** 2
               + (z[k0:k1] - cz) ** 2) < r * r
        inside[k0:k1, j0:j1, i0:i1] |= sub
    return np.where(inside, float(rho_in), float(rho_out))


def density_poisson_system(n, nspheres=64, seed=20240601, rho_in=1e-3, rho_out=1.0,
                           dt=1e-3, periodic=(False, False, False), shape=None):
    """S2/S3/S4: variable-density projection system (diag>0, off<0).

    a_f = h*dt/rho_f with harmonic rho_f = 2/(1/rho_- + 1/rho_+); walls are
    Neumann (zero face coefficient); e7 = sum_q outward(q)*v_f with
    v_f = (u.n_f)*h^2 at face centres,
        u = (sin(pi x) cos(2 pi y), sin(pi y) cos(2 pi z), sin(pi z) cos(2 pi x)):
    a predicted velocity with non-zero divergence whose wall-normal component
    vanishes (wall fluxes are set to exactly zero), so sum(e7) telescopes to 0 and
    the singular system is consistent.  (SURVEY.md 8d proposed the field
    (sin 2pi x cos 2pi y, -cos 2pi x sin 2pi y, 0); it is discretely
    divergence-free, i.e. its right-hand side is rounding noise, so it is not used.)
    Returns (system, rho).
    """
    nz, ny, nx = shape if shape is not None else (n, n, n)
    h = 1.0 / max(nx, ny, nz)
    rho = sphere_density((nz, ny, nx), random_spheres(nspheres, seed), rho_out, rho_in)
    a_lo = []
    for d in range(3):
        ax = {0: 2, 1: 1, 2: 0}[d]
        rho_m = np.roll(rho, 1, axis=ax)
        rho_f = 2.0 / (1.0 / rho_m + 1.0 / rho)
        a_lo.append(h * dt / rho_f)
    sys = _assemble(a_lo, periodic, 1.0, (nz, ny, nx))
    xc, yc, zc = _centres(nx, h), _centres(ny, h), _centres(nz, h)
    # face-normal fluxes; index i of v* is the face at coordinate i*h
    xf = np.arange(nx + 1, dtype=np.float64) * h
    yf = np.arange(ny + 1, dtype=np.float64) * h
    zf = np.arange(nz + 1, dtype=np.float64) * h
    hh = h * h
    vx = (np.sin(np.pi * xf)[None, :] * np.cos(2 * np.pi * yc)[:, None]) * hh   # (ny, nx+1)
    vy = (np.sin(np.pi * yf)[None, :] * np.cos(2 * np.pi * zc)[:, None]) * hh   # (nz, ny+1)
    vz = (np.sin(np.pi * zf)[None, :] * np.cos(2 * np.pi * xc)[:, None]) * hh   # (nx, nz+1)
    for v, per in ((vx, periodic[0]), (vy, periodic[1]), (vz, periodic[2])):
        if not per:
            v[:, 0] = 0.0
            v[:, -1] = 0.0   # no flux through the walls
        else:
            v[:, -1] = v[:, 0]
    dx = (vx[:, 1:] - vx[:, :-1])[None, :, :]                    # (1, ny, nx)
    dy = (vy[:, 1:] - vy[:, :-1])[:, :, None]                    # (nz, ny, 1)
    dz = (vz[:, 1:] - vz[:, :-1]).T[:, None, :]                  # (nz, 1, nx)
    sys[..., 7] = (dx + dy) + dz
    return sys, rho


def periodic_constant_system(n, shape=None, dt=1.0):
    """S5: constant density, triply periodic, projection sign (diag>0).

    RHS chosen so that t.linear's exact solution (main.cpp:47-52) solves it.
    Returns (system, exact).
    """
    nz, ny, nx = shape if shape is not None else (n, n, n)
    h = 1.0 / max(nx, ny, nz)
    a = np.full((nz, ny, nx), h * dt, dtype=np.float64)
    per = (True, True, True)
    sys = _assemble([a, a, a], per, 1.0, (nz, ny, nx))
    exact = _exact_tlinear(nz, ny, nx, h)
    sys[..., 7] = -_apply(sys, exact, per)
    return sys, exact


def projection_rows(rho, vx, vy, vz, source=None, dt=1e-3, periodic=(False, False, False),
                    h=None, volume=None):
    """Rows of the projection step's pressure system from a cell density and face volume
    fluxes, in the reference's own order of operations (Proj::GetFlux + GetFluxSum,
    src/solver/proj.ipp:343-383, on a uniform mesh without embedded boundaries; every
    non-periodic domain face a wall):

        rho_f = 1 / ((1/rho_+ + 1/rho_-) * 0.5)           (approx_eb.h:351-363, ipp:834-836)
        k_f   = (1/h) * (((V/h) / rho_f) * dt)             (approx_eb.ipp:1440-1444, proj.ipp:356;
                                                            face area = V / h, mesh.ipp:91)
        e0 = (((((k_x- + k_x+) + k_y-) + k_y+) + k_z-) + k_z+     (AppendExpr, mesh.h:575-579)
        e[1+q] = -k_f(q)
        e7 = (((((-v_x- + v_x+) - v_y-) + v_y+) - v_z-) + v_z+) - source*V   (proj.ipp:377-379)

    This is the host statement of what aphcg_assemble_projection computes on the device; it is
    pinned bit for bit to the reference's functions by tests/test_oracle.py
    (oracle/_ref/ref_assemble and tests/golden/assemble_*.npz).
    rho (nz,ny,nx); vx (nz,ny,nx+1), vy (nz,ny+1,nx), vz (nz+1,ny,nx).  Returns (nz,ny,nx,8)."""
    rho = np.asarray(rho, dtype=np.float64)
    shape = rho.shape
    nz, ny, nx = shape
    h = 1.0 / max(shape) if h is None else float(h)
    vol = cell_volume(shape) if volume is None else float(volume)
    area = vol / h
    inv = 1.0 / rho
    rows = np.zeros(shape + (8,), dtype=np.float64)
    diag = np.zeros(shape, dtype=np.float64)
    axis_of = {0: 2, 1: 1, 2: 0}
    for d in range(3):
        ax = axis_of[d]
        # lower face of every cell: between the cell (plus side) and its lower neighbour
        inv_m = np.roll(inv, 1, axis=ax)
        rho_f = 1.0 / ((inv + inv_m) * 0.5)
        k_lo = (1.0 / h) * ((area / rho_f) * dt)
        if not periodic[d]:
            idx = [slice(None)] * 3
            idx[ax] = 0
            k_lo[tuple(idx)] = 0.0
        k_hi = np.roll(k_lo, -1, axis=ax)
        rows[..., 1 + 2 * d] = -k_lo
        rows[..., 2 + 2 * d] = -k_hi
        diag = diag + k_lo
        diag = diag + k_hi
    rows[..., 0] = diag
    e7 = -vx[:, :, :-1] + vx[:, :, 1:]
    e7 = e7 - vy[:, :-1, :]
    e7 = e7 + vy[:, 1:, :]
    e7 = e7 - vz[:-1]
    e7 = e7 + vz[1:]
    if source is not None:
        e7 = e7 - np.asarray(source, dtype=np.float64) * vol
    else:
        e7 = e7 - 0.0 * vol
    rows[..., 7] = e7
    return rows


def projection_inputs(local_shape, spheres, rho_in=1e-3, rho_out=1.0, z0=0, nz_global=None):
    """What a caller of aphcg_assemble_projection holds for the S2..S4 bubble systems (walls on
    all sides): the cell density of a z-slab WITH its two ghost planes, and the face volume
    fluxes of density_poisson_system's velocity field.  Returns (rho (nzl+2,ny,nx),
    vx (nzl,ny,nx+1), vy (nzl,ny+1,nx), vz (nzl+1,ny,nx)); the caller may pass preallocated
    (e.g. pinned) arrays through `out=(rho, vx, vy, vz)`."""
    nzl, ny, nx = local_shape
    nzg = nz_global if nz_global is not None else nzl
    h = 1.0 / max(nx, ny, nzg)
    # planes z0-1 .. z0+nzl; outside the domain: a copy of the boundary plane (any finite
    # value does: the wall coefficient is zero)
    lo, hi = max(z0 - 1, 0), min(z0 + nzl + 1, nzg)
    core = sphere_density((hi - lo, ny, nx), spheres, rho_out, rho_in, z0=lo, nz_global=nzg)
    rho = np.concatenate(([core[:1]] if z0 == 0 else []) + [core]
                         + ([core[-1:]] if z0 + nzl == nzg else []))
    xc, yc = _centres(nx, h), _centres(ny, h)
    zc = (np.arange(nzl, dtype=np.float64) + z0 + 0.5) * h
    xf = np.arange(nx + 1, dtype=np.float64) * h
    yf = np.arange(ny + 1, dtype=np.float64) * h
    zf = (np.arange(nzl + 1, dtype=np.float64) + z0) * h
    hh = h * h
    sx, sy, sz = np.sin(np.pi * xf), np.sin(np.pi * yf), np.sin(np.pi * zf)
    sx[0] = sx[-1] = 0.0            # no flux through the walls
    sy[0] = sy[-1] = 0.0
    if z0 == 0:
        sz[0] = 0.0
    if z0 + nzl == nzg:
        sz[-1] = 0.0
    vx = np.broadcast_to(((sx[None, :] * np.cos(2 * np.pi * yc)[:, None]) * hh)[None], (nzl, ny, nx + 1))
    vy = np.broadcast_to(((sy[None, :] * np.cos(2 * np.pi * zc)[:, None]) * hh)[:, :, None], (nzl, ny + 1, nx))
    vz = np.broadcast_to(((sz[:, None] * np.cos(2 * np.pi * xc)[None, :]) * hh)[:, None, :], (nzl + 1, ny, nx))
    return rho, vx, vy, vz
