"""Synthetic inflow planes for the named benchmark / parity shapes (BASELINE.json configs 2-4).

A plane is a dict of INPUTS only -- geometry (yc, dy, dz per row), row tables, integral scales -- in
the form both the library (DFConfig.from_plane) and the reference object (oracle RefFilter.reshape)
accept; the half-widths are always derived from the geometry by whoever consumes the plane
(calculate_filter_properties, df.cpp:130-218), never supplied, so that product and checker each
run their own setup.  Pure numpy; no compute path lives here.
"""
import numpy as np

D_I = 0.0013      # df.cpp:7
U_E = 869.1       # df.cpp:9


def _scales(d_i, s=1.0, U_e=U_E):
    """per field (Iz_inn, Iz_out, Lt), df.cpp:35-45, integral lengths scaled by s"""
    d_v = d_i / 4500.0            # df.cpp:326
    return np.array([[s * 150 * d_v, s * 0.4 * d_i, 0.8 * d_i / U_e],
                     [s * 75 * d_v, s * 0.3 * d_i, 0.3 * d_i / U_e],
                     [s * 150 * d_v, s * 0.4 * d_i, 0.3 * d_i / U_e]])


def synthetic_rows(yc, d_i=D_I, U_e=U_E):
    """Smooth analytic stand-ins for the DNS row tables: [R11,R21,R22,R33,Us,Ts,rhos,Ms] x Ny with
    R11*R22 > R21^2 everywhere and Us > 0."""
    eta = np.asarray(yc) / d_i
    A = 100.0 * (0.2 + 4.0 * np.sqrt(eta) * np.exp(-2.0 * eta))
    Us = U_e * np.tanh(3.0 * eta) + 1.0
    Ts = 55.2 + 42.3 * (1.0 - np.tanh(2.0 * eta))
    rhos = 0.044 * 55.2 / Ts
    Ms = Us / np.sqrt(1.4 * 287.0 * Ts)
    return np.stack([A, -0.25 * A, 0.4 * A, 0.6 * A, Us, Ts, rhos, Ms])


def tanh_grid(Ny, y_max, a=2.0):
    """the reference's made-up wall-normal stretching (df.cpp:92-117) for any Ny"""
    j = np.arange(Ny, -1, -1, dtype=np.float64)
    y = y_max * (1.0 - np.tanh(a * (j / (Ny + 1))) / np.tanh(a))
    return 0.5 * (y[:-1] + y[1:]), np.diff(y)


def _N_from_geometry(yc, dy, dz, d_i, Iz_inn, Iz_out):
    Iz = Iz_inn + (Iz_out - Iz_inn) * 0.5 * (1 + np.tanh((yc / d_i - 0.2) / 0.03))
    return (2 * np.maximum(1.0, 0.67 * Iz / dy).astype(np.int64), 2 * np.maximum(1.0, Iz / dz).astype(np.int64))


def plane_profile(Ny, Nz, Ny_target, Nz_target, d_i=D_I):
    """Reference-shaped half-width profiles (tanh grid, tanh Iz blend) scaled so that the u field's
    largest N_y / N_z are exactly the targets ("3a profile")."""
    yc, dy = tanh_grid(Ny, 3.0 * d_i)
    lo, hi = 1e-3, 1e3
    for _ in range(200):                      # largest s with max N_y <= target
        s = np.sqrt(lo * hi)
        sc = _scales(d_i, s)
        ny, _ = _N_from_geometry(yc, dy, np.ones_like(yc), d_i, sc[0, 0], sc[0, 1])
        if ny.max() > Ny_target:
            hi = s
        else:
            lo = s
    sc = _scales(d_i, lo)
    dz = np.full(Ny, sc[0, 1] / (Nz_target // 2 + 0.5))
    ny, nz = _N_from_geometry(yc, dy, dz, d_i, sc[0, 0], sc[0, 1])
    assert ny.max() == Ny_target and nz.max() == Nz_target, (ny.max(), nz.max())
    return dict(name=f"{Ny}x{Nz}_profile_N{Ny_target}", Ny=Ny, Nz=Nz, d_i=d_i, U_e=U_E, yc=yc, dy=dy, dz=dz,
                rows=synthetic_rows(yc, d_i), scales=sc)


def plane_saturated(Ny, Nz, N, d_i=D_I):
    """Every cell of every field has N_y = N_z = N ("3b saturated": the pure compute-bound regime)."""
    I0 = 0.4 * d_i
    sc = _scales(d_i)
    sc[:, 0] = I0
    sc[:, 1] = I0
    dy = np.full(Ny, 0.67 * I0 / (N // 2 + 0.5))
    dz = np.full(Ny, I0 / (N // 2 + 0.5))
    yc = (np.arange(Ny) + 0.5) * dy
    ny, nz = _N_from_geometry(yc, dy, dz, d_i, I0, I0)
    assert ny.min() == ny.max() == N and nz.min() == nz.max() == N, (ny.min(), ny.max(), nz.min(), nz.max())
    return dict(name=f"{Ny}x{Nz}_saturated_N{N}", Ny=Ny, Nz=Nz, d_i=d_i, U_e=U_E, yc=yc, dy=dy, dz=dz,
                rows=synthetic_rows(yc, d_i), scales=sc)


def plane_ragged(Ny, Nz, Nmax, seed=0, d_i=D_I):
    """Per-cell geometry that is NOT uniform along z (exercises the general, non row-uniform path):
    dy and dz wobble from column to column."""
    rng = np.random.default_rng(seed)
    yc1, dy1 = tanh_grid(Ny, 3.0 * d_i)
    sc = _scales(d_i)
    wob = 1.0 + 0.6 * rng.random((Ny, Nz))
    yc = np.repeat(yc1[:, None], Nz, axis=1)
    dy = np.repeat(dy1[:, None], Nz, axis=1) * wob
    dz = np.full((Ny, Nz), sc[0, 1] / (Nmax // 2 + 0.5)) * (1.0 + 0.6 * rng.random((Ny, Nz)))
    s = 1.0
    for _ in range(60):
        ny, _ = _N_from_geometry(yc, dy * s, dz, d_i, sc[0, 0], sc[0, 1])
        if ny.max() <= Nmax:
            break
        s *= 1.1
    return dict(name=f"{Ny}x{Nz}_ragged_N{Nmax}", Ny=Ny, Nz=Nz, d_i=d_i, U_e=U_E, yc=yc, dy=dy * s, dz=dz,
                rows=synthetic_rows(yc1, d_i), scales=sc)


NAMED = {
    # BASELINE.json configs[1]: 512x512, half-widths <= 32
    "512x512_N32": lambda: plane_profile(512, 512, 32, 32),
    # configs[2]: 1024x2048, half-widths up to 128 -- both variants of SURVEY 8d
    "1024x2048_profile_N128": lambda: plane_profile(1024, 2048, 128, 128),
    "1024x2048_saturated_N128": lambda: plane_saturated(1024, 2048, 128),
    # configs[3]: 4096x8192 (sharded spanwise)
    "4096x8192_profile_N128": lambda: plane_profile(4096, 8192, 128, 128),
    "4096x8192_saturated_N128": lambda: plane_saturated(4096, 8192, 128),
}


def taps_per_step(N_y, N_z):
    """algorithmic tap-FMAs of one step: sum over fields and cells of (2N_y+1)+(2N_z+1)  """
    return int((2 * np.asarray(N_y, dtype=np.int64) + 1).sum() + (2 * np.asarray(N_z, dtype=np.int64) + 1).sum())
