"""digital_filtering_b200 -- host-side mirror of the reference's DIGITAL_FILTER interface over the C ABI.

The directory is named `digital-filtering_b200` (not importable as written); `_dfb_import.py` at the
repo root registers it as the module `digital_filtering_b200`.

The compute path is libdfb200.so (hand-written sm_100a CUDA, csrc/).  Nothing here computes:
this module only marshals arguments through ctypes, exactly like the C++ facade
(include/digital_filter.hpp) and the Fortran module (fortran/digital_filtering.f90) do.  If the
library is missing the import fails loudly; there is no Python / CPU fallback.

Reference surface mirrored (digital-filtering-c++/df/df.hpp):
    DFConfig                 df.hpp:38-49
    DIGITAL_FILTER(config)   df.hpp:89, df.cpp:4-66
    .filter(dt_input)        df.hpp:100, df.cpp:449-468
    .u/.v/.w (.fluc, .filt)  df.hpp:24-34, 87;  .T_fluc / .rho_fluc  df.hpp:59
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFB_LIB") or os.path.join(_HERE, "lib", "libdfb200.so")   # DFB_LIB: development builds side by side

OK, ERR_ARG, ERR_IO, ERR_CUDA, ERR_STATE, ERR_INTERP = range(6)
U_FLUC, V_FLUC, W_FLUC, T_FLUC, RHO_FLUC, U_FILT, V_FILT, W_FILT = range(8)
NOISE_GENERATE, NOISE_INJECT = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class DfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dfb200 error {code}: {msg}")
        self.code = code


class dfb_config(C.Structure):
    """ctypes image of `struct dfb_config` (include/dfb200.h)."""
    _fields_ = [
        ("d_i", C.c_double), ("rho_e", C.c_double), ("U_e", C.c_double), ("mu_e", C.c_double),
        ("vel_file_offset", C.c_int), ("vel_file_N_values", C.c_int),
        ("grid_file", C.c_char_p), ("grid_file_len", C.c_int),
        ("vel_fluc_file", C.c_char_p), ("vel_fluc_file_len", C.c_int),
        ("struct_bytes", C.c_int), ("honor_flow_config", C.c_int),
        ("line_file", C.c_char_p), ("line_file_len", C.c_int),
        ("Ny", C.c_int), ("Nz", C.c_int), ("geom_per_row", C.c_int),
        ("yc", c_dp), ("dy", c_dp), ("dz", c_dp), ("rows", c_dp), ("scales", c_dp),
        ("N_y", c_ip), ("N_z", c_ip),
        ("seed", C.c_uint64), ("noise_mode", C.c_int), ("device", C.c_int), ("plane_id", C.c_int),
        ("k_begin", C.c_int), ("k_end", C.c_int), ("skip_first_step", C.c_int), ("kernel_variant", C.c_int),
    ]


_lib = None


def lib():
    """Loads libdfb200.so.  Missing library = hard error (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C digital-filtering_b200/csrc). There is no CPU fallback.")
        # libdfb200.so links libnccl.so.2 (config 4's hand-off).  If PyTorch's bundled NCCL is installed, bring that copy in first so that
        # the process holds ONE NCCL whichever of torch / this library is imported first (same soname: the loader would otherwise hand
        # torch the system copy, which may be older than the one it was built against).
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            if spec and spec.submodule_search_locations:
                cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    C.CDLL(cand, mode=C.RTLD_GLOBAL)
        except Exception:
            pass
        L = C.CDLL(LIB_PATH)
        L.dfb_last_error.restype = C.c_char_p
        L.dfb_version.restype = C.c_char_p
        L.dfb_create.argtypes = [C.POINTER(dfb_config), C.POINTER(C.c_void_p)]
        L.dfb_create_batch.argtypes = [C.POINTER(dfb_config), C.c_int, C.POINTER(C.c_void_p)]
        L.dfb_num_planes.argtypes = [C.c_void_p, c_ip]
        L.dfb_get_field_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.dfb_device_ptr_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.dfb_stats_get_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, C.POINTER(C.c_int64)]
        L.dfb_get_rms.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.dfb_write_rms_csv.argtypes = [C.c_void_p, C.c_char_p]
        L.dfb_write_tecplot.argtypes = [C.c_void_p, C.c_char_p]
        L.dfb_face_map.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp, c_ip]
        L.dfb_comm_unique_id.argtypes = [C.c_void_p]
        L.dfb_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.dfb_comm_info.argtypes = [C.c_void_p, c_ip, c_ip, c_ip]
        L.dfb_gather_begin.argtypes = [C.c_void_p, C.c_int]
        L.dfb_gather_end.argtypes = [C.c_void_p]
        L.dfb_gathered_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.dfb_gathered_to_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.dfb_gather_wire_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.dfb_comm_destroy.argtypes = [C.c_void_p]
        L.dfb_comm_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        for name in ("dfb_destroy", "dfb_sync", "dfb_first_step"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.dfb_dims.argtypes = [C.c_void_p, c_ip, c_ip]
        L.dfb_info.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.dfb_get_table.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int]
        L.dfb_get_half_widths.argtypes = [C.c_void_p, C.c_int, C.c_int, c_ip]
        L.dfb_filter.argtypes = [C.c_void_p, C.c_double]
        L.dfb_scatter_to_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.dfb_filter_to_host.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 5
        L.dfb_filter_to_host_begin.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 5
        L.dfb_filter_to_host_end.argtypes = [C.c_void_p]
        L.dfb_filter_batch.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_void_p]
        L.dfb_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.dfb_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.dfb_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.dfb_set_noise.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.dfb_set_noise_ref_layout.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.dfb_get_noise.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.dfb_generate_noise.argtypes = [C.c_void_p, C.c_int64]
        L.dfb_get_state.argtypes = [C.c_void_p, c_dp, C.POINTER(C.c_int64)]
        L.dfb_set_state.argtypes = [C.c_void_p, c_dp, C.c_int64]
        L.dfb_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.dfb_last_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.dfb_measure_fp64_peak.argtypes = [C.c_int, c_dp, c_dp]
        L.dfb_stats_enable.argtypes = [C.c_void_p, C.c_int]
        L.dfb_stats_get.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.POINTER(C.c_int64)]
        L.dfb_write_csv.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise DfbError(rc, lib().dfb_last_error().decode(errors="replace"))


def _dptr(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def measure_fp64_peak(device=-1):
    """DFMA microbenchmark -> (TFLOP/s, implied SM MHz at 64 DFMA/clk/SM)."""
    t, m = C.c_double(), C.c_double()
    _check(lib().dfb_measure_fp64_peak(device, C.byref(t), C.byref(m)))
    return t.value, m.value


class DFConfig:
    """The reference's DFConfig (df.hpp:38-49) plus the extension fields of dfb_config.
    `dc_config` (the README's spelling, README.md:21) is an alias."""

    def __init__(self, **kw):
        self.d_i = self.rho_e = self.U_e = self.mu_e = 0.0
        self.vel_file_offset = self.vel_file_N_values = 0
        self.grid_file = ""
        self.vel_fluc_file = ""
        # extensions
        self.honor_flow_config = 0
        self.line_file = ""
        self.Ny = self.Nz = 0
        self.geom_per_row = 1
        self.yc = self.dy = self.dz = self.rows = self.scales = self.N_y = self.N_z = None
        self.seed = 0
        self.noise_mode = NOISE_GENERATE
        self.device = -1
        self.plane_id = 0
        self.k_begin = self.k_end = 0
        self.skip_first_step = 0
        self.kernel_variant = 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"DFConfig has no field {k!r}")
            setattr(self, k, v)

    @classmethod
    def from_plane(cls, plane, **kw):
        """Explicit-geometry config from a workload dict (workloads.py): Ny, Nz, d_i, yc, dy, dz, rows, scales
        [, N_y, N_z].  The half-widths are computed by the library from the geometry unless given."""
        c = cls(Ny=plane["Ny"], Nz=plane["Nz"], d_i=plane["d_i"], U_e=plane.get("U_e", 869.1), honor_flow_config=1,
                yc=plane["yc"], dy=plane["dy"], dz=plane["dz"], rows=plane["rows"], scales=plane["scales"],
                geom_per_row=1 if np.ndim(plane["yc"]) == 1 else 0)
        if plane.get("explicit_N"):
            c.N_y, c.N_z = plane["N_y"], plane["N_z"]
        for k, v in kw.items():
            setattr(c, k, v)
        return c


dc_config = DFConfig


class FilterField:
    """Host view of one velocity component (df.hpp:24-34): .fluc and .filt, Ny x Nz, row-major j*Nz+k."""

    def __init__(self, owner, idx):
        self._o, self._i = owner, idx
        self.fluc = np.zeros((owner.Ny, owner.Nz))
        self.filt = np.zeros((owner.Ny, owner.Nz))

    @property
    def Ny_max(self):
        return self._o.info(0, self._i)

    @property
    def Nz_max(self):
        return self._o.info(1, self._i)

    @property
    def N_ys(self):
        return self._o.half_widths(self._i, 0)

    @property
    def N_zs(self):
        return self._o.half_widths(self._i, 1)


class DIGITAL_FILTER:
    """DIGITAL_FILTER (df.hpp:52-125) on one B200.  Construction = df.cpp:4-66 (setup + first step);
    filter(dt) = df.cpp:449-468 without the print and the CSV."""

    def __init__(self, config=None, fetch=True, nplanes=1):
        """nplanes > 1: a batch of independent planes of this geometry advanced by one launch set per step (dfb_create_batch,
        BASELINE config 5); plane p draws from stream group config.plane_id + p.  The host members then describe plane 0;
        get(which, plane=p) / device_ptr(which, plane=p) reach the others."""
        config = config or DFConfig()
        L = lib()
        c = dfb_config()
        _check(L.dfb_config_init(C.byref(c)))
        self._keep = []

        def arr(a, dtype):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(a)
            return a

        c.d_i, c.rho_e, c.U_e, c.mu_e = config.d_i, config.rho_e, config.U_e, config.mu_e
        c.vel_file_offset, c.vel_file_N_values = config.vel_file_offset, config.vel_file_N_values
        c.grid_file = config.grid_file.encode() if config.grid_file else None
        c.vel_fluc_file = config.vel_fluc_file.encode() if config.vel_fluc_file else None
        c.line_file = config.line_file.encode() if config.line_file else None
        c.honor_flow_config = config.honor_flow_config
        c.Ny, c.Nz, c.geom_per_row = config.Ny, config.Nz, config.geom_per_row
        for name in ("yc", "dy", "dz", "rows", "scales"):
            a = arr(getattr(config, name), np.float64)
            setattr(c, name, _dptr(a))
        for name in ("N_y", "N_z"):
            a = arr(getattr(config, name), np.int32)
            setattr(c, name, a.ctypes.data_as(c_ip) if a is not None else None)
        c.seed, c.noise_mode, c.device, c.plane_id = config.seed, config.noise_mode, config.device, config.plane_id
        c.k_begin, c.k_end = config.k_begin, config.k_end
        c.skip_first_step, c.kernel_variant = config.skip_first_step, config.kernel_variant
        self._h = C.c_void_p()
        self.nplanes = int(nplanes)
        if self.nplanes == 1:
            _check(L.dfb_create(C.byref(c), C.byref(self._h)))
        else:
            _check(L.dfb_create_batch(C.byref(c), self.nplanes, C.byref(self._h)))
            fetch = False
        ny, nz = C.c_int(), C.c_int()
        _check(L.dfb_dims(self._h, C.byref(ny), C.byref(nz)))
        self.Ny, self.Nz, self.n_cells = ny.value, nz.value, ny.value * nz.value
        self.noise_mode = config.noise_mode
        self._device = self.info(6)
        self.u, self.v, self.w = (FilterField(self, i) for i in range(3))
        self.T_fluc = np.zeros((self.Ny, self.Nz))
        self.rho_fluc = np.zeros((self.Ny, self.Nz))
        self.dt = 0.0
        self._fetch = fetch
        if fetch and config.noise_mode == NOISE_GENERATE and not config.skip_first_step:
            self.fetch()

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            lib().dfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's entry point ----
    def filter(self, dt_input):
        """One timestep.  With fetch=True (default) the five outputs land in .u.fluc, .v.fluc, .w.fluc,
        .T_fluc, .rho_fluc (host), like the reference's vectors; with fetch=False they stay on the device."""
        self.dt = float(dt_input)
        if self._fetch:
            _check(lib().dfb_filter_to_host(self._h, self.dt, self.u.fluc.ctypes.data, self.v.fluc.ctypes.data,
                                            self.w.fluc.ctypes.data, self.T_fluc.ctypes.data, self.rho_fluc.ctypes.data))
        else:
            _check(lib().dfb_filter(self._h, self.dt))

    def first_step(self):
        _check(lib().dfb_first_step(self._h))

    def filter_batch(self, dts, out=None):
        dts = np.ascontiguousarray(dts, dtype=np.float64)
        _check(lib().dfb_filter_batch(self._h, len(dts), _dptr(dts), out.ctypes.data if out is not None else None))

    # ---- data access ----
    def get(self, which, out=None, plane=0):
        out = np.zeros((self.Ny, self.Nz)) if out is None else out
        _check(lib().dfb_get_field_plane(self._h, plane, which, out.ctypes.data, 0))
        return out

    def get_to_device(self, which, device_ptr):
        _check(lib().dfb_get_field(self._h, which, C.c_void_p(device_ptr), 1))

    def device_ptr(self, which, plane=0):
        p = C.c_void_p()
        _check(lib().dfb_device_ptr_plane(self._h, plane, which, C.byref(p)))
        return p.value

    def device_tensor(self, which):
        """zero-copy torch view (Ny, Nz) of a device-resident field (for NCCL hand-offs)"""
        import torch

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = dict(shape=(self.Ny, self.Nz), typestr="<f8", data=(self.device_ptr(which), False),
                                          version=3, strides=None)
        return torch.as_tensor(v, device=torch.device("cuda", self._device))

    def scatter_to_cells(self, which, plane_index, dst_index, dst, mean=None, scale=1.0):
        """CFD hand-off on the device (dfb_scatter_to_cells): dst[dst_index[i]] = (mean[i] | dst[dst_index[i]]) + scale * field[plane_index[i]].
        Arguments are CUDA torch tensors (int32 indices, float64 values) on this handle's device."""
        n = int(plane_index.numel())
        _check(lib().dfb_scatter_to_cells(self._h, which, n, C.c_void_p(plane_index.data_ptr()), C.c_void_p(dst_index.data_ptr()),
                                          C.c_void_p(mean.data_ptr()) if mean is not None else None, C.c_double(scale), C.c_void_p(dst.data_ptr())))

    def stream(self):
        p = C.c_void_p()
        _check(lib().dfb_stream(self._h, C.byref(p)))
        return p.value

    def fetch(self):
        """Copy all outputs (and filt) to the host members."""
        for f, F in enumerate((self.u, self.v, self.w)):
            self.get(U_FLUC + f, F.fluc)
            self.get(U_FILT + f, F.filt)
        self.get(T_FLUC, self.T_fluc)
        self.get(RHO_FLUC, self.rho_fluc)
        return self

    def sync(self):
        _check(lib().dfb_sync(self._h))

    def info(self, what, field=0):
        v = C.c_int64()
        _check(lib().dfb_info(self._h, what, field, C.byref(v)))
        return v.value

    @property
    def step(self):
        return self.info(2)

    @property
    def tuned(self):
        return bool(self.info(3))

    @property
    def taps_per_step(self):
        return self.info(4)

    def table(self, which, arg=0, n=None):
        n = n if n is not None else (3 if which == 10 else (2 * arg + 1 if which == 11 else self.Ny))
        out = np.zeros(n)
        _check(lib().dfb_get_table(self._h, which, arg, _dptr(out), n))
        return out

    def rows(self):
        return np.stack([self.table(i) for i in range(8)])

    def half_widths(self, field, direction):
        out = np.zeros((self.Ny, self.Nz), dtype=np.int32)
        _check(lib().dfb_get_half_widths(self._h, field, direction, out.ctypes.data_as(c_ip)))
        return out

    # ---- noise ----
    def set_noise(self, field, r_ys, halo):
        r_ys = np.ascontiguousarray(r_ys, dtype=np.float64)
        halo = np.ascontiguousarray(halo, dtype=np.float64) if halo is not None else None
        _check(lib().dfb_set_noise(self._h, field, _dptr(r_ys), _dptr(halo)))

    def set_noise_ref_layout(self, field, r_ys, r_zs):
        r_ys = np.ascontiguousarray(r_ys, dtype=np.float64)
        r_zs = np.ascontiguousarray(r_zs, dtype=np.float64)
        _check(lib().dfb_set_noise_ref_layout(self._h, field, _dptr(r_ys), _dptr(r_zs)))

    def get_noise(self, field):
        F = (self.u, self.v, self.w)[field]
        r_ys = np.zeros((self.Ny + 2 * F.Ny_max, self.Nz))
        halo = np.zeros((self.Ny, 2 * F.Nz_max))
        _check(lib().dfb_get_noise(self._h, field, _dptr(r_ys), _dptr(halo)))
        return r_ys, halo

    def generate_noise(self, step):
        _check(lib().dfb_generate_noise(self._h, int(step)))

    # ---- checkpoint ----
    def get_state(self):
        fo = np.zeros((3, self.Ny, self.Nz)) if self.nplanes == 1 else np.zeros((3, self.nplanes, self.Ny, self.Nz))
        s = C.c_int64()
        _check(lib().dfb_get_state(self._h, _dptr(fo), C.byref(s)))
        return fo, s.value

    def set_state(self, filt_old, step):
        fo = np.ascontiguousarray(filt_old, dtype=np.float64) if filt_old is not None else None
        _check(lib().dfb_set_state(self._h, _dptr(fo), int(step)))

    # ---- N2 running statistics / N4 CSV (both opt-in; filter() itself never writes or accumulates) ----
    def stats_enable(self, on=True):
        _check(lib().dfb_stats_enable(self._h, int(on)))

    def stats(self, which, rms=False, plane=0):
        """which: 0 u'^2, 1 v'^2, 2 w'^2, 3 T'^2, 4 rho'^2, 5 u'v' (per-cell sums, or sqrt(sum/count) with rms=True).
        Returns (array, accumulated steps)."""
        out = np.zeros((self.Ny, self.Nz))
        cnt = C.c_int64()
        _check(lib().dfb_stats_get_plane(self._h, plane, which, int(rms), _dptr(out), C.byref(cnt)))
        return out, cnt.value

    def get_rms(self, nsteps=500, dt=1e-5, path=None):
        """get_rms (df.cpp:584-611): 500 steps of dt = 1e-5 with on-device accumulation; plot_rms's file when `path` is given."""
        _check(lib().dfb_get_rms(self._h, int(nsteps), float(dt)))
        if path is not None:
            _check(lib().dfb_write_rms_csv(self._h, str(path).encode()))

    def write_csv(self, path):
        _check(lib().dfb_write_csv(self._h, str(path).encode()))

    def write_tecplot(self, path):
        _check(lib().dfb_write_tecplot(self._h, str(path).encode()))

    def face_map(self, yf, zf):
        """plane_index (j*Nz + k within this handle's slab, -1 outside it) of the cells containing the face centres (yf, zf)"""
        yf = np.ascontiguousarray(yf, dtype=np.float64)
        zf = np.ascontiguousarray(zf, dtype=np.float64)
        out = np.zeros(len(yf), dtype=np.int32)
        _check(lib().dfb_face_map(self._h, len(yf), _dptr(yf), _dptr(zf), out.ctypes.data_as(c_ip)))
        return out

    # ---- config 4: this handle as one spanwise slab of a plane shared with the other ranks (NCCL inside the library) ----
    @staticmethod
    def comm_unique_id():
        """rank 0: the 128-byte NCCL id every rank passes to comm_init (ship it with the job's own means)"""
        buf = C.create_string_buffer(128)
        _check(lib().dfb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, ident, rank, world):
        _check(lib().dfb_comm_init(self._h, C.c_char_p(ident), int(rank), int(world)))
        self._NzG = self.info(5)

    def comm_bounds(self):
        r, w = C.c_int(), C.c_int()
        _check(lib().dfb_comm_info(self._h, C.byref(r), C.byref(w), None))
        b = np.zeros(2 * w.value, dtype=np.int32)
        _check(lib().dfb_comm_info(self._h, None, None, b.ctypes.data_as(c_ip)))
        return [(int(b[2 * i]), int(b[2 * i + 1])) for i in range(w.value)]

    def gather_begin(self, dst=0):
        _check(lib().dfb_gather_begin(self._h, int(dst)))

    def gather_end(self):
        _check(lib().dfb_gather_end(self._h))

    def gathered(self, which, out=None):
        """destination rank: host copy (Ny, Nz_global) of a gathered field"""
        out = np.zeros((self.Ny, self._NzG)) if out is None else out
        _check(lib().dfb_gathered_to_host(self._h, which, out.ctypes.data))
        return out

    def gathered_ptr(self, which):
        p = C.c_void_p()
        _check(lib().dfb_gathered_ptr(self._h, which, C.byref(p)))
        return p.value

    def comm_stream(self):
        p = C.c_void_p()
        _check(lib().dfb_comm_stream(self._h, C.byref(p)))
        return p.value

    def gather_wire_bytes(self):
        v = C.c_int64()
        _check(lib().dfb_gather_wire_bytes(self._h, C.byref(v)))
        return v.value

    # ---- timing ----
    def set_timing(self, on=True):
        _check(lib().dfb_set_timing(self._h, int(on)))

    def last_ms(self):
        out = []
        for s in range(4):
            v = C.c_float()
            _check(lib().dfb_last_ms(self._h, s, C.byref(v)))
            out.append(v.value)
        return dict(noise=out[0], ysweep=out[1], zsweep_epilogue=out[2], step=out[3])
