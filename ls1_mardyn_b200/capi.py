"""ctypes binding of libls1b200.so (include/ls1b200.h) -- the host-side mirror, in Python, of the calls an ls1-MarDyn
adapter makes (INTEGRATION.md shows the C++ adapter). ``Context`` exposes the reference's phases by their reference names:

    update()            LinkedCells::update + DomainDecompBase::exchangeMolecules + updateMoleculeCaches
    traverse_cells()    ParticleContainer::traverseCells(VectorizedCellProcessor&) + Simulation::updateForces
    leapfrog_pre/post   Leapfrog::eventNewTimestep / eventForcesCalculated
    scale_velocities    VelocityScalingThermostat::apply
    simulate            Simulation::simulate (hot loop only)

There is no CPU fallback: loading fails loudly if the CUDA library is missing, creating a context fails without a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .components import CComponent, CSite, Ensemble

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LS1B200_LIB") or os.path.join(_HERE, "libls1b200.so")      # LS1B200_LIB: another build of the same library (kernel sweeps)
_LIB = None

FIELDS = dict(ID=0, CID=1, R=2, V=3, Q=4, L=5, F=6, M=7, VI=8, CELL=9)
_FIELD_SHAPE = {0: (np.uint64, ()), 1: (np.uint32, ()), 2: (np.float64, (3,)), 3: (np.float64, (3,)), 4: (np.float64, (4,)),
                5: (np.float64, (3,)), 6: (np.float64, (3,)), 7: (np.float64, (3,)), 8: (np.float64, (3,)), 9: (np.uint64, ())}

EXPORTS = ["ls1b200_last_error", "ls1b200_abi_version", "ls1b200_kernel_launch_count", "ls1b200_create", "ls1b200_destroy",
           "ls1b200_set_components", "ls1b200_upload", "ls1b200_num_molecules", "ls1b200_download", "ls1b200_set_field",
           "ls1b200_geometry", "ls1b200_update", "ls1b200_traverse_cells", "ls1b200_leapfrog_pre", "ls1b200_leapfrog_post",
           "ls1b200_scale_velocities", "ls1b200_selective_thermostat", "ls1b200_simulate", "ls1b200_halo_export", "ls1b200_halo_import",
           "ls1b200_halo_import_host", "ls1b200_phase_times", "ls1b200_count_flops", "ls1b200_measure_fp32_peak", "ls1b200_measure_fp64_peak",
           "ls1b200_measure_hbm_copy", "ls1b200_stream", "ls1b200_nccl_unique_id", "ls1b200_decomp_attach", "ls1b200_download_begin", "ls1b200_download_end", "ls1b200_peer_handle", "ls1b200_peer_attach", "ls1b200_peer_detach", "ls1b200_exchange",
           "ls1b200_allreduce_macro", "ls1b200_halo_export_host", "ls1b200_generate_cubic_grid", "ls1b200_rmm_upload", "ls1b200_rmm_start", "ls1b200_rmm_simulate", "ls1b200_rmm_download", "ls1b200_rmm_state_bytes", "ls1b200_exchange_overlapped_ms"]


class Ls1B200Error(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("box_min", C.c_double * 3), ("box_max", C.c_double * 3), ("cutoff", C.c_double), ("cutoff_lj", C.c_double),
                ("epsilon_rf", C.c_double), ("precision", C.c_int32), ("geometry", C.c_int32), ("periodic", C.c_int32 * 3),
                ("device", C.c_int32), ("reserved", C.c_int32 * 4)]


class Macro(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upot", "virial", "upot6lj", "upot_xpoles", "virial_raw", "my_rf")]


class Kinetic(C.Structure):
    _fields_ = [("summv2", C.c_double), ("sum_iw2", C.c_double), ("n_molecules", C.c_uint64), ("rot_dof", C.c_uint64)]


class StepRecord(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upot", "virial", "summv2", "sum_iw2", "temperature", "pressure", "beta_trans", "beta_rot")]


RECORD_FIELDS = [k for k, _ in StepRecord._fields_]


PEER_HANDLE_BYTES = 96      # LS1B200_PEER_HANDLE_BYTES


class Decomp(C.Structure):
    """ls1b200_decomp"""
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("grid", C.c_int32 * 3), ("coords", C.c_int32 * 3),
                ("global_min", C.c_double * 3), ("global_max", C.c_double * 3), ("nccl_unique_id", C.c_uint8 * 128),
                ("use_nccl", C.c_int32), ("reserved", C.c_int32 * 3)]


# ls1b200_particle (== ParticleDataFull): the record of the halo / leaving-molecule messages
PARTICLE = np.dtype([("id", "<u8"), ("cid", "<u4"), ("flags", "<u4"), ("r", "<f8", 3), ("v", "<f8", 3), ("q", "<f8", 4), ("L", "<f8", 3)])
assert PARTICLE.itemsize == 120


class FlopCount(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "mol_dist", "lj", "charge", "charge_dipole", "dipole", "charge_quadrupole", "dipole_quadrupole", "quadrupole",
        "lj_macro", "charge_macro", "charge_dipole_macro", "dipole_macro", "charge_quadrupole_macro",
        "dipole_quadrupole_macro", "quadrupole_macro", "flops")]


def lib():
    """Loads libls1b200.so. Raises if it has not been built (python -m ls1_mardyn_b200.build): there is no other path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise Ls1B200Error(f"{LIB_PATH} is missing: build it with `python -m ls1_mardyn_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.ls1b200_last_error.restype = C.c_char_p
        L.ls1b200_kernel_launch_count.restype = C.c_uint64
        L.ls1b200_num_molecules.restype = C.c_uint64
        L.ls1b200_num_molecules.argtypes = [C.c_void_p]
        L.ls1b200_stream.restype = C.c_void_p
        L.ls1b200_stream.argtypes = [C.c_void_p]
        L.ls1b200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.ls1b200_destroy.argtypes = [C.c_void_p]
        L.ls1b200_set_components.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ls1b200_upload.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6
        L.ls1b200_exchange_overlapped_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ls1b200_rmm_upload.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ls1b200_rmm_start.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.ls1b200_rmm_simulate.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.ls1b200_rmm_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ls1b200_rmm_state_bytes.argtypes = [C.c_void_p]; L.ls1b200_rmm_state_bytes.restype = C.c_uint64
        L.ls1b200_generate_cubic_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_int32]
        L.ls1b200_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.ls1b200_set_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.ls1b200_geometry.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ls1b200_update.argtypes = [C.c_void_p]
        L.ls1b200_traverse_cells.argtypes = [C.c_void_p, C.c_void_p]
        L.ls1b200_leapfrog_pre.argtypes = [C.c_void_p, C.c_double]
        L.ls1b200_leapfrog_post.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.ls1b200_scale_velocities.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ls1b200_selective_thermostat.argtypes = [C.c_void_p, C.c_double]
        L.ls1b200_simulate.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_uint64, C.c_double, C.c_void_p]
        L.ls1b200_phase_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ls1b200_count_flops.argtypes = [C.c_void_p, C.c_void_p]
        L.ls1b200_measure_fp32_peak.argtypes = [C.c_int, C.c_void_p]
        L.ls1b200_measure_fp64_peak.argtypes = [C.c_int, C.c_void_p]
        L.ls1b200_measure_hbm_copy.argtypes = [C.c_int, C.c_void_p]
        L.ls1b200_nccl_unique_id.argtypes = [C.c_void_p]
        L.ls1b200_decomp_attach.argtypes = [C.c_void_p, C.c_void_p]
        L.ls1b200_exchange.argtypes = [C.c_void_p]
        L.ls1b200_download_begin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.ls1b200_download_end.argtypes = [C.c_void_p]
        L.ls1b200_peer_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.ls1b200_peer_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.ls1b200_peer_detach.argtypes = [C.c_void_p]
        L.ls1b200_allreduce_macro.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ls1b200_halo_export_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ls1b200_halo_import_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ls1b200_halo_import.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.ls1b200_halo_export.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise Ls1B200Error(lib().ls1b200_last_error().decode())


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def kernel_launch_count() -> int:
    return int(lib().ls1b200_kernel_launch_count())


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId (rank 0 calls it and broadcasts the 128 bytes)."""
    b = (C.c_uint8 * 128)(); _check(lib().ls1b200_nccl_unique_id(b)); return bytes(b)


def measure_fp32_peak(device: int = 0) -> float:
    v = C.c_double(); _check(lib().ls1b200_measure_fp32_peak(device, C.byref(v))); return v.value


def measure_fp64_peak(device: int = 0) -> float:
    v = C.c_double(); _check(lib().ls1b200_measure_fp64_peak(device, C.byref(v))); return v.value


def measure_hbm_copy(device: int = 0) -> float:
    v = C.c_double(); _check(lib().ls1b200_measure_hbm_copy(device, C.byref(v))); return v.value


class Context:
    """One LinkedCells + VectorizedCellProcessor + Leapfrog on one GPU (ls1b200_create)."""

    def __init__(self, box_min, box_max, cutoff, cutoff_lj=None, epsilon_rf=1e10, precision="mixed", geometry="xml",
                 periodic=(1, 1, 1), device=0):
        cfg = Config()
        for d in range(3):
            cfg.box_min[d] = box_min[d]; cfg.box_max[d] = box_max[d]; cfg.periodic[d] = int(periodic[d])
        cfg.cutoff = cutoff; cfg.cutoff_lj = cutoff if cutoff_lj is None else cutoff_lj; cfg.epsilon_rf = epsilon_rf
        cfg.precision = {"mixed": 0, "double": 1}[precision]; cfg.geometry = {"xml": 0, "ctor": 1}[geometry]
        cfg.device = device
        self.cfg = cfg
        self.h = C.c_void_p()
        self.n = 0
        self.decomposed = False
        _check(lib().ls1b200_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().ls1b200_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_components(self, ens: Ensemble, tables=None):
        cc, cs, ns = ens.pack()
        e24, s2, sh6 = tables if tables is not None else ens.lj_tables(self.cfg.cutoff_lj)
        e24 = np.ascontiguousarray(e24, dtype=np.float64); s2 = np.ascontiguousarray(s2, dtype=np.float64); sh6 = np.ascontiguousarray(sh6, dtype=np.float64)
        _check(lib().ls1b200_set_components(self.h, len(ens.components), C.cast(cc, C.c_void_p), ns, C.cast(cs, C.c_void_p), e24.shape[0], _vp(e24), _vp(s2), _vp(sh6)))

    def upload(self, id, cid, r, v, q=None, L=None):
        n = len(id)
        id = np.ascontiguousarray(id, dtype=np.uint64); cid = np.ascontiguousarray(cid, dtype=np.uint32)
        r = np.ascontiguousarray(r, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
        q = None if q is None else np.ascontiguousarray(q, dtype=np.float64)
        L = None if L is None else np.ascontiguousarray(L, dtype=np.float64)
        _check(lib().ls1b200_upload(self.h, n, _vp(id), _vp(cid), _vp(r), _vp(v), _vp(q), _vp(L)))
        self.n = n

    def exchange_overlapped_ms(self, reset=False):
        v = C.c_double(); _check(lib().ls1b200_exchange_overlapped_ms(self.h, C.byref(v), int(reset))); return v.value

    # ---- reduced-memory mode (rmm.cu; the reference's RMM build) ----
    def rmm_upload(self, id, r, v):
        n = len(id)
        id = np.ascontiguousarray(id, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
        _check(lib().ls1b200_rmm_upload(self.h, n, _vp(id), _vp(r), _vp(v)))
        self.n = n

    def rmm_start(self, dt, target_temperature, upot_corr=0.0, virial_corr=0.0):
        rec = StepRecord()
        _check(lib().ls1b200_rmm_start(self.h, dt, target_temperature, upot_corr, virial_corr, C.byref(rec)))
        return np.array([getattr(rec, k) for k, _ in StepRecord._fields_])

    def rmm_simulate(self, nsteps, dt, target_temperature, upot_corr=0.0, virial_corr=0.0, records=True):
        rec = (StepRecord * nsteps)() if records else None
        _check(lib().ls1b200_rmm_simulate(self.h, nsteps, dt, target_temperature, upot_corr, virial_corr, rec))
        if not records:
            return None
        return np.array([[getattr(x, k) for k, _ in StepRecord._fields_] for x in rec])

    def rmm_download(self):
        """-> (id, r, v) of the owned molecules, sorted by id."""
        id = np.zeros(self.n, dtype=np.uint64); r = np.zeros((self.n, 3)); v = np.zeros((self.n, 3))
        _check(lib().ls1b200_rmm_download(self.h, _vp(id), _vp(r), _vp(v), self.n))
        o = np.argsort(id)
        return id[o], r[o], v[o]

    def generate_cubic_grid(self, mols_per_dim, sim_box, temperature, seed=0, remove_momentum=True, reference_velocities=False):
        """ls1b200_generate_cubic_grid: the reference's CubicGridGenerator on the device (replaces upload). Returns the number of
        molecules this context generated (those of the global lattice that lie inside its box). reference_velocities: the
        reference's own per-cell std::mt19937(0) sequence (LS1B200_GEN_REFERENCE_VELOCITIES) instead of the counter-based one."""
        m = np.ascontiguousarray(np.broadcast_to(np.asarray(mols_per_dim, dtype=np.uint64), (3,)))
        b = np.ascontiguousarray(sim_box, dtype=np.float64)
        flags = (1 if remove_momentum else 0) | (2 if reference_velocities else 0)
        _check(lib().ls1b200_generate_cubic_grid(self.h, _vp(m), _vp(b), float(temperature), int(seed), flags))
        self.n = int(lib().ls1b200_num_molecules(self.h))
        return self.n

    # ---- multi-GPU block decomposition (ls1_mardyn_b200.decomposition.DomainDecomposition) ----
    def decomp_attach(self, dd, unique_id: bytes | None = None):
        """ls1b200_decomp_attach; collective when unique_id is given (NCCL transport inside the library)."""
        d = Decomp()
        d.rank = dd.rank; d.n_ranks = dd.n_ranks
        for k in range(3):
            d.grid[k] = int(dd.grid[k]); d.coords[k] = int(dd.coords[k]); d.global_min[k] = 0.0; d.global_max[k] = float(dd.global_length[k])
        d.use_nccl = 0
        if unique_id is not None:
            d.use_nccl = 1
            for k in range(128):
                d.nccl_unique_id[k] = unique_id[k]
        _check(lib().ls1b200_decomp_attach(self.h, C.byref(d)))
        self.decomposed = True

    def peer_handle(self) -> bytes:
        """ls1b200_peer_handle: the CUDA IPC handle of this rank's receive buffers (after upload)."""
        buf = (C.c_uint8 * PEER_HANDLE_BYTES)()
        _check(lib().ls1b200_peer_handle(self.h, buf))
        return bytes(buf)

    def peer_attach(self, handles) -> None:
        """ls1b200_peer_attach: `handles` = the peer_handle() of every rank, in rank order. From here on exchange / simulate
        store the molecule records straight into the neighbours' memory over NVLink instead of calling NCCL send / recv."""
        blob = b"".join(handles)
        assert len(blob) % PEER_HANDLE_BYTES == 0
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(lib().ls1b200_peer_attach(self.h, arr, len(blob) // PEER_HANDLE_BYTES))

    def peer_detach(self) -> None:
        """ls1b200_peer_detach; every rank calls it, then a barrier, before any context is closed."""
        _check(lib().ls1b200_peer_detach(self.h))

    def peer_detach_collective(self, dist) -> None:
        self.peer_detach()
        dist.barrier()

    def peer_attach_collective(self, dist, device=None) -> None:
        """Gathers the handles of all ranks over torch.distributed (any backend) and attaches them."""
        import torch
        mine = torch.frombuffer(bytearray(self.peer_handle()), dtype=torch.uint8)
        if device is not None:
            mine = mine.to(device)
        out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
        self.peer_attach([bytes(t.cpu().numpy().tobytes()) for t in out])

    def num_molecules(self) -> int:
        return int(lib().ls1b200_num_molecules(self.h))

    def exchange(self):
        _check(lib().ls1b200_exchange(self.h))

    def allreduce_macro(self):
        m = Macro(); k = Kinetic(); _check(lib().ls1b200_allreduce_macro(self.h, C.byref(m), C.byref(k))); return m, k

    def halo_export(self, dim, side, cap=None):
        """Records (numpy, dtype PARTICLE) for the neighbour on `side` of dimension `dim`."""
        cap = int(cap or max(4096, self.n)) + 1
        buf = np.zeros(cap, dtype=PARTICLE); cnt = C.c_uint64()
        _check(lib().ls1b200_halo_export_host(self.h, dim, side, _vp(buf), cap, C.byref(cnt)))
        return buf[:cnt.value].copy()

    def halo_import(self, records):
        records = np.ascontiguousarray(records, dtype=PARTICLE)
        _check(lib().ls1b200_halo_import_host(self.h, _vp(records), len(records)))

    def stream(self) -> int:
        """cudaStream_t of this context as an integer (torch.cuda.ExternalStream(ctx.stream()))."""
        return int(lib().ls1b200_stream(self.h))

    def geometry(self):
        cpd = (C.c_int32 * 3)(); cl = (C.c_double * 3)(); hm = (C.c_double * 3)()
        _check(lib().ls1b200_geometry(self.h, cpd, cl, hm))
        return np.array(cpd[:]), np.array(cl[:]), np.array(hm[:])

    def update(self):
        _check(lib().ls1b200_update(self.h))
        if self.decomposed:
            self.n = self.num_molecules()

    def traverse_cells(self):
        m = Macro(); _check(lib().ls1b200_traverse_cells(self.h, C.byref(m))); return m

    def leapfrog_pre(self, dt):
        _check(lib().ls1b200_leapfrog_pre(self.h, dt))

    def leapfrog_post(self, dt):
        k = Kinetic(); _check(lib().ls1b200_leapfrog_post(self.h, dt, C.byref(k))); return k

    def scale_velocities(self, bt, br):
        _check(lib().ls1b200_scale_velocities(self.h, bt, br))

    def selective_thermostat(self, target_temperature):
        """ls1b200_selective_thermostat: the per-molecule energy limits of Domain's explosion heuristics (Domain.cpp:266-296)."""
        _check(lib().ls1b200_selective_thermostat(self.h, float(target_temperature)))

    def download(self, field, out=None):
        f = FIELDS[field] if isinstance(field, str) else field
        dt, shp = _FIELD_SHAPE[f]
        if self.decomposed:
            self.n = self.num_molecules()
        if out is None:
            out = np.zeros((self.n,) + shp, dtype=dt)
        _check(lib().ls1b200_download(self.h, f, _vp(out), self.n))
        return out

    def download_begin(self, field, out):
        """ls1b200_download_begin: starts the copy of a field to `out` (keep it alive, ideally pinned) on a side stream; call
        download_end() before reading it. Overlaps with whatever is enqueued next (e.g. traverse_cells)."""
        f = FIELDS[field] if isinstance(field, str) else field
        assert out.dtype == _FIELD_SHAPE[f][0] and out.flags["C_CONTIGUOUS"] and len(out) == self.n
        self._dl_keep = out
        _check(lib().ls1b200_download_begin(self.h, f, _vp(out), self.n))

    def download_end(self):
        _check(lib().ls1b200_download_end(self.h))
        self._dl_keep = None

    def set_field(self, field, a):
        f = FIELDS[field] if isinstance(field, str) else field
        a = np.ascontiguousarray(a, dtype=_FIELD_SHAPE[f][0])
        _check(lib().ls1b200_set_field(self.h, f, _vp(a), len(a)))

    def simulate(self, nsteps, dt, target_temperature, global_n=0, global_volume=0.0, records=True):
        rec = (StepRecord * nsteps)() if records else None
        _check(lib().ls1b200_simulate(self.h, nsteps, dt, target_temperature, global_n, global_volume, rec))
        if self.decomposed:
            self.n = self.num_molecules()
        if not records:
            return None
        return np.frombuffer(rec, dtype=np.float64).reshape(nsteps, len(RECORD_FIELDS)).copy()

    def phase_times(self, reset=False):
        ms = (C.c_double * 4)(); ln = (C.c_uint64 * 4)()
        _check(lib().ls1b200_phase_times(self.h, ms, ln, int(reset)))
        return dict(zip(("update", "force", "leapfrog_pre", "leapfrog_post"), ms[:])), dict(zip(("update", "force", "leapfrog_pre", "leapfrog_post"), ln[:]))

    def count_flops(self):
        f = FlopCount(); _check(lib().ls1b200_count_flops(self.h, C.byref(f))); return f
