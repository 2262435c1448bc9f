#!/usr/bin/env python
"""bench.py -- MMUPS (million molecule updates per second) of the short-range force path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n-per-dim n]

Workload at N = 1 (BASELINE.json configs[1]): 1CLJ Lennard-Jones liquid, 16 M molecules (BCC lattice 2*200^3, rho* = 0.7,
T* = 0.85, rc = 2.5 sigma, dt = 0.002), mixed precision, NVT velocity scaling. A "step" is one full time step of the hot
path: leapfrog_pre -> update (bin + halo + SoA) -> traverse_cells (forces) -> leapfrog_post -> global values.

  value     N*K / t: state resident in HBM, t = wall clock of ls1b200_simulate(K) between synchronisations (max over ranks),
            cross-checked with CUDA events on the library's stream;
  e2e       the same step driven phase by phase through the C ABI with HOST buffers: every step copies r, v from pinned host
            memory to the device and reads r, v and the step's scalars back (the host-coherent mode an ls1 plugin sees);
  roofline  dominant kernel = the LJ cell-pair force kernel, FP32-FMA bound (no tensor-core work exists on this path):
            achieved = FlopCounter flops of the state (reference's algorithmic count, Newton-3 counting) / avg launch time;
            peak = FP32 FMA rate measured in this run (MEASURED_PEAKS.json has no FP32 figure). The HBM-bound kernels
            (update, leapfrog) are reported under roofline.hbm against MEASURED_PEAKS.json's copy bandwidth;
  cpu_baseline  the reference (oracle/_ref MarDyn, OpenMP on all host cores) or the scalar port, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO, TEMP, RC, DT = 0.7, 0.85, 2.5, 0.002


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe). Every row carries nvidia-smi's own
    time stamp, so rows are attributed to the timed region by the clock of the sample, not by when the pipe delivered them."""

    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        import datetime
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = time.time()
            self.rows.append((ts, f[1:]))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the rows sampled inside [t_begin, t_end] (time.time() of the timed region's ends). If the region is shorter than
        the sampling period, the rows of the second before it (warm-up steps: the same load) are added and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.06)      # let the row of the last 20 ms arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t_begin is None or (t_begin - 0.01 <= t <= t_end + 0.03)]
        window = "timed region"
        if len(rows) < 2 and t_begin is not None:
            rows = [r for (t, r) in self.rows if t_begin - 1.0 <= t <= t_end + 0.25]; window = "timed region and the second before it (warm-up steps; the region is shorter than the sampling period)"
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window,
                "rows_total": len(self.rows)}


def make_system(n_per_dim, mixed=True, seed=0, wl=None):
    """BCC lattice like the reference's CubicGridGenerator (generators.lattice), positions rounded through the calc type
    as ParticleCellBase::initCubicGrid does (ParticleCellBase.cpp:97-99)."""
    from ls1_mardyn_b200 import generators
    if wl is not None:
        ps = generators.lattice(wl.ensemble, n_per_dim, wl.rho, wl.temperature, seed=seed, jitter=0.0, rc=wl.rc)
    else:
        ps = generators.lj_lattice(n_per_dim, rho=RHO, temperature=TEMP, rc=RC, seed=seed, jitter=0.0)
    if mixed:
        ps.r = ps.r.astype(np.float32).astype(np.float64)
    return ps


def grid_for(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference's own CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------------------------
XML = """<?xml version='1.0' encoding='UTF-8'?>
<mardyn version="20100525">
 <refunits type="SI"><length unit="nm">0.1</length><mass unit="u">1</mass><energy unit="K">1</energy></refunits>
 <simulation type="MD">
  <integrator type="Leapfrog"><timestep unit="reduced">{dt}</timestep></integrator>
  <run><currenttime>0</currenttime><production><steps>{steps}</steps></production></run>
  <ensemble type="NVT">
   <temperature unit="reduced">{T}</temperature>
   <domain type="box"><lx>{L}</lx><ly>{L}</ly><lz>{L}</lz></domain>
   <components>
    <moleculetype id="1" name="1CLJ">
     <site type="LJ126" id="1" name="LJfull"><coords><x>0.0</x><y>0.0</y><z>0.0</z></coords><mass>1.0</mass><sigma>1.0</sigma><epsilon>1.0</epsilon><shifted>false</shifted></site>
     <momentsofinertia rotaxes="xyz"><Ixx>0.0</Ixx><Iyy>0.0</Iyy><Izz>0.0</Izz></momentsofinertia>
    </moleculetype>
   </components>
   <phasespacepoint><generator name="CubicGridGenerator"><specification>density</specification><density>{rho}</density><binaryMixture>false</binaryMixture></generator></phasespacepoint>
  </ensemble>
  <algorithm>
   <parallelisation type="DomainDecomposition"></parallelisation>
   <datastructure type="LinkedCells"><cellsInCutoffRadius>1</cellsInCutoffRadius></datastructure>
   <cutoffs type="CenterOfMass"><radiusLJ unit="reduced">{rc}</radiusLJ></cutoffs>
   <electrostatic type="ReactionField"><epsilon>1.0e+10</epsilon></electrostatic>
  </algorithm>
  <output><outputplugin name="ResultWriter"><writefrequency>1</writefrequency><outputprefix>res</outputprefix><writeprecision>15</writeprecision></outputplugin></output>
 </simulation>
</mardyn>
"""


_REF_BUILD = None


def reference_build():
    """Which build of the unmodified reference the CPU legs time: where the host has AVX-512 and the AVX-512 build is in the snapshot
    (oracle/build_ref.sh mixed512, VCP_VEC_AVX512F) both builds run a short probe and the FASTER one is the baseline (on the Xeon of
    the build container the AVX2 build wins, 3.3 vs 3.1 MMUPS on 8 cores); otherwise the AVX2 build."""
    global _REF_BUILD
    if _REF_BUILD is not None:
        return _REF_BUILD
    _REF_BUILD = ("mixed", "AVX2+FMA")
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        flags = ""
    need = ("avx512f", "avx512cd", "avx512vl", "avx512bw", "avx512dq")
    if all(f in flags for f in need) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mixed512", "MarDyn")):
        a = run_reference_binary(40, 3, prec="mixed"); b = run_reference_binary(40, 3, prec="mixed512")
        if b is not None and (a is None or b[0] > a[0]):
            _REF_BUILD = ("mixed512", "AVX-512")
    return _REF_BUILD


LAST_REFERENCE_TRACE = None      # U_kin / N, U_pot / N, p after the last step of the most recent reference run (from its ResultWriter)


def run_reference_binary(n_per_dim, steps, threads=None, prec=None, wl=None):
    """Runs oracle/_ref/<prec>/MarDyn (the unmodified reference, built by oracle/build_ref.sh) on a lattice of 2*n^3 molecules of
    workload `wl` (default: the 1CLJ liquid), generated by the reference's own CubicGridGenerator.
    Returns (MMUPS from the reference's own main-loop timer, N, threads, main-loop seconds) or None if the binary is not there."""
    prec = prec or reference_build()[0]
    exe = os.path.join(ROOT, "oracle", "_ref", prec, "MarDyn")
    if not os.path.exists(exe):
        return None
    rho, temp, rc, dt = (RHO, TEMP, RC, DT) if wl is None else (wl.rho, wl.temperature, wl.rc, wl.dt)
    N = 2 * n_per_dim ** 3
    L = (N / rho) ** (1.0 / 3.0)
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        cfg = os.path.join(td, "config.xml")
        if wl is None or wl.name == "lj":
            text = XML.format(dt=dt, steps=steps, T=temp, L=repr(L), rho=repr(N / L ** 3 * (1 + 1e-9)), rc=rc)
        else:
            from ls1_mardyn_b200.xml_config import simulation_xml
            text = simulation_xml(wl.ensemble, [L, L, L], temp, dt, steps, rc, generator_density=N / L ** 3 * (1 + 1e-9), longrange="homogeneous")      # ResultWriter on: the trace
        if prec == "rmm":
            text = text.replace('integrator type="Leapfrog"', 'integrator type="LeapfrogRMM"')      # the reference's reduced-memory build
        open(cfg, "w").write(text)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        p = subprocess.run([exe, "--steps", str(steps), "--final-checkpoint=0", cfg], cwd=td, env=env, capture_output=True, text=True, timeout=3000)
        out = p.stdout + p.stderr
        global LAST_REFERENCE_TRACE
        LAST_REFERENCE_TRACE = None
        try:      # ResultWriter rows (1CLJ template only): step, time, U_pot, U_pot_avg, U_kin, U_kin_avg, p, ...
            rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
            last = [float(x) for x in rows[-1]]
            if int(last[0]) == steps:
                LAST_REFERENCE_TRACE = {"at_step": steps, "ukin_per_molecule": last[4] / last[11], "upot_per_molecule": last[2] / last[11], "p": last[6]}
        except (OSError, IndexError, ValueError):
            pass
    if p.returncode != 0:
        return None
    m = re.search(r"Computation in main loop took:\s*([0-9.eE+-]+)\s*sec", out)
    nm = re.search(r"[Nn]umber of (?:global )?molecules[^0-9]*([0-9]+)", out)
    if not m:
        return None
    t = float(m.group(1))
    n_run = int(nm.group(1)) if nm else N
    return n_run * steps / t / 1e6, n_run, threads, t


def run_port(n_per_dim, steps):
    """The scalar CPU port (oracle/liboracle.so), one core."""
    from oracle import oracle as orc
    ps = make_system(n_per_dim)
    s = orc.OracleSim([0, 0, 0], list(ps.length), RC, precision="mixed", geometry="xml", periodic=(1, 1, 1))
    s.set_components(ps.ensemble); s.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    s.update(); s.traverse_cells()
    t0 = time.perf_counter(); s.simulate(steps, DT, TEMP); t = time.perf_counter() - t0
    return len(ps.id) * steps / t / 1e6, len(ps.id), 1, t


# bounded samples of the CPU legs inside the default run: molecules-per-dimension and steps chosen for a few seconds of main loop on
# the GPU box's host cores (1CLJ: 4 M molecules x 5 steps = 20 M updates at ~12 MMUPS; the multi-site models are 10 - 40 x slower per update)
CPU_SAMPLE = {"lj": (126, 5), "water": (50, 25), "co2": (40, 30), "argon": (37, 150)}


def cpu_baseline(wl=None):
    name = "lj" if wl is None else wl.name
    sample_n, steps = CPU_SAMPLE[name]
    r = run_reference_binary(sample_n, steps, wl=wl)
    if r is not None:
        v, n, cores, t = r
        return {"value": round(v, 3), "unit": "MMUPS", "cores": cores, "kind": "reference",
                "sample": f"oracle/_ref/{reference_build()[0]}/MarDyn (unmodified reference, {reference_build()[1]}, OpenMP): {name} lattice N={n}, {steps} steps, main loop {t:.2f} s"}
    if name != "lj":
        return None
    v, n, cores, t = run_port(24, 2)
    return {"value": round(v, 3), "unit": "MMUPS", "cores": cores, "kind": "port",
            "sample": f"oracle/liboracle.so scalar port: 1CLJ lattice N={n}, 2 steps, {t:.2f} s"}


def workload_config(wl, n_per_dim, world):
    """The `config` object of the JSON line: the workload only (identical for our arm and the reference arm); how each arm executes
    it is reported under `execution`."""
    N = 2 * n_per_dim ** 3
    L1 = (N / wl.rho) ** (1.0 / 3.0)
    g = grid_for(world)
    return {"workload": wl.description, "molecules_per_gpu": N, "molecules": world * N, "box": [L1 * g[0], L1 * g[1], L1 * g[2]], "steps_size": "dt=%g" % wl.dt}


WL_LJ_DESCRIPTION = "BASELINE.json configs[1]: 1CLJ LJ liquid rho*=0.7 T*=0.85 rc=2.5 sigma dt=0.002, mixed precision, NVT velocity scaling"


def reference_arm(args):
    """--impl reference: the unmodified reference on the host cores, on THE benchmark configuration (16 M molecules at N = 1; for
    N > 1 the single-GPU size, MMUPS being size independent to first order -- 128 M molecules do not fit a few minutes of CPU time).
    Two processes of the reference binary: W warm-up steps (discarded), then exactly K steps timed by its own main-loop timer."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ls1_mardyn_b200 import workloads
    wl = workloads.get(args.workload)
    n = args.ref_n_per_dim or args.n_per_dim or wl.n_per_dim
    t0 = time.perf_counter()
    if args.warmup:
        run_reference_binary(n, args.warmup, wl=wl)
    r = run_reference_binary(n, args.steps, wl=wl)
    kind = "reference"
    if r is None:
        r = run_port(24, max(1, min(args.steps, 3))); kind = "port"
    v, nmol, cores, t = r
    cfg = workload_config(wl, args.n_per_dim or wl.n_per_dim, args.gpus)
    sample = "the whole workload" if nmol == cfg["molecules"] else "bounded sample: %d of the workload's %d molecules (MMUPS is size independent to first order)" % (nmol, cfg["molecules"])
    line = {"metric": "MMUPS", "value": round(v, 4), "unit": "MMUPS (1e6 molecule updates / s)", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * t / max(args.steps, 1), 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 pair arithmetic, f64 accumulation and state (MARDYN_SPDP)", "data": "synthetic (BCC lattice, CubicGridGenerator)", "impl": "reference",
            "config": cfg,
            "execution": {"molecules_run": nmol, "sample": sample, "host_threads": cores, "build": "oracle/_ref/%s/MarDyn (%s, OpenMP)" % reference_build()},
            "cpu_baseline": {"value": round(v, 4), "unit": "MMUPS", "cores": cores, "kind": kind,
                             "sample": f"oracle/_ref/{reference_build()[0]}/MarDyn ({reference_build()[1]}, OpenMP), N={nmol} ({sample}), {args.steps} steps, main loop {t:.2f} s, total wall {time.perf_counter() - t0:.1f} s"},
            "e2e": {"value": round(v, 4), "unit": "MMUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and LAST_REFERENCE_TRACE is not None:
        # the physical state after the K timed steps (ResultWriter, 15 digits): our arm generates the SAME initial state on the device
        # (ls1b200_generate_cubic_grid with the reference's velocity sequence) and reports the same three numbers at the same step
        line["trace"] = dict(LAST_REFERENCE_TRACE, molecules=nmol, note="ResultWriter of the timed run, started from CubicGridGenerator; compare with trace.at_reference_step of the other arm")
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the force kernel, from the `ncu --set full` captures under profiles/
# (keyed by workload, precision, molecules): reported, with its source, as roofline.traffic
FORCE_TRAFFIC = {("lj", "mixed", 16000000): (734425088 + 748010240, "profiles/r02w_force_full.md (ncu --set full of a launch inside the fused step at this size: dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: 256 MB of calc records + 384 MB of velocities and 64 MB of the lazy permutation for the kinetic-sum epilogue read, 768 MB of F, Vi written; the stand-alone traversal reads 280 MB, profiles/r02m_force_full.md)")}


REFERENCE_TRACES = {("lj", "mixed"): "reference_trace_lj16M.npz", ("lj", "double"): "reference_trace_lj16M_double.npz", ("argon", "mixed"): "reference_trace_argon_full.npz",
                    ("water", "mixed"): "reference_trace_water_full.npz", ("co2", "mixed"): "reference_trace_co2_full.npz"}


def vs_reference_fixture(wl_name, precision, n_molecules, series):
    """The step records of a run that started from the reference's CubicGridGenerator state (generated on the device) against the
    committed ResultWriter rows of the UNMODIFIED reference binary on the same full-size configuration
    (tests/golden/make_reference_traces_full_size.py): largest relative deviation over the steps both cover. None if no fixture fits."""
    f = REFERENCE_TRACES.get((wl_name, precision))
    fx = os.path.join(ROOT, "tests", "golden", f) if f else None
    if not fx or not os.path.exists(fx):
        return None
    g = np.load(fx)
    if int(g["molecules"]) != int(n_molecules):
        return None
    ref = g["trace"]; m = min(len(series), len(ref) - 1)
    if m < 1:
        return None
    b = ref[1:m + 1]
    return {"steps_compared": m, "fixture": "tests/golden/%s (unmodified reference binary, the same %d-molecule configuration)" % (f, n_molecules),
            "U_pot_rel_max": float(np.abs(series[:m, 0] - b[:, 2]).max() / np.abs(b[:, 2]).max()),
            "T_rel_max": float(np.abs(0.5 * (series[:m, 2] + series[:m, 3]) - b[:, 4]).max() / np.abs(b[:, 4]).max()),
            "p_rel_max": float(np.abs(series[:m, 5] - b[:, 6]).max() / np.abs(b[:, 6]).max()),
            "beta_trans_abs_max": float(np.abs(series[:m, 6] - b[:, 8]).max()), "beta_rot_abs_max": float(np.abs(series[:m, 7] - b[:, 9]).max()),
            "tolerance": "north_star: T and p within 1e-6 (mixed) / 1e-9 (double); U_pot and p of the rotating models' generated lattice are near-total cancellations (tests: 1e-5 of the terms)"}


def run_secondary(name, local, fp32_peak, capi, with_cpu=True, steps=10, warmup=3):
    """One of the other single-GPU configurations of BASELINE.json, device-resident loop only: MMUPS, force-kernel time and its
    fraction of the FP32 peak (FlopCounter flops of the state), CPU baseline on a bounded sample."""
    from ls1_mardyn_b200 import workloads
    wl = workloads.get(name)
    # the state the reference's CubicGridGenerator builds (what this workload's cpu_baseline and `--impl reference` start from),
    # generated on the device; for the rotating models that state is an "explosion" by Domain's definition -- both sides then run
    # through the same heuristics (Domain.cpp:259-314)
    n = wl.n_per_dim
    N = 2 * n ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    ctx = capi.Context([0, 0, 0], [L, L, L], wl.rc, precision="mixed", geometry="xml", periodic=(1, 1, 1), device=local)
    ctx.set_components(wl.ensemble)
    assert ctx.generate_cubic_grid(n, [L, L, L], wl.temperature, seed=0, remove_momentum=True, reference_velocities=True) == N
    ctx.update(); ctx.traverse_cells()
    fc0 = ctx.count_flops()
    rec_w = ctx.simulate(warmup, wl.dt, wl.temperature) if warmup else np.zeros((0, 8))
    ctx.phase_times(reset=True)
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = ctx.simulate(steps, wl.dt, wl.temperature)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    ph_ms, _ = ctx.phase_times(reset=True)
    fc1 = ctx.count_flops()
    force_ms = ph_ms["force"] / steps
    flops = 0.5 * (fc0.flops + fc1.flops)
    out = {"workload": wl.description, "molecules": N, "steps": steps, "warmup": warmup, "MMUPS": round(N * steps / t / 1e6, 2), "ms_per_step": round(1e3 * t / steps, 4),
           "phases_ms_per_step": {k: round(v / steps, 4) for k, v in ph_ms.items()},
           "kernel_ms": round(force_ms, 4), "flops_per_molecule": round(flops / N, 1), "achieved_tflops": round(flops / (force_ms * 1e-3) / 1e12, 3),
           "frac": round(flops / (force_ms * 1e-3) / 1e12 / fp32_peak, 4), "T_last": rec[-1, 4], "upot_last_per_molecule": rec[-1, 0] / N,
           "data": "the reference's CubicGridGenerator state, generated on the device",
           "vs_reference_fixture": vs_reference_fixture(wl.name, "mixed", N, np.concatenate([rec_w, rec]))}
    ctx.close()
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline(wl)
    return out


def run_adapter(n_per_dim, steps, host_sync_frequency=0, prec="mixed"):
    """The drop-in itself: adapter/_build/<prec>/MarDyn_b200 -- the UNMODIFIED ls1 Simulation loop, XML input, CubicGridGenerator, Domain,
    thermostat and long-range correction on the host, with <datastructure type="LinkedCellsB200"> + <integrator type="LeapfrogB200">
    driving libls1b200.so -- on the 1CLJ workload. MMUPS from ls1's own main-loop timer. host_sync_frequency = k refreshes the host
    Molecule objects (r, v, q, D, F, M, Vi: 7 downloads) every k steps for molecule-iterating plugins; 0 = only at the last step."""
    exe = os.path.join(ROOT, "adapter", "_build", prec, "MarDyn_b200")
    if not os.path.exists(exe):
        return None
    from ls1_mardyn_b200 import workloads
    from ls1_mardyn_b200.xml_config import simulation_xml
    wl = workloads.get("lj")
    N = 2 * n_per_dim ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "config.xml"), "w").write(simulation_xml(
            wl.ensemble, [L, L, L], wl.temperature, wl.dt, steps, wl.rc, generator_density=N / L ** 3 * (1 + 1e-9), datastructure="LinkedCellsB200", integrator="LeapfrogB200",
            longrange="homogeneous", result_prefix=None, datastructure_extra="<hostSyncFrequency>%d</hostSyncFrequency><syncAtLastStep>false</syncAtLastStep>" % host_sync_frequency))
        t0 = time.perf_counter()
        p = subprocess.run([exe, "--steps", str(steps), "--final-checkpoint=0", "config.xml"], cwd=td, capture_output=True, text=True, timeout=3000,
                           env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
        wall = time.perf_counter() - t0
    out = p.stdout + p.stderr
    m = re.search(r"Computation in main loop took:\s*([0-9.eE+-]+)\s*sec", out)
    if p.returncode != 0 or not m:
        return {"error": out[-400:]}
    t = float(m.group(1))
    return {"MMUPS": round(N * steps / t / 1e6, 2), "molecules": N, "steps": steps, "main_loop_s": round(t, 3), "process_wall_s": round(wall, 1),
            "host_sync_frequency": host_sync_frequency,
            "path": "stock ls1 Simulation loop (XML, CubicGridGenerator, Domain, thermostat, LRC on the host) -> LinkedCellsB200 / LeapfrogB200 -> libls1b200.so"}


def run_rmm(local, fp32_peak, capi, with_cpu=True, steps=10, warmup=3, n=200):
    """configs[1] in the reference's reduced-memory mode (all FP32, 32 B per molecule, force fused with the kick; csrc/rmm.cu),
    device-resident loop; CPU baseline = the reference built in RMM configuration (oracle/_ref/rmm/MarDyn)."""
    import torch
    from ls1_mardyn_b200 import workloads
    wl = workloads.get("lj")
    ps = make_system(n, mixed=True, seed=0, wl=wl)
    ctx = capi.Context([0, 0, 0], list(ps.length), wl.rc, precision="mixed", geometry="xml", periodic=(1, 1, 1), device=local)
    ctx.set_components(ps.ensemble)
    ctx.rmm_upload(ps.id, ps.r, ps.v)
    N = len(ps.id)
    ctx.rmm_start(wl.dt, wl.temperature)
    ctx.rmm_simulate(warmup, wl.dt, wl.temperature, records=False)
    ctx.phase_times(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = ctx.rmm_simulate(steps, wl.dt, wl.temperature)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    ph_ms, _ = ctx.phase_times(reset=True)
    out = {"workload": wl.description + " -- reduced-memory mode (FP32 state, 32 B per molecule)", "molecules": N, "steps": steps, "warmup": warmup,
           "MMUPS": round(N * steps / t / 1e6, 2), "ms_per_step": round(1e3 * t / steps, 4), "phases_ms_per_step": {k: round(v / steps, 4) for k, v in ph_ms.items()},
           "state_bytes_per_molecule": 32, "T_last": rec[-1, 4], "upot_last_per_molecule": rec[-1, 0] / N}
    ctx.close()
    if with_cpu:
        r = run_reference_binary(CPU_SAMPLE["lj"][0], CPU_SAMPLE["lj"][1], prec="rmm")
        if r is not None:
            v, nn, cores, tt = r
            out["cpu_baseline"] = {"value": round(v, 3), "unit": "MMUPS", "cores": cores, "kind": "reference",
                                   "sample": f"oracle/_ref/rmm/MarDyn (the reference built with ENABLE_REDUCED_MEMORY_MODE, PRECISION=SINGLE, AVX2, OpenMP): lj lattice N={nn}, {CPU_SAMPLE['lj'][1]} steps, main loop {tt:.2f} s"}
    return out


def our_arm(args):
    # exactly ONE line on stdout: libraries that chat on fd 1 (NCCL prints its version there) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from ls1_mardyn_b200 import capi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        assert world == 1 and args.gpus == 1, "for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N"
    torch.cuda.set_device(local)
    sampler = ClockSampler(local); sampler.start()      # started first: nvidia-smi takes up to a second to deliver its first row; rows are time-stamped and filtered to the timed region
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # weak scaling: the box grows with the number of GPUs; every rank owns one block of the 3D decomposition
    # with the same 2 n^3 molecules, generates only its own part of the lattice, and exchanges leaving molecules and halo
    # copies with its neighbours over NCCL inside the library every step
    from ls1_mardyn_b200 import generators, workloads
    from ls1_mardyn_b200.decomposition import DomainDecomposition
    wl = workloads.get(args.workload)
    RHO, TEMP, RC, DT = wl.rho, wl.temperature, wl.rc, wl.dt
    n = args.n_per_dim or wl.n_per_dim
    rot = any(c.Ipa()[1] for c in wl.ensemble.components)
    device_gen = (not rot or world == 1) and not args.host_gen      # generated ON the device (ls1b200_generate_cubic_grid): no host array of the system exists (rotating models on several GPUs: host blocks)
    grid = grid_for(world)
    L1 = (2 * n ** 3 / RHO) ** (1.0 / 3.0)             # box of one block
    box = np.array([L1 * grid[0], L1 * grid[1], L1 * grid[2]])
    if world == 1:
        dd = None
        ctx = capi.Context([0, 0, 0], list(box), RC, precision=args.precision, geometry="xml", periodic=(1, 1, 1), device=local)
        ctx.set_components(wl.ensemble)
    else:
        dd = DomainDecomposition(rank, world, box, grid)
        ctx = capi.Context(list(dd.box_min()), list(dd.box_max()), RC, precision=args.precision, geometry="xml", periodic=dd.periodic(), device=local)
        ctx.set_components(wl.ensemble)
        uid = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=torch.device("cuda", local))
        ctx.decomp_attach(dd, uid[0])
    if device_gen:
        # the reference's own generator state: per-cell std::mt19937(0) velocity sequence, momentum removed, rescaled to T -- at N = 1 this
        # is bit for bit (velocities to 1e-13) what `--impl reference` starts from (tests: test_device_generator_reference_velocities_*)
        N = ctx.generate_cubic_grid([n * grid[0], n * grid[1], n * grid[2]], box, TEMP, seed=0, remove_momentum=True, reference_velocities=True)
    else:
        if world == 1:
            ps = make_system(n, mixed=(args.precision == "mixed"), seed=rank, wl=wl)
        else:
            ps = generators.lattice_block(wl.ensemble, n, RHO, TEMP, dd.coords, grid, seed=0)
            ps.r = ps.r.astype(np.float32).astype(np.float64)
        ctx.upload(ps.id, ps.cid, ps.r, ps.v, ps.q if rot else None, ps.L if rot else None)
        N = len(ps.id)
    N_global = world * N
    if world > 1 and device_gen:
        tn = torch.tensor([N], dtype=torch.int64, device="cuda"); dist.all_reduce(tn); N_global = int(tn.item())
    V_global = float(np.prod(box))
    transport = "none"
    if dd is not None:
        # molecule exchange: NVLink stores into the neighbours' receive buffers (CUDA IPC); LS1B200_TRANSPORT=nccl keeps ncclSend / Recv
        transport = "nccl"
        if os.environ.get("LS1B200_TRANSPORT", "peer") != "nccl":
            # every rank must end up on the same transport: if mapping a neighbour's buffers fails anywhere (no P2P / IPC between
            # two of the GPUs), all ranks detach and the exchange goes over NCCL
            ok = torch.ones(1, device=torch.device("cuda", local))
            try:
                ctx.peer_attach_collective(dist, device=torch.device("cuda", local))
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("rank %d: peer-memory transport unavailable (%s), falling back to NCCL send/recv\n" % (rank, e))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) > 0:
                transport = "peer-memory"
            else:
                ctx.peer_detach_collective(dist)
        ctx.exchange()
    ctx.update(); ctx.traverse_cells()
    fc0 = ctx.count_flops()

    fp32_peak = capi.measure_fp32_peak(local)
    fp64_peak = capi.measure_fp64_peak(local) if args.precision == "double" else None
    hbm_peak, hbm_src = peaks()

    # ---- HBM-resident: K steps inside the library, no host round trip ----
    rec_w = ctx.simulate(args.warmup, DT, TEMP, global_n=N_global, global_volume=V_global) if args.warmup else np.zeros((0, 8))
    ctx.phase_times(reset=True)
    launches0 = capi.kernel_launch_count()
    ext = torch.cuda.ExternalStream(ctx.stream(), device=local)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter(); w0 = time.time()
    e0.record(ext)
    rec = ctx.simulate(args.steps, DT, TEMP, global_n=N_global, global_volume=V_global)
    e1.record(ext)
    barrier()
    t_wall = time.perf_counter() - t0
    clocks = sampler.stop(w0, w0 + t_wall)
    t_dev = e0.elapsed_time(e1) * 1e-3
    launches = capi.kernel_launch_count() - launches0
    ph_ms, ph_launch = ctx.phase_times(reset=True)
    xchg_ms = ctx.exchange_overlapped_ms(reset=True) if dd is not None else 0.0
    fc1 = ctx.count_flops()
    t = torch.tensor([t_wall, t_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_wall, t_dev = float(t[0]), float(t[1])
    value = N_global * args.steps / t_wall / 1e6

    # ---- end to end through the C ABI with host buffers ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    cap_h = int(N * 1.1) + 4096      # decomposed: the number of molecules a rank owns changes from step to step
    r_h = torch.empty((cap_h, 3), dtype=torch.float64).pin_memory(); v_h = torch.empty((cap_h, 3), dtype=torch.float64).pin_memory()
    n_now = ctx.num_molecules() if dd is not None else N
    ctx.download("R", out=r_h.numpy()[:n_now]); ctx.download("V", out=v_h.numpy()[:n_now])
    h2d = d2h = 0
    ctx.set_field("R", r_h.numpy()[:n_now]); ctx.set_field("V", v_h.numpy()[:n_now])      # untimed: first touch of the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.set_field("R", r_h.numpy()[:n_now]); ctx.set_field("V", v_h.numpy()[:n_now])        # host -> device: this step's inputs
        h2d += 48 * n_now
        ctx.leapfrog_pre(DT)
        if dd is not None:
            ctx.exchange()
        ctx.update()
        if dd is None:
            ctx.download_begin("R", r_h.numpy()[:n_now])        # positions are final after the update: they travel while the forces run
        m = ctx.traverse_cells(); k = ctx.leapfrog_post(DT)
        if dd is not None:
            m, k = ctx.allreduce_macro()
            n_now = ctx.num_molecules()
            ctx.download("R", out=r_h.numpy()[:n_now])
        ctx.download("V", out=v_h.numpy()[:n_now])              # device -> host: the step's results
        if dd is None:
            ctx.download_end()
        d2h += 48 * n_now + 24
        _ = (m.upot, m.virial, k.summv2)
    barrier()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N_global * e2e_steps / float(t[0]) / 1e6

    if rank == 0:
        force_ms = ph_ms["force"] / max(args.steps, 1)
        flops = 0.5 * (fc0.flops + fc1.flops)          # algorithmic flops / force launch (state drifts slowly off the lattice)
        achieved = flops / (force_ms * 1e-3) / 1e12
        # algorithmic bytes of the HBM-bound phases per molecule, SURVEY.md 8(d): update = read r 24 + key 4 + permute r, v, id, cid
        # 2 x 60 = 148 B, + 16 B for the calc-precision record the force kernel reads; leapfrog pre 120 B, post 72 B (FP64 state)
        hbm = {}
        for name, bpm in (("update", 164.0), ("leapfrog_pre", 120.0), ("leapfrog_post", 72.0)):
            ms = ph_ms[name] / max(args.steps, 1)
            hbm[name] = {"ms": round(ms, 4), "algorithmic_bytes_per_molecule": bpm, "achieved_gbs": round(bpm * N / (ms * 1e-3) / 1e9, 1) if ms > 0 else None,
                         "frac": round(bpm * N / (ms * 1e-3) / 1e9 / hbm_peak, 4) if ms > 0 else None}
        # Fused step of ls1b200_simulate (ensembles without rotation, tiled 1CLJ kernel; LS1B200_FUSED=0 turns it off): the second half
        # kick and the binning run inside the first kernel of the next step, the kinetic sum inside the force kernel -- the three
        # phases no longer map onto SURVEY 8(d)'s three byte counts one by one, so the roofline is taken over their SUM:
        # 356 B of algorithmic traffic per molecule (164 + 120 + 72) over the time of everything in the step that is not the force kernel.
        fused = os.environ.get("LS1B200_FUSED", "1") != "0" and wl.name in ("lj", "argon") and not os.environ.get("LS1B200_LJ_KERNEL") and args.steps >= 2
        if fused:
            ms = sum(ph_ms[k] for k in ("update", "leapfrog_pre", "leapfrog_post")) / max(args.steps, 1)
            hbm = {"fused_step": True,
                   "integrate_bin_sort": {"ms": round(ms, 4), "algorithmic_bytes_per_molecule": 356.0, "achieved_gbs": round(356.0 * N / (ms * 1e-3) / 1e9, 1),
                                          "frac": round(356.0 * N / (ms * 1e-3) / 1e9 / hbm_peak, 4),
                                          "note": "update + leapfrog_pre + leapfrog_post of SURVEY 8(d) = 164 + 120 + 72 B; the fused kernels move less than that (the state is read and written once)"},
                   "phase_ms": {k: round(ph_ms[k] / max(args.steps, 1), 4) for k in ("update", "leapfrog_pre", "leapfrog_post")}}
        double = args.precision == "double"
        peak = fp64_peak if double else fp32_peak
        traffic, traffic_src = FORCE_TRAFFIC.get((wl.name, args.precision, N), (None, None))
        roofline = {"kernel": ("k_force_1clj_tiled<%s, true>" if wl.name in ("lj", "argon") else "k_force_multi_tiled<%s> (site-signature specialisation)") % ("double" if double else "float"),
                    "bound": "fp64" if double else "fp32", "achieved": round(achieved, 3), "peak": round(peak, 2), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at this size from an `ncu --set full` capture (not measured in this run)
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": ("FP64" if double else "FP32") + " FMA micro-benchmark in this run (ls1b200_measure_fp%s_peak); MEASURED_PEAKS.json holds no CUDA-core figure; nominal FP32 74.4" % ("64" if double else "32"),
                    "algorithmic_flops_per_launch": flops, "flops_per_molecule": round(flops / N, 1), "kernel_ms": round(force_ms, 4),
                    "pair_interactions_per_s": round(0.5 * (fc0.lj + fc1.lj) / (force_ms * 1e-3), 1),
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "hbm": hbm}
        line = {
            "metric": "MMUPS", "value": round(value, 2), "unit": "MMUPS (1e6 molecule updates / s)", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * t_wall / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 pair arithmetic, f64 accumulation and state (MARDYN_SPDP)" if args.precision == "mixed" else "f64 throughout (MARDYN_DPDP, the reference's double build)",
            "data": "synthetic (BCC lattice like CubicGridGenerator, equal-magnitude velocities" + (", generated on the device by ls1b200_generate_cubic_grid)" if device_gen else ", host arrays uploaded)"),
            "config": workload_config(wl, n, world),
            "execution": {"cells_per_dim": [int(x) for x in ctx.geometry()[0]],
                          "parallelism": "1 GPU" if world == 1 else ("3D block decomposition %dx%dx%d, one process per GPU, leaving molecules + halo copies every step " % grid_for(world)) + ("stored by the packing kernel straight into the neighbours' memory over NVLink (CUDA IPC)" if transport == "peer-memory" else "over NCCL send/recv") + ", NCCL allreduce of 7 doubles",
                          "l2_policy": "inputs larger than L2: state %.1f GB per GPU" % (N * 150 / 1e9)},
            "ms_per_step_device_events": round(1e3 * t_dev / args.steps, 4),
            "phases_ms_per_step": {k: round(v / args.steps, 4) for k, v in ph_ms.items()},
            "exchange_overlapped_ms_per_step": round(xchg_ms / args.steps, 4),      # multi-GPU: the exchange on its own stream beside the inner slots' integration / binning (not part of the phases)
            "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 2), "unit": "MMUPS", "h2d_bytes_per_step": int(h2d / e2e_steps), "d2h_bytes_per_step": int(d2h / e2e_steps), "steps": e2e_steps,
                    "path": "set_field(R,V) from pinned host -> leapfrog_pre -> update -> [ls1b200_download_begin(R): the positions travel while the forces run] traverse_cells -> leapfrog_post -> download(V), download_end + scalars"},
            "roofline": roofline,
            "clocks": clocks,
            "trace": {"T_first": rec[0, 4], "T_last": rec[-1, 4], "upot_last_per_molecule": rec[-1, 0] / N_global, "p_last": rec[-1, 5]},
        }
        if device_gen and world == 1:
            # `--impl reference` times K steps from the same generated state: its ResultWriter row K is this run's step K (warm-up included)
            series = np.concatenate([rec_w, rec]); ks = args.steps
            line["trace"]["at_reference_step"] = {"at_step": ks, "ukin_per_molecule": 0.5 * (series[ks - 1, 2] + series[ks - 1, 3]) / N_global, "upot_per_molecule": series[ks - 1, 0] / N_global, "p": series[ks - 1, 5],
                                                  "note": "same initial state as the reference arm (its CubicGridGenerator sequence, generated on the device): compare with that line's trace"}
            # ... and against the committed ResultWriter rows of the unmodified reference on this very configuration (full size;
            # tests/golden/make_reference_traces_full_size.py): largest relative deviation over the steps both cover
            cmpf = vs_reference_fixture(wl.name, args.precision, N_global, series)
            if cmpf is not None:
                line["trace"]["vs_reference_fixture"] = cmpf
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(wl)
            if cb is not None:
                line["cpu_baseline"] = cb
        if world == 1 and wl.name == "lj" and args.precision == "mixed" and not args.no_secondary and not args.n_per_dim:
            # BASELINE.json configs[0], [2], [3] at their single-GPU sizes, measured in the same run (device-resident loop only)
            ctx.close()
            line["secondary"] = {k: run_secondary(k, local, fp32_peak, capi, with_cpu=not args.no_cpu_baseline) for k in ("argon", "water", "co2")}
            line["secondary"]["lj_rmm"] = run_rmm(local, fp32_peak, capi, with_cpu=not args.no_cpu_baseline)
            # the drop-in binary on the headline workload: device-resident through the unmodified ls1 loop, and with the host Molecule
            # objects refreshed every step (what a molecule-iterating plugin running every step would cost)
            line["secondary"]["adapter_MarDyn_b200"] = {"device_resident": run_adapter(wl.n_per_dim, args.steps, 0), "host_mirror_every_step": run_adapter(100, 5, 1)}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        # orderly teardown: nobody frees receive buffers a neighbour still has mapped
        torch.cuda.synchronize()
        if transport == "peer-memory":
            ctx.peer_detach_collective(dist)
        else:
            dist.barrier()
        ctx.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lj", choices=["lj", "water", "co2", "argon"], help="lj = BASELINE.json configs[1] (the headline); argon / water / co2 = configs[0] / [2] / [3]")
    ap.add_argument("--n-per-dim", type=int, default=0, help="BCC lattice cells per dimension: N = 2 n^3 (default: the workload's single-GPU size, lj 200 -> 16 M molecules)")
    ap.add_argument("--ref-n-per-dim", type=int, default=0, help="size of the reference arm's run (default: the workload's single-GPU size, 200 -> 16 M molecules)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the water-like 4 M and CO2 8 M runs reported under `secondary`")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--precision", default="mixed", choices=["mixed", "double"], help="mixed = the headline (BASELINE.json configs[1]); double = the full-FP64 variant of the reference's double build")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-gen", action="store_true", help="build the lattice in numpy and upload it instead of generating it on the device")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
