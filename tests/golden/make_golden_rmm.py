#!/usr/bin/env python
"""Generates tests/golden/rmm_1clj_2000.npz: the per-step ResultWriter output of the reference built in its reduced-memory mode
(oracle/build_ref.sh rmm -> oracle/_ref/rmm/MarDyn: all FP32, VCP1CLJRMM + LeapfrogRMM) for a 2000-molecule 1CLJ liquid read from an
ASCII phase-space file, together with that initial state. Build container only (needs /root/reference); the .npz is committed.

    python tests/golden/make_golden_rmm.py
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ls1_mardyn_b200 import generators  # noqa: E402
from ls1_mardyn_b200.ascii_reader import write_phase_space  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "rmm", "MarDyn")
RC, DT, T, STEPS = 2.5, 0.002, 0.85, 40

XML = """<?xml version='1.0' encoding='UTF-8'?>
<mardyn version="20100525">
 <refunits type="SI"><length unit="nm">0.1</length><mass unit="u">1</mass><energy unit="K">1</energy></refunits>
 <simulation type="MD">
  <integrator type="LeapfrogRMM"><timestep unit="reduced">{dt}</timestep></integrator>
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
   <phasespacepoint><file type="ASCII">input.inp</file></phasespacepoint>
  </ensemble>
  <algorithm>
   <parallelisation type="DomainDecomposition"></parallelisation>
   <datastructure type="LinkedCells"><cellsInCutoffRadius>1</cellsInCutoffRadius></datastructure>
   <cutoffs type="CenterOfMass"><radiusLJ unit="reduced">{rc}</radiusLJ></cutoffs>
   <electrostatic type="ReactionField"><epsilon>1.0e+10</epsilon></electrostatic>
  </algorithm>
  <output><outputplugin name="ResultWriter"><writefrequency>1</writefrequency><outputprefix>res</outputprefix><writeprecision>12</writeprecision></outputplugin></output>
 </simulation>
</mardyn>
"""


def main():
    if not os.path.exists(EXE):
        sys.exit("make_golden_rmm: %s not built (oracle/build_ref.sh rmm)" % EXE)
    ps = generators.lj_lattice(10, seed=21, jitter=0.08)          # 2000 molecules, rho* = 0.7, T* = 0.85
    ps.r = ps.r.astype(np.float32).astype(np.float64)            # what MoleculeRMM will hold
    ps.v = ps.v.astype(np.float32).astype(np.float64)
    with tempfile.TemporaryDirectory() as td:
        write_phase_space(os.path.join(td, "input.inp"), ps)
        open(os.path.join(td, "config.xml"), "w").write(XML.format(dt=DT, steps=STEPS, T=T, L=repr(float(ps.length[0])), rc=RC))
        p = subprocess.run([EXE, "--steps", str(STEPS), "--final-checkpoint=0", "config.xml"], cwd=td, capture_output=True, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=600)
        if p.returncode != 0:
            sys.exit((p.stdout + p.stderr)[-3000:])
        rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
    trace = np.array([[float(x) for x in r] for r in rows])
    # columns: step time U_pot U_pot_avg U_kin U_kin_avg p p_avg beta_trans beta_rot c_v N
    out = os.path.join(ROOT, "tests", "golden", "rmm_1clj_2000.npz")
    np.savez_compressed(out, id=ps.id, r=ps.r, v=ps.v, length=ps.length, cutoff=RC, dt=DT, temperature=T, steps=STEPS,
                        upot=trace[:, 2], ukin=trace[:, 4], p=trace[:, 6], beta_trans=trace[:, 8], n=trace[:, 11])
    print("wrote", out, trace.shape)


if __name__ == "__main__":
    main()
