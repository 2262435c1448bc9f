#!/usr/bin/env bash
# gpurun -- 'bash scripts/gpu_adapter_timers.sh <tag>': the drop-in binary on 2 M / 16 M molecules with ls1's own timer summary
set -u
TAG="${1:-adp}"; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python - "$OUT" <<'PY'
import os, sys, subprocess, tempfile, re
sys.path.insert(0, os.getcwd())
from ls1_mardyn_b200 import workloads
from ls1_mardyn_b200.xml_config import simulation_xml
out = sys.argv[1]
wl = workloads.get("lj")
for n, steps in ((100, 20), (200, 20)):
    N = 2 * n ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    td = tempfile.mkdtemp()
    open(td + "/config.xml", "w").write(simulation_xml(wl.ensemble, [L, L, L], wl.temperature, wl.dt, steps, wl.rc, generator_density=N / L ** 3 * (1 + 1e-9),
                                                       datastructure="LinkedCellsB200", integrator="LeapfrogB200", longrange="homogeneous", result_prefix=None))
    p = subprocess.run([os.path.join(os.getcwd(), "adapter/_build/mixed/MarDyn_b200"), "--steps", str(steps), "--final-checkpoint=0", "config.xml"], cwd=td, capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count())))
    open(os.path.join(out, "mardyn_b200_%d.log" % N), "w").write(p.stdout + p.stderr)
    print("N", N, "rc", p.returncode)
    for l in (p.stdout + p.stderr).splitlines():
        if "took" in l or "Timer" in l: print("   ", l[-150:])
PY
