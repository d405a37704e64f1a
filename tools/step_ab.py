#!/usr/bin/env python
"""A/B timing of differently built libpvd_b200.so variants on the headline step (H2O, discrete, 1e6 walkers).
usage: step_ab.py lib_a.so lib_b.so ...   (each variant runs in its own process; PVD_B200_LIB selects the library)"""
import json
import os
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def one():
    import numpy as np
    sys.path.insert(0, ROOT)
    from pyvibdmc_b200 import _capi, kernels as K
    from pyvibdmc_b200.simulation_utilities import Constants
    eq = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
    n = int(os.environ.get("AB_WALKERS", "1000000"))
    rng = {"fast": _capi.RNG_FAST, "fp64": _capi.RNG_FP64}.get(os.environ.get("AB_RNG"), _capi.RNG_DEFAULT)
    mH, mO = Constants.mass("H"), Constants.mass("O")
    sim = K.DeviceSim(3, 3, [mH, mH, mO], n, 5.0, _capi.POT_H2O_PS, seed=7, rng_mode=rng)
    if os.environ.get("AB_MODE"):
        sim.set_resident(int(os.environ["AB_MODE"]))       # 0 self-compacting step, 2 resident kernel, 3 deferred-compaction steps
    sim.upload(np.broadcast_to(eq * 1.01, (n, 3, 3)).copy())
    sim.run(300)
    sim.sync()
    best = 1e9
    for _ in range(4):
        sim.run(200)
        sim.sync()
        best = min(best, sim.last_run_ms() / 200)
    st = sim.state()
    print(json.dumps({"lib": os.environ.get("PVD_B200_LIB", "default"), "ms_per_step": best, "n": st["n"], "vref": st["vref"]}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one()
    else:
        for lib in sys.argv[1:] or [""]:
            env = dict(os.environ)
            if lib:
                env["PVD_B200_LIB"] = os.path.abspath(lib)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-400:])
