#!/usr/bin/env python
"""Golden vectors for ChainRuleHelper: every method of the UNMODIFIED reference class
(/root/reference/pyvibdmc/simulation_utilities/imp_samp_helper.py:10-209) on random water-like and 5-atom geometries.
Build container only:  python tests/golden/make_chain_rule_golden.py"""
import os, sys, warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path[:0] = [os.path.join(ROOT, "oracle", "refstubs"), "/root/reference"]
warnings.filterwarnings("ignore")
from pyvibdmc.simulation_utilities.imp_samp_helper import ChainRuleHelper      # noqa: E402  (the reference)

rng = np.random.default_rng(77)
EQ = np.array([[1.81005599, 0., 0.], [-0.45344658, 1.75233806, 0.], [0., 0., 0.]])
out = {}
for tag, cds, pairs, ang in (("w", EQ[None] + rng.normal(0, 0.15, (64, 3, 3)), ([0, 2], [2, 1]), [0, 2, 1]),
                             ("p", rng.normal(0, 1.5, (48, 5, 3)), ([0, 3], [3, 4]), [0, 3, 4])):
    h = ChainRuleHelper(cds.copy(), np)
    dr = [h.dr_dx(p) for p in pairs]
    d2r = [h.d2r_dx2(p) for p in pairs]
    dc, d2c = h.dcth_dx(ang), h.d2cth_dx2(ang)
    dth, d2th = h.dth_dx(ang), h.d2th_dx2(ang)
    dpsi = rng.normal(0, 1, (3, len(cds)))
    d2psi = rng.normal(0, 1, (3, len(cds)))
    stack1 = np.stack([dr[0], dr[1], dth])
    stack2 = np.stack([d2r[0], d2r[1], d2th])
    out.update({f"{tag}_cds": cds, f"{tag}_dr0": dr[0], f"{tag}_dr1": dr[1], f"{tag}_d2r0": d2r[0], f"{tag}_d2r1": d2r[1],
                f"{tag}_dc": dc, f"{tag}_d2c": d2c, f"{tag}_dth": dth, f"{tag}_d2th": d2th, f"{tag}_dpsi": dpsi, f"{tag}_d2psi": d2psi,
                f"{tag}_jac": h.dpsidx(dpsi, stack1), f"{tag}_lap": h.d2psidx2(d2psi, stack2, dpsi, stack1),
                f"{tag}_pairs": np.array(pairs), f"{tag}_ang": np.array(ang)})
np.savez_compressed(os.path.join(HERE, "chain_rule_golden.npz"), **out)
print("wrote chain_rule_golden.npz", sorted(out)[:5], "...")
