import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden, rel_l2_mag
from oracle.make_golden import NAMES, TANGO_CASES, case_inputs
from oracle import tango_f64
from disco_b200.tango import offline_tango
for name in sorted(TANGO_CASES):
    seed, chans, length, vads, mfz, keep = TANGO_CASES[name]
    if "ivad" in name: continue
    g = load_golden(name)
    y, s, n = case_inputs(seed, chans, length, vads)
    res = offline_tango(y, s, n, list(vads), [None, None], mfz)
    line = []
    for nm, val in zip(NAMES, res):
        for k in range(len(chans)):
            key = "%s_%d" % (nm, k)
            if key in g and not nm.startswith("mask"):
                line.append("%s=%.1e" % (key, rel_l2_mag(val[k], g[key])))
    print(name, " ".join(line))
    if len(set(chans)) == 1 and mfz in ("local", "distant") and vads == ("irm1", "irm1"):
        ref = tango_f64.offline_tango(np.array(y), np.array(s), np.array(n), mask_for_z=mfz)
        print("    vs f64:", " ".join("yf_%d=%.1e z_%d=%.1e" % (k, rel_l2_mag(res[0][k], ref["yf"][k]), k,
              rel_l2_mag(res[3][k], ref["z_y"][k])) for k in range(len(chans))),
              "| golden vs f64:", " ".join("%.1e" % rel_l2_mag(g["yf_%d" % k], ref["yf"][k]) for k in range(len(chans))))
