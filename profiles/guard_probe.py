#!/usr/bin/env python
"""How accurate is the doubling method on the sources its conditioning guard flags?  9x9 holey pair (151 freqs x 8 kx): forced
doubling with the guard off against the eigen method, for the flagged sources and for the rest."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads as wk
from khepri_b200 import Engine
eng = Engine(workspace_cap_bytes=40 << 30)
freqs = np.linspace(0.49, 0.6, 151); kx = np.linspace(0, 0.3 * np.pi, 64)
wl = np.tile(1 / freqs, 8); kp = np.stack([np.repeat(kx[:8], 151), np.zeros(151 * 8)], 1); pol = np.tile([[1.0, 0.0]], (wl.size, 1))
for pp in (9, 5):
    st = wk.holey_pair(pp, 128)
    cl = wk.build_crystal(st, eng)
    plan = cl._get_plan(False)
    ref = eng.solve_batch(plan, wl, kp, pol, want_flux=True, method="eig")["RT"].cpu().numpy()
    os.environ["KH_DBL_COND_LIMIT"] = "1e300"
    dbl = eng.solve_batch(plan, wl, kp, pol, want_flux=True, method="doubling")["RT"].cpu().numpy()
    os.environ.pop("KH_DBL_COND_LIMIT")
    res = eng.solve_batch(plan, wl, kp, pol, want_flux=True, method="doubling")
    flagged = ((res["info"] & 8) != 0).cpu().numpy()
    err = np.abs(dbl - ref).max(1)
    print(json.dumps({"basis": pp, "sources": int(wl.size), "flagged": int(flagged.sum()), "err_flagged": [float(e) for e in err[flagged]],
                      "max_err_unflagged": float(err[~flagged].max()), "limits": {lim: None for lim in ()}}))
    for lim in ("1e6", "1e7"):
        os.environ["KH_DBL_COND_LIMIT"] = lim
        r2 = eng.solve_batch(plan, wl, kp, pol, want_flux=True, method="doubling")
        f2 = ((r2["info"] & 8) != 0).cpu().numpy()
        print(json.dumps({"basis": pp, "limit": lim, "flagged": int(f2.sum()), "max_err_unflagged": float(err[~f2].max())}))
    os.environ.pop("KH_DBL_COND_LIMIT")
