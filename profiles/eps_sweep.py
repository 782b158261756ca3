#!/usr/bin/env python
"""profiles/eps_sweep.py [model] [ntraj] — splitting error of the windowed sSSA against the reference ensemble as a function of
rdme_epsilon (tau * max per-molecule jump rate): mean of every species total at the fixture's two taps vs the reference's, in
standard errors, and the KS p-value.  Run on the GPU box; reads only tests/golden fixtures."""
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_ens, load_model          # noqa: E402
from spatialpy_b200.engine import Engine       # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cdc42"
ntraj = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
fm, ens = load_model(name), load_ens(name)
steps = [int(s) for s in ens["steps"]]
for eps in (0.05, 0.025, 0.0125):
    tot = {s: [] for s in steps}
    with Engine(fm, rdme_epsilon=eps) as eng:
        for k in range(ntraj):
            eng.reset(int(ens["seed0"]) + k)
            done = 0
            for s in steps:
                eng.step(s - done)
                done = s
                tot[s].append(eng.get("xx").astype(np.int64).sum(axis=0))
        win = eng.counters()["windows"]
    for ti, s in enumerate(steps):
        g, r = np.array(tot[s]), ens[f"t{ti}_totals"]
        for j in range(g.shape[1]):
            if r[:, j].std() == 0 and g[:, j].std() == 0:
                continue
            se = np.sqrt(g[:, j].var() / len(g) + r[:, j].var() / len(r))
            print(f"eps {eps:7.4f} windows/traj {win:6d} step {s} species {j}: gpu mean {g[:, j].mean():9.3f} ref mean {r[:, j].mean():9.3f} "
                  f"diff {(g[:, j].mean() - r[:, j].mean()) / se:+6.2f} se  rel {(g[:, j].mean() / r[:, j].mean() - 1) * 100:+6.2f} %  KS p {stats.ks_2samp(g[:, j], r[:, j]).pvalue:.4f}", flush=True)
