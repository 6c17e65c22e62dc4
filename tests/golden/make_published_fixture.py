"""The published per-repeat results of the reference's imputation experiments, as fixtures.

ExperimentImpute/output/ holds the PSMF / rPSMF result files of the paper's imputation table; the input CSVs of 18 of them
ship with the reference (LondonAir_PM25, LondonAir_PM10, sp500_closing_prices x {20, 30, 40} % x {PSMF, rPSMF}; 100 repeats
each, seed 123).  This script writes

  tests/golden/published_results.json            per file: seed, percentage, parameters, the blake2b hashes of the inputs
                                                 of every repeat, the published error_predict / error_full / inside_sig /
                                                 runtime lists
  tests/golden/dataset_<name>.npz                the (d, n) input matrices of the two data sets that are not in
                                                 impute_pm25_30.npz already (NaN = originally missing)

`test_published_experiment_replayed_on_the_gpu` replays the PM2.5 / 30 % files inside `pytest -m gpu`;
`scratch/replay_published.py` replays all 18 (1800 fits) on a GPU box and writes profiles/r02_replay_published.log.
Run in the build container (needs /root/reference):

    python tests/golden/make_published_fixture.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ExperimentImpute"

out = dict(generated_by="tests/golden/make_published_fixture.py", source="ExperimentImpute/output/<data set>_<pct>_<method>.json", files={})
for ds in ("LondonAir_PM25", "LondonAir_PM10", "sp500_closing_prices"):
    if ds != "LondonAir_PM25":
        Y = np.genfromtxt(os.path.join(REF, "data", ds + ".csv"), delimiter=",")
        np.savez_compressed(os.path.join(HERE, "dataset_%s.npz" % ds), Yorig=Y)
    for pct in (20, 30, 40):
        for method in ("PSMF", "rPSMF"):
            name = "%s_%d_%s" % (ds, pct, method)
            pub = json.load(open(os.path.join(REF, "output", name + ".json")))
            rec = {k: pub[k] for k in ("seed", "missing_percentage", "parameters", "hashes", "results", "hostname")}
            rec["dataset"] = ds
            rec["method"] = method
            out["files"][name] = rec
with open(os.path.join(HERE, "published_results.json"), "w") as fp:
    json.dump(out, fp)
for f in sorted(os.listdir(HERE)):
    if f.startswith(("published_results", "dataset_")):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
