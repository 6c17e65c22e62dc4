"""The published per-repeat results of one full experiment of the reference, as a small fixture.

ExperimentImpute/output/LondonAir_PM25_30_{PSMF,rPSMF}.json (100 repeats, seed 123) -> tests/golden/published_pm25_30.json:
seed, percentage, parameters, the blake2b hashes of the inputs of every repeat and the published error_predict /
error_full / inside_sig / runtime lists.  The GPU test `test_published_experiment_replayed_on_the_gpu` regenerates the
inputs from the `Yorig` of tests/golden/impute_pm25_30.npz, filters all 100 repeats on the device and compares every
repeat with these numbers.  Run in the build container (needs /root/reference):

    python tests/golden/make_published_fixture.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ExperimentImpute/output"

out = dict(generated_by="tests/golden/make_published_fixture.py", source="ExperimentImpute/output/LondonAir_PM25_30_{PSMF,rPSMF}.json")
for method in ("PSMF", "rPSMF"):
    pub = json.load(open(os.path.join(REF, "LondonAir_PM25_30_%s.json" % method)))
    out[method] = {k: pub[k] for k in ("seed", "missing_percentage", "missing_ratio", "parameters", "hashes", "results", "hostname")}
with open(os.path.join(HERE, "published_pm25_30.json"), "w") as fp:
    json.dump(out, fp)
print("wrote published_pm25_30.json", os.path.getsize(os.path.join(HERE, "published_pm25_30.json")) // 1024, "KiB")
