"""profiles/traffic_r02.json from the `ncu --set full` raw pages of scratch/profile.sh (DRAM bytes per launch / per filter step
for every shard size): python scratch/make_traffic.py"""
import csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
STEPS = 100                      # --window 100 in scratch/profile.sh


def raw(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    return rows[h], rows[h + 1], rows[h + 2]


out = {"source": "ncu --set full --clock-control none, psmf_stream_kernel<16,double>, one launch of %d filter steps per shard size "
                 "(scratch/profile.sh, round 2, NaN-encoded mask; the shard of N GPUs measured on one GPU: the kernel's traffic "
                 "depends on its rows only)" % STEPS, "shards": {}}
for rows in (1000000, 500000, 250000, 125024):
    p = os.path.join(ROOT, "gpurun_out", "r02_stream_full_%d_raw.csv" % rows)
    if not os.path.exists(p):
        continue
    h, u, v = raw(p)
    g = lambda m: (float(v[h.index(m)].replace(",", "")), u[h.index(m)])
    rd, ru = g("dram__bytes_read.sum"); wr, wu = g("dram__bytes_write.sum")
    rd *= UNIT[ru]; wr *= UNIT[wu]
    resident = rows * 128 <= 147 * 112 * 1024
    out["shards"]["%d:16:f64" % rows] = dict(
        rows=rows, filter_steps_per_launch=STEPS, dram_bytes_read_per_launch=rd, dram_bytes_write_per_launch=wr,
        dram_bytes_per_filter_step=(rd + wr) / STEPS, gpu_time_ms="%s %s" % g("gpu__time_duration.sum"),
        algorithmic_hbm_bytes_per_filter_step=rows * 8 if resident else 2 * rows * 128 + rows * 8,
        l2_hit_rate="%s %s" % g("lts__t_sector_hit_rate.pct"))
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
