"""profiles/traffic.json from `ncu --set full` captures of ONE launch each of the SW extension and the
seed-search kernel inside a bench.py run (config 3, chunk 0 = a full 120 MiB chunk):

  python tools/make_traffic.py <sw_raw.csv> <search_raw.csv> <bench.json of the same command> <label>

bench.py scales the per-candidate / per-position DRAM bytes to the launches it times and names the
source file in `roofline.traffic_source`."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))

    def val(name):
        v, u = float(d[name].replace(",", "")), units[hdr.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), d["Kernel Name"], float(d["gpu__time_duration.sum"])


sw_csv, se_csv, bench_json, label = sys.argv[1:5]
b = json.loads([ln for ln in open(bench_json).read().split("\n") if ln.startswith("{")][-1])
cfg = b["config"]
share = (120 << 20) / cfg["db_bytes"]          # chunk 0 of 8: a full 120 MiB chunk
sw_bytes, sw_name, sw_ms = dram_bytes(sw_csv)
se_bytes, se_name, se_ms = dram_bytes(se_csv)
out = {
    "config3:sw_extend": {"dram_bytes_per_candidate": sw_bytes / (b["candidates_per_step"] * share),
                          "dram_bytes_per_launch": sw_bytes, "kernel": sw_name,
                          "source": f"profiles/{label}_sw_extend_ncu.md (ncu --set full, one launch, chunk 0)"},
    "config3:seed_search": {"dram_bytes_per_position": se_bytes / (b["seed_positions_per_step"] * share),
                            "dram_bytes_per_launch": se_bytes, "kernel": se_name,
                            "source": f"profiles/{label}_search_tile_bench_ncu.md (ncu --set full, one launch, chunk 0)"},
}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
