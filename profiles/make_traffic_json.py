"""profiles/r02_traffic.json from the ncu launch lists (profiles/r02_launches_<workload>.csv): per kernel class the
DRAM bytes ncu saw (dram__bytes_read.sum + dram__bytes_write.sum) over the launches of the class in one forward +
one inverse, per launch and in total.  bench.py reads `dram_bytes_per_launch` of its dominant class as
roofline.traffic.      python profiles/make_traffic_json.py
"""
import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import summarize_ncu_launches as S  # noqa: E402

CLASS_OF = [
    (r"k_duval|k_chunkmin|k_chunk_threshold|k_sufmin|k_tile_min|k_prefix_min", "lyndon"),
    (r"k_flag_|k_set_u32|k_factor_lmax|k_coarse_index", "factor_table"),
    (r"k_byte_presence|k_code_table|k_init_keys", "init_keys"),
    (r"k_radix_hist|k_digit_hists", "radix_hist"),
    (r"k_onesweep_pass<u64", "onesweep_pass"),
    (r"k_onesweep_pass<u32.*, 2>", "emit"),
    (r"k_onesweep_pass<u32", "rerank"),
    (r"k_build_keys", "build_keys"),
    (r"k_rerank|k_bin_bases|k_bin_count|k_scatter_pairs", "rerank"),
    (r"k_emit|k_scatter_bytes|k_scatter_packed", "emit"),
    (r"k_local_sort|k_ls_probe", "local_sort"),
    (r"k_tuple|k_sample_lcp", "tuple_round"),
    (r"k_inv_tile_hist", "inv_tile_hist"), (r"k_inv_lf_rank", "inv_lf_rank"),
    (r"k_inv_walk_stage|k_inv_spl_count|k_inv_spl_write|k_inv_resolve|k_inv_find|k_inv_verify|k_inv_self_walk|k_inv_walk_mark|k_inv_walk$", "inv_walk"),
    (r"k_inv_min_jump|k_inv_sum_|k_inv_origin", "inv_jump"),
    (r"k_inv_colsum|k_inv_chunk_scan|k_inv_tile_base|k_tile_sum|k_tile_scan|k_scan_excl", "inv_scan"),
    (r"k_inv_spl_record|k_inv_place|k_inv_walk_tail|k_inv_walk_place", "inv_place"),
]


def main():
    out = {}
    for wl in ("C4", "C3", "C2"):
        path = HERE / f"r02_launches_{wl}.csv"
        if not path.exists():
            continue
        agg = {}
        for x in S.load(str(path)):
            name = S.short(x["name"])
            cls = next((c for pat, c in CLASS_OF if re.search(pat, name)), None)
            if cls is None:
                continue
            a = agg.setdefault(cls, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
            a["launches"] += 1
            a["dram_bytes"] += x.get("rd", 0) + x.get("wr", 0)
            a["ms"] += x.get("ms", 0)
        out[wl] = {c: {"launches": a["launches"], "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
                       "dram_bytes_total": a["dram_bytes"], "ncu_ms": a["ms"],
                       "dram_gbs": a["dram_bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0.0} for c, a in agg.items()}
    out["_how"] = ("per kernel class: dram__bytes_read.sum + dram__bytes_write.sum over the launches of the class in one forward + "
                   "one inverse (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                   "none; profiles/r02_launches_<workload>.csv, tests/gpu_batch.sh), divided by the class's launches; bench.py "
                   "reads dram_bytes_per_launch of its dominant class as roofline.traffic")
    (HERE / "r02_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    for wl in ("C4", "C3"):
        print(wl)
        for c, a in sorted(out.get(wl, {}).items(), key=lambda kv: -kv[1]["ncu_ms"]):
            print("  %-14s %3d launches %8.2f ms  %7.1f GB  %6.0f GB/s (%.0f %% of 6552.6)" % (
                c, a["launches"], a["ncu_ms"], a["dram_bytes_total"] / 1e9, a["dram_gbs"], 100 * a["dram_gbs"] / 6552.6))


if __name__ == "__main__":
    main()
