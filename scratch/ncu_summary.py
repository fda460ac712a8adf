"""Summarise an `ncu --page raw --csv` dump: per-kernel table for profiles/ and the per-stage DRAM traffic json."""
import csv, json, sys
raw, out_csv, out_json, label = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
def col(name): return hdr.index(name)
def val(r, name):
    i = col(name)
    return float(r[i].replace(",", "")) * SC.get(units[i], 1.0)
names = ['k_list_merge_blocks', 'k_build_btab', 'k_pc_register', 'k_pc_walk', 'k_pc_scan', 'k_pc_apply', 'k_pc_undo', 'k_pc_free', 'k_clear_prev_blocks', 'k_merge_ogm', 'k_alloc_observed',
         'k_edt_ybits_clear', 'k_edt_ybits_blocks', 'k_edt_ybits', 'k_edt_ycols', 'k_edt_slices', 'k_edt_xsweep', 'k_edt_zsweep_banded', 'k_edt_zsweep', 'k_list_blocks',
         'k_mark_blocks', 'k_mark', 'k_frontiers', 'k_waves', 'k_commit', 'k_wave_stats']
per = {}
keep = ['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']
with open(out_csv, "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "time_us", "dram_read_bytes", "dram_write_bytes", "dram_pct_of_peak", "sm_pct_of_peak", "warps_active_pct", "regs"])
    for r in rows[2:]:
        full = r[col('Kernel Name')]
        short = next(n for n in names if n + "(" in full or n + "<" in full)
        t, rd, wr = val(r, 'gpu__time_duration.sum'), val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
        if short in per:   # a second launch of the same kernel in the capture window: keep the first
            continue
        per[short] = {"us": t, "dram_bytes": rd + wr, "dram_gbs": (rd + wr) / t / 1e3 if t > 0 else None}
        w.writerow([short, r[col('Grid Size')], r[col('Block Size')], f"{t:.2f}", int(rd), int(wr)] + [r[col(k)] for k in keep])
grp = {"batch_dt": ["k_edt_ybits_clear", "k_edt_ybits_blocks", "k_edt_ybits", "k_edt_ycols", "k_edt_slices", "k_edt_xsweep", "k_edt_zsweep_banded", "k_edt_zsweep"],
       "edt_pack": ["k_edt_ybits_clear", "k_edt_ybits_blocks", "k_edt_ybits", "k_edt_ycols", "k_edt_slices"], "edt_x": ["k_edt_xsweep"], "edt_z": ["k_edt_zsweep_banded", "k_edt_zsweep"],
       "ogm": ["k_pc_register", "k_pc_walk", "k_pc_scan", "k_pc_apply", "k_pc_undo"], "hash_merge": ["k_build_btab", "k_clear_prev_blocks", "k_list_merge_blocks", "k_merge_ogm"],
       "mark_frontier": ["k_list_blocks", "k_mark", "k_mark_blocks", "k_frontiers"], "waves": ["k_waves"], "commit": ["k_commit"]}
out = {g: sum(per[n]["dram_bytes"] for n in ns if n in per) for g, ns in grp.items()}
out["_source"] = label
out["_per_kernel"] = per
json.dump(out, open(out_json, "w"), indent=1)
print(json.dumps({g: out[g] for g in grp}))
