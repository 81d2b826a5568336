"""Summarise .ncu-rep captures into profiles/ (markdown table + JSON of per-launch DRAM traffic that bench.py reads)."""
import csv, json, subprocess, sys, io, os

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_size", "cluster"), ("smsp__inst_executed.sum", "inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait")]

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "ns": 1e-6, "second": 1e3, "s": 1e3}


def load(rep):
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` export made on the GPU box (reports with 100 full captures exceed the
    # size that travels back)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")}
        for m, key in WANT:
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if key in ("time",):
                    v *= UNIT.get(u, 1.0)            # -> ms
                elif key in ("dram_rd", "dram_wr", "l2_bytes"):
                    v *= UNIT.get(u, 1.0)            # -> bytes
                d[key] = v
        res.append(d)
    return res


def main():
    reps = sys.argv[1:-1]
    tag = sys.argv[-1]
    allk = []
    for rep in reps:
        allk += load(rep)
    lines = [f"# ncu --set full --clock-control none captures ({tag}); one row per captured launch",
             "", "| kernel | grid x block (cluster) | ms | DRAM rd MB | DRAM wr MB | L2 MB | DRAM % | SM % | tensor % | issue % | warps % | regs | stall long-sb | stall barrier |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for d in allk:
        lines.append("| {k} | {g:.0f} x {b:.0f} ({c:.0f}) | {t:.3f} | {r:.1f} | {w:.1f} | {l:.1f} | {dp:.1f} | {sp:.1f} | {tp:.2f} | {ip:.1f} | {wp:.1f} | {rg:.0f} | {s1:.2f} | {s2:.2f} |".format(
            k=d["kernel"][:60], g=d.get("grid", 0), b=d.get("block", 0), c=d.get("cluster", 0), t=d.get("time", 0), r=d.get("dram_rd", 0) / 1e6,
            w=d.get("dram_wr", 0) / 1e6, l=d.get("l2_bytes", 0) / 1e6, dp=d.get("dram_pct", 0), sp=d.get("sm_pct", 0), tp=d.get("tensor_pct", 0),
            ip=d.get("issue_pct", 0), wp=d.get("warps_pct", 0), rg=d.get("regs", 0), s1=d.get("stall_long_sb", 0), s2=d.get("stall_barrier", 0)))
    os.makedirs("profiles", exist_ok=True)
    open(f"profiles/{tag}.md", "w").write("\n".join(lines) + "\n")
    # per-kernel traffic of one launch per kernel family (what bench.py reports as roofline.traffic): the launch that
    # moves the most DRAM bytes for the GEMM families (bench.py reports the shape with the most algorithmic bytes),
    # the longest launch otherwise
    fam = {}
    for d in allk:
        name = "fps_kernel" if "fps" in d["kernel"] else "knn_grid_kernel" if "knn_grid" in d["kernel"] else \
            "gemm_wgrad_kernel" if "wgrad" in d["kernel"] else "gemm_rows_kernel" if "gemm_rows" in d["kernel"] else d["kernel"]
        gemm = name.startswith("gemm_")
        key = (lambda e: e["dram_bytes"]) if gemm else (lambda e: e["ms"])
        cand = {"ms": d.get("time", 0), "dram_bytes": d.get("dram_rd", 0) + d.get("dram_wr", 0)}
        if name not in fam or key(cand) > key(fam[name]):
            fam[name] = {"ms": d.get("time", 0), "dram_bytes": d.get("dram_rd", 0) + d.get("dram_wr", 0), "l2_bytes": d.get("l2_bytes", 0)}
    json.dump(fam, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
