"""Round-2 profile summary: from the .ncu-rep files and launch lists a gpurun call left in gpurun_out/ to the tracked
text / json files under profiles/.   python profiles/micro/r02_summarise.py"""
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarise(rep, title, lines, traffic_key=None, traffic=None):
    rows = ncu_csv(rep, "raw")
    hdr, units, r = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    lines.append("=" * 110)
    lines.append("%s   [%s]" % (title, os.path.basename(rep)))
    lines.append(r[ix["Kernel Name"]][:150])
    for w in WANT:
        if w in ix:
            lines.append("  %-70s %s %s" % (w, r[ix[w]], units[ix[w]]))
    rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
    mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd *= mul.get(units[ix["dram__bytes_read.sum"]], 1.0)
    wr *= mul.get(units[ix["dram__bytes_write.sum"]], 1.0)
    if traffic is not None and traffic_key:
        traffic[traffic_key] = rd + wr
    st = {h.split("stalled_")[-1]: float(r[i]) for i, h in enumerate(hdr)
          if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued") and r[i] not in ("0", "")}
    tot = sum(st.values()) or 1.0
    lines.append("  stall samples (%%): %s" % ", ".join("%s %.1f" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    # opcode mix from the source page (executed warp instructions)
    src = ncu_csv(rep, "source")
    h2 = src[1]
    jx = {h: i for i, h in enumerate(h2)}
    ops, T = collections.Counter(), 0
    for row in src[2:]:
        if len(row) < len(h2):
            continue
        t = row[jx["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        n = int(row[jx["Instructions Executed"]] or 0)
        ops[op] += n
        T += n
    lines.append("  executed warp instructions: %d; mix: %s" % (T, ", ".join("%s %.1f%%" % (k, 100 * v / T) for k, v in ops.most_common(12))))
    return rd + wr


def launch_share(path, lines, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]])
        except (ValueError, KeyError):
            continue
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = r[ix["Kernel Name"]].split("(")[0][:70]
        per[name][0] += 1
        per[name][1] += v
    tot = sum(v[1] for v in per.values()) or 1.0
    lines.append("=" * 110)
    lines.append(title)
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append("  %-72s launches %4d   total %10.1f us   share %5.1f%%   avg %8.1f us" % (k, n, us, 100 * us / tot, us / n))


def main():
    lines = ["Round 2 — ncu summaries (B200, sm_100a, --clock-control none).  Produced by profiles/micro/r02_summarise.py from",
             "the captures of profiles/micro/r02_ncu.sh (copy / one-step kernels) and r02_ncu_t2.sh (two-step sweep after the pillar march, launch lists); per-launch times under ncu are cold-cache and serialised: the kernel's",
             "SHARE of the step is what must agree with bench.py, not the absolute."]
    traffic = {}
    tp = os.path.join(OUT, "traffic.json")
    for rep, title, key in (("r02_t2.ncu-rep", "k_heat_t2<64>: two steps per pass, C2 level (512^3 in 64^3 boxes), one launch", "t2_dram_bytes_per_launch"),
                            ("r02_stream.ncu-rep", "k_heat_stream<64,64>: one-step sweep writing the outer edge strips, C2 level", "sweep_dram_bytes_per_launch"),
                            ("r02_copy.ncu-rep", "k_copy_items: stand-alone exchange of the C2 level, x faces from the edge strips, 16-byte path", "exchange_dram_bytes_per_launch")):
        p = os.path.join(G, rep)
        if os.path.exists(p):
            summarise(p, title, lines, key, traffic)
    for f, title in (("r02_launches.csv", "launch list: python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --e2e-steps 0 (C2)"),
                     ("r02_launches_c3.csv", "launch list: python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 (re-boxed loop)")):
        p = os.path.join(G, f)
        if os.path.exists(p):
            launch_share(p, lines, title)
            with open(os.path.join(OUT, f), "w") as o:
                o.write(open(p).read())
    with open(os.path.join(OUT, "r02_ncu_summary.txt"), "w") as o:
        o.write("\n".join(lines) + "\n")
    traffic["source"] = "ncu --set full, one launch each, profiles/micro/r02_ncu.sh + r02_ncu_t2.sh (round 2); t2/sweep on the 512^3/64^3 level"
    json.dump(traffic, open(tp, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
