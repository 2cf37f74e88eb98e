"""Turn the scratch ncu captures under gpurun_out/ into the tracked summaries under profiles/.

    python tools/make_profiles.py <tag> <launch list csv> <ncu --set full report>

writes profiles/<tag>_launches_raw.csv (ncu's per-launch durations), profiles/<tag>_launches_one_step.csv
(the last fused step of that run: every libmal_b200 launch with its duration and share),
profiles/<tag>_ncu_full_summary.txt (the metrics that decide what bounds each heavy kernel) and refreshes
profiles/dram_traffic.json (dram bytes per launch, read by bench.py for roofline.traffic).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

FULL_KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
             "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
             "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
             "l1tex__throughput.avg.pct_of_peak_sustained_active",
             "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(PROF, f"{tag}_launches_raw.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_us\n")
        for r in rows:
            f.write(f'{r[0]},"{r[4]}","{r[8]}","{r[7]}",{float(r[14]) / 1e3:.2f}\n')
    ours = [r for r in rows if "mal::" in r[4]]
    # the run executes 1 warm-up + 2 timed steps + reference extras: the last complete step ends with step_combine
    ends = [i for i, r in enumerate(ours) if "step_combine" in r[4]]
    start = ends[-2] + 1 if len(ends) > 1 else 0
    step = ours[start:ends[-1] + 1]
    total = sum(float(r[14]) for r in step) / 1e3
    with open(os.path.join(PROF, f"{tag}_launches_one_step.csv"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, one fused MAL step (B=12, 192x640): "
                f"{len(step)} launches, {total:.1f} us serialised\n")
        f.write("id,kernel,grid,block,duration_us,share\n")
        for r in step:
            us = float(r[14]) / 1e3
            f.write(f'{r[0]},"{r[4]}","{r[8]}","{r[7]}",{us:.2f},{us / total:.3f}\n')
    return len(step), total


def full(tag, rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(os.path.join(PROF, f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none summary, {tag} (profiles/{tag}_notes.md has the command)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n\n## {name}\n\n")
            for k in FULL_KEYS:
                if k in hdr:
                    f.write(f"{k:76s}{r[hdr.index(k)]} {units[hdr.index(k)]}\n")
            st = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) >= 0.15:
                    st.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            f.write("stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)) + "\n")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            traffic.setdefault(name, []).append(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]])
    return traffic, rows


def limiter_sentence(hdr, r):
    """One sentence from the counters that decide what bounds a kernel (no hand-written numbers)."""
    g = lambda k: float(r[hdr.index(k)]) if k in hdr and r[hdr.index(k)] else float("nan")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]:
            st.append((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
    top = ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:3] if n != "selected")
    return ("not HBM-bound (bit-exact fp32 arithmetic, DESIGN.md section 4): DRAM throughput %.1f%% of peak; issue slots "
            "%.0f%%, fma pipe %.0f%%, L1 data pipe %.0f%%, %.0f%% of the warp slots resident at %d registers; top stalls "
            "per issue: %s" % (g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                               g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                               g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                               g("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
                               g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                               int(g("launch__registers_per_thread")), top))


def main():
    tag, lcsv, rep = sys.argv[1:4]
    n, total = launches(tag, lcsv)
    print(f"{tag}: one step = {n} launches, {total:.1f} us serialised")
    traffic, rows = full(tag, rep)
    pick = lambda pat, idx=0: next((v[min(idx, len(v) - 1)] for k, v in traffic.items() if pat in k), None)
    photo = lambda pat: [v for k, v in traffic.items() if pat in k]
    dram = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, {tag} "
                     f"(profiles/{tag}_ncu_full_summary.txt); cost_volume = cv_pack_kernel (lookup) + cv_sweep_quad_kernel"}
    old = {}
    if os.path.exists(os.path.join(PROF, "dram_traffic.json")):
        old = json.load(open(os.path.join(PROF, "dram_traffic.json")))
    # cv_pack_kernel is a plain transpose whose traffic does not change: keep the last captured value
    sweep, pack = pick("cv_sweep"), pick("cv_pack") or old.get("cv_pack_kernel")
    if sweep and pack:
        dram.update(cost_volume=int(round(sweep + pack, -5)), cv_sweep_quad_kernel=int(round(sweep, -5)),
                    cv_pack_kernel=int(round(pack, -5)))
    # photo_kernel<WARP, GRAD, CONV, LOWRES, SYNG, NC>: the teacher pass is the 4-candidate gradient kernel, the
    # student the 2-candidate one; ensemble = WARP without gradient, identity = PRED without gradient
    import re
    def first(rx):
        return next((v[0] for k, v in traffic.items() if re.search(rx, k)), None)
    pats = (("photo_teacher", r"photo_kernel<1, 1, \d, \d, \d, 4"), ("photo_student", r"photo_kernel<1, 1, \d, \d, \d, 2"),
            ("photo_ensemble", r"photo_kernel<1, 0,"), ("photo_identity", r"photo_kernel<0, 0,"),
            ("smooth_kernel", r"smooth_kernel"), ("cost_volume", r"cv_sweep"))
    for key, rx in pats:
        v = first(rx)
        if v and key != "cost_volume":
            dram[key] = int(round(v, -5))
    json.dump(dram, open(os.path.join(PROF, "dram_traffic.json"), "w"), indent=1)
    print(json.dumps(dram, indent=1))
    # what bounds each heavy kernel, in words, from the same capture (bench.py prints it as roofline.limiter)
    hdr = rows[0]
    lim = {"_note": f"generated by tools/make_profiles.py from the ncu --set full capture {tag} (profiles/{tag}_ncu_full_summary.txt)"}
    for key, rx in pats:
        r = next((r for r in rows[2:] if re.search(rx, r[hdr.index("Kernel Name")])), None)
        if r is not None:
            lim[key] = limiter_sentence(hdr, r) + f" (ncu {tag})"
    json.dump(lim, open(os.path.join(PROF, "limiters.json"), "w"), indent=1)
    print(json.dumps(lim, indent=1))


if __name__ == "__main__":
    main()
