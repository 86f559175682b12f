"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.
    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches_by_kernel.txt
    python profiles/summarize.py full gpurun_out/prof_gemm_r1.ncu-rep > profiles/r1_gemm_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total")
    print("# (cold-cache, serialised replay: compare SHARES, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} n={v[0]:5d} total={v[1] / 1e3:9.3f} ms avg={v[1] / v[0]:9.1f} us share={100 * v[1] / tot:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        d = dict(zip(hdr, row))
        print("kernel:", d.get("Kernel Name", "?")[:120], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for i, h in enumerate(hdr):
            if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                print(f"   {h:95s} {row[i]:>16s} {units[i]}")
        print()


def sass_summary(lib_path):
    """per-kernel counts of the Blackwell-native SASS mnemonics (tcgen05 MMA = UTCHMMA / UTCQMMA..., TMEM ld/st = LDTM / STTM,
    TMA = UTMALDG / UTMASTG / UBLKCP, mbarrier = SYNCS) next to the legacy tensor-core HMMA, from `cuobjdump -sass`."""
    import collections
    import re
    import subprocess
    txt = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    pats = {"UTCHMMA": r"\bUTC[A-Z]*MMA\b", "LDTM": r"\bLDTM\b", "STTM": r"\bSTTM\b", "UTMALDG": r"\bUTMALDG\b", "UTMASTG": r"\bUTMASTG\b",
            "UBLKCP": r"\bUBLKCP\b", "SYNCS": r"\bSYNCS\b", "HMMA": r"\bHMMA\b", "MUFU.EX2": r"\bMUFU\.EX2\b", "LDGSTS": r"\bLDGSTS\b"}
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
    dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(counts, dem)) if len(dem) == len(counts) else {k: k for k in counts}
    cols = list(pats)
    print("# cuobjdump -sass %s : instruction counts per kernel" % lib_path)
    print("%-110s " % "kernel" + " ".join("%8s" % c for c in cols))
    tot = collections.Counter()
    for k, c in counts.items():
        full = names[k].replace("void ", "").replace("showo::", "")
        depth, cut = 0, len(full)
        for i, ch in enumerate(full):          # cut at the parameter list: the first "(" outside the template brackets
            depth += ch == "<"
            depth -= ch == ">"
            if ch == "(" and depth == 0:
                cut = i
                break
        label = full[:cut][:110]
        print("%-110s " % label + " ".join("%8d" % c[x] for x in cols))
        tot.update(c)
    print("%-110s " % "TOTAL" + " ".join("%8d" % tot[x] for x in cols))


if __name__ == "__main__":
    {"launches": launches, "full": full, "sass": sass_summary}[sys.argv[1]](sys.argv[2])
