"""Summarise an `ncu --set full` report (.ncu-rep) into the handful of numbers DESIGN.md / bench.py quote:
per-launch duration, tensor-pipe %, DRAM bytes, L2 throughput, issue activity, plus the hottest SASS lines of
the first launch with their stall reasons.  Runs here (no GPU): `python tools/ncu_summary.py X.ncu-rep > profiles/X.txt`"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"report: {rep}   launches captured: {len(data)}")
    for li, r in enumerate(data):
        print(f"\n== launch {li}: {r[col['Kernel Name']][:100]}")
        for k in KEYS:
            if k in col:
                print(f"   {k:75s} {r[col[k]]:>16s} {units[col[k]]}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-count", "1"]))))
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    shdr, sdata = src[h], src[h + 1:]
    sc = {n: i for i, n in enumerate(shdr)}
    seen, uniq = set(), []
    for r in sdata:
        if len(r) > sc["Source"] and r[sc["Address"]] not in seen:
            seen.add(r[sc["Address"]])
            uniq.append(r)

    def f(r, k):
        try:
            return float(r[sc[k]])
        except Exception:
            return 0.0
    stalls = [n for n in shdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(f(r, "# Samples") for r in uniq) or 1.0
    print(f"\n== launch 0: hottest SASS lines by warp-stall samples (total {tot:.0f})")
    for r in sorted(uniq, key=lambda r: -f(r, "# Samples"))[:top_n]:
        big = sorted(((k[6:], int(f(r, k))) for k in stalls), key=lambda kv: -kv[1])[:2]
        print(f"   {100 * f(r, '# Samples') / tot:5.1f}%  {r[sc['Source']][:86]:86s} {[b for b in big if b[1] > 0]}")
    agg = {k: sum(f(r, k) for r in uniq) for k in stalls}
    print("\n== launch 0: stall reasons, all warps:",
          ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
