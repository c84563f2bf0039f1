"""Summarises an .ncu-rep (no GPU needed): key throughput metrics + sampled stall reasons + hottest source lines.
    python tools/ncu_summary.py gpurun_out/prof_attn.ncu-rep [--lines 12]
"""
import csv
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[-1]))


def main():
    path = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 12
    m = raw(path)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
    print(f"== {path}  kernel: {m.get('Kernel Name', '?')[:80]}")
    for k in keys:
        if k in m:
            print(f"  {k:72s} {m[k]}")
    stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in m.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "n/a")}
    tot = sum(stalls.values()) or 1.0
    print("  sampled stall reasons:", ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if rows and rows[0] and rows[0][0] == "Kernel Name":
        rows = rows[1:]
    if len(rows) > 2:
        hdr = rows[0]
        try:
            i_src = hdr.index("Source")
            i_smp = hdr.index("# Samples")
            i_exe = hdr.index("Instructions Executed")
        except (ValueError, StopIteration):
            print("  (source page columns:", hdr[:12], ")")
            return
        items = []
        for r in rows[1:]:
            try:
                items.append((float(r[i_smp]), r[i_src], float(r[i_exe])))
            except (ValueError, IndexError):
                pass
        total = sum(t[0] for t in items) or 1.0
        texec = sum(t[2] for t in items) or 1.0
        ops = {}
        for v, s, e in items:
            op = s.strip().split()[0] if s.strip() else "?"
            if op.startswith("@"):
                op = s.strip().split()[1]
            op = op.split(".")[0]
            d = ops.setdefault(op, [0.0, 0.0])
            d[0] += v
            d[1] += e
        print("  by opcode (stall samples % | executed %):",
              ", ".join(f"{k} {100 * a / total:.0f}|{100 * b / texec:.0f}" for k, (a, b) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:12]))
        print("  hottest SASS lines by stall samples (idx: % samples, executed):")
        order = sorted(range(len(items)), key=lambda i: -items[i][0])[:nlines]
        for i in order:
            v, s, e = items[i]
            print(f"    [{i:5d}] {100 * v / total:5.1f}%  x{int(e):>9d}  {s.strip()[:100]}")


if __name__ == "__main__":
    main()
