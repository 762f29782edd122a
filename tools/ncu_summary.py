"""Condense an `ncu --csv --page raw` log into one row per kernel NAME (launches summed): launches, total time, DRAM bytes read /
written, DRAM throughput %, tensor-pipe %, SM throughput %, achieved occupancy (warps active %), registers, L2 bytes.

    python tools/ncu_summary.py gpurun_out/raw.csv > profiles/summary.csv
"""
import collections
import csv
import re
import sys


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main(path):
    rows = [ln for ln in open(path, errors="replace") if ln.startswith('"')]
    rd = csv.reader(rows)
    header = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(header)}

    def find(*pats):
        for p in pats:
            for h in header:
                if re.fullmatch(p, h):
                    return h
        return None

    want = {
        "time": find(r"gpu__time_duration\.sum"),
        "dram_rd": find(r"dram__bytes_read\.sum"),
        "dram_wr": find(r"dram__bytes_write\.sum"),
        "dram_pct": find(r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed", r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed"),
        "tensor_pct": find(r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active", r"sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_active",
                           r"sm__inst_executed_pipe_tensor.*pct_of_peak_sustained_active"),
        "sm_pct": find(r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed"),
        "warps_pct": find(r"sm__warps_active\.avg\.pct_of_peak_sustained_active"),
        "regs": find(r"launch__registers_per_thread"),
        "l2_bytes": find(r"lts__t_bytes\.sum"),
        "smem": find(r"launch__shared_mem_per_block_dynamic", r"launch__shared_mem_per_block"),
        "grid": find(r"launch__grid_size"),
    }
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rd:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        a = agg.setdefault(name, collections.defaultdict(float))
        a["n"] += 1
        for k, h in want.items():
            if not h:
                continue
            v = num(r[col[h]])
            if v is None:
                continue
            v *= scale.get(units[col[h]], 1.0) if k in ("time", "dram_rd", "dram_wr", "l2_bytes") else 1.0
            if k in ("time", "dram_rd", "dram_wr", "l2_bytes"):
                a[k] += v
            elif k in ("regs", "smem", "grid"):
                a[k] = max(a[k], v)
            else:
                a[k + "_tw"] += v * (num(r[col[want["time"]]]) or 0.0)  # time-weighted mean of the percentages
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "time_us", "share_pct", "dram_read_MB", "dram_write_MB", "dram_GBps", "dram_pct_of_peak", "tensor_pipe_pct",
                "sm_throughput_pct", "warps_active_pct", "registers", "smem_dyn_B", "max_grid", "l2_MB"])
    tot = sum(a["time"] for a in agg.values()) or 1.0
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time"]):
        tw = a["time"] / (scale.get(units[col[want["time"]]], 1.0)) if want["time"] else 0.0

        def pct(k):
            return f"{a[k + '_tw'] / tw:.1f}" if tw and (k + "_tw") in a else ""

        w.writerow([name, int(a["n"]), f"{a['time']:.1f}", f"{100 * a['time'] / tot:.1f}", f"{a['dram_rd'] / 1e6:.2f}", f"{a['dram_wr'] / 1e6:.2f}",
                    f"{(a['dram_rd'] + a['dram_wr']) / 1e3 / a['time']:.0f}" if a["time"] else "", pct("dram_pct"), pct("tensor_pct"), pct("sm_pct"),
                    pct("warps_pct"), int(a["regs"]), int(a["smem"]), int(a["grid"]), f"{a['l2_bytes'] / 1e6:.1f}"])


if __name__ == "__main__":
    main(sys.argv[1])
