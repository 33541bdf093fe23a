#!/usr/bin/env python
"""ncu --page raw --csv of a `--set full` capture -> one markdown row per launch: duration, tensor-pipe %, DRAM bytes and %, achieved
GB/s, registers, local-memory traffic, top two stall reasons.   python tools/ncu_table.py raw.csv out.md "title" """
import csv
import re
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(src)))
    while rows and not (rows[0] and rows[0][0] == "ID"):
        rows.pop(0)
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, key, scale=True):
        if key not in ix:
            return float("nan")
        v = num(r[ix[key]])
        u = units[ix[key]]
        if scale:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
                  "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return v
    stall_keys = [k for k in hdr if "issue_stalled" in k and "per_issue_active" in k]
    local_keys = [k for k in ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum") if k in ix]
    out = [f"# {title}\n", "Per-launch times under ncu are cold-cache and serialised (compare shares, not absolutes).  `tensor %` = "
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed; `dram MB` = dram__bytes_read.sum + dram__bytes_write.sum; stalls = "
           "smsp__average_warps_issue_stalled_*_per_issue_active.ratio (top two).\n",
           "| # | kernel | grid | us | tensor % | dram MB (r+w) | dram % | GB/s | regs | local ld+st MB | top stalls |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    agg = {}
    for n, r in enumerate(data):
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void |ydst::", "", name)
        dur = col(r, "gpu__time_duration.sum")
        rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
        tens = col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", False)
        dpct = col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False)
        regs = col(r, "launch__registers_per_thread", False)
        loc = sum(col(r, k, False) for k in local_keys) * 32.0          # 32-byte sectors
        stalls = sorted(((num(r[ix[k]]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys
                         if r[ix[k]] not in ("", "n/a")), reverse=True)[:2]
        out.append(f"| {n} | {name} | {r[ix['Grid Size']]} | {dur * 1e6:.1f} | {tens:.1f} | {(rd + wr) / 1e6:.2f} | {dpct:.1f} | {(rd + wr) / dur / 1e9:.0f} | {regs:.0f} | "
                   f"{loc / 1e6:.2f} | {', '.join(f'{k} {v:.1f}' for v, k in stalls)} |")
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += dur; a[2] += tens * dur if tens == tens else 0.0; a[3] += rd + wr
    tot = sum(a[1] for a in agg.values())
    out += ["", "## per kernel", "", "| kernel | launches | total us | share | time-weighted tensor % | dram GB/s |", "|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {a[0]} | {a[1] * 1e6:.1f} | {a[1] / tot:.3f} | {a[2] / a[1] if a[1] else 0:.1f} | {a[3] / a[1] / 1e9 if a[1] else 0:.0f} |")
    out.append(f"\n{len(data)} launches, {tot * 1e6:.1f} us total; local-memory columns found: {local_keys}\n")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-(len(agg) + 6):]))


if __name__ == "__main__":
    main()
