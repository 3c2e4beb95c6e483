"""Per-kernel summary of an ncu raw-page CSV (`ncu -i X.ncu-rep --page raw --csv > X_raw.csv`): launches, device time, DRAM
traffic and throughput, tensor-pipe activity, occupancy, registers / shared memory — one row per kernel name, sorted by time.
Usage: ncu_summary.py <raw.csv> [--md]"""
import csv, re, sys, collections

path = sys.argv[1]
md = "--md" in sys.argv
rows = list(csv.reader(open(path, newline="")))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, first = r, i + 2          # next row = units
        units = rows[i + 1]
        break
assert hdr is not None, "no header row"
col = {h: i for i, h in enumerate(hdr)}


def find(*subs):
    for h in hdr:
        if all(s in h for s in subs):
            return col[h]
    return None


C = dict(
    t=find("gpu__time_duration.sum"), rd=find("dram__bytes_read.sum"), wr=find("dram__bytes_write.sum"),
    rate=find("dram__bytes.sum.per_second"),
    dram=find("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    tensor=find("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or find("sm__inst_executed_pipe_tensor"),
    smthr=find("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    occ=find("sm__warps_active.avg.pct_of_peak_sustained_active"), regs=find("launch__registers_per_thread"),
    smem=find("launch__shared_mem_per_block_dynamic"), grid=find("launch__grid_size"), block=find("launch__block_size"))


def val(r, k, scale_unit=False):
    i = C[k]
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    if scale_unit:
        u = units[i]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6,
              "Gbyte/s": 1e9, "Tbyte/s": 1e12}.get(u, 1.0)
    return v


agg = collections.OrderedDict()
for r in rows[first:]:
    if len(r) <= col["Kernel Name"]:
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = name.replace("void ", "").replace("dvgr::", "").replace("v_bf16::", "")[:64]
    a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, dram=[], tensor=[], smthr=[], occ=[], regs=None, smem=None, grid=None))
    a["n"] += 1
    a["t"] += val(r, "t", True) or 0.0
    if C["rd"] is not None:
        a["rd"] += val(r, "rd", True) or 0.0
        a["wr"] += val(r, "wr", True) or 0.0
    elif C["rate"] is not None:          # sections without the byte counters: bytes = DRAM rate x duration
        a["rd"] += (val(r, "rate", True) or 0.0) * (val(r, "t", True) or 0.0) * 1e-6
    for k in ("dram", "tensor", "smthr", "occ"):
        v = val(r, k)
        if v is not None:
            a[k].append(v)
    a["regs"], a["smem"], a["grid"] = val(r, "regs"), val(r, "smem"), val(r, "grid")
tot = sum(a["t"] for a in agg.values())
mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")
sep = " | " if md else "  "
head = ["kernel", "n", "us/launch", "% time", "DRAM MB/launch", "GB/s", "DRAM %", "tensor %", "SM %", "occ %", "regs", "smem KB", "grid"]
if md:
    print("| " + " | ".join(head) + " |\n|" + "---|" * len(head))
else:
    print(f"total device time {tot / 1e3:.3f} ms over {sum(a['n'] for a in agg.values())} launches")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    us = a["t"] / a["n"]
    mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
    cells = [name, str(a["n"]), f"{us:.1f}", f"{100 * a['t'] / tot:.1f}", f"{mb:.1f}", f"{mb * 1e6 / (us * 1e-6) / 1e9:.0f}" if us > 0 else "-",
             f"{mean(a['dram']):.0f}", f"{mean(a['tensor']):.0f}", f"{mean(a['smthr']):.0f}", f"{mean(a['occ']):.0f}",
             f"{a['regs']:.0f}" if a["regs"] is not None else "-", f"{(a['smem'] or 0) / 1e3:.0f}", f"{a['grid']:.0f}" if a["grid"] is not None else "-"]
    print(("| " + " | ".join(cells) + " |") if md else sep.join(cells))
