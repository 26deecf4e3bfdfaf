#!/usr/bin/env python3
"""Turn an .ncu-rep (captured under gpurun with `ncu --set full --import-source on`) into the small text
summary committed under profiles/: headline metrics, stall reasons per issued instruction, and executed
instructions per device function (function boundaries from nvdisasm of the library that was profiled).

usage: python profiles/summarise.py <report.ncu-rep> <kernel-name-substring> [libitw_bcn.so [mangled-section-substring]] > profiles/<name>.txt
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        # which pipe binds: share of its peak for every issue pipe, and the shared-memory wavefronts behind the conflicts
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]


def ncu_csv(rep, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else None
    section = sys.argv[4] if len(sys.argv) > 4 else kern
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    print(f"# {os.path.basename(rep)} -- kernel filter '{kern}'")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        if kern not in d.get("Kernel Name", ""):
            continue
        print("kernel:", d["Kernel Name"])
        for k in KEYS:
            if k in d:
                print(f"  {k:66s} {d[k]:>18s} {units[hdr.index(k)]}")
        st = {h: float(v) for h, v in zip(hdr, vals) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
        print("  warp stall reasons (avg warps stalled per issued instruction):")
        for h, v in sorted(st.items(), key=lambda x: -x[1])[:9]:
            print(f"    {h.split('stalled_')[1].split('_per')[0]:22s} {v:7.3f}")
    if not lib:
        return
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = end = None
    for i, l in enumerate(dis):
        if l.startswith(".text.") and section in l and l.endswith(":"):
            start = i
        elif start is not None and l.startswith("//---------------------") and i > start:
            end = i
            break
    end = end or len(dis)
    labels, idx = [(0, kern + " (kernel body)")], 0
    for l in dis[start:end]:
        m = re.match(r"^\$.*\$(_ZN3itw\w+):$", l) or re.match(r"^(\$__internal_\d+_\$\w+):$", l)
        if m:
            labels.append((idx, m.group(1)))
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            idx += 1
    src = ncu_csv(rep, "--page", "source", "--print-source", "sass")
    h = src[1]
    ia, iex, ith, ismp, ino = (h.index(x) for x in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples", "stall_no_inst"))
    data = [r for r in src[2:] if len(r) > iex and r[ia].startswith("0x")]
    if len(data) != idx:
        print(f"  (SASS of the report has {len(data)} instructions, the library {idx}: function table skipped)")
        return
    agg, tot = {}, sum(int(r[iex]) for r in data)
    for i, r in enumerate(data):
        name = [n for s, n in labels if s <= i][-1]
        a = agg.setdefault(name, [0, 0, 0, 0, 0])
        a[0] += int(r[iex]); a[1] += int(r[ith]); a[2] += int(r[ismp]); a[3] += int(r[ino]); a[4] += 1
    print(f"  executed warp instructions per device function (total {tot}):")
    print(f"    {'function':40s} {'SASS':>6s} {'exec %':>7s} {'thr/inst':>9s} {'samples':>8s} {'no_inst %':>9s}")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        dn = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"    {dn[:40]:40s} {a[4]:6d} {100 * a[0] / tot:7.1f} {a[1] / max(a[0], 1):9.1f} {a[2]:8d} {100 * a[3] / max(a[2], 1):9.0f}")


if __name__ == "__main__":
    main()
