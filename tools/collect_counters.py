#!/usr/bin/env python3
"""Measure, on the GPU box, the per-launch DRAM traffic and executed warp instructions of the dominant kernel for every
config of bench.py's sweep, and write them as profiles/dram_traffic.json-style JSON (bench.py reads that file for
roofline.traffic / the issue-rate roofline).

    gpurun -- 'python tools/collect_counters.py gpurun_out/counters.json'       (then copy into profiles/dram_traffic.json)

One `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum` pass per config over
`bench.py --format F --profile P --size S --steps 1 --warmup 1 --no-cpu --no-extras`; the last WHOLE-SURFACE launch of the kernel is taken (most executed
instructions; the end-to-end leg's row-band launches are smaller)."""
import csv
import json
import subprocess
import sys

SWEEP = [("BC1", None, "bc1_bc3"), ("BC3", None, "bc1_bc3"), ("BC6H", "bc6h_slow", "bc6h_kernel"), ("BC7", "slow", "bc7_kernel"), ("BC7", "basic", "bc7_kernel")]
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/counters.json"
    out = {"_comment": "per launch of the dominant kernel at the named workload (tools/collect_counters.py: ncu --metrics " + METRICS +
                       " --clock-control none, last whole-surface launch of `bench.py --steps 1 --warmup 1`): dram = read + write bytes, warp_inst = smsp__inst_executed.sum",
           "warp_inst": {}, "_kernel": {}}
    for size in (4096, 8192):
        for fmt, prof, kern in SWEEP:
            cmd = ["ncu", "--clock-control", "none", "--metrics", METRICS, "-k", f"regex:{kern}", "--csv", sys.executable, "bench.py", "--format", fmt,
                   "--size", str(size), "--steps", "1", "--warmup", "1", "--no-cpu", "--no-extras"]
            if prof:
                cmd += ["--profile", prof]
            res = subprocess.run(cmd, capture_output=True, text=True)
            rows = [r for r in csv.reader(res.stdout.splitlines()) if len(r) > 10]
            while rows and "Metric Name" not in rows[0]:               # bench.py's own JSON line also parses as a long CSV row
                rows.pop(0)
            rows = [r for r in rows if len(r) == len(rows[0])]
            if len(rows) < 2:
                print("no rows for", fmt, prof, size, res.stderr[-300:], file=sys.stderr)
                continue
            hdr = rows[0]
            im, iv, ik, iid = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name"), hdr.index("ID")
            inst = {int(r[iid]): float(r[iv].replace(",", "")) for r in rows[1:] if r[im] == "smsp__inst_executed.sum"}
            last = max(inst, key=lambda i: (inst[i], i))               # the whole-surface launch (the end-to-end leg encodes in row bands): most instructions, latest
            vals = {r[im]: float(r[iv].replace(",", "")) for r in rows[1:] if int(r[iid]) == last}
            unit = {r[im]: r[hdr.index("Metric Unit")] for r in rows[1:] if int(r[iid]) == last}
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            dram = sum(vals[m] * scale.get(unit[m], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            key = f"{fmt}:{prof}:{size}"
            out[key] = int(dram)
            out["warp_inst"][key] = int(vals["smsp__inst_executed.sum"])
            out["_kernel"][key] = [r[ik] for r in rows[1:] if int(r[iid]) == last][0][:80]
            print(key, out[key], out["warp_inst"][key], flush=True)
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
