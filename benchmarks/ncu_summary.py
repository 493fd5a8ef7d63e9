#!/usr/bin/env python
"""Text summary of an .ncu-rep (run on the GPU box so that only small files travel back):
the details page, plus per kernel the instruction mix and the SASS lines with the most stall samples.

    python benchmarks/ncu_summary.py REPORT.ncu-rep OUT.txt [top_lines]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    with open(out, "w") as f:
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        f.write(det)
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        sections, cur = [], None
        for r in rows:
            if r and r[0] == "Kernel Name":
                cur = {"name": r[1], "rows": []}
                sections.append(cur)
            elif r and r[0] == "Address" and cur is not None:
                cur["hdr"] = r
            elif cur is not None and len(r) > 5:
                cur["rows"].append(r)
        for s in sections:
            h = s.get("hdr")
            if not h or not s["rows"]:
                continue
            ie, sc, sm = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
            tot = sum(int(r[ie]) for r in s["rows"]) or 1
            tots = sum(int(r[sm]) for r in s["rows"]) or 1
            f.write(f"\n==== {s['name'][:150]}\n  warp instructions executed: {tot}, stall samples: {tots}\n")
            mix, smp = collections.Counter(), collections.Counter()
            for r in s["rows"]:
                ops = [o for o in r[sc].split() if not o.startswith("@")]
                op = ops[0].split(".")[0] if ops else "?"
                mix[op] += int(r[ie])
                smp[op] += int(r[sm])
            f.write("  instruction mix: " + ", ".join(f"{k} {100 * v / tot:.1f}% ({100 * smp[k] / tots:.0f}% smp)" for k, v in mix.most_common(12)) + "\n")
            f.write(f"  top {top} SASS lines by stall samples:\n")
            for r in sorted(s["rows"], key=lambda r: -int(r[sm]))[:top]:
                f.write(f"    {100 * int(r[sm]) / tots:5.1f}%  exec {int(r[ie]):>9d}  {r[sc].strip()[:110]}\n")
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of the raw page): the `traffic` of a roofline
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        launches = []
        if len(rows) > 2:
            h, units = rows[0], rows[1]
            try:
                kn, rd, wr, du = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
                for r in rows[2:]:
                    if len(r) <= max(rd, wr, du):
                        continue
                    fl = lambda v: float(v.replace(",", ""))
                    launches.append({"kernel": r[kn][:120], "dram_read_bytes": fl(r[rd]) * scale.get(units[rd], 1.0),
                                     "dram_write_bytes": fl(r[wr]) * scale.get(units[wr], 1.0),
                                     "duration_us": fl(r[du]) * tscale.get(units[du], 1.0)})
            except ValueError:
                pass
        if launches:
            f.write("\n==== DRAM traffic per launch (ncu --set full, raw page)\n")
            for l in launches:
                f.write(f"  {l['kernel'][:90]}: read {l['dram_read_bytes'] / 1e6:.1f} MB, write {l['dram_write_bytes'] / 1e6:.1f} MB, "
                        f"{l['duration_us']:.1f} us under ncu\n")
            import json
            with open(out.rsplit(".", 1)[0] + "_traffic.json", "w") as jf:
                json.dump({"source": "ncu --set full --clock-control none, raw page: dram__bytes_read.sum + dram__bytes_write.sum per launch",
                           "launches": launches}, jf, indent=1)


if __name__ == "__main__":
    main()
