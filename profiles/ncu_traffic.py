"""ncu CSV (one row per kernel launch and metric, `--csv --log-file`, as produced by the command in
profiles/one_step.py) -> profiles/traffic.json: per-kernel DRAM bytes per launch and the total of the step, keyed by
the git SHA the capture was taken at.  bench.py reads the file for `roofline.traffic` / `roofline.step_traffic`.
  python profiles/ncu_traffic.py gpurun_out/traffic.csv [out.json]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short_name(full):
    """`void dpk::<unnamed>::ratspn_leaf_mma_kernel<0, 0, 1>(args)` -> ratspn_leaf_mma_kernel<conv>"""
    tail = full.split("unnamed>::", 1)[1] if "unnamed>::" in full else full
    m = re.match(r"\s*(?:void\s+)?(?:\w+::)*([A-Za-z_]\w*)\s*(<.*>)?\s*\(", tail)
    name = m.group(1) if m else full
    targs = (m.group(2) or "") if m else ""
    if name == "ratspn_leaf_mma_kernel":
        flags = re.findall(r"(?:\(bool\))?\s*(1|0|true|false)", targs)
        on = [f in ("1", "true") for f in flags] + [False] * 3
        return name + ("<prep>" if on[0] else "<conv>" if on[2] else "<main>")
    return name


def main():
    src = sys.argv[1]
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "traffic.json")
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    head = rows[0]
    ix = {h: i for i, h in enumerate(head)}
    launches = {}
    for r in rows[1:]:
        if len(r) < len(head):
            continue
        key = r[ix["ID"]]
        d = launches.setdefault(key, {"name": short_name(r[ix["Kernel Name"]]), "full": r[ix["Kernel Name"]]})
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]].lower()
        mname = r[ix["Metric Name"]]
        if "byte" in unit:
            val *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        if unit in ("us", "usecond"):
            val *= 1e3
        elif unit in ("ms", "msecond"):
            val *= 1e6
        elif unit in ("s", "second"):
            val *= 1e9
        d[mname] = val
    kernels, step_total, step_ns = {}, 0.0, 0.0
    for d in launches.values():
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        k = kernels.setdefault(d["name"], {"launches": 0, "dram_bytes": 0.0, "read": 0.0, "write": 0.0, "ns": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += b
        k["read"] += d.get("dram__bytes_read.sum", 0.0)
        k["write"] += d.get("dram__bytes_write.sum", 0.0)
        k["ns"] += d.get("gpu__time_duration.sum", 0.0)
        step_total += b
        step_ns += d.get("gpu__time_duration.sum", 0.0)
    for k in kernels.values():
        k["dram_bytes_per_launch"] = k["dram_bytes"] / k["launches"]
        k["us_per_launch_under_ncu"] = k["ns"] / k["launches"] / 1e3
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"git_sha": sha, "source": os.path.basename(src), "step_dram_bytes": step_total,
           "step_us_under_ncu_serialised": step_ns / 1e3, "kernels": kernels,
           "note": "one eager step of the bench workload (profiles/one_step.py), batch 65536; per-launch times are "
                   "cold-cache and serialised: compare shares, not absolutes"}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, "step traffic %.1f MB over %d launches" % (step_total / 1e6, len(launches)))


if __name__ == "__main__":
    main()
