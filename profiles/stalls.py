"""Top stalled SASS instructions of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python profiles/stalls.py gpurun_out/prof.ncu-rep regex:edge_bwd_kernel [launch_skip] [n_top]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--launch-skip", skip,
                          "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:150])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = {h: sum(int(r[ix[h]]) for r in data if r[ix[h]].isdigit()) for h in stall_cols}
    print("samples", tot, "sass instructions", len(data))
    print("stall totals:", [(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]:
        st = {h[6:]: int(r[ix[h]]) for h in stall_cols if r[ix[h]].isdigit() and int(r[ix[h]]) > 0}
        print(r[ix["# Samples"]].rjust(7), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:64].ljust(64),
              dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))


if __name__ == "__main__":
    main()
