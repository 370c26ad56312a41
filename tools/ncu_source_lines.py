"""Per-source-line share of executed warp instructions and stall samples of one kernel in an .ncu-rep captured with
--import-source on:  python tools/ncu_source_lines.py rep.ncu-rep kernel_regex [min_pct]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Line No" and "Instructions Executed" in r)
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    agg, cur, tot_i, tot_s = {}, None, 0, 0
    first = True
    for r in rows:
        if len(r) < len(hdr):
            continue
        if r[0] == "Line No":
            if not first:
                break  # only the first launch
            first = False
            continue
        if r[0]:
            try:
                cur = int(r[0])
            except ValueError:
                continue
            agg.setdefault(cur, [r[1], 0, 0, 0])
            continue
        if cur is None:
            continue
        try:
            inst, samp, thr = int(r[i_i] or 0), int(r[i_s] or 0), int(r[i_t] or 0)
        except ValueError:
            continue
        agg[cur][1] += inst
        agg[cur][2] += samp
        agg[cur][3] += thr
        tot_i += inst
        tot_s += samp
    print(f"warp instructions {tot_i}, samples {tot_s}")
    for ln, (src, i, s, t) in sorted(agg.items()):
        if i > tot_i * min_pct / 100 or s > tot_s * min_pct / 100:
            print(f"{ln:4d} {100 * i / max(tot_i, 1):5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% samp "
                  f"lanes {t / max(i, 1):4.1f} | {src.strip()[:110]}")


if __name__ == "__main__":
    main()
