"""Per-source-line share of executed warp instructions and stall samples of one kernel in an .ncu-rep captured with
--import-source on (all source files the kernel inlines):
    python tools/ncu_source_lines.py rep.ncu-rep kernel_regex [min_pct] [--by-samples]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 0.5
    by_samples = "--by-samples" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    agg, cur, cur_file, hdr, tot_i, tot_s = {}, None, "", None, 0, 0
    seen_funcs = 0
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index(
                "Thread Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0]:
            try:
                cur = (cur_file, int(r[0]))
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
    print(f"warp instructions {tot_i}, samples {tot_s} (summed over the captured launches)")
    items = sorted(agg.items(), key=(lambda kv: -kv[1][2]) if by_samples else (lambda kv: kv[0]))
    for (f, ln), (src, i, s, t) in items:
        if i > tot_i * min_pct / 100 or s > tot_s * min_pct / 100:
            print(f"{f}:{ln:4d} {100 * i / max(tot_i, 1):5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% samp "
                  f"lanes {t / max(i, 1):4.1f} | {src.strip()[:100]}")


if __name__ == "__main__":
    main()
