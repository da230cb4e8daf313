"""Instruction mix and stall-sample share per opcode of one kernel out of an ncu report captured with --import-source on:
    ncu -i REPORT.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_mix.py src.csv
(all launches of the kernel in the report are summed)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    i_src, i_exec, i_stall = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    mix, stall = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_exec or r[i_exec] == "Instructions Executed":
            continue
        toks = r[i_src].strip().split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")
        key = op[0]
        if key == "IMAD" and len(op) > 1:
            key += "." + op[1]
        if key in ("LDG", "STG", "LDS", "STS") and op[-1].isdigit():
            key += "." + op[-1]
        mix[key] += int(float(r[i_exec] or 0))
        stall[key] += int(float(r[i_stall] or 0))
    tot, tots = sum(mix.values()), max(1, sum(stall.values()))
    print(rows[0][1])
    print(f"warp instructions {tot}, stall samples {tots}")
    for key, e in mix.most_common(24):
        print(f"{key:12s} {e:10d} {100 * e / tot:5.1f}%   stall samples {100 * stall[key] / tots:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
