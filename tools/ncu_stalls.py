#!/usr/bin/env python
"""Source-level stall table of one kernel from an ncu report (`--import-source on`, built with -lineinfo):

    python tools/ncu_stalls.py gpurun_out/r02_prof.ncu-rep k_medoid_screen_sym profiles/r02a_stalls_screen_sym.txt

Writes the kernel's warp-stall samples by reason (all samples and not-issued samples), by SASS opcode, and the
25 instructions that collect the most samples."""
import collections
import csv
import io
import subprocess
import sys


def main(rep, kernel, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    blocks = raw.split('"Kernel Name",')
    text = ""
    for b in blocks[1:2]:                     # first matching launch
        lines = b.splitlines()
        name = lines[0]
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr = rows[0]
        reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
        ni = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" in h]
        col = {h: i for i, h in enumerate(hdr)}
        tot, tot_ni, by_op, by_op_inst = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
        insts = []
        for r in rows[1:]:
            if len(r) < len(hdr):
                continue
            op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
            if op.startswith("@"):
                op = r[col["Source"]].split()[1]
            op = op.split(".")[0]
            n = int(r[col["# Samples"]] or 0)
            by_op[op] += n
            by_op_inst[op] += int(r[col["Instructions Executed"]] or 0)
            for h in reasons:
                tot[h] += int(r[col[h]] or 0)
            for h in ni:
                tot_ni[h] += int(r[col[h]] or 0)
            insts.append((n, r[col["Address"]][-5:], r[col["Source"]].strip(), {h: int(r[col[h]] or 0) for h in reasons if int(r[col[h]] or 0)}))
        s_all, s_ni = sum(tot.values()), sum(tot_ni.values())
        text += f"# {name.strip(',')}\n# warp-stall samples of one launch: {s_all} (all), {s_ni} (not issued)\n\n"
        text += "reason, all samples, share, not-issued samples, share\n"
        for h, v in tot.most_common():
            w = tot_ni.get(h + " (Not Issued)", 0)
            text += f"{h}, {v}, {v / max(s_all, 1):.3f}, {w}, {w / max(s_ni, 1):.3f}\n"
        text += "\nSASS opcode, samples, share, warp-instructions executed\n"
        for op, v in by_op.most_common(14):
            text += f"{op}, {v}, {v / max(s_all, 1):.3f}, {by_op_inst[op]}\n"
        text += "\ntop instructions by samples: samples, address, SASS, stall reasons\n"
        for n, addr, src, rs in sorted(insts, key=lambda x: -x[0])[:25]:
            text += f"{n}, {addr}, {src}, {rs}\n"
    open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:4])
