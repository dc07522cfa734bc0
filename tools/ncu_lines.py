#!/usr/bin/env python
"""Top CUDA-C source lines of a profiled kernel by warp-stall samples.
Joins the per-SASS-instruction samples of an .ncu-rep (ncu --set full) with the line table of the cubin inside
libgigs_b200.so (nvdisasm -g; sources are compiled with -lineinfo).
usage: python tools/ncu_lines.py x.ncu-rep [topN] [launch_index]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gi-gs_b200", "lib", "libgigs_b200.so")


def sass_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    return blocks


def line_table(mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\s*\.text\.(\S*%s\S*):" % re.escape(mangled_hint), txt, re.M)
        if not m:
            continue
        body = txt[m.end():]
        end = re.search(r"^//-+ ", body, re.M)
        body = body[:end.start()] if end else body
        table, cur = {}, ("?", 0)
        for ln in body.splitlines():
            mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if mm:
                table[int(mm.group(1), 16)] = cur
        return table
    return {}


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    b = sass_rows(path)[which]
    name = b["name"]
    hint = re.sub(r"^(void )?(gigs::)?", "", name).split("(")[0].split("<")[0]
    table = line_table(hint)
    hdr = b["rows"][0]
    col = {h: i for i, h in enumerate(hdr)}
    si, ii = col["Warp Stall Sampling (All Samples)"], col["Instructions Executed"]
    base = None
    per_line = collections.defaultdict(lambda: [0, 0])
    tot = 0
    for r in b["rows"][1:]:
        try:
            addr = int(r[col["Address"]], 16); s = int(r[si]); n = int(r[ii])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = addr
        key = table.get(addr - base, ("?", 0))
        per_line[key][0] += s
        per_line[key][1] += n
        tot += s
    srcs = {}
    print(f"== {name[:110]}   total stall samples {tot}")
    by_inst = os.environ.get("BY_INST") is not None      # BY_INST=1: rank by executed instructions instead of stalls
    for (f, l), (s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][1 if by_inst else 0])[:top]:
        if f not in srcs:
            p = os.path.join(ROOT, "gi-gs_b200", "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
        print(f"{100.0 * s / max(tot, 1):5.1f}%  inst {n:>10d}  {f}:{l:<4d} {text[:110]}")


if __name__ == "__main__":
    main()
