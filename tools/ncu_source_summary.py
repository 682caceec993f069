"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: instructions by CUDA source line and by SASS opcode.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_source_summary.py src.csv [launch_index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# the dump repeats (file sections) per launch; a launch starts again when the first file path re-appears
sections, cur_file, launch, first_file = [], None, -1, None
hdr = None
by_line = collections.Counter()
by_line_thr = collections.Counter()
by_op = collections.Counter()
by_op_thr = collections.Counter()
stall_line = collections.Counter()
src_text = {}
seen = set()
seen_al = set()
tot_s = 0
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1]
        if first_file is None:
            first_file = cur_file
        if cur_file == first_file:
            launch += 1
        continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if launch != want_launch or hdr is None:
        continue
    if r[0] not in ("", "-"):
        cur_line = (cur_file.split("/")[-1], int(r[0]))
        src_text[cur_line] = r[1]
        continue
    sass = r[3].strip()
    if sass in ("...", "-", ""):
        continue
    try:
        n = int(r[i_inst]); t = int(r[i_thr]); s = int(r[i_samp]) if r[i_samp] not in ("-", "") else 0
    except ValueError:
        continue
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    if (r[2], cur_line) in seen_al:   # the cuda,sass view lists every row twice
        continue
    seen_al.add((r[2], cur_line))
    by_line[cur_line] += n; by_line_thr[cur_line] += t; stall_line[cur_line] += s
    if r[2] in seen:            # inlined code is listed under every frame of its inline stack: count an address once
        continue
    seen.add(r[2])
    by_op[op] += n; by_op_thr[op] += t; tot_s += s
tot = sum(by_op.values())
print(f"launch {want_launch}: {tot} warp instructions, {sum(by_op_thr.values())} thread instructions, {tot_s} stall samples")
print("-- by opcode (%% of warp instructions, avg threads)")
for op, n in by_op.most_common(28):
    print(f"  {op:10s} {100 * n / tot:5.1f}%  [{by_op_thr[op] / max(n, 1):4.1f}]")
print("-- by source line, INCLUSIVE of inlined callees (top 70): %inst  %stall-samples  avg-threads  text")
for ln, n in by_line.most_common(70):
    print(f"  {ln[0]}:{ln[1]:<4d} {100 * n / tot:5.1f}% {100 * stall_line[ln] / max(tot_s, 1):5.1f}% [{by_line_thr[ln] / max(n, 1):4.1f}] {src_text.get(ln, '')[:110]}")
