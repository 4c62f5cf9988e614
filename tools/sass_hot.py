"""Aggregate an `ncu --page source --csv --print-source sass` dump: executed warp-instructions by opcode
for one kernel, and the hottest stall sites."""
import csv, sys, collections, re
path, want = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
blocks, cur = {}, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; blocks[cur] = []
    elif cur is not None:
        blocks[cur].append(r)
for name, blk in blocks.items():
    if want not in name: continue
    hdr = blk[0]; iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
    by = collections.Counter(); tot = 0; stalls = []
    for r in blk[1:]:
        if len(r) <= iE: continue
        try: n = int(r[iE])
        except ValueError: continue
        op = re.sub(r"^@!?U?P\w+\s+", "", r[iS].strip()).split()[0]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD", "LDG", "SHFL", "RED", "ISETP")) and "." in op else "")
        by[op] += n; tot += n
        stalls.append((int(r[iW] or 0), r[iS].strip(), n))
    print(name, "total executed", tot)
    for op, n in by.most_common(40): print(f"  {op:16s} {n:12d} {100*n/tot:5.1f}%")
    print("  top stall sites:")
    for w, src, n in sorted(stalls, reverse=True)[:25]: print(f"   {w:6d} samples  exec {n:9d}  {src}")
