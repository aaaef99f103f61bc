#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS view) per CUDA source line.

usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex-for-ncu> [top] [mangled-substring-for-nvdisasm]
Line numbers come from `nvdisasm -g` of the library's current cubin, matched to the ncu rows by instruction order,
so the report must have been captured from the same build."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
mangled = sys.argv[4] if len(sys.argv) > 4 else kern
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200", "lib", "libspsg_raycast.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in sorted(os.listdir(tmp)) if f.endswith(".cubin") and f.startswith("spsg_raycast.")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn, inl = [], None, False, ""
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = mangled in l
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); inl = m.group(3); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the CSV holds one table per profiled launch: "Kernel Name",<name> / header / rows ...; take the first whose name matches
want = sys.argv[5] if len(sys.argv) > 5 else kern
starts = [i for i, r in enumerate(rows) if len(r) >= 2 and r[0] == "Kernel Name"]
sel = next(i for i in starts if re.search(want, rows[i][1]))
end = next((j for j in starts if j > sel), len(rows))
hi = next(i for i in range(sel, end) if "Source" in rows[i] and "Address" in rows[i])
hdr = rows[hi]; body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
print("kernel:", rows[sel][1][:100])
iI, iS, iSrc, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
print("sass rows %d, disasm instrs %d" % (len(body), len(lines)))
agg = {}
tot = totS = 0
for k, r in enumerate(body):
    ln = lines[k] if k < len(lines) else -1
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += int(r[iI]); a[1] += int(r[iS]); a[2] += int(r[iT]); tot += int(r[iI]); totS += int(r[iS])
CSRC = os.path.join(ROOT, "self-supervised-scene-generation-with-semantic-segmentation_b200", "csrc")
src = {f: open(os.path.join(CSRC, f)).read().splitlines() for f in os.listdir(CSRC)}
print("total warp instr %d, samples %d" % (tot, totS))
bysamples = os.environ.get("BY_SAMPLES") == "1"
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if bysamples else 0])[:top]:
    f, n = ln if isinstance(ln, tuple) else ("?", -1)
    text = src[f][n - 1].strip()[:100] if f in src and 0 < n <= len(src[f]) else "?"
    print("%6.2f%% instr  %6.2f%% samples  thr/instr %4.1f  %s:%-4s %s" % (100.0 * a[0] / tot, 100.0 * a[1] / max(totS, 1), a[2] / max(a[0], 1), f.replace("spsg_", "").replace(".cuh", ""), n, text))
