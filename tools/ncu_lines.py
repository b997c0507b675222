"""Per-source-line stall samples of one kernel in an .ncu-rep (developer tool, runs on the CPU box).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> sci_b200/lib/obj/lu.o [top]

ncu's CSV source page lists SASS only; nvdisasm -g gives the SASS -> source line mapping of the same object file.  The
two listings are matched by instruction order (the object must be the build that was profiled).
"""
import collections, csv, io, re, subprocess, sys, tempfile, os

rep, kre, obj = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
si, ns, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sass = []
for r in rows[2:]:
    if len(r) <= ns or not r[ns].isdigit():
        if len(r) == 2 and r[0] == "Kernel Name":
            break  # next kernel instance
        continue
    sass.append((r[si].strip(), int(r[ns]), int(r[ie]) if r[ie].isdigit() else 0, [int(r[i]) if r[i].isdigit() else 0 for i in stall_cols]))
# the mangled name: find the function in the object whose demangled template args match
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split per function
funcs = {}
cur = None
line = None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        cur = m.group(1); funcs[cur] = []; line = None; continue
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((m.group(1).strip(), line))
# choose function by instruction count match + name tokens
want = re.sub(r"[^A-Za-z0-9]", " ", kname).split()
best = None
for f, ins in funcs.items():
    if len(ins) and len(sass) % len(ins) == 0 and len(sass) // len(ins) <= 4:
        score = sum(1 for t in want if t in f)
        if best is None or score > best[0]:
            best = (score, f)
if best is None:
    print("no function with", len(sass), "instructions; candidates:", {f: len(i) for f, i in funcs.items()}); sys.exit(1)
ins = funcs[best[1]]
sass = sass[:len(ins)]  # ncu repeats the listing once per captured instance: keep the first
print("kernel:", kname[:120]); print("matched:", best[1], len(ins), "instructions")
tot = sum(s[1] for s in sass)
by_line = collections.defaultdict(lambda: [0, 0, [0] * len(stall_cols)])
for (txt, n, ex, st), (t2, ln) in zip(sass, ins):
    e = by_line[ln]
    e[0] += n; e[1] += ex
    for i, v in enumerate(st): e[2][i] += v
src_cache = {}
def src(ln):
    if not ln: return ""
    f, n = ln
    if f not in src_cache:
        for root in ("sci_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p): src_cache[f] = open(p).read().splitlines(); break
        else: src_cache[f] = []
    L = src_cache[f]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
print(f"total samples {tot}, total warp-instructions {sum(s[2] for s in sass)}")
names = [hdr[i].replace("stall_", "") for i in stall_cols]
for ln, (n, ex, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    top3 = sorted(zip(st, names), reverse=True)[:3]
    print(f"{n:6d} {100*n/max(tot,1):5.1f}% inst={ex:9d} {ln[0] if ln else '?'}:{ln[1] if ln else 0:4d} [{' '.join(f'{a}:{b}' for b,a in top3 if b)}] {src(ln)}")
