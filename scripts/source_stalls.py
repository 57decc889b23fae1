"""Development aid: per-function warp-stall attribution of sqp_kernel<cfg 2> from an `ncu --set full --import-source on`
capture.  The report's SASS-level samples are mapped to source lines with the line table of the shipped library
(`cuobjdump -xelf` + `nvdisasm -g`) and summed per function of csrc/*.cuh.  Run in the build container:
    python scripts/source_stalls.py gpurun_out/prof_r01_full.ncu-rep > profiles/r01_source_stalls.md"""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
# optional: the library the capture ran and the source tree it was built from (defaults: the in-tree build and csrc/)
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "mpc_planner_b200", "lib", "libmpcgpu.so")
SRC = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "mpc_planner_b200", "csrc")
KEY = "sqp_kernelIN6mpcgpu6ConfigILi0ELi0ELi20ELi5ELi0ELi8ELi1ELi8ELi0ELi0E"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and KEY in l][0]
cur, table = None, {}
for l in dis[start + 1:]:
    if l.startswith(".text.") and table:
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        table[int(m.group(1), 16)] = cur


def functions_of(path):
    out = []
    for n, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r"(?:MPC_DEV|MPC_HD)\s+[\w:<>\s\*&]+?\s+(\w+)\(", l)
        if m:
            out.append((n, m.group(1)))
    return out


F = {f: functions_of(os.path.join(SRC, f)) for f in ("sqp_core.cuh", "stage_functions.cuh")}


def fn(key):
    if not key:
        return "?"
    name = "(other)"
    for n, f in F.get(key[0], []):
        if n <= key[1]:
            name = f
    return name if key[0] in F else key[0]


raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, iex, ib = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_barrier")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and i != ib]
base = int(data[0][ia], 16)
tot, nb, st, executed = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
iti, iws, iwi = hdr.index("Thread Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
winst, tinst, wsh, wshi = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    f = fn(table.get(int(r[ia], 16) - base))
    winst[f] += int(r[iex]); tinst[f] += int(r[iti]); wsh[f] += int(r[iws] or 0); wshi[f] += int(r[iwi] or 0)
    tot[f] += int(r[isamp])
    nb[f] += int(r[isamp]) - int(r[ib])
    executed[f] += int(r[iex]) > 0
    for i in cols:
        st[f][hdr[i][6:]] += int(r[i])
T, NB = sum(tot.values()), sum(nb.values())
mix = collections.Counter()
for f in st:
    mix.update(st[f])
print(f"# Warp-stall attribution of `sqp_kernel<cfg 2>` ({os.path.basename(rep)}; `scripts/source_stalls.py`)\n")
print(f"{T} samples, {NB} outside barriers (a warp waiting at `__syncthreads` for the serial sweep is not work).")
print("Non-barrier stall mix: " + ", ".join(f"{k} {100 * v / sum(mix.values()):.1f} %" for k, v in mix.most_common(8)) + ".")
print(f"Executed instruction footprint: {sum(executed.values()) * 16 / 1024:.0f} KB of {len(table) * 16 / 1024:.0f} KB.\n")
print("| function | share of non-barrier samples | top stall reasons | executed code |\n|---|---|---|---|")
for f, s in nb.most_common(16):
    top = ", ".join(f"{k} {100 * v // max(sum(st[f].values()), 1)} %" for k, v in st[f].most_common(4))
    print(f"| `{f}` | {100 * s / NB:.1f} % | {top} | {executed[f] * 16 / 1024:.1f} KB |")

WI, WS = sum(winst.values()), sum(wsh.values())
print(f"\nWarp instructions executed: {WI} ({sum(tinst.values()) / max(WI, 1):.1f} of 32 lanes active on average); shared-memory wavefronts: {WS} "
      f"({100 * (WS - sum(wshi.values())) / max(WS, 1):.0f} % above the conflict-free count).\n")
print("| function | warp instructions | share | active lanes | shared wavefronts | share | excess over ideal |\n|---|---|---|---|---|---|---|")
for f, n in winst.most_common(18):
    print(f"| `{f}` | {n} | {100 * n / WI:.1f} % | {tinst[f] / max(n, 1):.1f} | {wsh[f]} | {100 * wsh[f] / max(WS, 1):.1f} % | {100 * (wsh[f] - wshi[f]) / max(wsh[f], 1):.0f} % |")
