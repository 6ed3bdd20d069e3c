"""Per-function SASS statistics of the decode kernels in libllmf90_b200.so: instruction count, highest register,
local-memory loads / stores (LDL / STL: spills and call-convention saves -- an L2 round trip each in this kernel),
and an opcode histogram.  usage: python tools/sass_stats.py [lib.so] [kernel-substring, default ILi0ELb0]"""
import re, subprocess, sys, tempfile, os, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "llm/f90_b200/libllmf90_b200.so"
key = sys.argv[2] if len(sys.argv) > 2 else "ILi0ELb0"
lib = os.path.abspath(lib)
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {lib} > /dev/null 2>&1", shell=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("stream") or "stream" in f][0]
sass = subprocess.run(f"cuobjdump -sass {tmp}/{cubin}", shell=True, capture_output=True, text=True).stdout
i = sass.index("Function : _ZN6llmf9020stream_decode_kernel" + key)
j = sass.find("Function :", i + 10)
ins = []
for l in sass[i:j if j > 0 else None].splitlines():
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
syms = subprocess.run(f"readelf -sW {tmp}/{cubin} 2>/dev/null | grep FUNC | grep {key}", shell=True, capture_output=True, text=True).stdout
funcs = []
for l in syms.splitlines():
    f = l.split()
    try:
        funcs.append((int(f[1], 16), int(f[2], 0), f[7]))
    except Exception:
        pass
funcs.sort()
end = max(a for a, _ in ins) + 16
bounds = [(0, funcs[0][0] if funcs and funcs[0][0] > 0 else end, "kernel body")] + [(a, sz, n.split("$")[-1]) for a, sz, n in funcs if "$" in n]
tot_l = 0
for a, sz, name in bounds:
    sub = [t for (x, t) in ins if a <= x < a + sz]
    if not sub:
        continue
    c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in sub)
    mx = max([int(r) for t in sub for r in re.findall(r"\bR(\d+)\b", t)] + [0])
    name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:60]
    print(f"{name:62s} {len(sub):5d} instrs {len(sub) * 16 / 1024:5.1f} KB  maxR{mx:<3d} LDL {c['LDL']:3d} STL {c['STL']:3d} CALL {c['CALL']:2d}")
