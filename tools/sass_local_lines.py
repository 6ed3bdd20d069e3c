"""Source lines of the local-memory instructions (LDL / STL) of a decode kernel, per device function.
usage: python tools/sass_local_lines.py [kernel-key, default ILi0ELb0] [function-substring] [lib.so]"""
import re, subprocess, sys, tempfile, os
from collections import Counter
key = sys.argv[1] if len(sys.argv) > 1 else "ILi0ELb0"
only = sys.argv[2] if len(sys.argv) > 2 else ""
lib = os.path.abspath(sys.argv[3] if len(sys.argv) > 3 else "llm/f90_b200/libllmf90_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {lib} > /dev/null 2>&1", shell=True)
cubin = ([f for f in os.listdir(tmp) if f.startswith("stream")] or [f for f in os.listdir(tmp) if f.endswith(".cubin")])[0]
lines = subprocess.run(f"nvdisasm -g -c {tmp}/{cubin}", shell=True, capture_output=True, text=True).stdout.splitlines()
start = end = None
for i, l in enumerate(lines):
    if l.startswith(".text._ZN6llmf9020stream_decode_kernel" + key):
        start = i
    elif start is not None and l.startswith(".text."):
        end = i
        break
cur, fn, out = None, "kernel", []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2 = re.match(r"^\s*(\$?_Z\S+):", l)
    if m2 and "$" in m2.group(1):
        fn = subprocess.run(["c++filt", m2.group(1).split("$")[-1]], capture_output=True, text=True).stdout.strip()[:48]
    m3 = re.search(r"\b(LDL|STL)\b", l)
    if m3 and cur and only in fn:
        out.append((fn, cur[0], cur[1], m3.group(1)))
c = Counter(out)
src = open("llm/f90_b200/csrc/stream.cu").read().splitlines()
for (fn, f, ln, op), v in sorted(c.items()):
    text = src[ln - 1].strip()[:90] if f == "stream.cu" and ln <= len(src) else ""
    print(f"{fn:48s} {f}:{ln:<5d} {op} x{v}  {text}")
