"""SASS of one device function inside a decode kernel.  usage: python tools/sass_fn.py <kernel-key e.g. ILi1ELb0> <function-substring> [lib.so]"""
import re, subprocess, sys, tempfile, os
key, fn = sys.argv[1], sys.argv[2]
lib = os.path.abspath(sys.argv[3] if len(sys.argv) > 3 else "llm/f90_b200/libllmf90_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {lib} > /dev/null 2>&1", shell=True)
cubin = ([f for f in os.listdir(tmp) if f.startswith("stream")] or [f for f in os.listdir(tmp) if f.endswith(".cubin")])[0]
sass = subprocess.run(f"cuobjdump -sass {tmp}/{cubin}", shell=True, capture_output=True, text=True).stdout
i = sass.index("Function : _ZN6llmf9020stream_decode_kernel" + key)
j = sass.find("Function :", i + 10)
syms = subprocess.run(f"readelf -sW {tmp}/{cubin} 2>/dev/null | grep FUNC | grep {key}", shell=True, capture_output=True, text=True).stdout
lo = hi = None
for l in syms.splitlines():
    f = l.split()
    if "$" in f[7] and fn in f[7].split("$")[-1]:
        lo, hi = int(f[1], 16), int(f[1], 16) + int(f[2], 0)
        break
for l in sass[i:j if j > 0 else None].splitlines():
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m and lo <= int(m.group(1), 16) < hi:
        print(f"{int(m.group(1), 16) - lo:05x} {m.group(2)}")
