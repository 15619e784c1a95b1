"""Developer tool: SASS instruction count (and bytes) per kernel of the product library."""
import re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "cpuvox_b200/libcpuvox_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, cnt = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f_]+raybuffer_kernels_cu_[0-9a-f]+", "", m.group(1)); cnt[name] = 0; continue
    if name and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
        cnt[name] += 1
for n, c in cnt.items():
    print(f"{c:6d} instr {c * 16 / 1024:6.1f} KB  {n[:80]}")
