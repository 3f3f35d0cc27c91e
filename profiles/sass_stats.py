"""Static SASS instruction mix per kernel of an object file: python profiles/sass_stats.py <file.o> [name filter]"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", f)
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:120]
    print(f"{len(ins):6d} instr {16 * len(ins) / 1024:6.1f} KB  {short}")
    if flt and flt in name:
        print("       ", collections.Counter(ins).most_common(24))
