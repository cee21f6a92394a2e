"""Per-kernel digest of liboea.so's SASS (addresses and encodings stripped): a host-side refactor or a moved helper must
leave the digests of the kernels it did not mean to touch unchanged.  `python scripts/sass_digest.py > new.txt` and diff
against profiles/r01_sass_digest.txt."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openea_b200", "_lib", "liboea.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if line.startswith("Fatbin elf code"):
        cur = None
    if cur is None or not line.strip() or re.match(r"^\s*/\*[0-9a-f]{4}\*/\s*$", line):
        continue
    funcs[cur].append(re.sub(r"/\*[0-9a-f]{16}\*/", "", line))
for name in sorted(funcs):
    print(hashlib.md5("\n".join(funcs[name]).encode()).hexdigest(), name)
