"""tools/sass_diff.py -- which kernels of two builds of libwoxel_b200.so differ in their SASS (addresses and encodings ignored).

  python tools/sass_diff.py OLD.so NEW.so

Used to show that a refactor or a new opt-in kernel leaves the kernels that were verified on the GPU untouched: identical SASS
means identical behaviour, so a build that cannot be re-run on a GPU can still be tied to one that was."""
import hashlib
import re
import subprocess
import sys


def kernels(so: str) -> dict:
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest(), len(v)) for k, v in d.items()}


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if k in b and a[k][0] == b[k][0]]
    print(f"{len(a)} kernels in {sys.argv[1]}, {len(b)} in {sys.argv[2]}: {len(same)} identical")
    for k in sorted(a):
        if k not in b:
            print("  removed:", k)
        elif a[k][0] != b[k][0]:
            print(f"  changed: {k}  ({a[k][1]} -> {b[k][1]} instructions)")
    for k in sorted(b):
        if k not in a:
            print(f"  new:     {k}  ({b[k][1]} instructions)")
    return 0 if all(k in b and a[k][0] == b[k][0] for k in a) else 1


if __name__ == "__main__":
    sys.exit(main())
