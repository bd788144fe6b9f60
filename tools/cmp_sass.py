"""Device-code regression check without a GPU: compile every .cu of two commits for sm_100a and compare the SASS of each
kernel (instruction streams, addresses and encodings stripped).  Used at the end of round 1, when host-side and additive
work continued after the GPU budget was spent, to show that the kernels verified on the B200 were not altered:
    python tools/cmp_sass.py <verified-commit> [HEAD]
Prints CHANGED / new / REMOVED kernels per source file."""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def checkout(commit, dst):
    os.makedirs(dst)
    tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "object_detection_tracking_b200/csrc", "include"],
                         check=True, capture_output=True).stdout
    subprocess.run(["tar", "-x", "-C", dst], input=tar, check=True)
    return os.path.join(dst, "object_detection_tracking_b200", "csrc")


def sass(csrc, f):
    if not os.path.exists(os.path.join(csrc, f)):
        return {}
    obj = os.path.join(csrc, f + ".o")
    subprocess.run(["nvcc", *FLAGS, "-c", os.path.join(csrc, f), "-o", obj], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    txt = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = part.split("\n", 1)
        name = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__X_", name.strip())
        out[name] = [m.group(1).strip() for m in (re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", l) for l in body.splitlines()) if m]
    return out


def main():
    base = sys.argv[1]
    head = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    with tempfile.TemporaryDirectory() as tmp:
        a, b = checkout(base, os.path.join(tmp, "a")), checkout(head, os.path.join(tmp, "b"))
        files = sorted(f for f in set(os.listdir(a)) | set(os.listdir(b)) if f.endswith(".cu"))
        changed = 0
        with ThreadPoolExecutor(8) as ex:
            for f, o, n in ex.map(lambda f: (f, sass(a, f), sass(b, f)), files):
                for k in sorted(set(o) | set(n)):
                    if k in o and k in n:
                        if o[k] != n[k]:
                            changed += 1
                            print("%-18s CHANGED  %s (%d -> %d instructions)" % (f, k[:100], len(o[k]), len(n[k])))
                    elif k in n:
                        print("%-18s new      %s (%d instructions)" % (f, k[:100], len(n[k])))
                    else:
                        twins = [j for j in n if j not in o and n[j] == o[k]]
                        if twins:      # e.g. a kernel that became a template instantiation
                            print("%-18s renamed  %s -> %s (instruction-identical)" % (f, k[:100], twins[0][:100]))
                        else:
                            changed += 1
                            print("%-18s REMOVED  %s" % (f, k[:100]))
                print("%-18s %d kernels in %s, %d in %s" % (f, len(o), base, len(n), head))
        print("kernels altered or removed: %d" % changed)
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
