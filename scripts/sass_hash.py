"""Per-kernel hash of the SASS in a built library: `python scripts/sass_hash.py lib.so > a.txt`, rebuild, run again and
diff the two listings to prove that a change left the code of the existing kernels untouched (used when a new variant
is added behind a switch and the default path must not move)."""
import hashlib
import re
import subprocess
import sys


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, lines, rows = None, [], []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, lines))
            name, lines = m.group(1), []
        elif name and re.match(r"\s*(/\*[0-9a-f]{4}\*/|\.)", line):  # instructions and labels only (not the next module's header)
            lines.append(line.strip())
    if name:
        rows.append((name, lines))
    demangled = subprocess.run(["c++filt"], input="\n".join(n for n, _ in rows), capture_output=True, text=True).stdout.splitlines()
    for (n, body), d in sorted(zip(rows, demangled), key=lambda r: r[1]):
        # label numbers are per module: rename them by order of appearance inside the function
        names = {}
        body = [re.sub(r"\.L_x_\d+", lambda m: names.setdefault(m.group(0), ".L%d" % len(names)), b) for b in body]
        n_instr = sum(1 for b in body if re.match(r"/\*[0-9a-f]{4}\*/", b))
        print(hashlib.sha1("\n".join(body).encode()).hexdigest()[:12], "%6d" % n_instr, d[:200])


if __name__ == "__main__":
    main(sys.argv[1])
