"""Static resource usage (registers, stack, shared, local) of every kernel in libdzb.so via ``cuobjdump -res-usage``;
writes profiles/r1b_static_resource_usage.txt.  Kernels with stack or local bytes index local arrays or spill."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    so = os.path.join(ROOT, "discretize_b200", "libdzb.so")
    text = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
    rows = []
    name = None
    for line in text.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((name,) + tuple(int(x) for x in m.groups()))
            name = None
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = ["# cuobjdump -res-usage discretize_b200/libdzb.so (sm_100a): registers / stack B / static shared B / local B per kernel",
           f"# {len(rows)} kernels; {sum(1 for r in rows if r[2] or r[4])} with a stack frame or local memory (listed first)", ""]
    table = sorted(zip(names, rows), key=lambda t: (-(t[1][2] + t[1][4]), t[0]))
    for nm, (_, reg, stack, shared, local) in table:
        out.append(f"reg {reg:>3}  stack {stack:>4}  smem {shared:>6}  local {local:>4}  {nm[:150]}")
    path = os.path.join(ROOT, "profiles", "r1b_static_resource_usage.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(out[1])
    for line in out[3:15]:
        print(line)


if __name__ == "__main__":
    main()
