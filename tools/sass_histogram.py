"""Opcode histogram of one kernel from the built objects (cuobjdump -sass): what the instruction counts quoted in
DESIGN.md / profiles can be re-derived from without a GPU.

    python tools/sass_histogram.py lag_corr_roll_kernel [--match "Lb1ELi16ELi2ELb0"] [--obj coreg_lag_roll.o]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "euispice_coreg_b200", "csrc", "build")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        yield name, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--match", default="", help="substring of the mangled name (template arguments)")
    ap.add_argument("--obj", default=None)
    args = ap.parse_args()
    objs = [os.path.join(BUILD, args.obj)] if args.obj else sorted(
        os.path.join(BUILD, f) for f in os.listdir(BUILD) if f.endswith(".o"))
    for obj in objs:
        for name, body in functions(obj):
            if args.kernel not in name or args.match not in name:
                continue
            ops = collections.Counter()
            for line in body:
                m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
                if m:
                    ops[m.group(1)] += 1
            total = sum(ops.values())
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            print(f"== {os.path.basename(obj)}: {demangled[:160]}")
            print(f"   {total} SASS instructions (static); FP64: "
                  f"{sum(v for k, v in ops.items() if k.split('.')[0] in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'))}, "
                  f"TMA / mbarrier: {sum(v for k, v in ops.items() if k.startswith(('UTMALDG', 'SYNCS', 'UBLKCP')))}")
            for k, v in ops.most_common(40):
                print(f"   {v:6d}  {k}")
            async_ops = {k: v for k, v in ops.items() if k.startswith(('UTMALDG', 'SYNCS', 'UBLKCP', 'UTMAPF', 'FENCE'))}
            if async_ops:
                print("   -- TMA / mbarrier / proxy-fence opcodes:")
                for k, v in sorted(async_ops.items()):
                    print(f"   {v:6d}  {k}")


if __name__ == "__main__":
    sys.exit(main())
