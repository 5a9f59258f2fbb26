"""Developer tool: registers / spills of the traversal and shade kernels from a build directory's ptxas log."""
import re
import sys

build = sys.argv[1] if len(sys.argv) > 1 else "/root/repo/raytracer-group27_b200/build"
s = open(build + "/csrc/rt_kernels.o.ptxas.log").read()
pat = r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers"
for m in re.finditer(pat, s, re.S):
    n = m.group(1)
    if any(k in n for k in sys.argv[2:] or ("k_extend", "k_shadow_", "k_shade", "k_trace")):
        print(f"{n[:64]:64s} stack {m.group(2):>4s} spill st {m.group(3):>4s} ld {m.group(4):>4s} regs {m.group(5)}")
