"""Opcode summary of every kernel in libmassiv_b200.so, from `cuobjdump -sass` (no GPU needed):

    python scripts/sass_summary.py > profiles/sass_r02.txt

One line per kernel with the counts that make the design claims auditable without the binary: TMA (UTMALDG), bulk
copies (UBLKCP), cp.async (LDGSTS), mbarrier (SYNCS), fused multiply-adds (FFMA / DFMA — must be 0 in every stencil
kernel except the exact-division helper and the tap-with-zero-coefficient trick of the known-kernel path), unfused
FP32 / FP64 multiplies and adds, shared-memory traffic (LDS / STS) and local-memory spills (LDL / STL)."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "massiv_b200", "libmassiv_b200.so")
OPS = ["UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "BAR", "FFMA", "DFMA", "FMUL", "FADD", "DMUL", "DADD", "IMAD", "LDS", "STS", "LDG", "STG",
       "LDL", "STL", "SHFL", "MUFU"]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.basename(lib)}: {len(kernels)} kernels; columns: total " + " ".join(OPS))
    for (mangled, cnt), name in sorted(zip(kernels.items(), names), key=lambda t: t[1]):
        name = re.sub(r"\(CUtensorMap_st.*|\(mb::.*|\(unsigned.*|\(const .*", "", name.replace("mb::", "")).replace("void ", "")
        print(f"{name:110s} {cnt['_total']:6d} " + " ".join(f"{op}={cnt[op]}" for op in OPS if cnt[op]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else LIB)
