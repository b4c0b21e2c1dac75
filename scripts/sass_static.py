#!/usr/bin/env python
"""Static facts about the shipped machine code (no GPU needed): per kernel the registers, stack, shared memory and the SASS
instruction mix, read with cuobjdump from lightmetrica-v2_b200/lib/liblmb200.so. Writes profiles/<tag>_sass_static.md.
Usage: python scripts/sass_static.py [tag]      (after ./build.sh)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "liblmb200.so")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

CLASSES = [
    ("global/const loads (LDG)", r"^LDG"), ("global stores / atomics (STG, ATOMG, RED)", r"^(STG|ATOMG|RED|ATOM)\b"),
    ("shared memory (LDS, STS, ATOMS)", r"^(LDS|STS|ATOMS|LDSM)"), ("local memory (LDL, STL)", r"^(LDL|STL)"),
    ("byte permute (PRMT)", r"^PRMT"), ("FP32 fused / mul / add (FFMA, FMUL, FADD)", r"^(FFMA|FMUL|FADD)"),
    ("FP32 min/max (FMNMX, FMNMX3)", r"^FMNMX"), ("FP32 compare / select (FSETP, FSEL, FSET)", r"^(FSETP|FSEL|FSET)\b"),
    ("warp votes / shuffles (VOTE, VOTEU, SHFL, MATCH, REDUX)", r"^(VOTE|VOTEU|SHFL|MATCH|REDUX)"),
    ("reciprocal etc. (MUFU)", r"^MUFU"), ("integer / logic (IADD3, IMAD, LOP3, SHF, LEA, ISETP, SEL, POPC, FLO, BREV ...)",
                                           r"^(IADD|IMAD|LOP3|SHF|LEA|ISETP|SEL|POPC|FLO|BREV|IABS|IMNMX|VIADD|VIMNMX|I2F|F2I|I2FP|F2FP|MOV|UMOV|UIADD|ULOP|USHF|ULEA|UISETP|UIMAD|USEL|UFLO|UPOPC|R2UR|S2R|S2UR|CS2R|R2P|P2R|PLOP3|UPLOP3|LDC|LDCU|UBREV|UPRMT|UFMUL|UFFMA)"),
    ("control (BRA, BSSY, BSYNC, CALL, RET, EXIT, WARPSYNC, BAR, NANOSLEEP, YIELD ...)", r"^(BRA|BRX|JMP|BSSY|BSYNC|BREAK|CALL|RET|EXIT|WARPSYNC|BAR|NANOSLEEP|YIELD|NOP|MEMBAR|ERRBAR|CCTL|DEPBAR|BMOV|ENDCOLLECTIVE|ACQBULK|UCGABAR)"),
]


def run(args):
    return subprocess.run(args, capture_output=True, text=True, check=True).stdout


def demangle(names):
    out = run(["c++filt"] + names).splitlines()
    return dict(zip(names, out))


def main():
    tool = "/usr/local/cuda/bin/cuobjdump"
    res = run([tool, "-res-usage", LIB])
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = run([tool, "-sass", LIB])
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        f = re.match(r"\s+Function : (\S+)", line)
        if f:
            cur = f.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    ours = [k for k in kernels if k.startswith("_ZN6lmb200")]
    names = demangle(ours)

    def short(k):
        n = names[k]
        n = n.replace("(anonymous namespace)::", "").replace("lmb200::", "")
        n = re.sub(r"\(.*", "", n).replace("void ", "")
        return n

    lines = ["# %s: static facts about the shipped SASS (`scripts/sass_static.py`, cuobjdump on `liblmb200.so`, sm_100a)" % TAG, "",
             "No GPU involved: this is what the compiler produced for the sources of this commit. Registers and stack per thread, static",
             "shared memory per block; `instr` = SASS instructions in the kernel (all paths, not a dynamic count).", "",
             "| kernel | regs | stack B | shared B | instr | LDG | LDG.256 | LDG.128 | STG/ATOMG | LDS/STS | LDL/STL | PRMT | FFMA+FMUL+FADD | FMNMX(3) | votes/shuffles |",
             "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]

    def count(ops, rx):
        r = re.compile(rx)
        return sum(1 for o in ops if r.match(o))

    order = sorted(ours, key=lambda k: (0 if "trace_kernel" in k else 1 if "trace_stream" in k else 2 if "k_extend" in k or "k_shadow" in k else 3 if "service" in k else 4 if "bvh_build" not in k else 5, short(k)))
    for k in order:
        ops = kernels[k]
        reg, stack, shared, local = usage.get(k, (0, 0, 0, 0))
        lines.append("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (
            short(k), reg, stack, shared, len(ops), count(ops, r"^LDG"), count(ops, r"^LDG.*\.256"), count(ops, r"^LDG.*\.128"),
            count(ops, r"^(STG|ATOMG|RED)"), count(ops, r"^(LDS|STS|ATOMS)"), count(ops, r"^(LDL|STL)"), count(ops, r"^PRMT"),
            count(ops, r"^(FFMA|FMUL|FADD)"), count(ops, r"^FMNMX"), count(ops, r"^(VOTE|VOTEU|SHFL|MATCH|REDUX)")))
    lines += ["", "What to read off it:", "",
              "* every traversal kernel (`trace_kernel`, `trace_stream_kernel`, `k_extend`, `k_shadow`) compiles to 56 registers with no stack and",
              "  no local-memory instruction: 9 blocks of 128 threads per SM (36 warps), the occupancy the ncu captures report;",
              "* the node fetch is the two `LDG.E.ENL2.256.CONSTANT` (64-byte unit = two 256-bit loads), the triangle record three",
              "  `LDG.E.128.CONSTANT`; the traversal stack and the octant table are the `LDS`/`STS`;",
              "* `k_nee` / `k_bsdf` / `k_logic` carry 32 bytes of stack (Philox block), `k_collapse` / `k_emit_dp` the builder's per-thread work lists;",
              "* the streaming gate costs `trace_stream_kernel` 256 bytes of shared memory and no register over `trace_kernel`.", ""]
    # instruction classes of the headline kernel
    head = [k for k in ours if "trace_kernelILb0ELb0ELb0E" in k][0]
    ops = kernels[head]
    lines += ["## `trace_kernel<false,false>`: instruction classes (static, %d instructions)" % len(ops), "", "| class | count |", "|---|---:|"]
    left = list(ops)
    for name, rx in CLASSES:
        r = re.compile(rx)
        n = sum(1 for o in left if r.match(o))
        left = [o for o in left if not r.match(o)]
        lines.append("| %s | %d |" % (name, n))
    lines.append("| other (%s) | %d |" % (", ".join(sorted(set(o.split(".")[0] for o in left))[:12]), len(left)))
    out = os.path.join(ROOT, "profiles", TAG + "_sass_static.md")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
