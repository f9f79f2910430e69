#!/usr/bin/env python3
"""Static instruction mix of the sm_100a kernels in libipt_b200.so (cuobjdump -sass), per kernel and per pipe class.
Static counts say what the compiler emitted, not what runs (loops, predication); the dynamic figures are the ncu
captures under profiles/.  Usage: python tools/sass_mix.py [kernel-substring ...] > profiles/rNN_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "improved-path-tracer_b200", "libipt_b200.so")

CLASSES = [
    ("fma pipe (FFMA FMUL FADD)", re.compile(r"^(FFMA|FMUL|FADD|FFMA2|FMUL2|FADD2)$")),
    ("fp64 (DFMA DMUL DADD DSETP)", re.compile(r"^D(FMA|MUL|ADD|SETP|MNMX)")),
    ("fp compare/select (FSETP FSEL FMNMX)", re.compile(r"^(FSETP|FSEL|FMNMX|FCHK|FSET)")),
    ("sfu (MUFU)", re.compile(r"^MUFU")),
    ("int mul (IMAD IMUL)", re.compile(r"^(IMAD|IMUL)")),
    ("alu int/logic (IADD3 LOP3 SHF LEA ISETP SEL ...)", re.compile(
        r"^(IADD|IADD3|LOP3|LOP|SHF|SHL|SHR|LEA|ISETP|SEL|PRMT|POPC|FLO|BREV|IABS|IMNMX|VIADD|VIMNMX|PLOP3|P2R|R2P|SGXT|BMSK|ICMP)")),
    ("convert (I2F F2I F2F I2I)", re.compile(r"^(I2F|F2I|F2F|I2I|F2FP|FRND|I2FP)")),
    ("move (MOV UMOV S2R CS2R ...)", re.compile(r"^(MOV|UMOV|S2R|S2UR|CS2R|R2UR|UR2R|MOVM)")),
    ("uniform datapath (U*)", re.compile(r"^U[A-Z]")),
    ("shared memory (LDS STS ATOMS LDSM)", re.compile(r"^(LDS|STS|ATOMS|LDSM)")),
    ("global/local memory (LDG STG LDL STL ATOMG RED LD ST)", re.compile(r"^(LDG|STG|LDL|STL|ATOMG|ATOM|RED|LD|ST|LDC|ULDC|LDCU|CCTL|MEMBAR|ERRBAR)")),
    ("warp collectives (SHFL VOTE MATCH REDUX)", re.compile(r"^(SHFL|VOTE|MATCH|REDUX|WARPSYNC|NANOSLEEP|ELECT)")),
    ("control (BRA BSSY BSYNC EXIT BAR CALL RET ...)", re.compile(r"^(BRA|BRX|BSSY|BSYNC|EXIT|BAR|CALL|RET|JMP|NOP|YIELD|BREAK|DEPBAR|BPT|KILL|ACQBULK|ENDCOLLECTIVE)")),
]


def classify(op):
    base = op.split(".")[0]
    for name, rx in CLASSES:
        if rx.match(base):
            return name
    return "other: " + base


def main():
    want = sys.argv[1:]
    sass = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE,
                           text=True, check=True).stdout.splitlines()
    bodies = re.split(r"\n\s*Function : \S+\n", "\n" + sass)[1:]
    inst = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    print(f"# static SASS mix of {os.path.relpath(SO, ROOT)} (cuobjdump -sass, sm_100a); see the docstring of tools/sass_mix.py")
    for name, body in zip(names, bodies):
        short = re.sub(r"^void ", "", name)
        short = re.sub(r"\(.*$", "", short)
        if want and not any(w in short for w in want):
            continue
        mix = collections.Counter()
        ops = collections.Counter()
        for line in body.splitlines():
            m = inst.match(line)
            if m:
                mix[classify(m.group(1))] += 1
                ops[m.group(1).split(".")[0]] += 1
        total = sum(mix.values())
        print(f"\n## {short}: {total} instructions")
        for cls, n in mix.most_common():
            print(f"  {n:6d}  {100.0 * n / total:5.1f} %  {cls}")
        print("  top opcodes: " + ", ".join(f"{o} {n}" for o, n in ops.most_common(12)))


if __name__ == "__main__":
    main()
