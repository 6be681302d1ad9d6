"""SASS evidence for profiles/: per kernel of libtloam_b200.so the instruction histogram, local-memory traffic
(LDL / STL = spills or stack arrays), and the instructions that prove the TMA path (UBLKCP + SYNCS mbarrier ops).

    python tools/sass_excerpt.py > profiles/r2_sass_excerpts.txt        # no GPU needed (cuobjdump)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tloam_b200", "libtloam_b200.so")
WANT = ["k_first", "k_eval", "k_correspond_dense", "k_correspond_fine", "k_correspond", "k_map_fine", "k_begin_frame", "k_ge_fit", "k_fe_sort", "k_fe_rank",
        "k_qbin_count", "k_os_seq", "k_os_union", "k_ee_section"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)\n")
    for b in blocks:
        name = b.split("\n", 1)[0].strip()
        short = next((w for w in WANT if re.search(r"\d+" + w + r"(I|E|P|N|$)", name) or ("5tloam" in name and w in name)), None)
        if short is None:
            continue
        ins = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b, flags=re.M)
        hist = collections.Counter(i.split(".")[0] for i in ins)
        full = collections.Counter(ins)
        print(f"## {name}")
        print(f"instructions: {len(ins)}   LDL: {hist.get('LDL', 0)}   STL: {hist.get('STL', 0)}   DFMA: {hist.get('DFMA', 0)}   "
              f"DADD: {hist.get('DADD', 0)}   DMUL: {hist.get('DMUL', 0)}   FFMA: {hist.get('FFMA', 0)}   LDG: {hist.get('LDG', 0)}   "
              f"LDS: {hist.get('LDS', 0)}   SHFL: {hist.get('SHFL', 0)}   BAR: {hist.get('BAR', 0)}")
        print("top mnemonics: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
        tma = [k for k in full if k.startswith(("UBLKCP", "SYNCS", "UTMA", "UCGABAR", "MEMBAR", "LDGSTS", "LDGDEPBAR", "DEPBAR", "ATOMS", "ATOMG", "VOTE", "MATCH"))]
        if tma:
            print("async-copy / barrier instructions: " + ", ".join(f"{k} x{full[k]}" for k in sorted(tma)))
            for line in b.split("\n"):
                if re.search(r"UBLKCP|SYNCS\.(ARRIVE|PHASECHK|EXCH)", line):
                    print("    " + line.strip()[:150])
        print()


if __name__ == "__main__":
    sys.exit(main())
