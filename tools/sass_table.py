"""Instruction-count table of the shipped library (cuobjdump -sass lib/libb200ctc.so): per kernel, how many
SASS instructions and how many of the mnemonics that show what the code is made of -- packed fp32
(FADD2/FMUL2/FFMA2), TMA bulk copies (UBLKCP) and their mbarriers (SYNCS), cp.async (LDGSTS), REDUX, named
barriers, shuffles, shared / global accesses, fp64 (safe lattice, beam search), and that no tensor-core
instruction (HMMA / UTCMMA / QMMA) is present, by design.   python tools/sass_table.py > profiles/rNN_sass_table.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pytorch_end2end_speech_recognition_b200", "lib", "libb200ctc.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["FADD2", "FMUL2", "FFMA2", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "BAR.SYNC", "SHFL", "LDS", "STS", "LDG", "STG",
        "RED.E", "ATOMG", "MUFU.EX2", "DADD", "DFMA", "BRA.DIV", "HMMA", "UTCMMA", "QMMA"]
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print("library:", os.path.relpath(lib, ROOT), " arch:", ", ".join(arch))
print("%-44s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in keys))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"b200ctc::|\(anonymous namespace\)::|void ", "", dem)
    dem = re.sub(r"\(int\)", "", dem)
    dem = dem[:dem.rfind("(")] if dem.endswith(")") else dem
    n = len(re.findall(r"/\*[0-9a-f]{4,5}\*/", f))
    cnt = [len(re.findall(r"\b" + re.escape(k) + r"\b", f)) for k in keys]
    print("%-44s %7d " % (dem[:44], n) + " ".join("%8d" % c for c in cnt))
