"""Developer aid (CPU box): build a variant of the library that differs from the product only in lattice.cu
(extra -D flags), reusing the product's other objects.   python tools/build_variant.py NAME -DFOO=1 ...
-> pytorch_end2end_speech_recognition_b200/lib/libb200ctc_NAME.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_end2end_speech_recognition_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
b.build_library()
obj = os.path.join(b.LIB_DIR, "var_%s_lattice.o" % name)
subprocess.run([b._nvcc()] + b.NVCC_FLAGS + flags + ["-I", os.path.join(b.REPO_DIR, "include"), "-I", b.CSRC, "-c",
                os.path.join(b.CSRC, "lattice.cu"), "-o", obj], check=True)
objs = [obj if s == "lattice.cu" else os.path.join(b.LIB_DIR, s.replace(".cu", ".o")) for s in b.SOURCES]
out = os.path.join(b.LIB_DIR, "libb200ctc_%s.so" % name)
subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-Xcompiler", "-fPIC"], check=True)
os.remove(obj)
print(out)
