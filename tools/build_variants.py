"""Build tuning variants of libsdirt_engine.so next to the product build (sdirt_b200/_lib/variants/<name>.so, git-ignored,
shipped to the GPU box by gpurun).  A variant is the same source with extra -D flags; a run selects it with
SDIRT_ENGINE_LIB=<path>.  Usage: python tools/build_variants.py name=-DFLAG=V[,-DFLAG2=V2] ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdirt_b200 import build as B


def main():
    out_dir = os.path.join(B.OUT_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    B.write_signatures()
    procs = []
    for arg in sys.argv[1:]:
        name, flags = arg.split("=", 1)
        out = os.path.join(out_dir, name + ".so")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + B.NVCC_FLAGS + flags.split(",") + ["-o", out, B.SRC]
        procs.append((name, out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, out, p in procs:
        log, _ = p.communicate()
        print(name, "->", out if p.returncode == 0 else "FAILED\n" + log)


if __name__ == "__main__":
    main()
