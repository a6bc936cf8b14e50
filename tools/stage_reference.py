"""Stage an UNMODIFIED copy of the reference's hot-path packages under the git-ignored `baseline/_ref/` so that the GPU box
(which has no /root/reference) can time the reference itself: `bench.py --impl reference`, the `cpu_baseline` leg and the
eager-GPU baseline run it through `baseline/ref_runner.py`.  Only Python sources and the two lens prescriptions are copied,
byte for byte; nothing under baseline/_ref/ is tracked by git (.gitignore) and nothing in the product imports it.
Usage: python tools/stage_reference.py [/root/reference]"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "deeplens")):
        return None
    n = 0
    for pkg in ("deeplens", "dfdp", "configs"):
        for dp, dn, fn in os.walk(os.path.join(src, pkg)):
            dn[:] = [d for d in dn if d != "__pycache__"]
            for f in fn:
                if not f.endswith((".py", ".yml", ".yaml", ".json")):
                    continue
                s = os.path.join(dp, f)
                d = os.path.join(DST, os.path.relpath(s, src))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                n += 1
    for lens in ("rf50mm", "rf35mm"):
        s = os.path.join(src, "lenses", lens, "lens_web.json")
        d = os.path.join(DST, "lenses", lens, "lens_web.json")
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
        n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(src + "\n")
    return n


if __name__ == "__main__":
    print("staged files:", stage(*(sys.argv[1:2])))
