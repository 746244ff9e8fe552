"""Build variant libraries of the fused SSD kernel under timeviper_b200/variants/ (git-ignored) for GPU sweeps:
each variant recompiles csrc/ssd_tc.cu with extra -D flags and links it with the objects of the default build.

    python tools/build_variants.py name1:-DTV_ABL=240 name2:-DFOO=1,-DBAR=2 ...
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from timeviper_b200 import build as B

B.build()
vdir = os.path.join(B.HERE, "variants"); os.makedirs(vdir, exist_ok=True)
for f in os.listdir(vdir):
    os.remove(os.path.join(vdir, f))
procs = []
for spec in sys.argv[1:]:
    name, flags = spec.split(":", 1)
    obj = os.path.join(B.HERE, "build", f"ssd_tc_{name}.o")
    cmd = ["/usr/local/cuda/bin/nvcc", *B.NVCC_FLAGS, *flags.split(","), "-c", os.path.join(B.CSRC, "ssd_tc.cu"), "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, pr in procs:
    out, _ = pr.communicate()
    if pr.returncode != 0:
        print(out); raise SystemExit(f"variant {name} failed")
    spills = [l.strip() for l in out.split("\n") if "spill" in l and " 0 bytes spill stores" not in l]
    objs = [os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "ssd_tc.cu"] + [obj]
    lib = os.path.join(vdir, f"{name}.so")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", lib, *objs, "-lcuda", "-L/usr/local/cuda/lib64/stubs"])
    print(name, "ok", ("SPILLS: " + "; ".join(spills[:2])) if spills else "")
