"""The reference's own C programs (unit_tests/*/problem.c, examples/*/problem.c), compiled UNMODIFIED against
include/ and linked with the product library -- the drop-in boundary seen from the reference's side
(reference unit_tests/Makefile.common:74-79).  TEST INFRASTRUCTURE.

build_all() needs /root/reference (this container); the binaries go to oracle/_ref/programs (git-ignored, travels to
the GPU box), where tests/test_gpu_ref_programs.py runs the data-independent ones on the synthetic ephemerides."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref", "programs")
# programs that do not build: none since the snapshot file (SimulationArchive surface) exists
NEED_ARCHIVE = set()


def sources():
    out = []
    for top in ("unit_tests", "examples"):
        base = os.path.join(REF, top)
        if not os.path.isdir(base):
            continue
        for name in sorted(os.listdir(base)):
            src = os.path.join(base, name, "problem.c")
            if os.path.exists(src):
                out.append((top + "_" + name if top == "examples" else name, src))
    return out


def build_all():
    """Returns {name: (ok, message)}."""
    os.makedirs(OUT, exist_ok=True)
    libdir = os.path.join(ROOT, "assist_b200")
    res = {}
    for name, src in sources():
        exe = os.path.join(OUT, name)
        cmd = ["gcc", "-std=c99", "-O3", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
               "-L" + libdir, "-lassist", "-lm", "-Wl,-rpath," + libdir, "-Wl,-rpath,$ORIGIN/../../../assist_b200"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        res[name] = (p.returncode == 0, (p.stderr or "").strip()[:400])
    return res
