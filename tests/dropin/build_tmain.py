"""tests/dropin/build_tmain.py -- TEST INFRASTRUCTURE ONLY: builds tests/dropin/_gen/ref_tmain_dropin (see ref_tmain_dropin.cpp)."""
import os
import subprocess
import sys

TMAIN_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(TMAIN_DIR))
sys.path.insert(0, ROOT)
TMAIN_EXE = os.path.join(TMAIN_DIR, "_gen", "ref_tmain_dropin")
REFERENCE_TREE = "/root/reference"


def build_tmain_dropin(force=False):
    """tests/dropin/_gen/ref_tmain_dropin: the reference's own `_tmain` (extracted verbatim at build time, never committed)
    compiled against include/lqro_dropin.hpp and linked with liblqro.so -- it proves that the reference's main loop
    (MAIN:1371-1418) links unmodified against the shims.  Only where /root/reference exists; the binary travels to the GPU
    box with the snapshot (tests/dropin/_gen is git-ignored)."""
    from lqr_obstacles_b200 import build as product

    if not os.path.isdir(os.path.join(REFERENCE_TREE, "QuadrotorHoverController")):
        return None
    so = product.build()
    libdir = os.path.dirname(so)
    src = os.path.join(TMAIN_DIR, "ref_tmain_dropin.cpp")
    if not force and os.path.exists(TMAIN_EXE) and os.path.getmtime(TMAIN_EXE) > max(
            os.path.getmtime(src), os.path.getmtime(so), os.path.getmtime(os.path.join(ROOT, "include", "lqro_dropin.hpp"))):
        return TMAIN_EXE
    sys.path.insert(0, TMAIN_DIR)
    import extract_tmain
    gen = os.path.join(TMAIN_DIR, "_gen")
    extract_tmain.main(os.path.join(gen, "tmain.inc"))
    shims = os.path.join(ROOT, "oracle", "shims")
    flags = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++14", "-fpermissive", "-w", "-include", os.path.join(shims, "pre.h"),
             "-I" + shims, "-I" + os.path.join(REFERENCE_TREE, "include"), "-I" + os.path.join(REFERENCE_TREE, "QuadrotorHoverController"),
             "-I" + os.path.join(ROOT, "include"), "-I" + TMAIN_DIR]
    try:
        subprocess.check_call(["/usr/bin/g++"] + flags + ["-c", src, "-o", os.path.join(gen, "ref_tmain_dropin.o")])
    finally:
        os.remove(os.path.join(gen, "tmain.inc"))  # the extracted reference text is only needed for this compile
    subprocess.check_call(["/usr/bin/g++"] + flags + ["-c", os.path.join(REFERENCE_TREE, "QuadrotorHoverController", "gjk.cpp"), "-o", os.path.join(gen, "gjk.o")])
    subprocess.check_call(["/usr/bin/g++"] + flags + ["-c", os.path.join(shims, "callisto_stubs.cpp"), "-o", os.path.join(gen, "callisto_stubs.o")])
    subprocess.check_call(["/usr/bin/g++", "-o", TMAIN_EXE, os.path.join(gen, "ref_tmain_dropin.o"), os.path.join(gen, "gjk.o"),
                           os.path.join(gen, "callisto_stubs.o"), "-L" + libdir, "-llqro", "-Wl,-rpath," + libdir,
                           "-Wl,-rpath,$ORIGIN/../../../lqr_obstacles_b200"])
    return TMAIN_EXE


if __name__ == "__main__":
    print(build_tmain_dropin(force=True))
