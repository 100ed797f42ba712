"""tests/dropin/extract_tmain.py -- TEST INFRASTRUCTURE ONLY.
Writes the text of the reference's `_tmain` (MAIN:1115 to the end of the file: scenario set-up, the time loop
MAIN:1371-1418, shutdown) into tests/dropin/_gen/tmain.inc, VERBATIM, at build time and only where /root/reference
exists.  The generated file is git-ignored (no reference source is copied into the repository); ref_tmain_dropin.cpp
#includes it so that the reference's own main loop is compiled against include/lqro_dropin.hpp unmodified."""
import os
import sys

MAIN = "/root/reference/QuadrotorHoverController/LQRObstaclesPositionControl(Working).cpp"


def main(out):
    with open(MAIN, encoding="latin-1") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.lstrip().startswith("int _tmain("))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w", encoding="latin-1") as f:
        f.write("// GENERATED from %s lines %d-%d by tests/dropin/extract_tmain.py -- not part of the repository\n" % (MAIN, start + 1, len(lines)))
        f.writelines(lines[start:])
    return start + 1, len(lines)


if __name__ == "__main__":
    print(main(sys.argv[1]))
