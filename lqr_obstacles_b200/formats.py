"""On-disk formats around the path, host side only.

* scenario file (input of tools/headless/lqro_headless --scenario file:PATH): text, first line N, then N lines
  "sx sy sz gx gy gz" -- start and goal positions in metres (the reference hard-codes these in _tmain, MAIN:1149-1337);
* trajectory dump (output of lqro_headless --out): little endian, 8 bytes magic "LQROTRJ1", int32 N, int32 steps, int32 16,
  int32 0, then per step N*16 doubles xTrue followed by N*16 doubles x (state layout [p v r w f], simulator2.h:4).  It stands
  in for the Callisto key frames, the reference's only trajectory record (MAIN:137-142);
* qhull's `n` / `Fv` text formats are written by tools/qconvex (MAIN:769-803 reads them).
"""
import numpy as np

TRAJ_MAGIC = b"LQROTRJ1"


def write_scenario(path, starts, goals):
    starts, goals = np.asarray(starts, dtype=np.float64), np.asarray(goals, dtype=np.float64)
    if starts.shape != goals.shape or starts.ndim != 2 or starts.shape[1] != 3 or len(starts) < 2:
        raise ValueError("starts and goals must both be (N, 3) with N >= 2")
    with open(path, "w") as f:
        f.write("%d\n" % len(starts))
        for s, g in zip(starts, goals):
            f.write(" ".join(repr(float(v)) for v in (*s, *g)) + "\n")


def read_scenario(path):
    with open(path) as f:
        tok = f.read().split()
    n = int(tok[0])
    v = np.array([float(t) for t in tok[1:1 + 6 * n]])
    if v.size != 6 * n:
        raise ValueError("scenario file %s is short: %d of %d numbers" % (path, v.size, 6 * n))
    v = v.reshape(n, 6)
    return v[:, :3].copy(), v[:, 3:].copy()


def write_trajectory(path, x_true, x_est):
    """x_true, x_est: (steps, N, 16)"""
    x_true, x_est = np.ascontiguousarray(x_true, dtype="<f8"), np.ascontiguousarray(x_est, dtype="<f8")
    if x_true.shape != x_est.shape or x_true.ndim != 3 or x_true.shape[2] != 16:
        raise ValueError("trajectories must both be (steps, N, 16)")
    steps, n, _ = x_true.shape
    with open(path, "wb") as f:
        f.write(TRAJ_MAGIC)
        f.write(np.array([n, steps, 16, 0], dtype="<i4").tobytes())
        f.write(np.stack([x_true, x_est], axis=1).tobytes())


def read_trajectory(path):
    """-> (x_true, x_est), each (steps, N, 16)"""
    raw = open(path, "rb").read()
    if raw[:8] != TRAJ_MAGIC:
        raise ValueError("%s is not a trajectory dump" % path)
    n, steps, sd, _ = (int(v) for v in np.frombuffer(raw[8:24], dtype="<i4"))
    if sd != 16 or len(raw) != 24 + 8 * 2 * steps * n * 16:
        raise ValueError("%s: header (N=%d, steps=%d, %d doubles per state) does not match the file size" % (path, n, steps, sd))
    tr = np.frombuffer(raw[24:], dtype="<f8").reshape(steps, 2, n, 16)
    return tr[:, 0].copy(), tr[:, 1].copy()
