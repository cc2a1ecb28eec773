"""Golden vectors of the VO front-end from the REFERENCE's own wrapper node: runs oracle/_ref/vo_pin (the reference's
stereo-pub-node.cpp compiled unmodified against stand-in headers, `make -C oracle vo_pin`; needs /root/reference) on a seeded script
of tracked camera poses and stores what the node published on orb/vo and orb/pos.  -> tests/golden/vo_frontend_golden.npz"""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def script(seed, n=60, spread=0.3):
    """Camera mounting + a trajectory of tracked poses; `spread` large enough that the quaternion extraction leaves the
    trace > 0 branch for some frames."""
    rng = np.random.default_rng(seed)
    q_ic = rng.normal(size=4)
    q_ic /= np.linalg.norm(q_ic)
    R_ic, p_ic = rot(q_ic), np.array([0.12, -0.03, 0.05])
    qs = rng.normal(size=(n, 4)) * spread + np.array([1.0, 0, 0, 0])
    qs[n // 2:] = rng.normal(size=(n - n // 2, 4))  # arbitrary orientations: every branch of Quaterniond(Matrix3d)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    ts = np.cumsum(rng.normal(size=(n, 3)) * 0.02, axis=0)
    recv = 0.5 + np.arange(n) / 30.0 + rng.uniform(0, 0.004, size=n)  # node clock when the image pair is handled
    img = recv - 0.021                                               # image header stamp
    return R_ic, p_ic, recv, img, qs, ts


def run_vo_pin(R_ic, p_ic, recv, img, qs, ts):
    exe = os.path.join(ROOT, "oracle", "_ref", "vo_pin")
    lines = [" ".join(f"{v:.17g}" for v in list(R_ic.reshape(-1)) + list(p_ic)), str(len(recv))]
    lines += [" ".join(f"{v:.17g}" for v in [recv[k], img[k], *qs[k], *ts[k]]) for k in range(len(recv))]
    out = subprocess.run([exe], input="\n".join(lines) + "\n", text=True, capture_output=True, check=True).stdout.splitlines()
    pick = lambda tag: np.array([[float(v) for v in ln.split()[1:]] for ln in out if ln.startswith(tag + " ")])
    return pick("in"), pick("ref"), pick("ours")


if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "vo_pin"])
    R_ic, p_ic, recv, img, qs, ts = script(20241017)
    fin, ref, ours = run_vo_pin(R_ic, p_ic, recv, img, qs, ts)
    assert fin.shape == (60, 8) and ref.shape == (59, 13)
    np.savez_compressed(os.path.join(HERE, "vo_frontend_golden.npz"), R_ic=R_ic, p_ic=p_ic, frames=fin, ref=ref, image_stamp=img)
    print("frames", fin.shape, "published", ref.shape, "max |ref - ours|", np.abs(ref[:, :12] - ours).max())
