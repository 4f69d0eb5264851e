"""CPU prototype of the batched RANSAC-P3P + Gauss-Newton pose solver of csrc/pnp.cu (development aid), checked here
against cv2.solvePnPRansac(EPNP) -- the call the reference makes (test_network_with_test_data.py:103-106)."""
import os, sys
import numpy as np
import cv2

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from checkerpose_b200 import synthetic as syn


def quartic_roots(A):            # A[0..4], real roots by Durand-Kerner on the monic polynomial
    a = np.array(A[:4], dtype=np.float64) / A[4]
    r = np.array([0.4 + 0.9j, -0.65 + 0.72j, 0.1 - 1.1j, -0.9 - 0.3j]) * (1 + np.max(np.abs(a)))
    for _ in range(60):
        for i in range(4):
            p = (((r[i] + a[3]) * r[i] + a[2]) * r[i] + a[1]) * r[i] + a[0]
            d = 1.0
            for j in range(4):
                if j != i:
                    d = d * (r[i] - r[j])
            r[i] = r[i] - p / d
    out = []
    for z in r:
        if abs(z.imag) < 1e-6 * (1 + abs(z.real)):
            x = z.real
            for _ in range(2):
                p = (((x + a[3]) * x + a[2]) * x + a[1]) * x + a[0]
                dp = ((4 * x + 3 * a[3]) * x + 2 * a[2]) * x + a[1]
                if dp != 0:
                    x -= p / dp
            out.append(x)
    return out


def p3p(X, f):
    """X (3,3) world points, f (3,3) unit bearings -> list of (R, t) with  s_i f_i = R X_i + t."""
    a2 = np.sum((X[1] - X[2]) ** 2); b2 = np.sum((X[0] - X[2]) ** 2); c2 = np.sum((X[0] - X[1]) ** 2)
    ca, cb, cg = f[1] @ f[2], f[0] @ f[2], f[0] @ f[1]
    q1, q2 = (a2 - c2) / b2, c2 / b2
    A0 = -4 * cg * cg * q1 - 4 * cg * cg * q2 + q1 * q1 + 2 * q1 + 1
    A1 = -4 * (-ca * cg * q1 - 2 * ca * cg * q2 + ca * cg - 2 * cb * cg * cg * q1 - 2 * cb * cg * cg * q2 + cb * q1 * q1 + cb * q1)
    A2 = 2 * (-2 * ca * ca * q2 + 2 * ca * ca - 4 * ca * cb * cg * q1 - 8 * ca * cb * cg * q2 + 2 * cb * cb * q1 * q1 - 2 * cg * cg * q1 - 2 * cg * cg * q2 + 2 * cg * cg + q1 * q1 - 1)
    A3 = -4 * (-2 * ca * ca * cb * q2 - ca * cg * q1 - 2 * ca * cg * q2 + ca * cg + cb * q1 * q1 - cb * q1)
    A4 = -4 * ca * ca * q2 + q1 * q1 - 2 * q1 + 1
    if abs(A4) < 1e-12:
        return []
    sols = []
    for v in quartic_roots([A0, A1, A2, A3, A4]):
        den = 2 * (cg - v * ca)
        if v <= 0 or abs(den) < 1e-12:
            continue
        u = (q1 * (1 + v * v - 2 * v * cb) - v * v + 1) / den
        w = 1 + v * v - 2 * v * cb
        if u <= 0 or w <= 0:
            continue
        s1 = np.sqrt(b2 / w)
        P = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])
        # absolute orientation from the two triangles (orthonormal frames)
        def frame(Q):
            e1 = Q[1] - Q[0]; e1 /= np.linalg.norm(e1)
            e3 = np.cross(e1, Q[2] - Q[0]); e3 /= np.linalg.norm(e3)
            return np.stack([e1, np.cross(e3, e1), e3], axis=1)
        R = frame(P) @ frame(X).T
        t = P[0] - R @ X[0]
        sols.append((R, t))
    return sols


def rodrigues(w):
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx


def refine(R, t, X, xn, fx, fy, iters=10):
    for _ in range(iters):
        Yr = X @ R.T
        Y = Yr + t
        iz = 1.0 / Y[:, 2]
        r = np.stack([fx * (Y[:, 0] * iz - xn[:, 0]), fy * (Y[:, 1] * iz - xn[:, 1])], 1)
        H = np.zeros((6, 6)); g = np.zeros(6)
        for k in range(len(X)):
            Jp = np.array([[fx * iz[k], 0, -fx * Y[k, 0] * iz[k] ** 2], [0, fy * iz[k], -fy * Y[k, 1] * iz[k] ** 2]])
            y = Yr[k]
            skew = np.array([[0, -y[2], y[1]], [y[2], 0, -y[0]], [-y[1], y[0], 0]])
            J = Jp @ np.concatenate([-skew, np.eye(3)], 1)
            H += J.T @ J; g += J.T @ r[k]
        d = np.linalg.solve(H + 1e-9 * np.eye(6), -g)
        R = rodrigues(d[:3]) @ R
        t = t + d[3:]
    return R, t


def solve(X, uv, K, thresh=2.0, iters=256, seed=0):
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xn = np.stack([(uv[:, 0] - cx) / fx, (uv[:, 1] - cy) / fy], 1)
    f = np.concatenate([xn, np.ones((len(xn), 1))], 1); f /= np.linalg.norm(f, axis=1, keepdims=True)
    rng = np.random.default_rng(seed)
    best = (-1, None, None)
    def inliers(R, t):
        Y = X @ R.T + t
        e = np.stack([fx * (Y[:, 0] / Y[:, 2] - xn[:, 0]), fy * (Y[:, 1] / Y[:, 2] - xn[:, 1])], 1)
        return (np.sum(e * e, 1) < thresh * thresh) & (Y[:, 2] > 0)
    for _ in range(iters):
        idx = rng.choice(len(X), 4, replace=False)
        cands = p3p(X[idx[:3]], f[idx[:3]])
        bestc, beste = None, 1e30
        for R, t in cands:
            Y = R @ X[idx[3]] + t
            if Y[2] <= 0:
                continue
            e = (fx * (Y[0] / Y[2] - xn[idx[3], 0])) ** 2 + (fy * (Y[1] / Y[2] - xn[idx[3], 1])) ** 2
            if e < beste:
                beste, bestc = e, (R, t)
        if bestc is None:
            continue
        n = int(inliers(*bestc).sum())
        if n > best[0]:
            best = (n, bestc[0], bestc[1])
    if best[1] is None:
        return np.eye(3), np.zeros(3), 0
    m = inliers(best[1], best[2])
    R, t = refine(best[1], best[2], X[m], xn[m], fx, fy)
    m = inliers(R, t)
    R, t = refine(R, t, X[m], xn[m], fx, fy, 5)
    return R, t, int(inliers(R, t).sum())


def scene(ds, obj, N, seed, outlier=0.2, invalid=0.1):
    rng = np.random.default_rng(seed)
    X = syn.load_fps_xyz(ds, obj, N).astype(np.float64)
    K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
    R, _ = cv2.Rodrigues(rng.normal(size=3) * 1.2)
    t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1200)])
    Y = X @ R.T + t
    uv = np.stack([K[0, 0] * Y[:, 0] / Y[:, 2] + K[0, 2], K[1, 1] * Y[:, 1] / Y[:, 2] + K[1, 2]], 1)
    lo, hi = uv.min(0), uv.max(0)
    side = float(np.ceil((hi - lo).max() * 1.2))
    c = (lo + hi) / 2
    bbox = np.array([np.floor(c[0] - side / 2), np.floor(c[1] - side / 2), side, side])
    xid = np.clip(np.floor((uv[:, 0] - bbox[0]) / (side / 64)), 0, 63).astype(np.int64)
    yid = np.clip(np.floor((uv[:, 1] - bbox[1]) / (side / 64)), 0, 63).astype(np.int64)
    bad = rng.random(N) < outlier
    xid[bad] = rng.integers(0, 64, bad.sum()); yid[bad] = rng.integers(0, 64, bad.sum())
    valid = rng.random(N) >= invalid
    return X, K, R, t, bbox, xid, yid, valid


def rot_err_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.degrees(np.arccos(np.clip(c, -1, 1))))


if __name__ == "__main__":
    for ds, obj, N in (("lmo", 1, 512), ("ycbv", 5, 512), ("lm", 9, 1024)):
        X, K, Rgt, tgt, bbox, xid, yid, valid = scene(ds, obj, N, 3 + obj)
        p2d = np.stack([bbox[0] + xid * bbox[2] / 64, bbox[1] + yid * bbox[3] / 64], 1)
        Xv, pv = X[valid], p2d[valid]
        ok, rvec, tcv, inl = cv2.solvePnPRansac(Xv, pv, K, None, reprojectionError=2, iterationsCount=150, flags=cv2.SOLVEPNP_EPNP)
        Rcv, _ = cv2.Rodrigues(rvec)
        R, t, n = solve(Xv, pv, K)
        print(f"{ds}/{obj} N={N}: valid {valid.sum()}  cv2 inliers {0 if inl is None else len(inl)}  ours {n};  "
              f"R err vs gt: cv2 {rot_err_deg(Rcv, Rgt):.3f} ours {rot_err_deg(R, Rgt):.3f} deg, ours vs cv2 {rot_err_deg(R, Rcv):.3f};  "
              f"t err vs gt: cv2 {np.linalg.norm(tcv.ravel() - tgt):.2f} ours {np.linalg.norm(t - tgt):.2f} mm, ours vs cv2 {np.linalg.norm(t - tcv.ravel()):.2f}")
