"""Numerical experiment (CPU, numpy): can the 36-entry fp64 Gram of the fit be replaced by an fp32 Gram plus an
iterative refinement whose residual is formed from the constraint ROWS (g = X^T (X f), 9 accumulators), at the
reference's own accuracy class?

For every synthetic scene (the oracle's generator, all weight modes) it reports the sign-aligned relative error of f
against the fp64 truth for: the reference's arithmetic (fp32 SVD of X), an fp32 Gram alone, and the fp32 Gram followed
by k refinement steps  f <- normalise(f - (G32 - mu I)^+ (g - rho f)),  g = X^T (X f) with fp32 rows / products and
either fp32 or fp64 accumulation of the 9 sums.  This is test-side research code: nothing here is on the product path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")]
import numpy as np
import torch
from fepe_b200 import synth
from oracle import fepe_oracle as O


def rows_fp32(d):
    """X [B,N,9] in fp32 exactly as the reference builds it (Hartley in fp32, normalised rows times w)."""
    m = torch.from_numpy(d["matches_xy_ori"])
    p1, p2, _ = O.norm_hw(m, d["image_size"])
    w = torch.from_numpy(d["weights"]).squeeze(1).unsqueeze(2)
    p1n, _ = O.hartley(p1)
    p2n, _ = O.hartley(p2)
    return (O.constraint_rows(p1n, p2n) * w).numpy().astype(np.float32)


def null_vec(G):
    lam, V = np.linalg.eigh(G)
    return V[:, 0], lam


def err(f, ft):
    f = f / np.linalg.norm(f)
    return min(np.linalg.norm(f - ft), np.linalg.norm(f + ft))


def gram32(X, chunks=64):
    """fp32 products and fp32 accumulation, `chunks` partial sums (lanes) combined at the end like the kernel would."""
    N = X.shape[0]
    acc = np.zeros((chunks, 9, 9), np.float32)
    for c in range(chunks):
        Xc = X[c::chunks]
        for i in range(Xc.shape[0]):
            acc[c] += np.outer(Xc[i], Xc[i]).astype(np.float32)
    return acc.astype(np.float64).sum(0)       # cross-lane combine in fp64 (36 values, once per pair)


def refine(X32, G32, f, steps, acc64):
    lam, V = np.linalg.eigh(G32)
    out = []
    for _ in range(steps):
        f32 = f.astype(np.float32)
        r = (X32 * f32).sum(1, dtype=np.float32)                       # r_i = x_i . f, fp32
        if acc64:
            g = (X32.astype(np.float64) * r.astype(np.float64)[:, None]).sum(0)
        else:
            g = (X32 * r[:, None]).astype(np.float32).sum(0, dtype=np.float32).astype(np.float64)
        rho = float(f @ g)
        res = g - rho * f
        # correction in the eigenbasis of the fp32 Gram, excluding its own smallest direction
        coef = V.T @ res
        shift = lam[0]
        delta = V[:, 1:] @ (coef[1:] / (lam[1:] - shift))
        f = f - delta
        f = f / np.linalg.norm(f)
        out.append(f.copy())
    return out


def f_to_F(f, T1, T2):
    """rank-2 projection + de-normalisation in fp64 (what the kernel does after the eigen-solve)."""
    U, S, Vt = np.linalg.svd(f.reshape(3, 3))
    S[2] = 0.0
    return T2.T @ (U @ np.diag(S) @ Vt) @ T1


def parity_against_reference():
    """The number the GPU tests assert: relative Frobenius distance (sign aligned) between OUR F and the reference's
    F (oracle = fp32 torch.svd path) -- for the fp64 Gram (today's kernel), the fp32 Gram alone and fp32 Gram + 1 step."""
    print("\nparity of F against the reference path (bar 1e-4): max over 12 scenes per case")
    print(f"{'mode':8s} {'N':>5s} | {'fp64 Gram':>10s} {'fp32 Gram':>10s} {'fp32 + 1 step':>13s}")
    for mode, N in [("uniform", 1000), ("softmax", 1000), ("inlier", 1000), ("peaked", 1000), ("inlier", 2000), ("inlier", 200)]:
        d = synth.make_batch(12, N, seed=11 + N, weight_mode=mode)
        m = torch.from_numpy(d["matches_xy_ori"])
        p1, p2, _ = O.norm_hw(m, d["image_size"])
        Fref, _ = O.fit_weighted_svd(p1, p2, torch.from_numpy(d["weights"]))
        _, T1 = O.hartley(p1)
        _, T2 = O.hartley(p2)
        X = rows_fp32(d)
        worst = [0.0, 0.0, 0.0]
        for b in range(12):
            Xb = X[b]
            t1, t2 = T1[b].numpy().astype(np.float64), T2[b].numpy().astype(np.float64)
            G64 = Xb.astype(np.float64).T @ Xb.astype(np.float64)
            G32 = gram32(Xb)
            f64, _ = null_vec(G64)
            f32, _ = null_vec(G32)
            f1 = refine(Xb, G32, f32, 1, acc64=False)[0]
            for k, f in enumerate((f64, f32, f1)):
                Fo = torch.from_numpy(f_to_F(f, t1, t2)).float().unsqueeze(0)
                worst[k] = max(worst[k], float(O.sign_aligned_rel_err(Fo, Fref[b:b + 1]).max()))
        print(f"{mode:8s} {N:5d} | {worst[0]:10.2e} {worst[1]:10.2e} {worst[2]:13.2e}")


def floors():
    """Which precision does g = X^T (X f0) need?  Scenes of tests/test_math_host.py (30 % outliers that the weights do
    NOT suppress in the softmax / uniform modes, so the residuals are large): worst error of f after one step with
    (fp32 r, fp32 products and sums), (fp32 r, exact products summed in fp64) and (everything fp64)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_math_host import scene_gram
    print("\nprecision of g (scenes of tests/test_math_host.py): worst of 12 scenes")
    print(f"{'mode':8s} | {'fp32 Gram':>10s} {'r32,g32':>10s} {'r32,g64':>10s} {'r64,g64':>10s}")
    for mode in ("softmax", "peaked", "inlier", "uniform"):
        worst = [0.0] * 4
        for seed in range(12):
            _, X = scene_gram(seed, mode, N=1000, noise=0.5 if seed % 2 else 0.1)
            X32 = X.astype(np.float32)
            X64 = X32.astype(np.float64)
            ft = np.linalg.eigh(X64.T @ X64)[1][:, 0]
            G32 = gram32(X32)
            l32, V32 = np.linalg.eigh(G32)
            f0 = V32[:, 0]

            def step(g):
                res = g - float(f0 @ g) * f0
                coef = V32.T @ res
                f = f0 - V32[:, 1:] @ (coef[1:] / (l32[1:] - l32[0]))
                return f / np.linalg.norm(f)
            r32 = (X32 * f0.astype(np.float32)).sum(1, dtype=np.float32)
            g_a = (X32 * r32[:, None]).astype(np.float32).sum(0, dtype=np.float32).astype(np.float64)
            g_b = (X64 * r32.astype(np.float64)[:, None]).sum(0)
            g_c = (X64 * (X64 @ f0)[:, None]).sum(0)
            es = [err(f0, ft), err(step(g_a), ft), err(step(g_b), ft), err(step(g_c), ft)]
            worst = [max(a, b) for a, b in zip(worst, es)]
        print(f"{mode:8s} | " + " ".join(f"{w:10.2e}" for w in worst))


def main():
    rng_cases = [("uniform", 1000), ("softmax", 1000), ("inlier", 1000), ("peaked", 1000), ("inlier", 2000), ("inlier", 200)]
    print(f"{'mode':8s} {'N':>5s} | {'ref fp32 SVD':>12s} {'fp32 Gram':>10s} | {'+1 (fp32 g)':>11s} {'+2 (fp32 g)':>11s} | {'+1 (fp64 g)':>11s} {'+2 (fp64 g)':>11s} | gap_rel")
    worst = {}
    for mode, N in rng_cases:
        d = synth.make_batch(12, N, seed=11 + N, weight_mode=mode)
        X = rows_fp32(d)
        for b in range(X.shape[0]):
            Xb = X[b]
            G64 = Xb.astype(np.float64).T @ Xb.astype(np.float64)
            ft, lam = null_vec(G64)
            gap = (lam[1] - lam[0]) / lam[-1]
            _, _, Vt = np.linalg.svd(Xb)                                # LAPACK sgesdd on fp32 rows = the reference's arithmetic
            e_ref = err(Vt[-1].astype(np.float64), ft)
            G32 = gram32(Xb)
            f0, _ = null_vec(G32)
            e0 = err(f0, ft)
            a = refine(Xb, G32, f0, 2, acc64=False)
            c = refine(Xb, G32, f0, 2, acc64=True)
            row = (e_ref, e0, err(a[0], ft), err(a[1], ft), err(c[0], ft), err(c[1], ft))
            key = (mode, N)
            worst[key] = tuple(max(x, y) for x, y in zip(worst.get(key, (0,) * 6), row))
            if b < 3:
                print(f"{mode:8s} {N:5d} | {row[0]:12.2e} {row[1]:10.2e} | {row[2]:11.2e} {row[3]:11.2e} | {row[4]:11.2e} {row[5]:11.2e} | {gap:.1e}")
    print("\nworst over 12 scenes per case:")
    for (mode, N), row in worst.items():
        print(f"{mode:8s} {N:5d} | {row[0]:12.2e} {row[1]:10.2e} | {row[2]:11.2e} {row[3]:11.2e} | {row[4]:11.2e} {row[5]:11.2e}")


if __name__ == "__main__":
    main()
    parity_against_reference()
    floors()
