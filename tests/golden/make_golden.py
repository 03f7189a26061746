"""Regenerates tests/golden/*.npz.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

1. step3_fixtures.npz   -- the reference's own five step-3 golden vectors
   (Code/CUDA/FinalProject/build/step3/{1..5}/{input,output}.txt; format step3.cu:59,79-81).
2. ref_steps_<case>.npz, ref_solve_<case>.npz -- outputs of the reference's OWN compiled
   seq_functions.cpp (oracle/_ref/libgpad_ref.so, built by oracle/Makefile) on seeded inputs,
   steps 1/2/4 in isolation and the 100-iteration loop in main.cu:160-175 order.  Operators are
   stored with the vectors so the fixtures do not depend on the host's BLAS rounding.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P            # noqa: E402
from oracle import RefLib, schedule  # noqa: E402

REF_FIX = "/root/reference/Code/CUDA/FinalProject/build/step3"
CASES = {"b3x4": (3, 4), "b4x3": (4, 3), "b10x15": (10, 15)}


def step3_fixtures():
    out = {}
    for k in range(1, 6):
        tok = open(f"{REF_FIX}/{k}/input.txt").read().split()
        n_u, N, m, theta = int(tok[0]), int(tok[1]), int(tok[2]), float(tok[3])
        n = n_u * N
        vals = np.array(tok[4:], np.float64)
        assert vals.size == 2 * n
        exp = np.array(open(f"{REF_FIX}/{k}/output.txt").read().split(), np.float64)
        assert exp.size == n
        out[f"dims{k}"] = np.array([n_u, N, m])
        out[f"theta{k}"] = np.array(theta)
        out[f"z_prev{k}"] = vals[:n]
        out[f"zhat{k}"] = vals[n:]
        out[f"z{k}"] = exp
    np.savez_compressed(os.path.join(HERE, "step3_fixtures.npz"), **out)


def ref_runs():
    ref = RefLib()
    theta, beta = schedule(100)
    for name, (n_u, N) in CASES.items():
        rng = np.random.default_rng(1234 + n_u * 100 + N)
        pb = P.battery(n_u, N)
        n, m = pb.n, pb.m
        x0 = rng.random(n_u) - 0.5
        g_P, p_D, f = pb.instance(x0)
        # isolated steps on random state
        y = np.maximum(rng.standard_normal(m), 0).astype(np.float32)
        y_prev = np.maximum(rng.standard_normal(m), 0).astype(np.float32)
        w = ref.step_one(y, y_prev, beta[7])
        zhat = ref.step_two(pb.M_G, w, g_P, n_u, N)
        y_next = ref.step_four(pb.G_L, w, p_D, zhat, n_u, N)
        np.savez_compressed(os.path.join(HERE, f"ref_steps_{name}.npz"), dims=np.array([n_u, N, m]),
                            M_G=pb.M_G, G_L=pb.G_L, g_P=g_P, p_D=p_D, y=y, y_prev=y_prev,
                            beta=beta[7], w=w, zhat=zhat, y_next=y_next)
        sol = ref.solve(n_u, N, m, pb.M_G, pb.G_L, g_P, p_D, theta, beta)
        np.savez_compressed(os.path.join(HERE, f"ref_solve_{name}.npz"), dims=np.array([n_u, N, m]),
                            L=np.float32(pb.L), x0=x0, f=f, M_G=pb.M_G, G_L=pb.G_L, g_P=g_P, p_D=p_D,
                            theta=theta, beta=beta,
                            **{k: sol[k] for k in ("y_next", "y", "z", "zhat", "w")})


if __name__ == "__main__":
    step3_fixtures()
    ref_runs()
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))
