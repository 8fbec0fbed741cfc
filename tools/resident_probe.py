#!/usr/bin/env python
"""Probe of the resident mode's managed-memory behaviour: time of a per-function call when its (managed) arguments
are device-resident vs just written by the CPU.  python tools/resident_probe.py [n_particles]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chimera_b200.fimera as f  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
f.resident(True)
rng = np.random.default_rng(0)
p = np.asfortranarray(rng.standard_normal((3, n)))
fld = np.asfortranarray(rng.standard_normal((6, n)))


def t(fn, rep=3):
    out = []
    for _ in range(rep):
        t0 = time.perf_counter()
        fn()
        out.append((time.perf_counter() - t0) * 1e3)
    return ["%.2f" % v for v in out]


state = {"p": p}


def call():
    state["p"] = f.push_velocs(state["p"], fld, 0.1)


print("bytes per call: p %.1f MB + fld %.1f MB" % (p.nbytes / 1e6, fld.nbytes / 1e6))
print("first calls (arguments written by the CPU at creation):", t(call))
print("device-resident:", t(call))


def touch_and_call():
    np.asarray(state["p"])[0, ::512] += 1e-9  # CPU touches one double per 4 KB page of p (migrates it to the host)
    call()


print("after the CPU touched every page of p:", t(touch_and_call))


def write_and_call():
    np.asarray(state["p"])[...] = 1.0  # CPU rewrites p completely
    call()


print("after the CPU rewrote p:", t(write_and_call))
print("host write of p alone:", t(lambda: np.asarray(state["p"]).__setitem__(Ellipsis, 1.0)))
