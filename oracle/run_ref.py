#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY: launch the prebuilt reference binaries under the MPI stand-in.

    python oracle/run_ref.py -np P [--threads T] -- oracle/_ref/nts_ref_driver cfg outdir dump 16

Spawns P processes with NTS_SHIM_SIZE / NTS_SHIM_RANK / NTS_SHIM_DIR set (see oracle/shim/mpi.h),
waits for all of them, returns the first non-zero exit status.  P=1 needs no scratch directory.
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile


def launch(nprocs, argv, threads=None, timeout=1800, quiet=False):
    env = dict(os.environ)
    if threads:
        env["NTS_THREADS"] = str(threads)
        env["OMP_NUM_THREADS"] = str(threads)
    scratch = None
    if nprocs > 1:
        scratch = tempfile.mkdtemp(prefix="nts_shim_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        env["NTS_SHIM_DIR"] = scratch
    procs = []
    try:
        for r in range(nprocs):
            e = dict(env)
            e["NTS_SHIM_SIZE"] = str(nprocs)
            e["NTS_SHIM_RANK"] = str(r)
            out = subprocess.DEVNULL if (quiet and r != 0) else None
            procs.append(subprocess.Popen(argv, env=e, stdout=out, stderr=None))
        rc = 0
        for p in procs:
            try:
                p.wait(timeout=timeout)
            except subprocess.TimeoutExpired:
                p.kill()
                rc = rc or 124
            rc = rc or p.returncode
        return rc
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        if scratch:
            shutil.rmtree(scratch, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-np", type=int, default=1)
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--timeout", type=int, default=1800)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    if not cmd:
        ap.error("missing command")
    sys.exit(launch(a.np, cmd, a.threads, a.timeout))


if __name__ == "__main__":
    main()
