#!/usr/bin/env python
"""mpirun.py -np N prog [args...] -- launcher for programs linked against the mini-MPI
(include/compat/mpi.h): starts N processes on this host with P3DFFT_RANK / P3DFFT_NRANKS / P3DFFT_SESSION set
(and LOCAL_RANK so that each rank binds its own GPU).  Exit code = first non-zero child exit code."""
import os
import subprocess
import sys
import uuid


def main():
    a = sys.argv[1:]
    if len(a) < 3 or a[0] != "-np":
        sys.exit(__doc__)
    n = int(a[1])
    cmd = a[2:]
    session = uuid.uuid4().hex[:12]
    procs = []
    for r in range(n):
        env = dict(os.environ, P3DFFT_RANK=str(r), P3DFFT_NRANKS=str(n), P3DFFT_SESSION=session, LOCAL_RANK=str(r))
        env.pop("RANK", None)
        env.pop("WORLD_SIZE", None)
        procs.append(subprocess.Popen(cmd, env=env))
    rc = 0
    for p in procs:
        c = p.wait()
        if c and not rc:
            rc = c
    sys.exit(rc)


if __name__ == "__main__":
    main()
