#!/usr/bin/env python
"""mvrun: single-node rank forker (there is no mpirun in this image).

    python tools/mvrun.py -n 4 [--port 41000] -- build/bin/mv_test matrix
    python tools/mvrun.py -n 2 -- python my_script.py

Every child gets MV_RANK / MV_SIZE / MV_PORT / MV_MASTER_ADDR (host runtime bootstrap) and
RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (torch.distributed bootstrap).
Exit code = first non-zero child exit code; all children are killed on timeout.
"""
import argparse
import os
import random
import subprocess
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=2)
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("--timeout", type=float, default=300)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    if not cmd:
        ap.error("no command")
    # below the kernel's ephemeral range (32768-60999): a rank's outgoing connection must never be
    # handed the port another rank is about to listen on
    port = a.port or random.randrange(10000, 30000, 16)
    procs = []
    for r in range(a.n):
        env = dict(os.environ, MV_RANK=str(r), MV_SIZE=str(a.n), MV_PORT=str(port),
                   MV_MASTER_ADDR="127.0.0.1", RANK=str(r), WORLD_SIZE=str(a.n), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port + 200))
        procs.append(subprocess.Popen(cmd, env=env))
    deadline = time.time() + a.timeout
    rc = 0
    while procs:
        for p in list(procs):
            r = p.poll()
            if r is not None:
                procs.remove(p)
                if r != 0 and rc == 0:
                    rc = r
        if time.time() > deadline or (rc != 0 and procs):
            time.sleep(1.0 if rc else 0)
            for p in procs:
                p.kill()
            rc = rc or 124
            break
        time.sleep(0.02)
    sys.exit(rc)


if __name__ == "__main__":
    main()
