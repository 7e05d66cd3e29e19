"""tools/mvrun.py, the rank forker that stands in for mpirun: environment of the children, exit code
propagation, kill on timeout, ports below the kernel's ephemeral range."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MVRUN = os.path.join(ROOT, "tools", "mvrun.py")


def mvrun(n, code, *extra, timeout=60):
    return subprocess.run([sys.executable, MVRUN, "-n", str(n), *extra, "--", sys.executable, "-c", code],
                          capture_output=True, text=True, timeout=timeout)


def test_children_get_both_bootstrap_environments():
    r = mvrun(3, "import os, sys; e = os.environ; sys.stdout.write(' '.join([e['MV_RANK'], e['MV_SIZE'], e['RANK'], "
                 "e['WORLD_SIZE'], e['LOCAL_RANK'], e['MV_PORT'], e['MASTER_PORT'], e['MASTER_ADDR'], e['MV_MASTER_ADDR']]) + '\\n')")
    assert r.returncode == 0, r.stderr
    rows = sorted(line.split() for line in r.stdout.splitlines())
    assert [row[0] for row in rows] == ["0", "1", "2"]
    for row in rows:
        assert row[0] == row[2] == row[4] and row[1] == row[3] == "3"
        port = int(row[5])
        assert 10000 <= port and port + 200 + 64 + 8 < 32768       # every derived port stays below the ephemeral range
        assert int(row[6]) == port + 200 and row[7] == row[8] == "127.0.0.1"
    assert len({row[5] for row in rows}) == 1


def test_explicit_port_and_exit_code_propagation():
    r = mvrun(2, "import os, sys; sys.exit(7 if os.environ['MV_RANK'] == '1' else 0)", "--port", "12340")
    assert r.returncode == 7
    r = mvrun(2, "import os; print(os.environ['MV_PORT'])", "--port", "12340")
    assert r.returncode == 0 and r.stdout.count("12340") == 2      # (the two ranks' lines may interleave)


def test_failure_of_one_rank_stops_the_others():
    t0 = time.time()
    r = mvrun(2, "import os, sys, time\nif os.environ['MV_RANK'] == '0': sys.exit(3)\ntime.sleep(60)")
    assert r.returncode == 3 and time.time() - t0 < 30


def test_timeout_kills_all_ranks():
    t0 = time.time()
    r = mvrun(2, "import time; time.sleep(60)", "--timeout", "2")
    assert r.returncode == 124 and time.time() - t0 < 30
