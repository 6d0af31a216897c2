"""The harness logic of the NFM / NGCF GPU tests (and the host logic of the two classes, dropout included) checked on the CPU:
tests/dryrun_on_oracle.py runs the GPU test bodies with oracle-backed stand-ins for the CUDA ops (own process: it patches torch)."""
import os
import subprocess
import sys


def test_gpu_test_bodies_pass_on_the_oracle_stand_in():
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "dryrun_on_oracle.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert r.stdout.count("DRYRUN OK") == 8, r.stdout
