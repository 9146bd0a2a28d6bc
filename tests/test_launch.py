"""The `mpirun -n N` stand-in (``/root/reference/Makefile:2``): CLI launcher, spawn error propagation, clocks helper."""
import os
import subprocess
import sys
import textwrap

import pytest

from pytorch_ps_mpi_b200.launch import free_port, spawn
from pytorch_ps_mpi_b200.utils import ClockSampler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_launcher_sets_rank_env(tmp_path):
    script = tmp_path / "who.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        open(os.path.join(sys.argv[1], "rank%s.txt" % os.environ["RANK"]), "w").write(
            "rank %s of %s %s" % (os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["MASTER_ADDR"]))
    """))
    out = subprocess.run([sys.executable, "-m", "pytorch_ps_mpi_b200.launch", "-n", "3", str(script), str(tmp_path)],
                         cwd=ROOT, stdout=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0
    lines = [(tmp_path / f"rank{r}.txt").read_text() for r in range(3)]
    assert lines == [f"rank {r} of 3 127.0.0.1" for r in range(3)]


def _boom(rank, size):
    if rank == 1:
        raise ValueError("rank 1 exploded")


def test_spawn_propagates_rank_failures():
    with pytest.raises(RuntimeError) as e:
        spawn(_boom, 2, timeout=60)
    assert "rank 1 exploded" in str(e.value)


def _boom_while_peer_waits(rank, size):
    import time
    if rank == 1:
        raise ValueError("rank 1 exploded early")
    time.sleep(600)                      # a peer that would sit in a barrier / device spin for its whole time-out


def test_spawn_fails_fast_when_one_rank_dies():
    """One deadline for the group and no waiting for healthy-but-stuck peers (a failing rank used to cost nprocs x timeout —
    20 minutes of a 4-GPU box in round 2)."""
    import time
    t0 = time.time()
    with pytest.raises(RuntimeError) as e:
        spawn(_boom_while_peer_waits, 2, timeout=300)
    assert "rank 1 exploded early" in str(e.value)
    assert time.time() - t0 < 60


def test_nvml_clock_sampler_degrades_without_gpu():
    from pytorch_ps_mpi_b200.utils import NvmlClockSampler
    with NvmlClockSampler(0) as c:
        pass
    assert set(c.summary()) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}


def test_free_port_and_clock_sampler_without_gpu():
    assert 1024 < free_port() < 65536
    with ClockSampler(0) as c:       # no nvidia-smi / no GPU here: must degrade to an empty summary, not raise
        pass
    s = c.summary()
    assert set(s) >= {"sm_mhz", "sm_max_mhz", "reasons"}
