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


def test_numa_binding_in_a_child_process():
    """``runtime.bind_to_gpu_numa_node``: intersect the allowed CPUs with NVML's GPU-local set, apply it to EVERY thread, refuse
    sets that are too small, never raise (run in a child: it changes the process's CPU affinity)."""
    code = textwrap.dedent("""
        import os, threading, time
        from pytorch_ps_mpi_b200 import runtime
        allowed = sorted(os.sched_getaffinity(0))
        half = allowed[: max(1, len(allowed) // 2)]

        class Nvml:
            def __init__(self, cpus): self.cpus = cpus
            def nvmlDeviceGetHandleByIndex(self, i): return i
            def nvmlDeviceGetCpuAffinity(self, h, n):
                words = [0] * n
                for c in self.cpus:
                    if c // 64 < n: words[c // 64] |= 1 << (c % 64)
                return words

        stop = threading.Event()
        seen = {}
        def other():
            while not stop.is_set(): time.sleep(0.01)
            seen["mask"] = sorted(os.sched_getaffinity(0))
        t = threading.Thread(target=other, daemon=True); t.start()
        r0 = runtime.bind_to_gpu_numa_node(0, min_cpus=len(half) + 1, nvml=Nvml(half))      # too few: refused
        assert not r0["bound"] and sorted(os.sched_getaffinity(0)) == allowed, r0
        r1 = runtime.bind_to_gpu_numa_node(0, min_cpus=1, nvml=Nvml(half + [max(allowed) + 1]))   # CPUs we may not use are ignored
        assert r1["bound"] and r1["cpus"] == len(half) and sorted(os.sched_getaffinity(0)) == half, r1
        stop.set(); t.join()
        assert seen["mask"] == half, seen            # the thread that already existed was moved too
        r2 = runtime.bind_to_gpu_numa_node(0, min_cpus=1, nvml=Nvml(half))
        assert r2["bound"] and r2["why"] == "already local", r2
        class Broken:
            def nvmlDeviceGetHandleByIndex(self, i): raise RuntimeError("no driver")
        r3 = runtime.bind_to_gpu_numa_node(0, nvml=Broken())
        assert not r3["bound"] and "no driver" in r3["why"], r3
        os.environ["PSB200_NUMA_BIND"] = "0"
        assert runtime.bind_to_gpu_numa_node(0, nvml=Nvml(half))["why"] == "disabled"
        print("numa-ok")
    """)
    if len(os.sched_getaffinity(0)) < 2:
        pytest.skip("one CPU")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and "numa-ok" in out.stdout, out.stderr[-2000:]
