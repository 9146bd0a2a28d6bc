"""``__graft_entry__.smoke()`` — the function the driver calls on the GPU box before the bench — executed UNMODIFIED on the CPU
emulator (``torch.device("cuda", 0)`` is redirected to the host for the duration of the test): ResNet-18 (bf16, channels-last, device
engine, fused stem / BN / pool ops) forward + backward + two PS steps, then top-k + Adam on an MLP."""
import pytest
import torch

from tests import test_bench_dry_run_emulation as B
from tests import test_model_integration_emulation as MI
from tests import test_multirank_engine_emulation as H


def test_smoke_runs_on_the_emulator(bench_env, monkeypatch, capsys):
    _, extm = bench_env
    real_device = torch.device

    class Meta(type):                                      # keeps ``isinstance(x, torch.device)`` working
        def __instancecheck__(cls, inst):
            return isinstance(inst, real_device)

    # ``torch.device(...)`` whose requests (``torch.device("cuda", 0)`` in smoke()) land on the host
    HostDevice = Meta("device", (), {"__new__": lambda cls, *a, **k: real_device("cpu")})
    monkeypatch.setattr(torch, "device", HostDevice)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 1)
    cluster = H.Cluster(extm.emu, 1)
    H._tls.world, H._tls.m = H.World(cluster, 0), MI.ModelM(cluster, extm)
    import __graft_entry__ as g
    g.smoke()
    out = capsys.readouterr().out
    assert "smoke ok" in out and "engine=device" in out


bench_env = B.bench_env
