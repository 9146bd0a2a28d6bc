"""BASELINE config 1: 2-layer MLP on MNIST-shaped synthetic data, world_size 2, CPU plumbing.

Every mode is checked against a single-process oracle (sum of all ranks' coded gradients →
``torch.optim``), and all ranks must end bit-identical.
"""
import pytest

from pytorch_ps_mpi_b200.launch import spawn
from tests import _mp


@pytest.mark.parametrize("mode,optim,coding,transport", [
    ("ps", "sgd", "identity", "shm"),
    ("ps", "adam", "cast", "shm"),
    ("ps", "sgd", "topk", "gloo"),
    ("allgather", "sgd", "identity", "shm"),
    ("allgather", "adam", "scale", "gloo"),
    ("allgather", "sgd", "topk", "shm"),
    ("ps", "sgd", "svd", "shm"),          # a host-path-only (user-style) coding through the generic wire format
])
def test_mlp_sync(mode, optim, coding, transport):
    spawn(_mp.mlp_train, 2, (mode, optim, coding, transport))


@pytest.mark.parametrize("mode", ["ps", "allgather"])
def test_mlp_sync_coalesced(mode):
    """coalesce=True: one message per step instead of one collective per parameter — same numerics."""
    spawn(_mp.mlp_train, 2, (mode, "sgd", "cast", "shm", True))


def test_mlp_sync_three_ranks():
    spawn(_mp.mlp_train, 3, ("ps", "sgd", "identity", "shm"))


@pytest.mark.parametrize("transport", ["shm", "gloo"])
def test_mlp_async(transport):
    spawn(_mp.mlp_async, 3, (transport,))


def test_shm_transport_detects_dead_peer():
    from pytorch_ps_mpi_b200.launch import spawn as _spawn
    import multiprocessing
    # rank 1 exits with os._exit(0) on purpose; rank 0 must notice instead of hanging
    _spawn(_mp.shm_dead_peer, 2, timeout=90)


@pytest.mark.parametrize("transport", ["shm", "gloo"])
def test_average_and_param_groups(transport):
    spawn(_mp.mlp_average_and_groups, 3, (transport,))


def test_async_consistent_reads():
    spawn(_mp.mlp_async_consistent, 3)


@pytest.mark.parametrize("n", [2, 3])
def test_reference_equivalent_comparator(n):
    """The comparator bench.py runs in the same invocation as the product arm: right numerics, no hang, on gloo."""
    spawn(_mp.comparator_host, n, env={"PSB200_TRANSPORT": "gloo"}, timeout=180)
